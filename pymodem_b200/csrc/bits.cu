// bits.cu -- bit-level stages over packed bitstreams (integer kernels).
//
//   gather : (sign, rollover mask) -> packed symbol bitstream + per-byte sample
//            address  == the AddressedData list of slicer.py:92-97 / 209-224
//   lfsr   : lfsr.py:22-52 as a GF(2) FIR over 32-bit words
//   ax25   : ax25.py:25-93 -- stateless flag detection, then one thread per
//            inter-flag gap replays the HDLC machine exactly
//   crc    : crc_functions.py:9-61, packet_meta.py:21-41
//
// Stream bit g lives in word g>>5 at bit g&31 (LSB-first), so "earlier" bits
// are at lower positions; the reference's bytes are MSB-first groups of 8.
#include <algorithm>
#include "pm_common.cuh"

#define GB_WORDS 1024          // mask words per gather block (32768 samples)
#define GB_THREADS 256



__device__ __forceinline__ unsigned int warp_incl_scan(unsigned int v)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned int t = __shfl_up_sync(0xffffffffu, v, d);
		if ((threadIdx.x & 31) >= d) v += t;
	}
	return v;
}

// exclusive scan over the block (blockDim.x multiple of 32, <= 1024); returns
// the exclusive prefix of v and the block total.
__device__ __forceinline__ unsigned int block_excl_scan(unsigned int v, unsigned int *s_warp, unsigned int &total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	unsigned int inc = warp_incl_scan(v);
	if (lane == 31) s_warp[wid] = inc;
	__syncthreads();
	if (wid == 0) {
		unsigned int t = (lane < nw) ? s_warp[lane] : 0;
		unsigned int ti = warp_incl_scan(t);
		s_warp[lane] = ti - t;
		if (lane == 31) s_warp[32] = ti;
	}
	__syncthreads();
	unsigned int ex = inc - v + s_warp[wid];
	total = s_warp[32];
	__syncthreads();
	return ex;
}

__device__ __forceinline__ uint32_t valid_mask_word(uint32_t m, long long w, long long nout)
{
	const long long first = w << 5;
	if (first >= nout) return 0;
	const long long remain = nout - first;
	return remain >= 32 ? m : (m & ((1u << (int)remain) - 1u));
}

// --- gather step 1: symbols per block of GB_WORDS mask words -----------------
__global__ void __launch_bounds__(GB_THREADS)
gather_count_kernel(const BitChain *__restrict__ chains, const uint32_t *__restrict__ mask, long long mask_stride,
                    unsigned int *__restrict__ block_count, int n_blocks, long long w_origin)
{
	__shared__ unsigned int s_warp[33];
	const int ch = blockIdx.y;
	const long long nout = chains[ch].nout;
	const uint32_t *mk = mask + (long long)ch * mask_stride;
	const long long w0 = w_origin + (long long)blockIdx.x * GB_WORDS + threadIdx.x * 4;
	unsigned int cnt = 0;
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const long long w = w0 + q;
		if ((w << 5) < nout) cnt += __popc(valid_mask_word(mk[w], w, nout));
	}
	unsigned int total;
	block_excl_scan(cnt, s_warp, total);
	if (threadIdx.x == 0) block_count[(long long)ch * n_blocks + blockIdx.x] = total;
}

// --- generic per-chain exclusive scan of a row of counters (one CTA per row) --
__global__ void __launch_bounds__(1024)
row_scan_kernel(const unsigned int *__restrict__ in, unsigned int *__restrict__ out, int n, int row_stride,
                unsigned int *__restrict__ totals)
{
	__shared__ unsigned int s_warp[33];
	const unsigned int *src = in + (long long)blockIdx.x * row_stride;
	unsigned int *dst = out + (long long)blockIdx.x * row_stride;
	unsigned int carry = 0;
	for (int base = 0; base < n; base += 1024) {
		const int i = base + threadIdx.x;
		unsigned int v = (i < n) ? src[i] : 0;
		unsigned int total;
		unsigned int ex = block_excl_scan(v, s_warp, total);
		if (i < n) dst[i] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// --- gather step 3: write bits + byte addresses ------------------------------
// symbol_base: symbols emitted before this shard (byte alignment is global,
// slicer.py:92-97); init_state: quadrature state_register before the shard.
__global__ void __launch_bounds__(GB_THREADS)
gather_write_kernel(const BitChain *__restrict__ chains, const uint32_t *__restrict__ sign, long long sign_stride,
                    const uint32_t *__restrict__ mask, long long mask_stride,
                    const unsigned int *__restrict__ block_base, int n_blocks,
                    uint32_t *__restrict__ bits, long long bits_stride,
                    uint32_t *__restrict__ byte_addr, long long addr_stride,
                    const unsigned int *__restrict__ init_state, long long w_origin,
                    const ShardBits *__restrict__ sb)
{
	__shared__ unsigned int s_warp[33];
	__shared__ uint32_t s_stage[2 * GB_WORDS + 2];
	const int ch = blockIdx.y;
	const BitChain C = chains[ch];
	const uint32_t *mk = mask + (long long)ch * mask_stride;
	const uint32_t *sg = sign + (long long)C.sign_row * sign_stride;
	const uint32_t *sq = C.quadrature ? sign + (long long)C.sign_q_row * sign_stride : nullptr;
	uint32_t *out = bits + (long long)ch * bits_stride;
	uint32_t *oaddr = byte_addr + (long long)ch * addr_stride;
	const long long w0 = w_origin + (long long)blockIdx.x * GB_WORDS + threadIdx.x * 4;
	const long long bit_off = sb[ch].bit_off;

	// everything the block needs from global memory is requested up front, in one round trip: the four mask words and
	// the four sign words of the thread as 128-bit loads (rows and the thread's first word are 16-byte aligned), and the
	// block's symbol base -- the scan and the scatter below then run out of registers
	uint32_t m[4], sv[4] = {0u, 0u, 0u, 0u}, qv[4] = {0u, 0u, 0u, 0u};
	unsigned int cnt = 0;
	const bool vec = ((w0 & 3) == 0) && (((w0 + 3) << 5) < C.nout) && ((mask_stride & 3) == 0) && ((sign_stride & 3) == 0);
	if (vec) {
		const uint4 mv = *reinterpret_cast<const uint4 *>(mk + w0);
		const uint4 s4 = *reinterpret_cast<const uint4 *>(sg + w0);
		m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
		sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w;
		if (sq) {
			const uint4 q4 = *reinterpret_cast<const uint4 *>(sq + w0);
			qv[0] = q4.x; qv[1] = q4.y; qv[2] = q4.z; qv[3] = q4.w;
		}
#pragma unroll
		for (int q = 0; q < 4; q++) {
			m[q] = valid_mask_word(m[q], w0 + q, C.nout);
			cnt += __popc(m[q]);
		}
	} else {
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const long long w = w0 + q;
			const bool in = (w << 5) < C.nout;
			m[q] = in ? valid_mask_word(mk[w], w, C.nout) : 0;
			if (in) { sv[q] = sg[w]; if (sq) qv[q] = sq[w]; }
			cnt += __popc(m[q]);
		}
	}
	const unsigned int sym_block = block_base[(long long)ch * n_blocks + blockIdx.x];   // symbols before block
	unsigned int total;
	const unsigned int ex = block_excl_scan(cnt, s_warp, total);
	const long long bit_block = (long long)sym_block * C.bps + bit_off;                  // stream bits before block
	const unsigned int nbits_block = total * C.bps;
	const int stage_shift = (int)(bit_block & 31);
	const unsigned int stage_words = (stage_shift + nbits_block + 31) >> 5;
	for (unsigned int i = threadIdx.x; i < stage_words; i += GB_THREADS) s_stage[i] = 0;
	__syncthreads();

	if (!C.quadrature && C.bps == 1) {
		// binary slicer (uniform per block): one stream bit per symbol, everything in 32-bit arithmetic relative to the
		// block's first symbol -- stream bit of block-local symbol k = bit_block + k
		if (cnt) {
			unsigned int k = ex;                                         // block-local index of this thread's next symbol
			const unsigned int gl = (unsigned int)(bit_block & 7);        // where block-local symbol 0 sits in its stream byte
			uint32_t *oa = oaddr + (bit_block >> 3);
#pragma unroll
			for (int q = 0; q < 4; q++) {
				uint32_t mm = m[q];
				const uint32_t s = sv[q];
				const uint32_t a0 = (uint32_t)((w0 + q) << 5) + 1u;        // 1-based address of the word's first sample (slicer.py:75)
				while (mm) {
					const int i = __ffs(mm) - 1;
					mm &= mm - 1;
					const unsigned int sp = k + stage_shift;
					if ((s >> i) & 1u) atomicOr(&s_stage[sp >> 5], 1u << (sp & 31));
					const unsigned int gb = k + gl;
					if ((gb & 7u) == 7u) oa[gb >> 3] = a0 + i;
					k++;
				}
			}
		}
	} else if (cnt) {
		// quadrature: IQ signs of the symbol before this thread's first one
		unsigned int prev_cur = 0;
		if (C.quadrature && C.state_mask > 3u) {
			long long w = w0;
			uint32_t mm = 0;
			int steps = 0;
			// search backwards for the previous symbol
			while (true) {
				w -= 1;
				if (w < w_origin) break;
				mm = mk[w];
				if (mm) break;
				if (++steps > (1 << 20)) break;
			}
			if (w >= w_origin && mm) {
				const int i = 31 - __clz(mm);
				prev_cur = (((sg[w] >> i) & 1u) << 1) | ((sq[w] >> i) & 1u);
			} else {
				prev_cur = init_state ? (init_state[ch] & 3u) : 0u;
			}
		}
		long long sym = (long long)sym_block + ex;          // shard-local symbol index
#pragma unroll
		for (int q = 0; q < 4; q++) {
			uint32_t mm = m[q];
			const long long w = w0 + q;
			const uint32_t s = sv[q];
			const uint32_t sqw = qv[q];
			while (mm) {
				const int i = __ffs(mm) - 1;
				mm &= mm - 1;
				unsigned int val;
				if (C.quadrature) {
					const unsigned int cur = (((s >> i) & 1u) << 1) | ((sqw >> i) & 1u);
					const unsigned int state = ((prev_cur << 2) | cur) & C.state_mask;
					val = C.demap[state];
					prev_cur = cur;
				} else {
					val = (s >> i) & 1u;
				}
				// stream bits of this symbol, MSB of val first (slicer.py:215-216)
				for (int t = 0; t < C.bps; t++) {
					const long long g = sym * C.bps + t + bit_off;
					const unsigned int b = (val >> (C.bps - 1 - t)) & 1u;
					const unsigned int sp = (unsigned int)(g - bit_block) + stage_shift;
					if (b) atomicOr(&s_stage[sp >> 5], 1u << (sp & 31));
					if ((g & 7) == 7) oaddr[g >> 3] = (uint32_t)((w << 5) + i + 1);   // 1-based (slicer.py:75)
				}
				sym++;
			}
		}
	}
	__syncthreads();
	const long long word0 = bit_block >> 5;
	for (unsigned int i = threadIdx.x; i < stage_words; i += GB_THREADS) {
		const uint32_t v = s_stage[i];
		if (i == 0 || i == stage_words - 1) {
			if (v) atomicOr(&out[word0 + i], v);
		} else {
			out[word0 + i] = v;
		}
	}
}

// --- finalize per-chain counters ----------------------------------------------
__global__ void finalize_counts_kernel(const BitChain *__restrict__ chains, const unsigned int *__restrict__ sym_totals,
                                       ChainCounters *__restrict__ cc, int n_chains, const ShardBits *__restrict__ sb)
{
	const int ch = blockIdx.x * blockDim.x + threadIdx.x;
	if (ch >= n_chains) return;
	const long long nbits = (long long)sym_totals[ch] * chains[ch].bps + sb[ch].bit_off;
	cc[ch].nbits = nbits;
	cc[ch].nbytes = nbits >> 3;
	cc[ch].nflags = 0;
	cc[ch].seq_needed = 0;
	cc[ch].tail_short = 0;
	cc[ch].n_emit = 0;
	cc[ch].n_emit_bytes = 0;
}

// --- hand-off tail between shards ---------------------------------------------------
// extract: the last k_bits own bits of every chain (local positions [own_hi - k_bits, own_hi))
// inject : the previous shard's tail into local positions [bit_off - k_bits, bit_off)
__device__ __forceinline__ uint32_t read_bits32(const uint32_t *__restrict__ src, long long pos)
{
	const long long w = pos >> 5;
	const int r = (int)(pos & 31);
	const uint32_t lo = src[w];
	if (r == 0) return lo;
	return (lo >> r) | (src[w + 1] << (32 - r));
}

__global__ void __launch_bounds__(256)
tail_extract_kernel(const uint32_t *__restrict__ bits, long long bits_stride, const ShardBits *__restrict__ sb,
                    int k_words, uint32_t *__restrict__ out)
{
	const int ch = blockIdx.y;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= k_words) return;
	const long long pos = sb[ch].own_hi - 32ll * k_words + 32ll * i;
	out[(long long)ch * k_words + i] = (pos >= 0) ? read_bits32(bits + (long long)ch * bits_stride, pos) : 0u;
}

__global__ void __launch_bounds__(256)
tail_inject_kernel(uint32_t *__restrict__ bits, long long bits_stride, const ShardBits *__restrict__ sb,
                   int k_words, const uint32_t *__restrict__ in)
{
	const int ch = blockIdx.y;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= k_words || sb[ch].first) return;
	const uint32_t v = in[(long long)ch * k_words + i];
	if (!v) return;
	const long long pos = sb[ch].bit_off - 32ll * k_words + 32ll * i;     // >= 0 by construction
	uint32_t *dst = bits + (long long)ch * bits_stride;
	const long long w = pos >> 5;
	const int r = (int)(pos & 31);
	atomicOr(&dst[w], v << r);
	if (r) atomicOr(&dst[w + 1], v >> (32 - r));
}

// --- LFSR: out[g] = XOR_{k in poly} in[g-k], in[<0] = 0 (lfsr.py:22-52) ----------
__global__ void __launch_bounds__(256)
lfsr_kernel(const BitChain *__restrict__ chains, const ChainCounters *__restrict__ cc,
            const uint32_t *__restrict__ in, uint32_t *__restrict__ out, long long bits_stride, int max_words)
{
	const int ch = blockIdx.y;
	const long long nb = cc[ch].nbytes * 8;
	const long long nw = (nb + 31) >> 5;
	const uint32_t *src = in + (long long)ch * bits_stride;
	uint32_t *dst = out + (long long)ch * bits_stride;
	const unsigned long long poly = chains[ch].lfsr_poly;
	const bool inv = chains[ch].lfsr_invert != 0;
	for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < max_words;
	     w += (long long)gridDim.x * blockDim.x) {
		uint32_t acc = 0;
		if (w < nw) {
			unsigned long long p = poly;
			while (p) {
				const int k = __ffsll((long long)p) - 1;
				p &= p - 1;
				const int q = k >> 5, r = k & 31;
				const long long wh = w - q;
				const uint32_t hi = (wh >= 0) ? src[wh] : 0u;
				const uint32_t lo = (wh - 1 >= 0) ? src[wh - 1] : 0u;
				acc ^= r ? ((hi << r) | (lo >> (32 - r))) : hi;
			}
			if (inv) acc = ~acc;
			const long long rem = nb - (w << 5);
			if (rem < 32) acc &= (1u << (int)rem) - 1u;
		}
		dst[w] = acc;
	}
}

// --- AX.25 flag detection -------------------------------------------------------
// A flag event is a 0 that follows exactly six 1s (ax25.py:73): stream pattern
// 0,1,1,1,1,1,1,0 ending at bit g (a virtual 0 precedes the stream: one_count
// starts at 0).
__device__ __forceinline__ uint32_t flag_word(const uint32_t *__restrict__ d, long long w, long long nb)
{
	const uint32_t cur = d[w];
	const uint32_t prev = (w > 0) ? d[w - 1] : 0u;
	const unsigned long long V = ((unsigned long long)cur << 32) | prev;
	const unsigned long long A = V >> 25;
	unsigned long long F = ~A & (A >> 1) & (A >> 2) & (A >> 3) & (A >> 4) & (A >> 5) & (A >> 6) & ~(A >> 7);
	uint32_t f = (uint32_t)F;
	const long long rem = nb - (w << 5);
	if (rem <= 0) return 0;
	if (rem < 32) f &= (1u << (int)rem) - 1u;
	return f;
}

// IL2P sync word (il2p.py:369-373) ending at each bit of word w, evaluated on the true stream bits
// (the reference's partially filled shift register after a resume is handled by il2p_resolve_kernel).
__device__ __forceinline__ uint32_t il2p_sync_word(const uint32_t *__restrict__ d, long long w, long long nb, int tol)
{
	const long long rem = nb - (w << 5);
	if (rem <= 0) return 0;
	const uint32_t cur = d[w];
	const uint32_t prev = (w > 0) ? d[w - 1] : 0u;
	const unsigned long long V = ((unsigned long long)cur << 32) | prev;
	uint32_t f = 0;
#pragma unroll
	for (int i = 0; i < 32; i++) {
		const uint32_t ww = __brev((uint32_t)(V >> (i + 1)));       // newest bit in bit 0
		if (__popc((ww & 0xFFFFFFu) ^ 0xF15E48u) <= tol || __popc(ww ^ 0x5D57DF7Fu) <= tol) f |= 1u << i;
	}
	if (rem < 32) f &= (1u << (int)rem) - 1u;
	return f;
}

__device__ __forceinline__ uint32_t event_word(const BitChain &C, const uint32_t *__restrict__ d, long long w, long long nb)
{
	return C.codec == 2 ? il2p_sync_word(d, w, nb, C.il2p_sync_tol) : flag_word(d, w, nb);
}

#define FL_WORDS 1024
__global__ void __launch_bounds__(256)
flag_count_kernel(const BitChain *__restrict__ chains, const ChainCounters *__restrict__ cc,
                  const uint32_t *__restrict__ d, long long bits_stride,
                  unsigned int *__restrict__ block_count, int n_blocks)
{
	__shared__ unsigned int s_warp[33];
	const int ch = blockIdx.y;
	const BitChain C = chains[ch];
	const long long nb = cc[ch].nbytes * 8;
	const uint32_t *src = d + (long long)ch * bits_stride;
	const long long w0 = (long long)blockIdx.x * FL_WORDS + threadIdx.x * 4;
	unsigned int cnt = 0;
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const long long w = w0 + q;
		if ((w << 5) < nb) cnt += __popc(event_word(C, src, w, nb));
	}
	unsigned int total;
	block_excl_scan(cnt, s_warp, total);
	if (threadIdx.x == 0) block_count[(long long)ch * n_blocks + blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
flag_write_kernel(const BitChain *__restrict__ chains, const ChainCounters *__restrict__ cc,
                  const uint32_t *__restrict__ d, long long bits_stride,
                  const unsigned int *__restrict__ block_base, int n_blocks,
                  unsigned int *__restrict__ flag_pos, long long flag_stride)
{
	__shared__ unsigned int s_warp[33];
	const int ch = blockIdx.y;
	const BitChain C = chains[ch];
	const long long nb = cc[ch].nbytes * 8;
	const uint32_t *src = d + (long long)ch * bits_stride;
	unsigned int *fp = flag_pos + (long long)ch * flag_stride;
	const long long w0 = (long long)blockIdx.x * FL_WORDS + threadIdx.x * 4;
	uint32_t f[4];
	unsigned int cnt = 0;
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const long long w = w0 + q;
		f[q] = ((w << 5) < nb) ? event_word(C, src, w, nb) : 0u;
		cnt += __popc(f[q]);
	}
	unsigned int total;
	unsigned int idx = block_excl_scan(cnt, s_warp, total) + block_base[(long long)ch * n_blocks + blockIdx.x];
	for (int q = 0; q < 4; q++) {
		uint32_t ff = f[q];
		while (ff) {
			const int i = __ffs(ff) - 1;
			ff &= ff - 1;
			if ((long long)idx < flag_stride) fp[idx] = (unsigned int)(((w0 + q) << 5) + i);
			idx++;
		}
	}
}

// --- AX.25 gap replay -----------------------------------------------------------

__device__ __forceinline__ unsigned int crc16_x25_dev(const uint8_t *p, unsigned int n)
{
	unsigned int crc = 0xFFFF;
	for (unsigned int k = 0; k < n; k++) {
		crc ^= p[k];
#pragma unroll
		for (int i = 0; i < 8; i++) crc = (crc & 1u) ? ((crc >> 1) ^ 0x8408u) : (crc >> 1);
	}
	return crc ^ 0xFFFFu;
}

// The HDLC machine of ax25.py:25-93 over stream bits [start, end]; `end` is a
// flag position.  Bytes are appended to `dst`.  Returns true when a packet is
// emitted at the flag.  `aborted` reports whether an abort (seven or more ones,
// ax25.py:35-38) was seen, i.e. whether the emission decision is independent of
// anything before `start`.
//
// Four stream bits per step where nothing special happens inside them -- no seventh one in a row, no flag: 97 % of
// the nibbles of noise, all but two per frame of a clean signal -- from a table indexed by (run of ones so far,
// nibble): the run after the nibble, how many data bits it holds (stuffed zeros dropped) and the bits themselves.
// `hist` is the reference's working_byte as it stands when a byte is appended (the newest bit in bit 7, ax25.py:32/44:
// OR in 0x80, append, then shift), i.e. the last eight bits shifted in; the nibbles the table marks as complex and the
// last bits before the flag go through the bit-by-bit form of the reference's statements.
//   table entry: bits 0-2 run of ones after the nibble, 3-5 data bits in it, 6-9 the bits (first one lowest), bit 10 complex
#define HDLC_COMPLEX 0x400u
__device__ __forceinline__ unsigned int hdlc_replay_entry(unsigned int oc, unsigned int nib)
{
	unsigned int n = 0, bits = 0, complex_ = 0;
	for (int b = 0; b < 4; b++) {
		if ((nib >> b) & 1u) {
			oc++;
			if (oc > 6u) complex_ = 1;                     // abort: bit_index / byte_index reset in the middle
			bits |= 1u << n; n++;
		} else {
			if (oc < 5u) n++;                              // a data zero
			else if (oc >= 6u) complex_ = 1;               // flag (or the zero that ends an abort run)
			oc = 0;                                        // oc == 5: stuffed zero, dropped
		}
	}
	return (oc & 7u) | (n << 3) | (bits << 6) | (complex_ ? HDLC_COMPLEX : 0u);
}

__device__ bool ax25_replay(const uint32_t *__restrict__ d, const unsigned short *__restrict__ tab, long long start,
                            long long end, uint8_t *dst, unsigned int &len_out, int &overflow, bool &aborted)
{
	unsigned int hist = 0, one_count = 0, bit_index = 0, byte_index = 0, len = 0;
	bool emit = false;
	aborted = false;
	long long g = start;
	long long nw = (g >> 5) + 1;                          // next word to fetch
	unsigned long long buf = (unsigned long long)(d[g >> 5] >> (g & 31));
	int avail = 32 - (int)(g & 31);
	while (g <= end) {
		if (avail <= 32) {                                 // keep at least 32 bits in the window while the gap lasts
			buf |= (unsigned long long)d[nw++] << avail;   // (the rows are padded: reading one word past `end` is in bounds)
			avail += 32;
		}
		if (one_count <= 6u && end - g >= 8) {
			const unsigned int e = tab[one_count * 16u + ((unsigned int)buf & 15u)];
			if (!(e & HDLC_COMPLEX)) {
				const unsigned int nb = (e >> 3) & 7u, bits = (e >> 6) & 15u;
				const unsigned int tot = bit_index + nb;
				if (tot >= 8u) {
					const unsigned int j = 8u - bit_index;     // the byte completes with the first j data bits of the nibble
					dst[len++] = (uint8_t)((hist >> j) | ((bits & ((1u << j) - 1u)) << (8u - j)));
					byte_index++;
					if (byte_index > 1023u) { overflow = 1; len_out = len; return false; }   // replayed sequentially by the caller's fallback
					bit_index = tot - 8u;
				} else {
					bit_index = tot;
				}
				hist = ((hist >> nb) | (bits << (8u - nb))) & 0xFFu;
				one_count = e & 7u;
				g += 4; buf >>= 4; avail -= 4;
				continue;
			}
		}
		// inside a long run of ones nothing changes any more (bit_index = byte_index = 0, hist = 0xFF): skip 32 at once
		if (one_count >= 8u && end - g >= 32 && (unsigned int)buf == 0xFFFFFFFFu) {
			g += 32; buf >>= 32; avail -= 32;
			if (one_count < (1u << 30)) one_count += 32;
			continue;
		}
		const unsigned int bit = (unsigned int)buf & 1u;
		buf >>= 1; avail--;
		if (bit) {
			hist = (hist >> 1) | 0x80u;
			one_count++;
			bit_index++;
			if (one_count > 6u) { bit_index = 0; byte_index = 0; aborted = true; }   // abort: data not cleared
			if (bit_index == 8u) {
				bit_index = 0;
				dst[len++] = (uint8_t)hist;
				byte_index++;
				if (byte_index > 1023u) { byte_index = 0; one_count = 0; overflow = 1; }
			}
		} else {
			if (one_count < 5u) {
				hist >>= 1;
				bit_index++;
				if (bit_index == 8u) {
					bit_index = 0;
					dst[len++] = (uint8_t)hist;
					byte_index++;
					if (byte_index > 1023u) { byte_index = 0; overflow = 1; }
				}
			} else if (one_count == 6u) {
				if (g == end) emit = (byte_index >= 18u && bit_index == 7u);
				// (a flag strictly inside the range cannot happen: ranges end at the first flag)
				byte_index = 0;
				bit_index = 0;
			}
			one_count = 0;
		}
		g++;
	}
	len_out = len;
	return emit;
}

// ---- which gaps can emit at all? ------------------------------------------------------------------------------
// Only about one gap in fourteen ends in a packet (>= 18 bytes and a whole number of bytes + 7 bits at the closing
// flag, ax25.py:76), and that follows from the bit COUNT alone: data bits since the gap start or since the zero that
// ended the last abort run (seven or more ones zero bit_index and byte_index, ax25.py:35-38), minus the stuffed
// zeros.  The count is taken four stream bits at a time from a table indexed by (run of ones so far, nibble);
// only the gaps that pass (and the rare ones that need care: open at a shard start, or long enough to overflow
// max_packet_length) are replayed bit by bit, densely, by a second kernel.
//   table entry: bits 0-2 = ones run after the nibble (capped at 7), bit 3 = an abort happened inside the nibble,
//   bits 4-6 = data bits counted (after the last abort inside the nibble, if any)
__device__ __forceinline__ unsigned int hdlc_nibble_entry(unsigned int oc, unsigned int nib)
{
	unsigned int cnt = 0, reset = 0;
	for (int b = 0; b < 4; b++) {
		if ((nib >> b) & 1u) {
			oc = min(oc + 1u, 7u);
			cnt++;
			if (oc >= 7u) { cnt = 0; reset = 1; }              // one_count > 6: abort
		} else {
			if (oc < 5u) cnt++;                                // oc == 5: stuffed zero; oc >= 7: the zero ending an abort run
			oc = 0;
		}
	}
	return oc | (reset << 3) | (cnt << 4);
}

__device__ __forceinline__ void hdlc_count_step(const unsigned char *__restrict__ tab, unsigned int &oc, unsigned int &cnt,
                                                unsigned int nib)
{
	const unsigned int e = tab[oc * 16u + nib];
	cnt = (e & 8u) ? (e >> 4) : cnt + (e >> 4);
	oc = e & 7u;
}

// data-bit count at the closing flag of the gap [start, end] (end = the flag's final zero, not counted)
__device__ unsigned int hdlc_gap_count(const uint32_t *__restrict__ d, const unsigned char *__restrict__ tab,
                                       long long start, long long end)
{
	unsigned int oc = 0, cnt = 0;
	long long p = start;
	while (p + 32 <= end) {
		const long long w = p >> 5;
		const int r = (int)(p & 31);
		uint32_t v = r ? __funnelshift_r(d[w], d[w + 1], r) : d[w];
#pragma unroll
		for (int q = 0; q < 8; q++) { hdlc_count_step(tab, oc, cnt, v & 15u); v >>= 4; }
		p += 32;
	}
	if (p < end) {
		const long long w = p >> 5;
		const int r = (int)(p & 31);
		const int left = (int)(end - p);                           // 1..31 bits
		uint32_t v = d[w] >> r;
		if (r && r + left > 32) v |= d[w + 1] << (32 - r);
		int done = 0;
		for (; done + 4 <= left; done += 4) { hdlc_count_step(tab, oc, cnt, v & 15u); v >>= 4; }
		for (; done < left; done++) {                              // last 1..3 bits one at a time
			if (v & 1u) { oc = min(oc + 1u, 7u); cnt++; if (oc >= 7u) cnt = 0; }
			else { if (oc < 5u) cnt++; oc = 0; }
			v >>= 1;
		}
	}
	return cnt;
}

__global__ void __launch_bounds__(128)
ax25_gap_filter_kernel(const BitChain *__restrict__ chains, ChainCounters *__restrict__ cc,
                       const uint32_t *__restrict__ d, long long bits_stride,
                       const unsigned int *__restrict__ flag_pos, long long flag_stride,
                       const unsigned int *__restrict__ flag_totals,
                       GapRec *__restrict__ gaps, long long gap_stride, const ShardBits *__restrict__ sb,
                       unsigned int *__restrict__ cand, unsigned int *__restrict__ ncand)
{
	__shared__ unsigned char s_tab[128];
	if (threadIdx.x < 128) s_tab[threadIdx.x] = (unsigned char)hdlc_nibble_entry(threadIdx.x >> 4, threadIdx.x & 15u);
	__syncthreads();
	const int ch = blockIdx.y;
	if (chains[ch].codec != 1) return;
	const unsigned int nfl = flag_totals[ch];
	if (blockIdx.x == 0 && threadIdx.x == 0) cc[ch].nflags = (int)nfl;
	const unsigned int *fp = flag_pos + (long long)ch * flag_stride;
	const uint32_t *bits = d + (long long)ch * bits_stride;
	const ShardBits B = sb[ch];
	for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < nfl; j += gridDim.x * blockDim.x) {
		const long long end = fp[j];
		const long long start = (j == 0) ? 0 : (long long)fp[j - 1] + 1;
		GapRec r;
		r.emit = 0; r.len = 0; r.scratch_off = (unsigned int)(start >> 3); r.addr = 0; r.corrected = 0;
		gaps[(long long)ch * gap_stride + j] = r;
		// a frame needs >= 18 bytes + the 7 leading flag bits before the closing 0, and is emitted by the shard
		// that holds its closing bit
		const bool mine = end >= B.own_lo && end < B.own_hi;
		const bool open_start = !B.first && (j == 0 || (long long)fp[j - 1] < B.valid_from + 8);
		const long long len = end - start + 1;
		if (!mine || !(len >= 18 * 8 + 8 || open_start)) continue;
		bool replay = open_start || len >= 8192;                    // unknown history / max_packet_length overflow possible
		if (!replay) {
			const unsigned int cnt = hdlc_gap_count(bits, s_tab, start, end);
			replay = (cnt & 7u) == 7u && (cnt >> 3) >= 18u;
		}
		if (replay) cand[(long long)ch * flag_stride + atomicAdd(&ncand[ch], 1u)] = j;
	}
}

// the gaps that passed the filter, replayed bit by bit (ax25.py:25-93)
__global__ void __launch_bounds__(128)
ax25_gap_kernel(const BitChain *__restrict__ chains, ChainCounters *__restrict__ cc,
                const uint32_t *__restrict__ d, long long bits_stride,
                const unsigned int *__restrict__ flag_pos, long long flag_stride,
                const uint32_t *__restrict__ byte_addr, long long addr_stride,
                uint8_t *__restrict__ scratch, long long scratch_stride,
                GapRec *__restrict__ gaps, long long gap_stride, const ShardBits *__restrict__ sb,
                const unsigned int *__restrict__ cand, const unsigned int *__restrict__ ncand)
{
	__shared__ unsigned short s_rep[7 * 16];
	if (threadIdx.x < 7 * 16) s_rep[threadIdx.x] = (unsigned short)hdlc_replay_entry(threadIdx.x >> 4, threadIdx.x & 15u);
	__syncthreads();
	const int ch = blockIdx.y;
	if (chains[ch].codec != 1) return;
	const unsigned int n = ncand[ch];
	const unsigned int *fp = flag_pos + (long long)ch * flag_stride;
	const ShardBits B = sb[ch];
	for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		const unsigned int j = cand[(long long)ch * flag_stride + k];
		const long long end = fp[j];
		long long start = (j == 0) ? 0 : (long long)fp[j - 1] + 1;
		GapRec r;
		r.emit = 0; r.len = 0; r.scratch_off = (unsigned int)(start >> 3); r.addr = 0; r.corrected = 0;
		// On a later shard the stream starts in the middle of the recording: a gap whose opening flag is
		// not a reliably detected one (all 8 bits of its pattern at valid positions) reaches back past the
		// hand-off tail.  It is replayed from the first valid bit with unknown history.
		const bool open_start = !B.first && (j == 0 || (long long)fp[j - 1] < B.valid_from + 8);
		if (open_start && start < B.valid_from) start = B.valid_from;
		int overflow = 0;
		unsigned int len = 0;
		bool aborted = false;
		const bool emit = ax25_replay(d + (long long)ch * bits_stride, s_rep, start, end,
			scratch + (long long)ch * scratch_stride + r.scratch_off, len, overflow, aborted);
		if (overflow) atomicExch(&cc[ch].seq_needed, 1);
		if (open_start && (emit || !aborted)) {
			// either junk bytes from before the tail would be part of the frame, or the
			// emission decision itself depends on bits we do not have (an abort -- seven ones, all
			// of them valid bits -- resets the machine whatever came before)
			atomicExch(&cc[ch].tail_short, 1);
		} else {
			r.emit = emit ? 1u : 0u;
			r.len = len;
			r.addr = byte_addr[(long long)ch * addr_stride + (end >> 3)];
		}
		gaps[(long long)ch * gap_stride + j] = r;
	}
}

// Sequential replay of one whole chain (only when a gap overflowed
// max_packet_length, ax25.py:46-51, which can desynchronise the stateless flag
// detector).  One thread per chain; rewrites the chain's gap records as a
// compact list of emitted packets.  (Unsharded runs only: the engine refuses a
// sharded run that needs it.)
__global__ void ax25_sequential_kernel(const BitChain *__restrict__ chains, ChainCounters *__restrict__ cc,
                                       const uint32_t *__restrict__ dall, long long bits_stride,
                                       const uint32_t *__restrict__ byte_addr, long long addr_stride,
                                       uint8_t *__restrict__ scratch, long long scratch_stride,
                                       GapRec *__restrict__ gaps, long long gap_stride)
{
	const int ch = blockIdx.x;
	if (threadIdx.x != 0 || chains[ch].codec != 1 || !cc[ch].seq_needed) return;
	const uint32_t *d = dall + (long long)ch * bits_stride;
	uint8_t *dst = scratch + (long long)ch * scratch_stride;
	GapRec *out = gaps + (long long)ch * gap_stride;
	const long long nb = cc[ch].nbytes * 8;
	unsigned int wb = 0, one_count = 0, bit_index = 0, byte_index = 0;
	// bytes of the current working packet are written at `base`, which only
	// moves forward when a new PacketMeta starts; appended bytes never outrun
	// the stream position (each needs >= 8 stream bits).
	long long base = 0, len = 0;
	unsigned int nrec = 0;
	for (long long g = 0; g < nb; g++) {
		const unsigned int bit = (d[g >> 5] >> (g & 31)) & 1u;
		if (bit) {
			wb |= 0x80; one_count++; bit_index++;
			if (one_count > 6) { bit_index = 0; byte_index = 0; }
			if (bit_index == 8) {
				bit_index = 0; dst[base + len++] = (uint8_t)wb; byte_index++;
				if (byte_index > 1023) { byte_index = 0; one_count = 0; }
			}
			wb >>= 1;
		} else {
			if (one_count < 5) {
				bit_index++;
				if (bit_index == 8) {
					bit_index = 0; dst[base + len++] = (uint8_t)wb; byte_index++;
					if (byte_index > 1023) byte_index = 0;
				}
				wb >>= 1;
			} else if (one_count == 6) {
				if (byte_index >= 18 && bit_index == 7) {
					GapRec r;
					r.emit = 1; r.len = (unsigned int)len; r.scratch_off = (unsigned int)base;
					r.addr = byte_addr[(long long)ch * addr_stride + (g >> 3)];
					r.corrected = 0;
					out[nrec++] = r;
					base += len;
				}
				// new PacketMeta(): later bytes may overwrite a non-emitted packet's bytes
				len = 0;
				byte_index = 0; bit_index = 0;
			}
			one_count = 0;
		}
	}
	cc[ch].nflags = (int)nrec;     // gap list now holds exactly the emitted packets
}

// --- compaction: gap records -> ordered packet records ---------------------------
// Chains in order, gaps in order => records ordered like the reference's
// per-chain decode() lists.  Step 1: emitted packets / bytes per chain.
__global__ void __launch_bounds__(1024)
packet_count_kernel(ChainCounters *__restrict__ cc, const GapRec *__restrict__ gaps, long long gap_stride)
{
	__shared__ unsigned int s_warp[33];
	const int ch = blockIdx.x;
	const int n = cc[ch].nflags;
	const GapRec *g = gaps + (long long)ch * gap_stride;
	unsigned int ne = 0, nb = 0;
	for (int j = threadIdx.x; j < n; j += 1024) {
		const GapRec r = g[j];
		if (r.emit) { ne++; nb += r.len; }
	}
	unsigned int te, tb;
	block_excl_scan(ne, s_warp, te);
	block_excl_scan(nb, s_warp, tb);
	if (threadIdx.x == 0) { cc[ch].n_emit = (int)te; cc[ch].n_emit_bytes = tb; }
}

// Step 2: one CTA per chain; its base is the sum over the chains before it.
__global__ void __launch_bounds__(1024)
packet_index_kernel(const ChainCounters *__restrict__ cc, int n_chains,
                    const GapRec *__restrict__ gaps, long long gap_stride,
                    PacketRecDev *__restrict__ recs, unsigned int *__restrict__ rec_src,
                    unsigned long long rec_cap, PacketTotals *__restrict__ totals, long long sample_base)
{
	__shared__ unsigned int s_warp[33];
	__shared__ unsigned long long s_base[2];
	const int ch = blockIdx.x;
	if (threadIdx.x == 0) {
		unsigned long long r = 0, b = 0;
		for (int c = 0; c < ch; c++) { r += (unsigned int)cc[c].n_emit; b += (unsigned long long)cc[c].n_emit_bytes; }
		s_base[0] = r; s_base[1] = b;
	}
	__syncthreads();
	unsigned long long n_rec = s_base[0], n_bytes = s_base[1];
	const int n = cc[ch].nflags;
	const GapRec *g = gaps + (long long)ch * gap_stride;
	// four consecutive gaps per thread and round: a chain of noise has ~10^5 gaps and a round costs two block scans
	for (int base = 0; base < n; base += 4096) {
		const int j0 = base + 4 * threadIdx.x;
		GapRec r[4];
		unsigned int e_sum = 0, b_sum = 0;
#pragma unroll
		for (int q = 0; q < 4; q++) {
			r[q].emit = 0; r[q].len = 0; r[q].scratch_off = 0; r[q].addr = 0; r[q].corrected = 0;
			if (j0 + q < n) r[q] = g[j0 + q];
			e_sum += r[q].emit ? 1u : 0u;
			b_sum += r[q].emit ? r[q].len : 0u;
		}
		unsigned int tot_e, tot_b;
		unsigned int ex_e = block_excl_scan(e_sum, s_warp, tot_e);
		unsigned int ex_b = block_excl_scan(b_sum, s_warp, tot_b);
#pragma unroll
		for (int q = 0; q < 4; q++) {
			if (r[q].emit) {
				const unsigned long long ri = n_rec + ex_e;
				if (ri < rec_cap) {
					PacketRecDev p;
					p.chain = ch; p.len = r[q].len; p.offset = n_bytes + ex_b;
					p.streamaddress = sample_base + (long long)r[q].addr;
					p.bytes_corrected = r[q].corrected; p.calculated_crc = 0; p.carried_crc = 0;
					p.valid_crc = 0; p.valid_header = 0;
					for (int t = 0; t < 6; t++) p.pad[t] = 0;
					recs[ri] = p;
					rec_src[ri] = r[q].scratch_off;
				}
				ex_e += 1u;
				ex_b += r[q].len;
			}
		}
		n_rec += tot_e;
		n_bytes += tot_b;
	}
	if (ch == n_chains - 1 && threadIdx.x == 0) { totals->n_packets = n_rec; totals->n_bytes = n_bytes; }
}

// one warp per packet: copy its bytes into the arena, CRC + header check
__global__ void __launch_bounds__(256)
packet_copy_kernel(PacketRecDev *__restrict__ recs, const unsigned int *__restrict__ rec_src,
                   const PacketTotals *__restrict__ totals, unsigned long long first_rec, unsigned long long rec_cap,
                   const uint8_t *__restrict__ scratch, long long scratch_stride,
                   uint8_t *__restrict__ arena, unsigned long long arena_cap)
{
	const unsigned long long n = min(totals->n_packets, rec_cap);
	const int lane = threadIdx.x & 31;
	for (unsigned long long ri = first_rec + (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	     ri < n; ri += (unsigned long long)gridDim.x * (blockDim.x >> 5)) {
		PacketRecDev p = recs[ri];
		if (p.offset + p.len > arena_cap) continue;
		const uint8_t *src = scratch + (long long)p.chain * scratch_stride + rec_src[ri];
		uint8_t *dst = arena + p.offset;
		for (unsigned int i = lane; i < p.len; i += 32) dst[i] = src[i];
		if (lane == 0) {
			// PacketMeta.CalcCRC (packet_meta.py:197-203) + Validate (:205-208)
			const unsigned int carried = (unsigned int)src[p.len - 1] * 256u + src[p.len - 2];
			const unsigned int calc = crc16_x25_dev(src, p.len - 2);
			bool hdr = p.len > 15;
			if (hdr)
				for (int i = 0; i < 7; i++) {
					const unsigned int wc = src[i] >> 1;
					if ((wc < 32 || wc > 126) && wc != 0) hdr = false;
				}
			recs[ri].calculated_crc = (unsigned short)calc;
			recs[ri].carried_crc = (unsigned short)carried;
			recs[ri].valid_crc = (calc == carried);
			recs[ri].valid_header = hdr;
		}
	}
}

// --- export of AddressedData streams for parity tests ----------------------------
__global__ void __launch_bounds__(256)
stream_export_kernel(const ChainCounters *__restrict__ cc, int ch, const uint32_t *__restrict__ bits,
                     long long bits_stride, const uint32_t *__restrict__ byte_addr, long long addr_stride,
                     uint8_t *__restrict__ out_bytes, long long *__restrict__ out_addr, long long sample_base)
{
	const long long nbytes = cc[ch].nbytes;
	const uint32_t *src = bits + (long long)ch * bits_stride;
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nbytes;
	     b += (long long)gridDim.x * blockDim.x) {
		const unsigned int v = (src[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
		out_bytes[b] = (uint8_t)(__brev(v) >> 24);            // first stream bit is the byte's MSB
		out_addr[b] = sample_base + (long long)byte_addr[(long long)ch * addr_stride + b];
	}
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
extern "C" {

// (sign, mask) words [w_origin, w_origin + n_words) -> packed bits placed at sb[ch].bit_off
cudaError_t pm_launch_gather(const BitChain *chains, int n_chains, ChainCounters *cc, const uint32_t *sign,
	long long sign_stride, const uint32_t *mask, long long mask_stride, long long w_origin, long long n_words,
	unsigned int *blk_count, unsigned int *blk_base, unsigned int *sym_totals,
	uint32_t *bits, long long bits_stride, uint32_t *byte_addr, long long addr_stride,
	const unsigned int *init_state, const ShardBits *sb, cudaStream_t st)
{
	const int n_blocks = (int)((n_words + GB_WORDS - 1) / GB_WORDS);
	dim3 grid(n_blocks, n_chains);
	pm_kt_mark("gather_count_kernel", st);
	gather_count_kernel<<<grid, GB_THREADS, 0, st>>>(chains, mask, mask_stride, blk_count, n_blocks, w_origin);
	pm_kt_mark("row_scan_kernel", st);
	row_scan_kernel<<<n_chains, 1024, 0, st>>>(blk_count, blk_base, n_blocks, n_blocks, sym_totals);
	pm_kt_mark("finalize_counts_kernel", st);
	finalize_counts_kernel<<<(n_chains + 31) / 32, 32, 0, st>>>(chains, sym_totals, cc, n_chains, sb);
	pm_kt_mark("memset bits", st);
	cudaMemsetAsync(bits, 0, sizeof(uint32_t) * (size_t)bits_stride * n_chains, st);
	pm_kt_mark("gather_write_kernel", st);
	gather_write_kernel<<<grid, GB_THREADS, 0, st>>>(chains, sign, sign_stride, mask, mask_stride, blk_base,
		n_blocks, bits, bits_stride, byte_addr, addr_stride, init_state, w_origin, sb);
	return cudaGetLastError();
}

cudaError_t pm_launch_tail_extract(const uint32_t *bits, long long bits_stride, const ShardBits *sb, int n_chains,
	int k_words, uint32_t *out, cudaStream_t st)
{
	dim3 grid((k_words + 255) / 256, n_chains);
	pm_kt_mark("tail_extract_kernel", st);
	tail_extract_kernel<<<grid, 256, 0, st>>>(bits, bits_stride, sb, k_words, out);
	return cudaGetLastError();
}

cudaError_t pm_launch_tail_inject(uint32_t *bits, long long bits_stride, const ShardBits *sb, int n_chains,
	int k_words, const uint32_t *in, cudaStream_t st)
{
	dim3 grid((k_words + 255) / 256, n_chains);
	pm_kt_mark("tail_inject_kernel", st);
	tail_inject_kernel<<<grid, 256, 0, st>>>(bits, bits_stride, sb, k_words, in);
	return cudaGetLastError();
}

cudaError_t pm_launch_lfsr(const BitChain *chains, int n_chains, const ChainCounters *cc, const uint32_t *in,
	uint32_t *out, long long bits_stride, cudaStream_t st)
{
	int bx = (int)((bits_stride + 255) / 256);
	if (bx > 1024) bx = 1024;
	dim3 grid(bx, n_chains);
	pm_kt_mark("lfsr_kernel", st);
	lfsr_kernel<<<grid, 256, 0, st>>>(chains, cc, in, out, bits_stride, (int)bits_stride);
	return cudaGetLastError();
}

cudaError_t pm_launch_ax25(const BitChain *chains, int n_chains, ChainCounters *cc, const uint32_t *d,
	long long bits_stride, unsigned int *blk_count, unsigned int *blk_base, unsigned int *flag_totals,
	unsigned int *flag_pos, long long flag_stride, const uint32_t *byte_addr, long long addr_stride,
	uint8_t *scratch, long long scratch_stride, GapRec *gaps, long long gap_stride, const ShardBits *sb,
	int allow_sequential, unsigned int *gap_cand, unsigned int *gap_ncand, cudaStream_t st)
{
	const int n_blocks = (int)((bits_stride + FL_WORDS - 1) / FL_WORDS);
	dim3 grid(n_blocks, n_chains);
	pm_kt_mark("flag_count_kernel", st);
	flag_count_kernel<<<grid, 256, 0, st>>>(chains, cc, d, bits_stride, blk_count, n_blocks);
	pm_kt_mark("row_scan_kernel", st);
	row_scan_kernel<<<n_chains, 1024, 0, st>>>(blk_count, blk_base, n_blocks, n_blocks, flag_totals);
	pm_kt_mark("flag_write_kernel", st);
	flag_write_kernel<<<grid, 256, 0, st>>>(chains, cc, d, bits_stride, blk_base, n_blocks, flag_pos, flag_stride);
	// gaps: at most flag_stride per chain
	dim3 ggrid((unsigned int)std::min<long long>((flag_stride + 127) / 128, 148 * 16), n_chains);
	cudaMemsetAsync(gap_ncand, 0, sizeof(unsigned int) * n_chains, st);
	pm_kt_mark("ax25_gap_filter_kernel", st);
	ax25_gap_filter_kernel<<<ggrid, 128, 0, st>>>(chains, cc, d, bits_stride, flag_pos, flag_stride, flag_totals, gaps,
		gap_stride, sb, gap_cand, gap_ncand);
	pm_kt_mark("ax25_gap_kernel", st);
	ax25_gap_kernel<<<dim3(148 * 2, n_chains), 128, 0, st>>>(chains, cc, d, bits_stride, flag_pos, flag_stride, byte_addr,
		addr_stride, scratch, scratch_stride, gaps, gap_stride, sb, gap_cand, gap_ncand);
	if (allow_sequential) {
		pm_kt_mark("ax25_sequential_kernel", st);
		ax25_sequential_kernel<<<n_chains, 32, 0, st>>>(chains, cc, d, bits_stride, byte_addr, addr_stride, scratch,
			scratch_stride, gaps, gap_stride);
	}
	return cudaGetLastError();
}

cudaError_t pm_launch_packets(int n_chains, ChainCounters *cc, const GapRec *gaps,
	long long gap_stride, PacketRecDev *recs, unsigned int *rec_src, unsigned long long rec_cap,
	PacketTotals *totals, const uint8_t *scratch, long long scratch_stride, uint8_t *arena,
	unsigned long long arena_cap, long long sample_base, cudaStream_t st)
{
	pm_kt_mark("packet_count_kernel", st);
	packet_count_kernel<<<n_chains, 1024, 0, st>>>(cc, gaps, gap_stride);
	pm_kt_mark("packet_index_kernel", st);
	packet_index_kernel<<<n_chains, 1024, 0, st>>>(cc, n_chains, gaps, gap_stride, recs, rec_src, rec_cap, totals,
		sample_base);
	pm_kt_mark("packet_copy_kernel", st);
	packet_copy_kernel<<<296, 256, 0, st>>>(recs, rec_src, totals, 0, rec_cap, scratch, scratch_stride, arena,
		arena_cap);
	return cudaGetLastError();
}

cudaError_t pm_launch_stream_export(const ChainCounters *cc, int ch, const uint32_t *bits, long long bits_stride,
	const uint32_t *byte_addr, long long addr_stride, uint8_t *out_bytes, long long *out_addr,
	long long sample_base, cudaStream_t st)
{
	pm_kt_mark("stream_export_kernel", st);
	stream_export_kernel<<<296, 256, 0, st>>>(cc, ch, bits, bits_stride, byte_addr, addr_stride, out_bytes, out_addr,
		sample_base);
	return cudaGetLastError();
}

}  // extern "C"

// il2p.cu -- IL2P decode (reference il2p.py:360-519) over packed bitstreams.
//
// The reference is one state machine per chain: search the 24/32-bit sync word with a Hamming
// tolerance, then consume a header (15 bytes, RS(15,13)), payload blocks (RS with 16 parity
// bytes each) and the Hamming(7,4)-protected trailing CRC, and go back to searching.  What is
// sequential is only the question "was the decoder searching at this bit?"; everything else
// about a sync position is a pure function of the bits that follow it.  So:
//
//   1. sync positions are detected statelessly for every bit in parallel (event kernels in
//      bits.cu with the IL2P detector below);
//   2. every candidate is decoded speculatively, one thread per candidate (RS syndromes,
//      Berlekamp-Massey, Chien, Forney exactly as rs_functions.py:33-150, descrambler, header
//      translation) -> status, last consumed bit, bytes, corrected-byte count;
//   3. one warp per chain walks the candidates in stream order: accept the first sync at or after
//      the resume point, jump to the bit after its frame, repeat.  The walk also reproduces the
//      reference's quirks: the shift register is 0xFFFFFF at the start and holds only the last
//      collected byte (zero-filled above) after a frame (il2p.py:119, 147-153), so for the 32 bits
//      after a resume point the sync comparison is evaluated explicitly; corrected-byte counts of
//      failed frames leak into the next emitted frame (il2p.py:200-211).
#include "pm_common.cuh"

__constant__ unsigned char c_hamming[128];  // il2p.py:19-42
// GF(2^8) tables back to back in global memory: exp (gf_functions.py table[], 255 entries) | log (index[], index[0] = 0)
// | inverse[].  Not __constant__: the speculative decode runs one thread per candidate, every lane looks up a different
// entry, and the constant cache serialises that -- il2p_decode_kernel copies them into shared memory, the resolve walk
// (one lane) reads them through L1; both pass the base pointer down (T below).
__device__ unsigned char d_gf_tab[768];

// GF(2^8) tables exactly as gf_functions.py:47-74 builds them (Galois LFSR over genpoly 0x11D walking
// the exponents downwards, inverse by definition) and the Hamming(7,4) table of il2p.py:19-42.
extern "C" cudaError_t pm_il2p_init_tables(void)
{
	static const unsigned char hamming[128] = {
		0x0, 0x0, 0x0, 0x3, 0x0, 0x5, 0xe, 0x7, 0x0, 0x9, 0xe, 0xb, 0xe, 0xd, 0xe, 0xe,
		0x0, 0x3, 0x3, 0x3, 0x4, 0xd, 0x6, 0x3, 0x8, 0xd, 0xa, 0x3, 0xd, 0xd, 0xe, 0xd,
		0x0, 0x5, 0x2, 0xb, 0x5, 0x5, 0x6, 0x5, 0x8, 0xb, 0xb, 0xb, 0xc, 0x5, 0xe, 0xb,
		0x8, 0x1, 0x6, 0x3, 0x6, 0x5, 0x6, 0x6, 0x8, 0x8, 0x8, 0xb, 0x8, 0xd, 0x6, 0xf,
		0x0, 0x9, 0x2, 0x7, 0x4, 0x7, 0x7, 0x7, 0x9, 0x9, 0xa, 0x9, 0xc, 0x9, 0xe, 0x7,
		0x4, 0x1, 0xa, 0x3, 0x4, 0x4, 0x4, 0x7, 0xa, 0x9, 0xa, 0xa, 0x4, 0xd, 0xa, 0xf,
		0x2, 0x1, 0x2, 0x2, 0xc, 0x5, 0x2, 0x7, 0xc, 0x9, 0x2, 0xb, 0xc, 0xc, 0xc, 0xf,
		0x1, 0x1, 0x2, 0x1, 0x4, 0x1, 0x6, 0xf, 0x8, 0x1, 0xa, 0xf, 0xc, 0xf, 0xf, 0xf};
	unsigned char exp_t[256] = {0}, log_t[256] = {0}, inv_t[256] = {0};
	unsigned int lfsr = 1;
	for (int i = 254; i >= 0; i--) {
		const unsigned int fb = lfsr & 1u;
		lfsr >>= 1;
		if (fb) lfsr ^= 0x11Du >> 1;
		exp_t[i] = (unsigned char)lfsr;
		log_t[lfsr] = (unsigned char)i;
	}
	for (int a = 1; a < 256; a++) inv_t[a] = exp_t[(255 - log_t[a]) % 255];
	cudaError_t e;
	if ((e = cudaMemcpyToSymbol(d_gf_tab, exp_t, 256, 0)) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(d_gf_tab, log_t, 256, 256)) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(d_gf_tab, inv_t, 256, 512)) != cudaSuccess) return e;
	return cudaMemcpyToSymbol(c_hamming, hamming, 128);
}

// T: exp[256] | log[256] | inverse[256] (shared or global memory)
__device__ __forceinline__ int gf_mul(const unsigned char *__restrict__ T, int a, int b)   // gf_functions.py:18-24
{
	if (a == 0 || b == 0) return 0;
	int r = T[256 + a] + T[256 + b];
	if (r > 254) r -= 255;
	return T[r];
}

// rs_functions.py:33-150, first_root 0.  NR = number of roots (2 header, 16 blocks).
// Returns the number of corrected bytes or -1; corrects data in place.
template <int NR>
__device__ int rs_decode_dev(const unsigned char *__restrict__ T, unsigned char *data, int block_size, int min_distance)
{
	int syn[NR], loc[NR], nxt[NR], where[NR];
	int cor[NR + 1];
	int error_count = 0;
	for (int i = 0; i < NR; i++) {                                  // :36-42
		int s = 0;
		const int x = T[i];
		for (int j = 0; j < block_size - 1; j++) s = gf_mul(T, s ^ data[j], x);
		syn[i] = s ^ data[block_size - 1];
		loc[i] = 0; nxt[i] = 0; where[i] = 0;
	}
	for (int i = 0; i <= NR; i++) cor[i] = 0;
	loc[0] = 1;
	cor[1] = 1;
	int order = 0;
	for (int step = 1; step <= NR; step++) {                        // :60-82
		const int y = step - 1;
		int e = syn[y];
		for (int i = 1; i <= order; i++) e ^= gf_mul(T, loc[i], syn[y - i]);
		if (e != 0) {
			for (int i = 0; i <= order; i++) nxt[i] = loc[i] ^ gf_mul(T, e, cor[i]);
			e = T[512 + e];
			for (int i = 0; i < NR / 2 + 1; i++) cor[i] = gf_mul(T, loc[i], e);
			for (int i = 0; i < NR / 2 + 1; i++) loc[i] = nxt[i];
		}
		if (2 * order < step) order = step - order;
		for (int i = NR; i > 0; i--) cor[i] = cor[i - 1];
		cor[0] = 0;
	}
	for (int j = 0; j < block_size; j++) {                          // :85-98 Chien
		int x = 0;
		const int y = j + 256 - block_size;
		for (int i = 1; i < NR / 2 + 1; i++)
			if (loc[i]) x ^= T[(y * i + T[256 + loc[i]]) % 255];
		x ^= loc[0];
		if (x == 0) where[error_count++] = j;
	}
	if (error_count <= NR / 2 - min_distance) {                     // :99-140 Forney
		for (int i = 0; i < error_count; i++) {
			cor[i] = syn[i];
			for (int j = 1; j <= i; j++) cor[i] ^= gf_mul(T, syn[i - j], loc[j]);
		}
		for (int i = 0; i < error_count; i++) {
			const int e = block_size - where[i] - 1;
			int z = cor[0];
			for (int j = 1; j < error_count; j++) {
				int x = (e * j) % 255;
				x = (255 - x) % 255;
				z ^= gf_mul(T, cor[j], T[x]);
			}
			z = gf_mul(T, z, T[e]);
			int y = loc[1];
			for (int j = 3; j < NR / 2 + 1; j += 2) {
				int x = (e * (j - 1)) % 255;
				x = (255 - x) % 255;
				y ^= gf_mul(T, loc[j], T[x]);
			}
			y = T[256 + y];
			y = 256 - y - 1;
			if (y == 255) y = 0;
			y = T[y];
			data[where[i]] ^= (unsigned char)gf_mul(T, y, z);
		}
	}
	for (int i = 0; i < NR; i++) {                                  // :142-149
		int s = 0;
		const int x = T[i];
		for (int j = 0; j < block_size - 1; j++) s = gf_mul(T, s ^ data[j], x);
		if ((s ^ data[block_size - 1]) != 0) return -1;
	}
	return error_count;
}

// 8 stream bits starting at p, first bit = MSB (il2p.py:147-153 with mask 0xFF)
__device__ __forceinline__ unsigned int il2p_byte(const uint32_t *__restrict__ d, long long p)
{
	const long long w = p >> 5;
	const int r = (int)(p & 31);
	unsigned int v = d[w] >> r;
	if (r > 24) v |= d[w + 1] << (32 - r);
	return __brev(v & 0xFFu) >> 24;
}

// block_unscramble il2p.py:160-163 (LFSRnoaddr lfsr.py:62-92: poly 0x211, state 0x1F0, no invert)
__device__ void il2p_unscramble(unsigned char *buf, int n)
{
	unsigned int sr = 0x1F0;
	for (int k = 0; k < n; k++) {
		unsigned int in = buf[k], out = 0;
#pragma unroll
		for (int b = 0; b < 8; b++) {
			if (in & 0x80u) sr ^= 0x211u;
			out = (out << 1) | (sr & 1u);
			in <<= 1;
			sr >>= 1;
		}
		buf[k] = (unsigned char)out;
	}
}

__device__ unsigned int crc16_x25_il2p(const unsigned char *p, unsigned int n)   // crc_functions.py:63-76
{
	unsigned int crc = 0xFFFF;
	for (unsigned int k = 0; k < n; k++) {
		crc ^= p[k];
#pragma unroll
		for (int i = 0; i < 8; i++) crc = (crc & 1u) ? ((crc >> 1) ^ 0x8408u) : (crc >> 1);
	}
	return crc ^ 0xFFFFu;
}

// Decode the frame that follows a sync word whose last bit is stream bit g.
// out: packet bytes (the rebuilt AX.25 frame + FCS).  nb = stream length in bits.
__device__ void il2p_try(const unsigned char *__restrict__ T, const uint32_t *__restrict__ d, long long nb, long long g,
                         const BitChain &C, unsigned char *out, Il2pRes &res)
{
	unsigned char buf[256];
	long long p = g + 1;
	unsigned int len = 0;
	int corrected = 0;
	res.status = IL2P_INCOMPLETE; res.len = 0; res.corrected = 0; res.end_bit = nb - 1;
	if (p + 120 > nb) return;
	for (int i = 0; i < 15; i++) buf[i] = (unsigned char)il2p_byte(d, p + 8 * i);
	p += 120;
	bool fail = false;
	{
		const int r = C.il2p_disable_rs ? 0 : rs_decode_dev<2>(T, buf, 15, C.il2p_min_dist);      // il2p.py:188-207
		if (r < 0) fail = true; else corrected += r;
	}
	il2p_unscramble(buf, 13);
	// unpack_il2p_header il2p.py:214-290
	const int type_subfield = (buf[1] & 0x80) >> 7;
	int count = 0, pid = 0, control = 0;
	for (int i = 0; i < 10; i++) if (buf[i + 2] & 0x80) count |= 0x200 >> i;
	for (int i = 0; i < 4; i++) if (buf[i + 1] & 0x40) pid |= 0x8 >> i;
	for (int i = 0; i < 7; i++) if (buf[i + 5] & 0x40) control |= 0x40 >> i;
	if (type_subfield == 1) {                                       // construct_ax25_header il2p.py:292-344
		int type;                                                   // 0 UI, 1 S, 2 U, 3 I
		if (buf[0] & 0x40) type = 0;
		else if (pid == 0) type = 1;
		else if (pid == 1) type = 2;
		else type = 3;
		const int pf = (control & 0x40) != 0;
		int cbit = 0, nr = 0, ns = 0, opcode = 0;
		if (type == 3) { ns = control & 0x7; nr = (control >> 3) & 0x7; cbit = 1; }
		else if (type == 1) { nr = (control >> 3) & 0x7; if (control & 0x4) cbit = 1; opcode = control & 0x3; }
		else { if (control & 0x4) cbit = 1; opcode = (control >> 3) & 0x7; }
		for (int i = 0; i < 6; i++) out[len++] = (unsigned char)(((buf[i] & 0x3F) + 0x20) << 1);
		out[len++] = (unsigned char)(((buf[12] >> 4) << 1) + 0x60 + (cbit ? 0x80 : 0));
		for (int i = 0; i < 6; i++) out[len++] = (unsigned char)(((buf[i + 6] & 0x3F) + 0x20) << 1);
		out[len++] = (unsigned char)(((buf[12] & 0xF) << 1) + 0x60 + (cbit ? 0 : 0x80) + 1);
		int cb;                                                     // reform_control_byte il2p.py:89-107
		if (type == 0 || type == 2) {
			const unsigned char u_control[8] = {0x2F, 0x43, 0x0F, 0x63, 0x87, 0x03, 0xAF, 0xE3};
			cb = u_control[opcode];
		} else if (type == 1) cb = 0x1 | (opcode << 2) | (nr << 5);
		else cb = (ns << 1) | (nr << 5);
		if (pf) cb |= 0x10;
		out[len++] = (unsigned char)cb;
		const unsigned char pid_table[16] = {0, 0, 0x10, 0x01, 0x06, 0x07, 0x08, 0xC3, 0xC4, 0xCA, 0xCB, 0xCC, 0xCD, 0xCE, 0xCF, 0xF0};
		if (pid_table[pid] != 0) out[len++] = pid_table[pid];
	}
	if (fail) {                                                     // il2p.py:405-409
		res.status = IL2P_FAIL; res.end_bit = p - 1; res.corrected = (unsigned int)corrected;
		return;
	}
	if (count > 0) {                                                // calc_big_small_blocks il2p.py:346-358
		const int block_count = (count + 238) / 239;
		const int small = count / block_count;
		const int big_blocks = count - block_count * small;
		for (int b = 0; b < block_count; b++) {
			const int size = (b < big_blocks) ? small + 1 : small;
			const int total = size + 16;
			if (p + 8ll * total > nb) { res.corrected = (unsigned int)corrected; return; }    // stream ends inside the frame
			for (int i = 0; i < total; i++) buf[i] = (unsigned char)il2p_byte(d, p + 8 * i);
			p += 8ll * total;
			const int r = C.il2p_disable_rs ? 0 : rs_decode_dev<16>(T, buf, total, C.il2p_min_dist);
			if (r < 0) fail = true; else corrected += r;
			il2p_unscramble(buf, total);
			for (int i = 0; i < size; i++) out[len++] = buf[i];
			if (fail) {                                             // il2p.py:456-460 / 487-491
				res.status = IL2P_FAIL; res.end_bit = p - 1; res.corrected = (unsigned int)corrected;
				return;
			}
		}
	}
	if (C.il2p_crc) {                                               // rx_trailing_crc il2p.py:502-518
		if (p + 32 > nb) { res.corrected = (unsigned int)corrected; return; }
		unsigned int crc = 0;
		for (int i = 0; i < 4; i++) crc += (unsigned int)c_hamming[il2p_byte(d, p + 8 * i) & 0x7F] << (12 - 4 * i);
		p += 32;
		out[len++] = (unsigned char)(crc & 0xFF);
		out[len++] = (unsigned char)(crc >> 8);
	} else {
		const unsigned int crc = crc16_x25_il2p(out, len);
		out[len++] = (unsigned char)(crc & 0xFF);
		out[len++] = (unsigned char)(crc >> 8);
	}
	res.status = IL2P_OK; res.end_bit = p - 1; res.len = len; res.corrected = (unsigned int)corrected;
}

// --- 2. speculative decode of every candidate ----------------------------------------
__global__ void __launch_bounds__(64)
il2p_decode_kernel(const BitChain *__restrict__ chains, const ChainCounters *__restrict__ cc,
                   const uint32_t *__restrict__ d, long long bits_stride,
                   const unsigned int *__restrict__ cand_pos, long long cand_stride,
                   const unsigned int *__restrict__ cand_totals, int cand_cap,
                   unsigned char *__restrict__ cand_scratch, long long cand_scratch_stride,
                   Il2pRes *__restrict__ results)
{
	__shared__ unsigned char s_tab[768];
	const int ch = blockIdx.y;
	const BitChain C = chains[ch];
	if (C.codec != 2) return;
	const unsigned int n = min(cand_totals[ch], (unsigned int)cand_cap);
	if (blockIdx.x * blockDim.x >= n) return;                  // (uniform per block: nobody waits at the barrier below)
	for (int i = threadIdx.x; i < 768; i += blockDim.x) s_tab[i] = d_gf_tab[i];
	__syncthreads();
	const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const long long nb = cc[ch].nbytes * 8;
	Il2pRes r;
	il2p_try(s_tab, d + (long long)ch * bits_stride, nb, cand_pos[(long long)ch * cand_stride + j], C,
		cand_scratch + (long long)ch * cand_scratch_stride + (long long)j * IL2P_SLOT, r);
	results[(long long)ch * cand_cap + j] = r;
}

// the reference's working_word when the decoder resumed searching at bit `pos` and has since consumed bits
// pos..g: true stream bits for indices >= lo, the fill pattern below (il2p.py:119 / 147-153)
__device__ unsigned int il2p_window(const uint32_t *__restrict__ d, long long g, long long pos, bool at_start)
{
	unsigned int ww = 0;
	const long long lo = at_start ? 0 : pos - 8;
	for (int k = 0; k < 32; k++) {
		const long long idx = g - k;
		unsigned int bit;
		if (idx >= lo) bit = (d[idx >> 5] >> (idx & 31)) & 1u;
		else bit = at_start ? (idx >= -24 ? 1u : 0u) : 0u;
		ww |= bit << k;
	}
	return ww;
}

__device__ __forceinline__ bool il2p_sync_match(unsigned int ww, int tol)     // il2p.py:369-373
{
	return __popc((ww & 0xFFFFFFu) ^ 0xF15E48u) <= tol || __popc(ww ^ 0x5D57DF7Fu) <= tol;
}

// --- 3. sequential resolution, one warp per chain -------------------------------------
// hand_in / hand_out: where the walk starts and where it stands when it reaches the end of the shard's own range
// (sb[ch].own_hi; the whole stream on an unsharded run).  A frame is processed by the shard that holds its last bit:
// a sync whose frame runs past own_hi (or past the local stream) is left, with the state in front of it, to the next
// shard, which sees the same bits through its hand-off tail.
__global__ void __launch_bounds__(32)
il2p_resolve_kernel(const BitChain *__restrict__ chains, ChainCounters *__restrict__ cc,
                    const uint32_t *__restrict__ dall, long long bits_stride,
                    const unsigned int *__restrict__ cand_pos, long long cand_stride,
                    const unsigned int *__restrict__ cand_totals, int cand_cap,
                    unsigned char *__restrict__ cand_scratch, long long cand_scratch_stride,
                    const Il2pRes *__restrict__ results,
                    const uint32_t *__restrict__ byte_addr, long long addr_stride,
                    uint8_t *__restrict__ scratch, long long scratch_stride,
                    GapRec *__restrict__ gaps, long long gap_stride,
                    const ShardBits *__restrict__ sb, const Il2pHand *__restrict__ hand_in, Il2pHand *__restrict__ hand_out)
{
	const int ch = blockIdx.x;
	const BitChain C = chains[ch];
	if (C.codec != 2) return;
	const int lane = threadIdx.x;
	const uint32_t *d = dall + (long long)ch * bits_stride;
	const long long nb = cc[ch].nbytes * 8;
	const long long own_hi = sb[ch].own_hi;
	const unsigned int *cp = cand_pos + (long long)ch * cand_stride;
	const unsigned int ncand = cand_totals[ch];
	if ((long long)ncand > cand_stride) {                           // sync tolerance so loose the list overflowed
		if (lane == 0) { cc[ch].seq_needed = 2; cc[ch].nflags = 0; }
		return;
	}
	unsigned char *slots = cand_scratch + (long long)ch * cand_scratch_stride;
	unsigned char *tmp = slots + (long long)cand_cap * IL2P_SLOT;        // slot for inline decodes
	uint8_t *dst = scratch + (long long)ch * scratch_stride;
	GapRec *out = gaps + (long long)ch * gap_stride;
	long long pos = hand_in[ch].pos;
	unsigned int mode = hand_in[ch].mode;
	unsigned int ci = 0, nrec = 0, leak = hand_in[ch].leak, out_off = 0;
	while (true) {
		// lane 0 finds the next accepted sync and its decode result
		long long found = -1;
		Il2pRes r;
		const unsigned char *src = nullptr;
		r.status = IL2P_INCOMPLETE; r.len = 0; r.corrected = 0; r.end_bit = nb - 1;
		// right after a frame (mode 1, the state of nearly every iteration): the 32 windows ending at pos .. pos+31 all lie
		// inside stream bits [pos-31, pos+31] -- fetched once (bit j of span = stream bit pos-31+j), what the reference's
		// register no longer holds (everything before the last collected byte, i.e. before pos-8) blanked; lane l tests the
		// window ending at pos+l (pos and mode are the same on every lane)
		long long found1 = -1;
		if (mode == 1 && pos < nb) {
			unsigned long long span = 0;
			const long long b0 = pos - 31;
			for (int q = 0; q < 3; q++) {
				const long long w = (b0 >> 5) + q;               // floor division: b0 may be negative
				if (w < 0 || (w << 5) >= nb) continue;
				const unsigned long long word = d[w];
				const long long sh = (w << 5) - b0;              // position of the word's bit 0 inside span
				if (sh >= 64) continue;
				span |= sh >= 0 ? (word << sh) : (word >> (-sh));
			}
			span &= ~((1ull << 23) - 1ull);                      // indices below pos-8 read as 0
			const bool hit = pos + lane < nb && il2p_sync_match(__brev((unsigned int)(span >> lane)), C.il2p_sync_tol);
			const unsigned int hits = __ballot_sync(0xffffffffu, hit);
			if (hits) found1 = pos + (__ffs(hits) - 1);
		}
		if (lane == 0 && pos < nb) {
			const long long lim = min(pos + 32, nb);
			if (mode == 0) {
				for (long long g = pos; g < lim; g++)
					if (il2p_sync_match(il2p_window(d, g, pos, true), C.il2p_sync_tol)) { found = g; break; }
			} else if (mode == 1) {
				found = found1;
			}
			if (found < 0) {
				const long long from = (mode == 2) ? pos : pos + 32;
				while (ci < ncand && (long long)cp[ci] < from) ci++;
				if (ci < ncand) found = cp[ci];
			}
			if (found >= 0 && found < own_hi) {
				// a speculative result exists when the position is in the candidate list
				while (ci < ncand && (long long)cp[ci] < found) ci++;
				if (ci < ncand && (long long)cp[ci] == found && ci < (unsigned int)cand_cap) {
					r = results[(long long)ch * cand_cap + ci];
					src = slots + (long long)ci * IL2P_SLOT;
				} else {
					il2p_try(d_gf_tab, d, nb, found, C, tmp, r);
					src = tmp;
				}
			}
		}
		found = __shfl_sync(0xffffffffu, found, 0);
		const unsigned int status = __shfl_sync(0xffffffffu, r.status, 0);
		const long long end_bit = __shfl_sync(0xffffffffu, r.end_bit, 0);
		if (found < 0 || found >= own_hi) {
			// nothing more in the own range: the next shard searches on; a resume point closer than 32 bits keeps its fill
			if (lane == 0 && !(mode != 2 && own_hi - pos < 32)) { pos = min(own_hi, nb); mode = 2; }
			break;
		}
		if (status == IL2P_INCOMPLETE || end_bit >= own_hi) break;  // the frame runs past this shard (or past the recording)
		const unsigned int len = __shfl_sync(0xffffffffu, r.len, 0);
		const unsigned int corrected = __shfl_sync(0xffffffffu, r.corrected, 0);
		if (status == IL2P_OK) {
			const unsigned long long sp = __shfl_sync(0xffffffffu, (unsigned long long)src, 0);
			const unsigned char *s = (const unsigned char *)sp;
			for (unsigned int i = lane; i < len; i += 32) dst[out_off + i] = s[i];
			if (lane == 0) {
				GapRec rec;                                         // write_n_search il2p.py:200-211
				rec.emit = 1; rec.len = len; rec.scratch_off = out_off;
				rec.addr = byte_addr[(long long)ch * addr_stride + (end_bit >> 3)];      // il2p.py:364
				rec.corrected = leak + corrected;
				out[nrec] = rec;
			}
			nrec++;
			out_off += len;
			leak = 0;
		} else {
			leak += corrected;                                      // a failed frame's corrections are never cleared
		}
		pos = end_bit + 1;
		mode = 1;
		__syncwarp();
	}
	if (lane == 0) {
		cc[ch].nflags = (int)nrec;
		Il2pHand h;
		h.pos = pos; h.mode = mode; h.leak = leak;
		hand_out[ch] = h;
	}
}

extern "C" cudaError_t pm_launch_il2p(const BitChain *chains, int n_chains, ChainCounters *cc, const uint32_t *d,
	long long bits_stride, const unsigned int *cand_pos, long long cand_stride, const unsigned int *cand_totals,
	int cand_cap, unsigned char *cand_scratch, long long cand_scratch_stride, Il2pRes *results,
	const uint32_t *byte_addr, long long addr_stride, uint8_t *scratch, long long scratch_stride,
	GapRec *gaps, long long gap_stride, const ShardBits *sb, const Il2pHand *hand_in, Il2pHand *hand_out, cudaStream_t st)
{
	dim3 grid((cand_cap + 63) / 64, n_chains);
	pm_kt_mark("il2p_decode_kernel", st);
	il2p_decode_kernel<<<grid, 64, 0, st>>>(chains, cc, d, bits_stride, cand_pos, cand_stride, cand_totals, cand_cap,
		cand_scratch, cand_scratch_stride, results);
	pm_kt_mark("il2p_resolve_kernel", st);
	il2p_resolve_kernel<<<n_chains, 32, 0, st>>>(chains, cc, d, bits_stride, cand_pos, cand_stride, cand_totals,
		cand_cap, cand_scratch, cand_scratch_stride, results, byte_addr, addr_stride, scratch, scratch_stride,
		gaps, gap_stride, sb, hand_in, hand_out);
	return cudaGetLastError();
}

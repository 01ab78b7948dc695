// slicer.cu -- symbol-timing recovery (reference slicer.py:59-107 BinarySlicer,
// slicer.py:193-242 QuadratureSlicer) as a segmented, verified parallel scan.
//
// The reference loop is sequential over every sample:
//     clock += 1.0
//     if clock >= sps/2 - 0.5: clock -= sps; take a bit (sample >= 0)
//     if sign(sample) != sign(previous): clock *= lock_rate
// Its only inputs are the SIGNS of the soft samples, so this kernel reads the
// packed sign bitstream written by the front end.  Each thread owns one
// (chain, segment): it first runs the loop over `warm` samples before its
// segment from a cold state -- every zero crossing contracts the clock error by
// lock_rate, so the state converges to the true one -- and records the state it
// reached at the segment start (S_k), then runs its segment, recording a
// checkpoint state every chk_words words and the end state (E_k).  A second
// kernel checks S_k == E_{k-1} BIT FOR BIT (IEEE double pattern + last signs); a
// segment that fails is re-run from the true state, but only until its state
// equals a stored checkpoint again (from there on the first run was already
// right).  By induction from segment 0 the result is exactly the sequential
// loop's.  All clock arithmetic is IEEE double, one rounding per operation, so
// it is bit-identical to CPython's floats.
//
// Output: one rollover mask bit per sample (bit i of word w = "a symbol was
// taken at sample 32w+i"); bits and byte addresses are produced from
// (sign, mask) by the gather kernels in bits.cu.
#include <algorithm>
#include "pm_common.cuh"

__device__ __forceinline__ bool seg_state_equal(const SegState &a, const SegState &b)
{
	return __double_as_longlong(a.clock) == __double_as_longlong(b.clock) && a.last == b.last &&
	       a.last_q == b.last_q;
}

// One sample of the loop (slicer.py:77-81, 104), arranged so that the compiler cannot put selects of float64 VALUES
// on the dependency chain: the roll-over selects the ADDEND (1.0 or -(sps - 1), resp. 0.0 or sps in the plain form)
// and a zero crossing selects the FACTOR (1.0 or lock_rate) -- x + 1.0, x - 0.0 and x * 1.0 are exact, so the clock
// is bit-identical to the reference's conditional statements -- and the chain per sample is compare, add,
// multiply.  Six or seven instructions instead of eleven (compare x2, add x2, select x2, multiply, select x2, bit
// tests); the kernel is bound by that stream (ncu r01e: ALU pipe 63 % busy at 9 warps per SM).
// A0 / F0: the low word of the roll-over addend (-(sps - 1) in the FAST form, -sps in the plain form) resp. of
// lock_rate is zero (40 or 36.75 samples per symbol, lock_rate 0.75 ...): one 32-bit select of the high word makes
// the operand instead of two.  The roll-over mask is collected from the SIGN of the addend (negative <=> roll) with
// one funnel shift per sample, most recent sample in bit 0: the caller bit-reverses the finished word.
// QUIET: a step of a word without zero crossings (the multiply by 1.0 left out).
template <bool WRITE, bool FAST, bool A0, bool F0, bool QUIET = false>
__device__ __forceinline__ void slicer_step(double &c, uint32_t &m, uint32_t z, uint32_t bit, const SlicerChain &C,
                                            double cs, double a_roll, double a_keep)
{
	bool roll;
	double base;
	if (FAST) {
		// both candidates straight from the old clock (SlicerChain::c_star_bits): roll <=> bits(c) >= c_star <=> c >= cs
		roll = c >= cs;
		base = c;
	} else {
		base = __dadd_rn(c, 1.0);                        // slicer.py:77
		roll = base >= C.thr;                            // slicer.py:79
	}
	// FAST: c + 1.0 or c - (sps - 1); plain: up + 0.0 (exact: up is never -0) or up - sps (slicer.py:81)
	const double add = A0 ? __hiloint2double(roll ? __double2hiint(a_roll) : __double2hiint(a_keep), 0)
	                      : (roll ? a_roll : a_keep);
	const double t = __dadd_rn(base, add);
	if (QUIET) {
		c = t;                                           // no crossing in this word: t * 1.0 == t, bit for bit
	} else {
		const bool cross = (z & bit) != 0;               // slicer.py:99-104
		const double f = F0 ? __hiloint2double(cross ? __double2hiint(C.lock) : 0x3FF00000, 0) : (cross ? C.lock : 1.0);
		c = __dmul_rn(t, f);
	}
	if (WRITE) m = __funnelshift_l((uint32_t)__double2hiint(add), m, 1);
}

// Advance the slicer over samples [w0*32, min(w1*32, nout)) of one chain.
// The sign word of the next iteration is requested before the 32 dependent steps of this one (the loads were 1.2
// stall cycles per issue, ncu r02k).
// SPARSE: the caller runs with few threads of a warp active (verify / sweep: a repair is one thread re-running up to a
// whole segment, and the repairs that do not merge early sit in stretches WITHOUT zero crossings): words without a
// crossing take steps without the lock multiply, a quarter off the dependent chain.  Not for the segments kernel, where
// all lanes are active and would run both forms.
template <bool WRITE, bool FAST, bool A0, bool F0, bool SPARSE>
__device__ __forceinline__ void run_words_t(const SlicerChain &C, const uint32_t *__restrict__ sg,
                                            const uint32_t *__restrict__ sgq, uint32_t *__restrict__ mk,
                                            long long w0, long long w1, SegState &st, QuietState *Q)
{
	const int q_words = (SPARSE && Q) ? C.quiet_words : 0, q_lead = C.quiet_lead;
	double c = st.clock;
	unsigned int last = st.last, last_q = st.last_q;
	const double thr = C.thr, sps = C.sps, lam = C.lock;
	const double cs = __longlong_as_double(C.c_star_bits);
	const double a_roll = FAST ? -C.sps_m1 : -C.sps, a_keep = FAST ? 1.0 : 0.0;
	const long long w_last = (C.nout + 31) >> 5;           // words with at least one valid sample
	if (w1 > w_last) w1 = w_last;
	if (w0 >= w1) return;
	uint32_t s_next = sg[w0], q_next = sgq ? sgq[w0] : 0u;
	for (long long w = w0; w < w1; w++) {
		const long long first = w << 5;
		const uint32_t s = s_next, q = q_next;
		if (w + 1 < w1) {
			s_next = sg[w + 1];
			if (sgq) q_next = sgq[w + 1];
		}
		uint32_t z = s ^ ((s << 1) | last);          // zero crossings (slicer.py:99-102)
		last = s >> 31;
		if (sgq) {
			z |= q ^ ((q << 1) | last_q);
			last_q = q >> 31;
		}
		uint32_t m = 0;
		const long long remain = C.nout - first;
		const int cnt = remain >= 32 ? 32 : (int)remain;
		if (cnt == 32) {
			if (SPARSE && z == 0u) {
				const int slot = q_words ? (int)(w % q_words) : 0;
				if (q_words && Q->run >= q_lead + q_words) {
					// deep inside a stretch without crossings: this word repeats the one quiet_words earlier, clock and mask
					// (SlicerChain::quiet_words)
					c = Q->c[slot];
					m = Q->m[slot];
				} else {
#pragma unroll
					for (int i = 0; i < 32; i++) slicer_step<WRITE, FAST, A0, F0, true>(c, m, z, 1u << i, C, cs, a_roll, a_keep);
					if (WRITE) m = __brev(m);
					if (q_words) { Q->c[slot] = c; Q->m[slot] = m; }
				}
				if (q_words) Q->run++;
			} else {
#pragma unroll
				for (int i = 0; i < 32; i++) slicer_step<WRITE, FAST, A0, F0>(c, m, z, 1u << i, C, cs, a_roll, a_keep);
				if (WRITE) m = __brev(m);
				if (q_words) Q->run = 0;
			}
		} else {
			if (q_words) Q->run = 0;
			for (int i = 0; i < cnt; i++) {
				// one rounding per operation, as in CPython: the compiler must not contract c * lam with the next + 1.0 (--fmad)
				c = __dadd_rn(c, 1.0);                                    // slicer.py:77
				if (c >= thr) { c = __dsub_rn(c, sps); m |= (1u << i); }   // slicer.py:79-81
				if ((z >> i) & 1u) c = __dmul_rn(c, lam);                 // slicer.py:104
			}
			// state after a partial word: last signs are those of sample cnt-1
			last = (s >> (cnt - 1)) & 1u;
			if (sgq) last_q = (q >> (cnt - 1)) & 1u;
		}
		if (WRITE) mk[w] = m;
	}
	st.clock = c;
	st.last = last;
	st.last_q = last_q;
}

// The far end of a warm-up crossing by crossing (no outputs, no claim of exactness: whatever error remains is
// contracted by the exact tail).  Between two crossings the loop only counts -- d steps of +1 with a roll-over
// whenever the clock reaches thr -- which has the closed form below, so a word costs one short iteration per zero
// crossing (0-3 in a 32-sample word of 1200 Bd audio) instead of 32 unrolled steps.
// T = float: round 1's form; 2^-24 of error needs ~75 crossings (16384 samples) of exact tail to vanish.
// T = double [r2]: the closed form differs from the sample-by-sample loop only in how often it rounds (once per run
// instead of once per binade the clock climbs through), i.e. by an ulp or two, which the exact tail loses within a
// few crossings: tools/slicer_warm_sim.py measures 0.4 % of hand-offs not bit-identical with a 4096-sample tail
// (0.2 % with 16384; 28 % for the FP32 form with 8192).
// ONE_WRAP: at least 32 samples per symbol, so the clock rolls over at most once between two crossings of a word.
template <typename T, bool ONE_WRAP>
__device__ __forceinline__ void run_words_far_t(const SlicerChain &C, const uint32_t *__restrict__ sg,
                                                const uint32_t *__restrict__ sgq, long long w0, long long w1, SegState &st)
{
	T c = (T)st.clock;
	unsigned int last = st.last, last_q = st.last_q;
	const T thr = (T)C.thr, sps = (T)C.sps, lam = (T)C.lock, inv_sps = (T)1 / sps;
	auto advance = [&](int d) {                      // d samples without a crossing
		const T u = c + (T)d;
		if (ONE_WRAP) c = u >= thr ? u - sps : u;
		else c = u >= thr ? u - sps * (floor((u - thr) * inv_sps) + (T)1) : u;
	};
	const long long w_full = C.nout >> 5;            // whole words only; the exact tail handles the rest
	if (w1 > w_full) w1 = w_full;
	if (w0 >= w1) return;
	// (a flat loop -- one iteration per crossing or per word, lanes drifting apart -- was tried and is slower: 1.17 ms
	// against 1.08 for the segments kernel, profiles/r02ae_slicer_split.txt)
	uint32_t s_next = sg[w0], q_next = sgq ? sgq[w0] : 0u;
	for (long long w = w0; w < w1; w++) {
		const uint32_t s = s_next, q = q_next;
		if (w + 1 < w1) {
			s_next = sg[w + 1];
			if (sgq) q_next = sgq[w + 1];
		}
		uint32_t z = s ^ ((s << 1) | last);
		last = s >> 31;
		if (sgq) {
			z |= q ^ ((q << 1) | last_q);
			last_q = q >> 31;
		}
		int pos = 0;
		while (z) {
			const int t = __ffs(z);                      // crossing at sample t - 1: count up to and including it, then lock
			z &= z - 1;
			advance(t - pos);
			c *= lam;
			pos = t;
		}
		advance(32 - pos);
	}
	st.clock = (double)c;
	st.last = last;
	st.last_q = last_q;
}

__device__ __forceinline__ void run_words_far(const SlicerChain &C, const uint32_t *__restrict__ sg,
                                              const uint32_t *__restrict__ sgq, long long w0, long long w1, SegState &st,
                                              bool f64)
{
	const bool one = C.sps >= 32.0;                  // uniform per chain
	if (f64) {
		if (one) run_words_far_t<double, true>(C, sg, sgq, w0, w1, st);
		else run_words_far_t<double, false>(C, sg, sgq, w0, w1, st);
	} else {
		if (one) run_words_far_t<float, true>(C, sg, sgq, w0, w1, st);
		else run_words_far_t<float, false>(C, sg, sgq, w0, w1, st);
	}
}

template <bool WRITE, bool SPARSE = false>
__device__ __forceinline__ void run_words(const SlicerChain &C, const uint32_t *__restrict__ sg,
                                          const uint32_t *__restrict__ sgq, uint32_t *__restrict__ mk,
                                          long long w0, long long w1, SegState &st, QuietState *Q = nullptr)
{
	// uniform per chain (blockIdx.y)
	const bool a0 = __double2loint(C.fast ? C.sps_m1 : C.sps) == 0, f0 = __double2loint(C.lock) == 0;
	if (C.fast) {
		if (a0) {
			if (f0) run_words_t<WRITE, true, true, true, SPARSE>(C, sg, sgq, mk, w0, w1, st, Q);
			else run_words_t<WRITE, true, true, false, SPARSE>(C, sg, sgq, mk, w0, w1, st, Q);
		} else {
			if (f0) run_words_t<WRITE, true, false, true, SPARSE>(C, sg, sgq, mk, w0, w1, st, Q);
			else run_words_t<WRITE, true, false, false, SPARSE>(C, sg, sgq, mk, w0, w1, st, Q);
		}
	} else {
		if (a0 && f0) run_words_t<WRITE, false, true, true, SPARSE>(C, sg, sgq, mk, w0, w1, st, Q);
		else run_words_t<WRITE, false, false, false, SPARSE>(C, sg, sgq, mk, w0, w1, st, Q);
	}
}

// grid: (ceil(k_count / 128), n_chains); block 128: segments [k_first, k_first + k_count)
// Segment k covers words [origin_w + k*seg_words, origin_w + (k+1)*seg_words).
// true_start != 0: local sample 0 is the true start of the recording, so a
// warm-up window clipped at 0 starts from init[] (no speculation).
__global__ void __launch_bounds__(128)
slicer_segments_kernel(const SlicerChain *__restrict__ chains, const uint32_t *__restrict__ sign,
                       long long sign_stride, uint32_t *__restrict__ mask, long long mask_stride,
                       SegState *__restrict__ S, SegState *__restrict__ E, SegState *__restrict__ chk,
                       const SegState *__restrict__ init, SlicerGeom G)
{
	const int k = G.k_first + blockIdx.x * blockDim.x + threadIdx.x;
	const int ch = blockIdx.y;
	if (k >= G.k_first + G.k_count || k >= G.n_seg) return;
	const SlicerChain C = chains[ch];
	const uint32_t *sg = sign + (long long)C.sign_row * sign_stride;
	const uint32_t *sgq = C.quadrature ? sign + (long long)C.sign_q_row * sign_stride : nullptr;
	uint32_t *mk = mask + (long long)ch * mask_stride;
	const long long w_begin = G.origin_w + (long long)k * G.seg_words;
	SegState st;
	long long w_warm = w_begin - G.warm_words;
	if (w_warm <= 0 && G.true_start) {
		st = init[ch];                     // the true start state: no speculation
		w_warm = 0;
	} else {
		if (w_warm < 0) w_warm = 0;
		st.clock = 0.0; st.last = 1u; st.last_q = 1u;   // cold start (slicer.py:50,55)
	}
	if (G.warm_f32_words > 0 && w_begin - w_warm > G.warm_f32_words / 4) {
		const long long w_mid = min(w_begin, w_warm + G.warm_f32_words);
		if (((w_mid) << 5) <= C.nout) {
			run_words_far(C, sg, sgq, w_warm, w_mid, st, G.warm_far_f64 != 0);
			w_warm = w_mid;
		}
	}
	run_words<false>(C, sg, sgq, mk, w_warm, w_begin, st);
	const long long idx = (long long)ch * G.n_seg + k;
	S[idx] = st;
	SegState *ck = chk + idx * G.n_chk;
	for (int j = 0; j < G.n_chk; j++) {
		const long long a = w_begin + (long long)j * G.chk_words;
		run_words<true>(C, sg, sgq, mk, a, a + G.chk_words, st);
		ck[j] = st;
	}
	E[idx] = st;
}

// One verification / repair pass.  E_in -> E_out (double buffered so that a thread never reads a neighbour's
// half-written state), in two kernels: slicer_verify_kernel compares every hand-off (one thread per segment) and lists
// the segments that have to be re-run; slicer_repair_kernel re-runs them, ONE WARP per listed segment with lane 0
// working.  A repair is a long sequential chain: 32 of them in one warp would execute each other's paths, and the cheap
// paths of a repair (words without crossings, SlicerChain::quiet_words) only pay when no neighbour is stepping -- a
// recording whose quiet gaps defeat the warm-up (a chain with a large space gain sees hardly a zero crossing in noise)
// has hundreds of repairs per pass in runs of consecutive segments (profiles/r02al_*: verify passes 3.0 -> 2.0 ms on
// afsk_1200.json / IL2P audio).
__global__ void __launch_bounds__(128)
slicer_verify_kernel(const SlicerChain *__restrict__ chains, SegState *__restrict__ S, const SegState *__restrict__ E_in,
                     SegState *__restrict__ E_out, const SegState *__restrict__ init, SlicerGeom G,
                     unsigned int *repairs, const unsigned int *__restrict__ skip_if_zero, unsigned int *__restrict__ list)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	const int ch = blockIdx.y;
	if (k >= G.n_seg) return;
	const long long idx = (long long)ch * G.n_seg + k;
	// passes are enqueued back to back without asking the host: a pass whose predecessor repaired nothing has
	// nothing to do either (every hand-off already verified) and only keeps the double buffer in step
	if (skip_if_zero && *skip_if_zero == 0u) {
		E_out[idx] = E_in[idx];
		return;
	}
	if (k < G.k_init) {                 // history segments before a repaired hand-off: already final
		E_out[idx] = E_in[idx];
		return;
	}
	const SegState prev = (k == G.k_init) ? init[ch] : E_in[idx - 1];
	const SegState mine = S[idx];
	if (seg_state_equal(prev, mine)) {
		E_out[idx] = E_in[idx];
		return;
	}
	const long long w_begin = G.origin_w + (long long)k * G.seg_words;
	S[idx] = prev;
	if ((w_begin << 5) >= chains[ch].nout) {     // empty segment past the end
		E_out[idx] = prev;
		return;
	}
	list[atomicAdd(repairs, 1u)] = (unsigned int)idx;      // at most one entry per segment: the list holds n_chains * n_seg
}

__global__ void __launch_bounds__(128)
slicer_repair_kernel(const SlicerChain *__restrict__ chains, const uint32_t *__restrict__ sign,
                     long long sign_stride, uint32_t *__restrict__ mask, long long mask_stride,
                     const SegState *__restrict__ S, const SegState *__restrict__ E_in, SegState *__restrict__ E_out,
                     SegState *__restrict__ chk, SlicerGeom G, const unsigned int *__restrict__ repairs,
                     const unsigned int *__restrict__ list)
{
	if ((threadIdx.x & 31) != 0) return;
	const unsigned int n = *repairs;
	const unsigned int n_warps = gridDim.x * (blockDim.x >> 5);
	for (unsigned int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += n_warps) {
		const long long idx = list[e];
		const int ch = (int)(idx / G.n_seg), k = (int)(idx - (long long)ch * G.n_seg);
		const SlicerChain C = chains[ch];
		const long long w_begin = G.origin_w + (long long)k * G.seg_words;
		const uint32_t *sg = sign + (long long)C.sign_row * sign_stride;
		const uint32_t *sgq = C.quadrature ? sign + (long long)C.sign_q_row * sign_stride : nullptr;
		uint32_t *mk = mask + (long long)ch * mask_stride;
		SegState st = S[idx];                        // the true start state (slicer_verify_kernel put it there)
		SegState *ck = chk + idx * G.n_chk;
		QuietState Q;
		Q.run = 0;
		bool merged = false;
		for (int j = 0; j < G.n_chk && !merged; j++) {
			const long long a = w_begin + (long long)j * G.chk_words;
			run_words<true, true>(C, sg, sgq, mk, a, a + G.chk_words, st, &Q);
			if (seg_state_equal(ck[j], st)) merged = true;     // merged with the first run: the rest is already right
			else ck[j] = st;
		}
		E_out[idx] = merged ? E_in[idx] : st;
	}
}

// Sequential fallback: one thread per chain walks its segments in order and repairs
// every hand-off that does not verify.  Used when the parallel verify passes do not
// converge quickly (stretches without zero crossings -- e.g. digital silence -- never
// contract the state error, so every speculated start state in them is wrong).
__global__ void slicer_sweep_kernel(const SlicerChain *__restrict__ chains, const uint32_t *__restrict__ sign,
                                    long long sign_stride, uint32_t *__restrict__ mask, long long mask_stride,
                                    SegState *__restrict__ S, SegState *__restrict__ E, SegState *__restrict__ chk,
                                    const SegState *__restrict__ init, SlicerGeom G, unsigned int *repairs)
{
	const int ch = blockIdx.x;
	if (threadIdx.x != 0) return;
	const SlicerChain C = chains[ch];
	const uint32_t *sg = sign + (long long)C.sign_row * sign_stride;
	const uint32_t *sgq = C.quadrature ? sign + (long long)C.sign_q_row * sign_stride : nullptr;
	uint32_t *mk = mask + (long long)ch * mask_stride;
	SegState prev = init[ch];
	unsigned int fixed = 0;
	QuietState Q;                          // (segments that verify are skipped: their words were not seen)
	Q.run = 0;
	for (int k = G.k_init; k < G.n_seg; k++) {
		const long long idx = (long long)ch * G.n_seg + k;
		const long long w_begin = G.origin_w + (long long)k * G.seg_words;
		if (!seg_state_equal(S[idx], prev)) {
			S[idx] = prev;
			if ((w_begin << 5) >= C.nout) {
				E[idx] = prev;
			} else {
				fixed++;
				SegState st = prev;
				SegState *ck = chk + idx * G.n_chk;
				bool merged = false;
				for (int j = 0; j < G.n_chk && !merged; j++) {
					const long long a = w_begin + (long long)j * G.chk_words;
					run_words<true, true>(C, sg, sgq, mk, a, a + G.chk_words, st, &Q);
					if (seg_state_equal(ck[j], st)) merged = true; else ck[j] = st;
				}
				if (!merged) E[idx] = st;
				else Q.run = 0;                    // the rest of the segment was not walked
			}
		} else {
			Q.run = 0;                             // nor was this segment
		}
		prev = E[idx];
	}
	if (fixed) atomicAdd(repairs, fixed);
}

// Symbols (mask bits) of every chain in samples [w0*32, min(w1*32, nout)); out[] must be zeroed.
__global__ void __launch_bounds__(256)
slicer_count_kernel(const SlicerChain *__restrict__ chains, const uint32_t *__restrict__ mask, long long mask_stride,
                    long long w0, long long w1, unsigned long long *__restrict__ out)
{
	const int ch = blockIdx.y;
	const long long nout = chains[ch].nout;
	const uint32_t *mk = mask + (long long)ch * mask_stride;
	unsigned int cnt = 0;
	for (long long w = w0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += (long long)gridDim.x * blockDim.x) {
		const long long first = w << 5;
		if (first >= nout) break;
		uint32_t m = mk[w];
		const long long remain = nout - first;
		if (remain < 32) m &= (1u << (int)remain) - 1u;
		cnt += __popc(m);
	}
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&out[ch], (unsigned long long)cnt);
}

// soft values -> the packed sign stream the slicer reads (bit = sample >= 0.0: slicer.py:85, 99-102; NaN counts as negative,
// as it does in the reference's comparisons).  For pm_engine_slice_soft.
__global__ void __launch_bounds__(256) soft_sign_kernel(const double *__restrict__ x, long long n, uint32_t *__restrict__ out)
{
	const long long words = (n + 31) >> 5;
	const int lane = threadIdx.x & 31;
	for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < words;
	     w += (long long)gridDim.x * (blockDim.x >> 5)) {
		const long long i = (w << 5) + lane;
		const unsigned int m = __ballot_sync(0xffffffffu, i < n && x[i] >= 0.0);
		if (lane == 0) out[w] = m;
	}
}

extern "C" cudaError_t pm_launch_soft_signs(const double *x, long long n, uint32_t *out, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	const long long words = (n + 31) >> 5;
	const int blocks = (int)std::min<long long>((words + 7) / 8, 148 * 8);
	pm_kt_mark("soft_sign_kernel", st);
	soft_sign_kernel<<<blocks, 256, 0, st>>>(x, n, out);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_slicer_segments(const SlicerChain *chains, int n_chains, const uint32_t *sign,
	long long sign_stride, uint32_t *mask, long long mask_stride, SegState *S, SegState *E, SegState *chk,
	const SegState *init, SlicerGeom G, cudaStream_t st)
{
	if (G.k_count <= 0) return cudaSuccess;
	dim3 grid((G.k_count + 127) / 128, n_chains);
	pm_kt_mark("slicer_segments_kernel", st);
	slicer_segments_kernel<<<grid, 128, 0, st>>>(chains, sign, sign_stride, mask, mask_stride, S, E, chk, init, G);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_slicer_verify(const SlicerChain *chains, int n_chains, const uint32_t *sign,
	long long sign_stride, uint32_t *mask, long long mask_stride, SegState *S, const SegState *E_in,
	SegState *E_out, SegState *chk, const SegState *init, SlicerGeom G, unsigned int *repairs,
	const unsigned int *skip_if_zero, unsigned int *list, cudaStream_t st)
{
	dim3 grid((G.n_seg + 127) / 128, n_chains);
	pm_kt_mark("slicer_verify_kernel", st);
	slicer_verify_kernel<<<grid, 128, 0, st>>>(chains, S, E_in, E_out, init, G, repairs, skip_if_zero, list);
	cudaError_t ce = cudaGetLastError();
	if (ce != cudaSuccess) return ce;
	pm_kt_mark("slicer_repair_kernel", st);
	slicer_repair_kernel<<<148 * 4, 128, 0, st>>>(chains, sign, sign_stride, mask, mask_stride, S, E_in, E_out, chk, G,
		repairs, list);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_slicer_sweep(const SlicerChain *chains, int n_chains, const uint32_t *sign,
	long long sign_stride, uint32_t *mask, long long mask_stride, SegState *S, SegState *E, SegState *chk,
	const SegState *init, SlicerGeom G, unsigned int *repairs, cudaStream_t st)
{
	pm_kt_mark("slicer_sweep_kernel", st);
	slicer_sweep_kernel<<<n_chains, 32, 0, st>>>(chains, sign, sign_stride, mask, mask_stride, S, E, chk, init, G,
		repairs);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_slicer_count(const SlicerChain *chains, int n_chains, const uint32_t *mask,
	long long mask_stride, long long w0, long long w1, unsigned long long *out, cudaStream_t st)
{
	cudaMemsetAsync(out, 0, sizeof(unsigned long long) * n_chains, st);
	long long nb = (w1 - w0 + 256 * 16 - 1) / (256 * 16);
	if (nb < 1) nb = 1;
	if (nb > 1184) nb = 1184;
	dim3 grid((unsigned int)nb, n_chains);
	pm_kt_mark("slicer_count_kernel", st);
	slicer_count_kernel<<<grid, 256, 0, st>>>(chains, mask, mask_stride, w0, w1, out);
	return cudaGetLastError();
}

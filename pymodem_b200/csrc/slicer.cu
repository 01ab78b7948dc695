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
// reached at the segment start (S_k), then runs its segment and records the end
// state (E_k).  A second kernel checks S_k == E_{k-1} BIT FOR BIT (IEEE double
// pattern + last sign); a segment that fails is re-run from the true state.
// By induction from segment 0 the result is exactly the sequential loop's.
// All clock arithmetic is IEEE double, one rounding per operation, so it is
// bit-identical to CPython's floats.
//
// Output: one rollover mask bit per sample (bit i of word w = "a symbol was
// taken at sample 32w+i"); bits and byte addresses are produced from
// (sign, mask) by the gather kernels in bits.cu.
#include "pm_common.cuh"



__device__ __forceinline__ bool seg_state_equal(const SegState &a, const SegState &b)
{
	return __double_as_longlong(a.clock) == __double_as_longlong(b.clock) && a.last == b.last &&
	       a.last_q == b.last_q;
}

// Advance the slicer over samples [w0*32, min(w1*32, nout)) of one chain.
template <bool WRITE>
__device__ __forceinline__ void run_words(const SlicerChain &C, const uint32_t *__restrict__ sg,
                                          const uint32_t *__restrict__ sgq, uint32_t *__restrict__ mk,
                                          long long w0, long long w1, SegState &st)
{
	double c = st.clock;
	unsigned int last = st.last, last_q = st.last_q;
	const double thr = C.thr, sps = C.sps, lam = C.lock;
	for (long long w = w0; w < w1; w++) {
		const long long first = w << 5;
		if (first >= C.nout) break;
		const uint32_t s = sg[w];
		uint32_t z = s ^ ((s << 1) | last);          // zero crossings (slicer.py:99-102)
		last = s >> 31;
		if (sgq) {
			const uint32_t q = sgq[w];
			z |= q ^ ((q << 1) | last_q);
			last_q = q >> 31;
		}
		uint32_t m = 0;
		const long long remain = C.nout - first;
		if (remain >= 32) {
#pragma unroll
			for (int i = 0; i < 32; i++) {
				c += 1.0;                                // slicer.py:77
				if (c >= thr) { c -= sps; m |= (1u << i); }   // slicer.py:79-81
				if ((z >> i) & 1u) c *= lam;             // slicer.py:104
			}
		} else {
			const int cnt = (int)remain;
			for (int i = 0; i < cnt; i++) {
				c += 1.0;
				if (c >= thr) { c -= sps; m |= (1u << i); }
				if ((z >> i) & 1u) c *= lam;
			}
			// state after a partial word: last signs are those of sample cnt-1
			last = (s >> (cnt - 1)) & 1u;
			if (sgq) last_q = (sgq[w] >> (cnt - 1)) & 1u;
		}
		if (WRITE) mk[w] = m;
	}
	st.clock = c;
	st.last = last;
	st.last_q = last_q;
}

// grid: (ceil(n_seg / 128), n_chains); block 128
__global__ void __launch_bounds__(128)
slicer_segments_kernel(const SlicerChain *__restrict__ chains, const uint32_t *__restrict__ sign,
                       long long sign_stride, uint32_t *__restrict__ mask, long long mask_stride,
                       SegState *__restrict__ S, SegState *__restrict__ E, const SegState *__restrict__ init,
                       int n_seg, int seg_words, int warm_words)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	const int ch = blockIdx.y;
	if (k >= n_seg) return;
	const SlicerChain C = chains[ch];
	const uint32_t *sg = sign + (long long)C.sign_row * sign_stride;
	const uint32_t *sgq = C.quadrature ? sign + (long long)C.sign_q_row * sign_stride : nullptr;
	uint32_t *mk = mask + (long long)ch * mask_stride;
	const long long w_begin = (long long)k * seg_words;
	const long long w_end = w_begin + seg_words;
	SegState st;
	long long w_warm = w_begin - warm_words;
	if (w_warm <= 0) {
		st = init[ch];                     // the true start state: no speculation
		w_warm = 0;
	} else {
		st.clock = 0.0; st.last = 1u; st.last_q = 1u;   // cold start (slicer.py:50,55)
	}
	run_words<false>(C, sg, sgq, mk, w_warm, w_begin, st);
	S[(long long)ch * n_seg + k] = st;
	run_words<true>(C, sg, sgq, mk, w_begin, w_end, st);
	E[(long long)ch * n_seg + k] = st;
}

// One verification / repair pass.  E_in -> E_out (double buffered so that a
// thread never reads a neighbour's half-written state).
__global__ void __launch_bounds__(128)
slicer_verify_kernel(const SlicerChain *__restrict__ chains, const uint32_t *__restrict__ sign,
                     long long sign_stride, uint32_t *__restrict__ mask, long long mask_stride,
                     SegState *__restrict__ S, const SegState *__restrict__ E_in, SegState *__restrict__ E_out,
                     const SegState *__restrict__ init, int n_seg, int seg_words, unsigned int *repairs)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	const int ch = blockIdx.y;
	if (k >= n_seg) return;
	const long long idx = (long long)ch * n_seg + k;
	const SegState prev = (k == 0) ? init[ch] : E_in[idx - 1];
	const SegState mine = S[idx];
	if (seg_state_equal(prev, mine)) {
		E_out[idx] = E_in[idx];
		return;
	}
	const SlicerChain C = chains[ch];
	if ((long long)k * seg_words * 32 >= C.nout) {     // empty segment past the end
		S[idx] = prev;
		E_out[idx] = prev;
		return;
	}
	atomicAdd(repairs, 1u);
	const uint32_t *sg = sign + (long long)C.sign_row * sign_stride;
	const uint32_t *sgq = C.quadrature ? sign + (long long)C.sign_q_row * sign_stride : nullptr;
	uint32_t *mk = mask + (long long)ch * mask_stride;
	SegState st = prev;
	S[idx] = prev;
	run_words<true>(C, sg, sgq, mk, (long long)k * seg_words, (long long)(k + 1) * seg_words, st);
	E_out[idx] = st;
}

extern "C" cudaError_t pm_launch_slicer_segments(const SlicerChain *chains, int n_chains, const uint32_t *sign,
	long long sign_stride, uint32_t *mask, long long mask_stride, SegState *S, SegState *E, const SegState *init,
	int n_seg, int seg_words, int warm_words, cudaStream_t st)
{
	dim3 grid((n_seg + 127) / 128, n_chains);
	slicer_segments_kernel<<<grid, 128, 0, st>>>(chains, sign, sign_stride, mask, mask_stride, S, E, init, n_seg,
		seg_words, warm_words);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_slicer_verify(const SlicerChain *chains, int n_chains, const uint32_t *sign,
	long long sign_stride, uint32_t *mask, long long mask_stride, SegState *S, const SegState *E_in,
	SegState *E_out, const SegState *init, int n_seg, int seg_words, unsigned int *repairs, cudaStream_t st)
{
	dim3 grid((n_seg + 127) / 128, n_chains);
	slicer_verify_kernel<<<grid, 128, 0, st>>>(chains, sign, sign_stride, mask, mask_stride, S, E_in, E_out, init,
		n_seg, seg_words, repairs);
	return cudaGetLastError();
}

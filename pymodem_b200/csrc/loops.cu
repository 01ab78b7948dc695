// loops.cu -- the float64 pipeline: modems whose demod contains a recursion.
//
//   BPSKModem.demod      psk.py:162-195       BPF FIR -> AGC -> Costas loop -> RRC FIR
//   MPSKModem.demod      psk.py:705-773       BPF FIR -> AGC -> Hilbert FIR / delay -> decision-directed loop
//                                             -> RRC FIR on I and on Q
//   AFSKPLLModem.demod   afsk_pll.py:140-170  BPF FIR -> AGC -> PLL -> LPF FIR
//   AFSKModem.demod      afsk.py:148-167      (only for tone pairs so close that |mark| - |space| cancels
//                                             below what FP32 resolves, e.g. the '300' preset 1695/1705 Hz)
//
// The loops feed quantised decisions back into themselves (256-entry NCO wavetable index, 64x64 integer
// phase-error table, round() of the frequency control), so a sample that differs from the reference's
// float64 value in the 7th digit would flip an index a few times per million samples and leave a visibly
// different trajectory behind.  Everything that feeds a loop is therefore float64, and inside a loop every
// operation is a separately rounded IEEE double operation in the reference's evaluation order (explicit
// __dmul_rn/__dadd_rn, no FMA contraction), which is what CPython floats do.  The loops are sequential per
// chain: one warp per chain, lanes stage samples through shared memory with coalesced loads/stores and lane
// 0 runs the recurrence ("parity mode" of SURVEY.md Appendix C: a segmented loop never becomes bit-identical).
// The FIR stages around them are ordinary data-parallel float64 FIRs.
#include "pm_common.cuh"

#define KIND_AFSK 1
#define KIND_BPSK 3
#define KIND_MPSK 4
#define KIND_PLL 5

// ---- float64 FIR tile: out[o] = sum_j hr[j] * x[o + j] ------------------------------------------------
// A CTA produces P64_TILE consecutive outputs; thread t owns outputs t, t+256, t+512, t+768 of the tile so
// that shared-memory reads are conflict-free and a warp's 32 results of one row pack into one sign word.
template <typename IN>
__device__ __forceinline__ void p64_stage_tile(const IN *__restrict__ in, long long in_len, long long base, int count,
                                               double *__restrict__ s_x)
{
	for (int i = threadIdx.x; i < count; i += P64_THREADS) {
		const long long k = base + i;
		s_x[i] = (k < in_len) ? (double)in[k] : 0.0;
	}
}

__device__ __forceinline__ void p64_stage_taps(const double *__restrict__ taps, int n, double *__restrict__ s_h)
{
	for (int i = threadIdx.x; i < n; i += P64_THREADS) s_h[i] = taps[i];
}

__device__ __forceinline__ void p64_fir4(const double *__restrict__ s_x, const double *__restrict__ s_h, int M,
                                         double acc[4])
{
	acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
	const double *x = s_x + threadIdx.x;
	for (int j = 0; j < M; j++) {
		const double h = s_h[j];
#pragma unroll
		for (int r = 0; r < 4; r++) acc[r] = fma(h, x[r * P64_THREADS + j], acc[r]);
	}
}

// input band-pass: int16 audio -> A                                      psk.py:165 / 710, afsk_pll.py:143, afsk.py:151
__global__ void __launch_bounds__(P64_THREADS)
p64_bpf_kernel(const P64Chain *__restrict__ chains, const int16_t *__restrict__ audio)
{
	extern __shared__ __align__(16) double sm[];
	const P64Chain C = chains[blockIdx.y];
	const long long base = (long long)blockIdx.x * P64_TILE;
	if (base >= C.L1) return;
	double *s_h = sm, *s_x = sm + C.n_bpf;
	p64_stage_taps(C.bpf, C.n_bpf, s_h);
	p64_stage_tile(audio + C.audio_off, C.n_audio, base, P64_TILE + C.n_bpf - 1, s_x);
	__syncthreads();
	double acc[4];
	p64_fir4(s_x, s_h, C.n_bpf, acc);
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const long long o = base + r * P64_THREADS + threadIdx.x;
		if (o < C.L1) C.A[o] = acc[r];
	}
}

// max(buffer) for AGC.normal (agc.py:67): doubles ordered as unsigned integers
__device__ __forceinline__ unsigned long long p64_order_key(double d)
{
	const long long b = __double_as_longlong(d);
	return b >= 0 ? ((unsigned long long)b | 0x8000000000000000ull) : ~(unsigned long long)b;
}

__device__ __forceinline__ double p64_order_value(unsigned long long k)
{
	return (k & 0x8000000000000000ull) ? __longlong_as_double((long long)(k & 0x7FFFFFFFFFFFFFFFull))
	                                    : __longlong_as_double((long long)~k);
}

__global__ void __launch_bounds__(P64_THREADS)
p64_max_kernel(const P64Chain *__restrict__ chains)
{
	__shared__ unsigned long long s_k[P64_THREADS / 32];
	const P64Chain C = chains[blockIdx.y];
	if (C.kind == KIND_AFSK) return;
	unsigned long long k = 0;
	for (long long i = (long long)blockIdx.x * P64_THREADS + threadIdx.x; i < C.L1; i += (long long)gridDim.x * P64_THREADS)
		k = max(k, p64_order_key(C.A[i]));
	for (int o = 16; o; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
	if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = k;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < P64_THREADS / 32; w++) k = max(k, s_k[w]);
		if (k) atomicMax(C.max_slot, k);
	}
}

// middle stage.  AFSK: B = sqrt(mi^2 + mq^2) - sqrt(si^2 + sq^2) (afsk.py:153-162, each numpy op rounds once);
// MPSK: B = Hilbert FIR of A (psk.py:714); the real branch is the pure delay A[d + k] (psk.py:715-716).
__global__ void __launch_bounds__(P64_THREADS)
p64_mid_kernel(const P64Chain *__restrict__ chains)
{
	extern __shared__ __align__(16) double sm[];
	const P64Chain C = chains[blockIdx.y];
	if (C.kind != KIND_AFSK && C.kind != KIND_MPSK) return;
	const long long base = (long long)blockIdx.x * P64_TILE;
	if (base >= C.L2) return;
	const int M = C.n_mid;
	double *s_x = sm, *s_h = sm + (P64_TILE + M - 1);
	p64_stage_tile(C.A, C.L1, base, P64_TILE + M - 1, s_x);
	if (C.kind == KIND_MPSK) {
		p64_stage_taps(C.mid0, M, s_h);
		__syncthreads();
		double acc[4];
		p64_fir4(s_x, s_h, M, acc);
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const long long o = base + r * P64_THREADS + threadIdx.x;
			if (o < C.L2) C.B[o] = acc[r];
		}
		return;
	}
	p64_stage_taps(C.mid0, M, s_h);
	p64_stage_taps(C.mid1, M, s_h + M);
	p64_stage_taps(C.mid2, M, s_h + 2 * M);
	p64_stage_taps(C.mid3, M, s_h + 3 * M);
	__syncthreads();
	double mi[4] = {0, 0, 0, 0}, mq[4] = {0, 0, 0, 0}, si[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
	const double *x = s_x + threadIdx.x;
	for (int j = 0; j < M; j++) {
		const double h0 = s_h[j], h1 = s_h[M + j], h2 = s_h[2 * M + j], h3 = s_h[3 * M + j];
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const double v = x[r * P64_THREADS + j];
			mi[r] = fma(h0, v, mi[r]); mq[r] = fma(h1, v, mq[r]);
			si[r] = fma(h2, v, si[r]); sq[r] = fma(h3, v, sq[r]);
		}
	}
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const long long o = base + r * P64_THREADS + threadIdx.x;
		if (o < C.L2) {
			const double m = __dsqrt_rn(__dadd_rn(__dmul_rn(mi[r], mi[r]), __dmul_rn(mq[r], mq[r])));
			const double s = __dsqrt_rn(__dadd_rn(__dmul_rn(si[r], si[r]), __dmul_rn(sq[r], sq[r])));
			C.B[o] = __dsub_rn(m, s);
		}
	}
}

// output FIR (RRC / low-pass) -> sign bits (+ float soft values for the parity tests).  blockIdx.z = 1 is the
// Q branch of an MPSK chain (psk.py:750-751).
__global__ void __launch_bounds__(P64_THREADS)
p64_out_kernel(const P64Chain *__restrict__ chains, uint32_t *__restrict__ sign, long long sign_stride,
               float *__restrict__ soft, long long soft_stride)
{
	extern __shared__ __align__(16) double sm[];
	const P64Chain C = chains[blockIdx.y];
	const int comp = blockIdx.z;
	if (comp == 1 && C.kind != KIND_MPSK) return;
	const long long base = (long long)blockIdx.x * P64_TILE;
	if (base >= ((C.L3 + P64_TILE - 1) / P64_TILE) * P64_TILE || (base >> 5) >= sign_stride) return;
	const double *in = (C.kind == KIND_MPSK) ? (comp ? C.D : C.C) : C.B;
	const int M = C.n_out;
	double *s_h = sm, *s_x = sm + M;
	p64_stage_taps(C.out_taps, M, s_h);
	p64_stage_tile(in, C.L2, base, P64_TILE + M - 1, s_x);
	__syncthreads();
	double acc[4];
	p64_fir4(s_x, s_h, M, acc);
	const int row = comp ? C.sign_q_row : C.sign_row;
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const long long o = base + r * P64_THREADS + threadIdx.x;
		const bool valid = o < C.L3;
		const uint32_t word = __ballot_sync(0xffffffffu, valid && acc[r] >= 0.0);   // slicer.py:85 sample >= 0
		if ((threadIdx.x & 31) == 0 && (o >> 5) < sign_stride) sign[(long long)row * sign_stride + (o >> 5)] = word;
		if (soft && valid) soft[(long long)row * soft_stride + o] = (float)acc[r];
	}
}

// ---- the recursions ------------------------------------------------------------------------------------
struct LoopState {
	double envelope, sustain_count;                 // agc.py:18-24
	double attack_step, decay_step;                 // scaled rate * normal (agc.py:29, 34)
	double phase, control, sine, cosine;            // nco.py:22-26
	double x1, y1;                                  // iir.py:33-34
	double integral, proportional;                  // pi_control.py:12-13
};

// AGC.peak_detect -- agc.py:26-37: the envelope after this sample.  Only this part is a recurrence; the scaling
// line of AGC.apply (agc.py:72-76, a division) is applied afterwards by all 32 lanes at once (agc_scale).
__device__ __forceinline__ double agc_envelope(const LoopConst &L, LoopState &s, double sample)
{
	const double compare_value = fabs(sample);
	if (compare_value > s.envelope) {
		s.envelope = __dadd_rn(s.envelope, s.attack_step);
		if (s.envelope > compare_value) s.envelope = compare_value;
		s.sustain_count = 0.0;
	}
	if (s.sustain_count >= L.agc_sustain_time) {
		s.envelope = __dsub_rn(s.envelope, s.decay_step);
		if (s.envelope < 0.0) s.envelope = 0.0;
	}
	s.sustain_count = __dadd_rn(s.sustain_count, L.agc_sustain_increment);
	return s.envelope;
}

__device__ __forceinline__ double agc_scale(const LoopConst &L, double sample, double envelope)
{
	return envelope != 0.0 ? __ddiv_rn(__dmul_rn(L.agc_target, sample), envelope) : sample;
}

// NCO.update -- nco.py:34-53 (an index of wavetable_size raises IndexError there: the old sine is kept)
__device__ __forceinline__ void nco_step(const LoopConst &L, LoopState &s, const double *__restrict__ wt, int wt_size)
{
	s.phase = __dadd_rn(s.phase, __dmul_rn(L.nco_phase_scale, __dadd_rn(L.nco_set_frequency, s.control)));
	while (s.phase >= L.nco_two_pi) s.phase = __dsub_rn(s.phase, L.nco_two_pi);
	while (s.phase < 0.0) s.phase = __dadd_rn(s.phase, L.nco_two_pi);
	const int si = __double2int_rz(__dmul_rn(s.phase, L.nco_index_scale));
	if (si < wt_size) s.sine = wt[si];
	int ci = __double2int_rz(__dadd_rn((double)si, L.nco_quarter));
	while (ci >= wt_size) ci -= wt_size;
	while (ci < 0) ci += wt_size;
	s.cosine = wt[ci];
}

// IIR_1.update -- iir.py:38-54
__device__ __forceinline__ double iir_step(const LoopConst &L, LoopState &s, double sample)
{
	double v = __dadd_rn(0.0, __dmul_rn(sample, L.iir_b0));
	v = __dadd_rn(v, __dmul_rn(s.x1, L.iir_b1));
	v = __dadd_rn(v, __dmul_rn(s.y1, L.iir_a1));
	s.x1 = sample;
	s.y1 = v;
	return v;
}

// PI_control.update_saturate -- pi_control.py:25-33
__device__ __forceinline__ double pi_step(const LoopConst &L, LoopState &s, double gain_p, double sample)
{
	s.proportional = __dmul_rn(gain_p, sample);                                  // (gain * p_rate) * sample
	s.integral = __dadd_rn(s.integral, __dmul_rn(L.pi_gain, __dmul_rn(L.pi_i, sample)));
	if (s.integral > L.pi_limit) s.integral = L.pi_limit;
	if (s.integral < -L.pi_limit) s.integral = -L.pi_limit;
	return __dadd_rn(s.proportional, s.integral);
}

// PhaseDetector.get_qpsk_angle_error -- phase_detector.py:124-149.  floor(x * granularity * 0.5): both factors are
// powers of two, so one multiplication by granularity/2 gives the identical double; the floor and the conversion are
// one saturating F2I (the clips at +-granularity follow anyway).  The table is kept as doubles: its value goes
// straight into the loop filter.
__device__ __forceinline__ double pd_qpsk_error(const double *__restrict__ tab, int g, double half_g, double re, double im)
{
	int real = __double2int_rd(__dmul_rn(re, half_g)), imag = __double2int_rd(__dmul_rn(im, half_g));
	real = min(real, g - 1);
	imag = min(imag, g - 1);
	if (real <= -g) real = -(g - 1);
	if (imag <= -g) imag = -(g - 1);
	if (real >= 0) return imag >= 0 ? tab[real * g + imag] : tab[(-imag) * g + real];
	return imag >= 0 ? tab[imag * g + (-real)] : tab[(-real) * g + (-imag)];
}

#define SEQ_CHUNK 128
#define SEQ_MAX_WT 1024
#define SEQ_MAX_PD (64 * 64)

// pass 0: BPSK: AGC + Costas loop, A -> B.  PLL: AGC + PLL, A -> B.  MPSK: AGC in place on A.
// pass 1: MPSK decision-directed loop, (A[d + k], B[k]) -> (C[k], D[k]).
// Two warps per chain.  Warp 1 feeds: it loads the next chunk and, in pass 0, runs the AGC on it (envelope recurrence
// on its lane 0, divisions on all lanes).  Warp 0 consumes the previous chunk: its lane 0 runs the carrier loop, its
// lanes store the results.  The chunks are double buffered, so the AGC costs nothing on top of the loop.
__global__ void __launch_bounds__(64)
p64_seq_kernel(const P64Chain *__restrict__ chains, int pass)
{
	__shared__ double s_wt[SEQ_MAX_WT];
	__shared__ double s_pd[SEQ_MAX_PD];
	__shared__ double s_feed0[2][SEQ_CHUNK], s_feed1[2][SEQ_CHUNK], s_env[SEQ_CHUNK], s_res0[SEQ_CHUNK], s_res1[SEQ_CHUNK];
	const P64Chain C = chains[blockIdx.x];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (C.kind == KIND_AFSK) return;
	if (pass == 1 && C.kind != KIND_MPSK) return;
	const LoopConst L = C.lc;
	for (int i = threadIdx.x; i < C.wt_size; i += 64) s_wt[i] = C.wavetable[i];
	if (C.kind == KIND_MPSK && pass == 1)
		for (int i = threadIdx.x; i < C.pd_g * C.pd_g; i += 64) s_pd[i] = (double)C.pd_table[i];
	LoopState s;
	s.envelope = 0.0; s.sustain_count = 0.0;
	const double normal = p64_order_value(*C.max_slot);
	s.attack_step = __dmul_rn(L.agc_scaled_attack, normal);
	s.decay_step = __dmul_rn(L.agc_scaled_decay, normal);
	s.phase = 0.0; s.control = 0.0; s.sine = 0.0; s.cosine = 0.0;
	s.x1 = 0.0; s.y1 = 0.0;
	s.integral = L.pi_integral0; s.proportional = 0.0;
	const double gain_p = __dmul_rn(L.pi_gain, L.pi_p);
	const double half_g = 0.5 * (double)C.pd_g;
	const long long n = (pass == 0) ? C.L1 : C.L2;
	const double *src0 = (pass == 0) ? C.A : C.A + C.mid_delay;
	const double *src1 = C.B;
	double *dst0 = (pass == 0) ? (C.kind == KIND_MPSK ? C.A : C.B) : C.C;
	double *dst1 = C.D;
	const bool two = (pass == 1);
	const long long n_chunks = (n + SEQ_CHUNK - 1) / SEQ_CHUNK;

	auto feed = [&](long long k) {            // warp 1: chunk k -> buffer k & 1
		const int bsel = (int)(k & 1);
		const long long c0 = k * SEQ_CHUNK;
		const int cnt = (int)min((long long)SEQ_CHUNK, n - c0);
		for (int i = lane; i < cnt; i += 32) {
			s_feed0[bsel][i] = src0[c0 + i];
			if (two) s_feed1[bsel][i] = src1[c0 + i];
		}
		if (pass == 0) {
			__syncwarp();
			if (lane == 0)
				for (int i = 0; i < cnt; i++) s_env[i] = agc_envelope(L, s, s_feed0[bsel][i]);
			__syncwarp();
			for (int i = lane; i < cnt; i += 32) s_feed0[bsel][i] = agc_scale(L, s_feed0[bsel][i], s_env[i]);
		}
	};

	__syncthreads();
	if (wid == 1 && n_chunks > 0) feed(0);
	__syncthreads();
	for (long long k = 0; k < n_chunks; k++) {
		if (wid == 1) {
			if (k + 1 < n_chunks) feed(k + 1);
		} else {
			const int bsel = (int)(k & 1);
			const long long c0 = k * SEQ_CHUNK;
			const int cnt = (int)min((long long)SEQ_CHUNK, n - c0);
			const double *in0 = s_feed0[bsel], *in1 = s_feed1[bsel];
			if (pass == 0 && C.kind == KIND_MPSK) {
				for (int i = lane; i < cnt; i += 32) dst0[c0 + i] = in0[i];          // AGC only (psk.py:713)
			} else {
				if (lane == 0) {
					if (pass == 0 && C.kind == KIND_BPSK) {                         // psk.py:173-189
						for (int i = 0; i < cnt; i++) {
							const double sample = in0[i];
							nco_step(L, s, s_wt, C.wt_size);
							const double i_mixer = __dmul_rn(sample, s.cosine);
							const double q_mixer = __dmul_rn(sample, -s.sine);
							const double f = iir_step(L, s, __dmul_rn(i_mixer, q_mixer));
							s.control = pi_step(L, s, gain_p, f);
							s_res0[i] = i_mixer;
						}
					} else if (pass == 0) {                                         // afsk_pll.py:152-165
						for (int i = 0; i < cnt; i++) {
							const double sample = in0[i];
							nco_step(L, s, s_wt, C.wt_size);
							const double f = iir_step(L, s, __dmul_rn(sample, s.sine));
							s.control = pi_step(L, s, gain_p, f);
							s_res0[i] = s.proportional;
						}
					} else {                                                        // psk.py:733-746
						for (int i = 0; i < cnt; i++) {
							nco_step(L, s, s_wt, C.wt_size);
							const double c_re = s.cosine, c_im = -s.sine;
							const double re = in0[i], im = in1[i];
							const double real = __dsub_rn(__dmul_rn(re, c_re), __dmul_rn(im, c_im));     // complexmath.py:15-19
							const double imag = __dadd_rn(__dmul_rn(c_re, im), __dmul_rn(re, c_im));
							const double f = iir_step(L, s, pd_qpsk_error(s_pd, C.pd_g, half_g, real, imag));
							s.control = rint(pi_step(L, s, gain_p, f));             // python round(): half to even
							s_res0[i] = real;
							s_res1[i] = imag;
						}
					}
				}
				__syncwarp();
				for (int i = lane; i < cnt; i += 32) {
					dst0[c0 + i] = s_res0[i];
					if (two) dst1[c0 + i] = s_res1[i];
				}
			}
		}
		__syncthreads();
	}
}

// Latency of one dependent float64 operation (the floor under a carrier loop: its ~25 operations per sample form one
// chain): a single thread alternates __dadd_rn / __dmul_rn on its own result and reports cycles per operation.
__global__ void fp64_chain_kernel(double *out, long long *cycles, int iters, double a, double b)
{
	double x = out[0];
	const long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 8; k++) {
			x = __dadd_rn(x, a);
			x = __dmul_rn(x, b);
		}
	}
	cycles[0] = clock64() - t0;
	out[0] = x;
}

extern "C" cudaError_t pm_launch_fp64_chain(double *out, long long *cycles, int iters, cudaStream_t st)
{
	fp64_chain_kernel<<<1, 1, 0, st>>>(out, cycles, iters, 1.0e-9, 0.999999999);
	return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
extern "C" cudaError_t pm_launch_p64(const P64Chain *d_chains, const P64Chain *h_chains, int n_chains,
	const int16_t *audio, uint32_t *sign, long long sign_stride, float *soft, long long soft_stride,
	unsigned long long *max_slots, cudaStream_t st)
{
	if (n_chains <= 0) return cudaSuccess;
	long long L1 = 0, L2 = 0, L3 = 0;
	int m_bpf = 0, m_mid = 0, m_out = 0;
	bool any_loop = false, any_mid = false, any_mpsk = false;
	for (int c = 0; c < n_chains; c++) {
		const P64Chain &C = h_chains[c];
		L1 = max(L1, C.L1); L2 = max(L2, C.L2); L3 = max(L3, C.L3);
		m_bpf = max(m_bpf, C.n_bpf); m_out = max(m_out, C.n_out);
		if (C.kind == KIND_AFSK) { any_mid = true; m_mid = max(m_mid, 4 * C.n_mid + C.n_mid); }
		else any_loop = true;
		if (C.kind == KIND_MPSK) { any_mid = true; any_mpsk = true; m_mid = max(m_mid, 2 * C.n_mid); }
	}
	auto tiles = [](long long n) { return (unsigned int)((n + P64_TILE - 1) / P64_TILE); };
	static bool attr_done = false;
	if (!attr_done) {
		const int big = (P64_TILE + 5 * P64_MAX_TAPS) * (int)sizeof(double);
		cudaFuncSetAttribute(p64_bpf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
		cudaFuncSetAttribute(p64_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
		cudaFuncSetAttribute(p64_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
		attr_done = true;
	}
	if (L1 <= 0) L1 = 1;
	cudaMemsetAsync(max_slots, 0, sizeof(unsigned long long) * n_chains, st);
	pm_kt_mark("p64_bpf_kernel", st);
	p64_bpf_kernel<<<dim3(tiles(L1), n_chains), P64_THREADS, sizeof(double) * (P64_TILE + 2 * m_bpf), st>>>(d_chains, audio);
	if (any_loop) {
		pm_kt_mark("p64_max_kernel", st);
		p64_max_kernel<<<dim3(min(tiles(L1) * 4u, 1184u), n_chains), P64_THREADS, 0, st>>>(d_chains);
		pm_kt_mark("p64_seq_kernel", st);
		p64_seq_kernel<<<n_chains, 64, 0, st>>>(d_chains, 0);
	}
	if (any_mid && L2 > 0) {
		pm_kt_mark("p64_mid_kernel", st);
		p64_mid_kernel<<<dim3(tiles(L2), n_chains), P64_THREADS, sizeof(double) * (P64_TILE + m_mid), st>>>(d_chains);
	}
	if (any_mpsk) {
		pm_kt_mark("p64_seq_kernel", st);
		p64_seq_kernel<<<n_chains, 64, 0, st>>>(d_chains, 1);
	}
	pm_kt_mark("p64_out_kernel", st);
	p64_out_kernel<<<dim3(max(tiles(L3), 1u), n_chains, any_mpsk ? 2 : 1), P64_THREADS,
		sizeof(double) * (P64_TILE + 2 * m_out), st>>>(d_chains, sign, sign_stride, soft, soft_stride);
	return cudaGetLastError();
}

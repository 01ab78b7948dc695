// front.cu -- FIR front ends (FP32 FFMA, shared-memory staged, fused).
//
// AFSK (reference afsk.py:148-167):
//     x1 = audio (*) BPF
//     m_j = sqrt((x1 (*) tone_i)^2 + (x1 (*) tone_q)^2)      per tone set j
//     y_c = (m_mark (*) LPF) - space_gain_c * (m_space (*) LPF)
// All convolutions are numpy 'valid' (y[n] = sum_k h[k] x[n+M-1-k]); taps are
// stored reversed so every stage is a correlation y[n] = sum_j hr[j] x[n+j].
// The reference computes LPF(m_mark - m_space_gained); LPF is linear, so the
// LPF of the unit-gain space magnitude is shared by every chain that differs
// only in space_gain (afsk_1200_ax25_super_opt.json chains 2-8).
// The tone correlators of the reference are rotations over a rectangular window
// (afsk.py:134-144) of which only the magnitude is used, so m_j is computed as a
// sliding window sum (SlideUnit / SlidePair below) unless a caller supplies
// other taps (FirUnitPair).
//
// Only the SIGN of y reaches the rest of the chain (slicer.py:85, 99-102), so
// the kernel's product is one bit per chain-sample.  A sample whose |y| is
// within guard_eps of the magnitude scale |L_mark| + g|L_space| is queued for
// FP64 re-evaluation (guard_fixup_kernel) so that the sign matches the
// reference's float64 arithmetic.
#include "pm_common.cuh"

__device__ __forceinline__ float4 lds4(const float *s, int i)
{
	return *reinterpret_cast<const float4 *>(s + pm_phys(i));
}

__device__ __forceinline__ void sts4(float *s, int i, float a, float b, float c, float d)
{
	*reinterpret_cast<float4 *>(s + pm_phys(i)) = make_float4(a, b, c, d);
}

// One thread = 16 consecutive outputs of one FIR (NSET = 1) or of two FIRs
// that share their input window (NSET = 2: the I and Q correlators).
// Register-blocked sliding window: 16 new inputs (4 x LDS.128) feed 256 (512)
// FFMAs; taps come from the constant bank.
template <int NSET>
struct FirUnit {
	float a[16];
	float b[16];

	__device__ __forceinline__ void run(const float *__restrict__ s, int base,
	                                    const float *__restrict__ tA, const float *__restrict__ tB, int ntaps)
	{
		float w[32];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			float4 v = lds4(s, base + 4 * q);
			w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
		}
#pragma unroll
		for (int r = 0; r < 16; r++) { a[r] = 0.f; b[r] = 0.f; }
		int j0 = 0;
		for (; j0 + 16 <= ntaps; j0 += 16) {
#pragma unroll
			for (int q = 0; q < 4; q++) {
				float4 v = lds4(s, base + j0 + 16 + 4 * q);
				w[16 + 4 * q] = v.x; w[17 + 4 * q] = v.y; w[18 + 4 * q] = v.z; w[19 + 4 * q] = v.w;
			}
#pragma unroll
			for (int k = 0; k < 16; k++) {
				float hA = tA[j0 + k];
				float hB = (NSET == 2) ? tB[j0 + k] : 0.f;
#pragma unroll
				for (int r = 0; r < 16; r++) {
					a[r] = fmaf(hA, w[r + k], a[r]);
					if (NSET == 2) b[r] = fmaf(hB, w[r + k], b[r]);
				}
			}
#pragma unroll
			for (int r = 0; r < 16; r++) w[r] = w[r + 16];
		}
		for (; j0 < ntaps; j0 += 4) {
			float4 v = lds4(s, base + j0 + 16);
			w[16] = v.x; w[17] = v.y; w[18] = v.z; w[19] = v.w;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				float hA = tA[j0 + k];
				float hB = (NSET == 2) ? tB[j0 + k] : 0.f;
#pragma unroll
				for (int r = 0; r < 16; r++) {
					a[r] = fmaf(hA, w[r + k], a[r]);
					if (NSET == 2) b[r] = fmaf(hB, w[r + k], b[r]);
				}
			}
#pragma unroll
			for (int r = 0; r < 16; r++) w[r] = w[r + 4];
		}
	}
};

// The same register-blocked FIR on PAIRS: the window holds float2 samples (the mark and the space magnitude of one
// tone pair), the tap is the same for both halves, so one packed FFMA2 (fma.rn.f32x2, sm_100) does both low-pass
// filters.  FFMA2 alone is no faster than two FFMAs, but it halves the issue slots the multiply-adds take and
// leaves room for the loads, tap fetches and register moves around them (tools/ubench/ffma2.cu: 57.9 vs 49.2 TFLOP/s
// in a FIR-shaped loop).
__device__ __forceinline__ void fma2(unsigned long long &c, unsigned long long a, unsigned long long b)
{
	asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}

struct FirUnitPair {
	unsigned long long a[16];

	__device__ __forceinline__ static void load2(const float *__restrict__ s, int i, unsigned long long &x, unsigned long long &y)
	{
		const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(s + 2 * pm_phys2(i));
		x = v.x; y = v.y;
	}

	// s: one pair stream; base: first output (multiple of 16); tdup: taps stored twice each; ntaps multiple of 4
	__device__ __forceinline__ void run(const float *__restrict__ s, int base, const float *__restrict__ tdup, int ntaps)
	{
		unsigned long long w[32];
#pragma unroll
		for (int q = 0; q < 8; q++) load2(s, base + 2 * q, w[2 * q], w[2 * q + 1]);
#pragma unroll
		for (int r = 0; r < 16; r++) a[r] = 0ull;
		int j0 = 0;
		for (; j0 + 16 <= ntaps; j0 += 16) {
#pragma unroll
			for (int q = 0; q < 8; q++) load2(s, base + j0 + 16 + 2 * q, w[16 + 2 * q], w[17 + 2 * q]);
#pragma unroll
			for (int k = 0; k < 16; k++) {
				const unsigned long long hh = *reinterpret_cast<const unsigned long long *>(tdup + 2 * (j0 + k));
#pragma unroll
				for (int r = 0; r < 16; r++) fma2(a[r], hh, w[r + k]);
			}
#pragma unroll
			for (int r = 0; r < 16; r++) w[r] = w[r + 16];
		}
		for (; j0 < ntaps; j0 += 4) {
			load2(s, base + j0 + 16, w[16], w[17]);
			load2(s, base + j0 + 18, w[18], w[19]);
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const unsigned long long hh = *reinterpret_cast<const unsigned long long *>(tdup + 2 * (j0 + k));
#pragma unroll
				for (int r = 0; r < 16; r++) fma2(a[r], hh, w[r + k]);
			}
#pragma unroll
			for (int r = 0; r < 16; r++) w[r] = w[r + 4];
		}
	}
	__device__ __forceinline__ float lo(int r) const { return __uint_as_float((unsigned int)a[r]); }
	__device__ __forceinline__ float hi(int r) const { return __uint_as_float((unsigned int)(a[r] >> 32)); }
};

// Sliding tone correlator.  The reference's correlator taps are a rotation, cos/sin(w k) over a rectangular window
// (afsk.py:134-144), and only the magnitude of the (I, Q) output is used (afsk.py:153-160).  That magnitude equals
// |sum_{m in window} x[m] e^{i w (m - m0)}| for ANY phase origin m0, so a thread that owns 16 consecutive outputs
// takes its own first sample as the origin, sums the first window directly (N FFMA2: (x, x) * (cos, sin)) and then
// slides: one FFMA2 adds the sample that enters, one removes the sample that leaves.  (N + 30) instead of 16 N
// packed multiply-adds per 16 outputs; the rounding error of the 30 extra updates stays ~1e-7 of the window sum.
struct SlideUnit {
	float m[16];

	__device__ __forceinline__ static float mag(unsigned long long S)
	{
		const float ci = __uint_as_float((unsigned int)S), cq = __uint_as_float((unsigned int)(S >> 32));
		float r;
		asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(ci, ci, cq * cq)));
		return r;
	}
	__device__ __forceinline__ static unsigned long long add2(unsigned long long a, unsigned long long b)
	{
		unsigned long long r;
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
		return r;
	}

	// s: the (x, x) stream; base: first output (multiple of 16); E: rotation table (see AfskPlan::mag_e_off); N >= 16.
	// The first window is summed in four interleaved partial sums (shorter dependency chains, and each partial sum
	// rounds at a quarter of the magnitude); the slides accumulate in D, which stays small while the tone is steady,
	// and every output is S0 + D: about one rounding at the magnitude of the result instead of N of them.
	// pg points at sample `base` (a multiple of 16, so the padding of base + off is that of base plus that of off):
	// the offset part depends on uniform values only and stays in the uniform datapath
	__device__ __forceinline__ static void ldp(const float *__restrict__ pg, int off, unsigned long long &x, unsigned long long &y)
	{
		const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(pg + 2 * pm_phys2(off));
		x = v.x; y = v.y;
	}

	__device__ __forceinline__ void run(const float *__restrict__ s, int base, const float *__restrict__ E, int N)
	{
		const float *__restrict__ pg = s + 2 * pm_phys2(base);
		const unsigned long long *__restrict__ E2 = reinterpret_cast<const unsigned long long *>(E);
		const unsigned long long *__restrict__ nE2 = E2 + N + 16;
		unsigned long long h[16];
#pragma unroll
		for (int q = 0; q < 8; q++) FirUnitPair::load2(s, base + 2 * q, h[2 * q], h[2 * q + 1]);
		unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
		for (int j = 0; j < 16; j++) fma2(acc[j & 3], h[j], E2[j]);
		int j = 16;
		for (; j + 8 <= N; j += 8) {
			unsigned long long v[8];
#pragma unroll
			for (int q = 0; q < 4; q++) FirUnitPair::load2(s, base + j + 2 * q, v[2 * q], v[2 * q + 1]);
#pragma unroll
			for (int k = 0; k < 8; k++) fma2(acc[k & 3], v[k], E2[j + k]);
		}
		for (; j + 2 <= N; j += 2) {
			unsigned long long v0, v1;
			FirUnitPair::load2(s, base + j, v0, v1);
			fma2(acc[0], v0, E2[j]);
			fma2(acc[1], v1, E2[j + 1]);
		}
		const unsigned long long *__restrict__ En = E2 + N;
		unsigned long long D = 0ull;
		if (N & 1) {                          // j == N - 1 (even): the pair's second half is the first sample to enter
			unsigned long long v0, carry;
			FirUnitPair::load2(s, base + j, v0, carry);
			fma2(acc[3], v0, E2[j]);
			const unsigned long long S0 = add2(add2(acc[0], acc[1]), add2(acc[2], acc[3]));
			m[0] = mag(S0);
			fma2(D, carry, En[0]);
			fma2(D, h[0], nE2[0]);
			m[1] = mag(add2(S0, D));
#pragma unroll
			for (int q = 0; q < 7; q++) {
				unsigned long long a, b;
				FirUnitPair::load2(s, base + N + 1 + 2 * q, a, b);
				fma2(D, a, En[1 + 2 * q]);
				fma2(D, h[1 + 2 * q], nE2[1 + 2 * q]);
				m[2 + 2 * q] = mag(add2(S0, D));
				fma2(D, b, En[2 + 2 * q]);
				fma2(D, h[2 + 2 * q], nE2[2 + 2 * q]);
				m[3 + 2 * q] = mag(add2(S0, D));
			}
		} else {
			const unsigned long long S0 = add2(add2(acc[0], acc[1]), add2(acc[2], acc[3]));
			m[0] = mag(S0);
#pragma unroll
			for (int q = 0; q < 8; q++) {
				unsigned long long a, b;
				FirUnitPair::load2(s, base + N + 2 * q, a, b);
				fma2(D, a, En[2 * q]);
				fma2(D, h[2 * q], nE2[2 * q]);
				m[1 + 2 * q] = mag(add2(S0, D));
				if (q < 7) {
					fma2(D, b, En[1 + 2 * q]);
					fma2(D, h[1 + 2 * q], nE2[1 + 2 * q]);
					m[2 + 2 * q] = mag(add2(S0, D));
				}
			}
		}
	}
};

// The same for the mark and the space tone of one pair at once (equal window lengths): the two window sums are
// independent dependency chains over the same samples, so one thread interleaves them -- twice the instruction-level
// parallelism of a stage that is bound by the latency of its chains, not by their number of operations (stage
// tracing: as many cycles per CTA as the band-pass with a tenth of its multiply-adds) -- and the magnitudes leave
// as (mark, space) pairs, two samples per 128-bit store, instead of one 32-bit store per tone and sample.
// Where the magnitudes of a SlidePair go.  EmitSmem: the (mark, space) pair stream in shared memory that the FFMA2
// low-pass reads (even outputs wait for their odd neighbour, then both leave in one 128-bit store).
struct EmitSmem {
	float *dst;               // the pair stream at the unit's first sample (16 (mark, space) float2 in a row, 16-byte aligned)
	float pa, pb;
	__device__ __forceinline__ void put(int r, float ma, float mb)
	{
		if (r & 1) *reinterpret_cast<float4 *>(dst + 2 * (r - 1)) = make_float4(pa, pb, ma, mb);
		else { pa = ma; pb = mb; }
	}
};

// two floats -> two bf16 in one word (the second operand in the low half: the lower address)
__device__ __forceinline__ uint32_t bf16x2_rn(float hi, float lo)
{
	uint32_t d;
	asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
	return d;
}

// Eight consecutive FP32 samples -> three 16-byte chunks of bf16 pieces, v = p0 + p1 + p2 exactly (each residual is
// exactly representable: it has at most 16, then 8 significant bits left)
__device__ __forceinline__ void split_store8(const float *v, unsigned char *b0, unsigned char *b1, unsigned char *b2, long long off)
{
	uint32_t a[4], b[4], c[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const float lo = v[2 * j], hi = v[2 * j + 1];
		a[j] = bf16x2_rn(hi, lo);
		float rl = lo - __uint_as_float(a[j] << 16), rh = hi - __uint_as_float(a[j] & 0xFFFF0000u);
		b[j] = bf16x2_rn(rh, rl);
		rl -= __uint_as_float(b[j] << 16); rh -= __uint_as_float(b[j] & 0xFFFF0000u);
		c[j] = bf16x2_rn(rh, rl);
	}
	*reinterpret_cast<uint4 *>(b0 + off) = make_uint4(a[0], a[1], a[2], a[3]);
	*reinterpret_cast<uint4 *>(b1 + off) = make_uint4(b[0], b[1], b[2], b[3]);
	*reinterpret_cast<uint4 *>(b2 + off) = make_uint4(c[0], c[1], c[2], c[3]);
}

// EmitPieces: the tensor-core route -- both magnitude streams leave for global memory as bf16 pieces in the operand
// layout of csrc/lpf_tc.cu (MagOut), eight samples = one 16-byte chunk per piece at a time.
struct EmitPieces {
	unsigned char *ma[3], *mb[3];     // piece arrays of the mark and of the space tone
	long long off[2];                 // byte offsets of the unit's two chunks
	float va[8], vb[8];
	__device__ __forceinline__ void put(int r, float a, float b)
	{
		va[r & 7] = a; vb[r & 7] = b;
		if ((r & 7) == 7) {
			split_store8(va, ma[0], ma[1], ma[2], off[r >> 3]);
			split_store8(vb, mb[0], mb[1], mb[2], off[r >> 3]);
		}
	}
};

struct SlidePair {
	template <typename EMIT>
	__device__ __forceinline__ static void run(const float *__restrict__ s, int base, const float *__restrict__ EA,
	                                           const float *__restrict__ EB, int N, EMIT &em)
	{
		const unsigned long long *__restrict__ A2 = reinterpret_cast<const unsigned long long *>(EA);
		const unsigned long long *__restrict__ B2 = reinterpret_cast<const unsigned long long *>(EB);
		const unsigned long long *__restrict__ nA2 = A2 + N + 16, *__restrict__ nB2 = B2 + N + 16;
		unsigned long long h[16];
#pragma unroll
		for (int q = 0; q < 8; q++) FirUnitPair::load2(s, base + 2 * q, h[2 * q], h[2 * q + 1]);
		unsigned long long a[4] = {0ull, 0ull, 0ull, 0ull}, b[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
		for (int j = 0; j < 16; j++) { fma2(a[j & 3], h[j], A2[j]); fma2(b[j & 3], h[j], B2[j]); }
		int j = 16;
		for (; j + 8 <= N; j += 8) {
			unsigned long long v[8];
#pragma unroll
			for (int q = 0; q < 4; q++) FirUnitPair::load2(s, base + j + 2 * q, v[2 * q], v[2 * q + 1]);
#pragma unroll
			for (int k = 0; k < 8; k++) { fma2(a[k & 3], v[k], A2[j + k]); fma2(b[k & 3], v[k], B2[j + k]); }
		}
		for (; j + 2 <= N; j += 2) {
			unsigned long long v0, v1;
			FirUnitPair::load2(s, base + j, v0, v1);
			fma2(a[0], v0, A2[j]); fma2(b[0], v0, B2[j]);
			fma2(a[1], v1, A2[j + 1]); fma2(b[1], v1, B2[j + 1]);
		}
		const unsigned long long *__restrict__ An = A2 + N, *__restrict__ Bn = B2 + N;
		unsigned long long DA = 0ull, DB = 0ull, SA, SB;
		auto emit = [&](int r) {
			em.put(r, SlideUnit::mag(SlideUnit::add2(SA, DA)), SlideUnit::mag(SlideUnit::add2(SB, DB)));
		};
		auto slide = [&](int r, unsigned long long xin) {          // window r -> r + 1
			fma2(DA, xin, An[r]); fma2(DB, xin, Bn[r]);
			fma2(DA, h[r], nA2[r]); fma2(DB, h[r], nB2[r]);
		};
		if (N & 1) {                          // j == N - 1 (even): the pair's second half is the first sample to enter
			unsigned long long v0, carry;
			FirUnitPair::load2(s, base + j, v0, carry);
			fma2(a[3], v0, A2[j]); fma2(b[3], v0, B2[j]);
			SA = SlideUnit::add2(SlideUnit::add2(a[0], a[1]), SlideUnit::add2(a[2], a[3]));
			SB = SlideUnit::add2(SlideUnit::add2(b[0], b[1]), SlideUnit::add2(b[2], b[3]));
			emit(0);
			slide(0, carry);
			emit(1);
#pragma unroll
			for (int q = 0; q < 7; q++) {
				unsigned long long x0, x1;
				FirUnitPair::load2(s, base + N + 1 + 2 * q, x0, x1);
				slide(1 + 2 * q, x0);
				emit(2 + 2 * q);
				slide(2 + 2 * q, x1);
				emit(3 + 2 * q);
			}
		} else {
			SA = SlideUnit::add2(SlideUnit::add2(a[0], a[1]), SlideUnit::add2(a[2], a[3]));
			SB = SlideUnit::add2(SlideUnit::add2(b[0], b[1]), SlideUnit::add2(b[2], b[3]));
			emit(0);
#pragma unroll
			for (int q = 0; q < 8; q++) {
				unsigned long long x0, x1;
				FirUnitPair::load2(s, base + N + 2 * q, x0, x1);
				slide(2 * q, x0);
				emit(1 + 2 * q);
				if (q < 7) {
					slide(1 + 2 * q, x1);
					emit(2 + 2 * q);
				}
			}
		}
	}
};

// sqrt.approx.f32: one MUFU instead of the IEEE sequence with its slow path; maximum relative error 2^-23, far
// inside the FP32 front end's error budget (the sign guard is 2^-18)
__device__ __forceinline__ float fast_sqrt(float x)
{
	float r;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// Stage a_len int16 samples starting at n0 into shared memory as floats
// (zero beyond the end of the recording).
// Returns the largest |sample| this thread staged (the raw-input term of the sign guard, see afsk_front_kernel).
__device__ __forceinline__ float stage_audio(float *s_a, const int16_t *__restrict__ audio,
                                             long long n0, long long n_audio, int a_len)
{
	float amax = 0.f;
	for (int i = threadIdx.x * 8; i < a_len; i += blockDim.x * 8) {
		long long g = n0 + i;
		float f[8];
		if (g + 8 <= n_audio) {
			uint4 v = __ldg(reinterpret_cast<const uint4 *>(audio + g));
			unsigned int u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
			for (int q = 0; q < 4; q++) {
				f[2 * q] = (float)(short)(u[q] & 0xFFFFu);
				f[2 * q + 1] = (float)(short)(u[q] >> 16);
			}
		} else {
#pragma unroll
			for (int q = 0; q < 8; q++) f[q] = (g + q < n_audio) ? (float)audio[g + q] : 0.f;
		}
		sts4(s_a, i, f[0], f[1], f[2], f[3]);
		sts4(s_a, i + 4, f[4], f[5], f[6], f[7]);
#pragma unroll
		for (int q = 0; q < 8; q++) amax = fmaxf(amax, fabsf(f[q]));
	}
	return amax;
}

// The same for the route whose band-pass runs two half tiles side by side (MAGS): sample i and sample i + half travel as
// one float2, the window operand of the packed FFMA2.  cnt = pairs to stage (a multiple of 8).
__device__ __forceinline__ float stage_audio_pairs(float *s_a2, const int16_t *__restrict__ audio, long long n0,
                                                   long long n_audio, int half, int cnt)
{
	float amax = 0.f;
	for (int i = threadIdx.x * 8; i < cnt; i += blockDim.x * 8) {
		float f[2][8];
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const long long g = n0 + i + (h ? half : 0);
			if (g + 8 <= n_audio) {
				const uint4 v = __ldg(reinterpret_cast<const uint4 *>(audio + g));
				const unsigned int u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (int q = 0; q < 4; q++) {
					f[h][2 * q] = (float)(short)(u[q] & 0xFFFFu);
					f[h][2 * q + 1] = (float)(short)(u[q] >> 16);
				}
			} else {
#pragma unroll
				for (int q = 0; q < 8; q++) f[h][q] = (g + q < n_audio) ? (float)audio[g + q] : 0.f;
			}
		}
		float *dst = s_a2 + 2 * pm_phys2(i);
#pragma unroll
		for (int q = 0; q < 4; q++)
			*reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(f[0][2 * q], f[1][2 * q], f[0][2 * q + 1], f[1][2 * q + 1]);
#pragma unroll
		for (int q = 0; q < 8; q++) amax = fmaxf(amax, fmaxf(fabsf(f[0][q]), fabsf(f[1][q])));
	}
	return amax;
}

template <bool WRITE_SOFT, bool MAGS>
__global__ void __launch_bounds__(PM_FRONT_THREADS, 2)
afsk_front_kernel(const __grid_constant__ AfskPlan P, const int16_t *__restrict__ audio, long long n_audio,
                  long long tile_first, uint32_t *__restrict__ sign, long long sign_stride,
                  float *__restrict__ soft, long long soft_stride, GuardList guard, MagOut mags)
{
	extern __shared__ __align__(16) float smem[];
	float *s_a = smem;
	float *s_x1 = smem + P.s_x1_off;
	float *s_m = smem + P.s_m_off;
	const long long n0 = (tile_first + blockIdx.x) * (long long)P.tile;
	const int tid = threadIdx.x;
	// stage tracing: thread 0 adds the cycles between barriers (its own work plus the wait for the slowest warp)
	const bool trace = guard.stage_clk != nullptr && tid == 0;
	long long t_prev = trace ? clock64() : 0;
	auto stage_done = [&](int i) {
		if (trace) {
			const long long t = clock64();
			atomicAdd(&guard.stage_clk[i], (unsigned long long)(t - t_prev));
			t_prev = t;
		}
	};

	// Largest raw |sample| of the tile: the band-pass rounds at the magnitude of its RAW input (DC, hum and anything else
	// out of band included), so the sign guard carries a term proportional to it (epilogue below)
	__shared__ float s_wmax[PM_FRONT_THREADS / 32];
	{
		float amax = MAGS ? stage_audio_pairs(s_a, audio, n0, n_audio, 16 * P.bpf_half, (16 * P.bpf_half + P.n_bpf + 16 + 7) & ~7)
		                  : stage_audio(s_a, audio, n0, n_audio, P.a_len);
		amax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(amax)));      // non-negative floats order like their bit patterns
		if ((tid & 31) == 0) s_wmax[tid >> 5] = amax;
	}
	__syncthreads();
	stage_done(0);
	if (MAGS && tid == 0) {
		float m = s_wmax[0];
#pragma unroll
		for (int q = 1; q < PM_FRONT_THREADS / 32; q++) m = fmaxf(m, s_wmax[q]);
		mags.tile_amax[tile_first + blockIdx.x] = m;
	}

	// input band-pass (afsk.py:151); the result is stored as (x, x) pairs: the window operand of the packed correlators
	if (MAGS) {
		// two half tiles side by side: unit u and unit u + bpf_half share one packed FFMA2 per tap (the taps stored twice,
		// a uniform operand) -- half the issue slots of the scalar form, which leaves room for the window loads
		for (int ub = tid - (tid & 31); ub < P.bpf_half; ub += PM_FRONT_THREADS) {
			const int u = ub + (tid & 31);
			if (u >= P.bpf_half) continue;
			FirUnitPair f;
			f.run(s_a, 16 * u, P.taps + P.bpf2_off, P.n_bpf);
			float *d0 = s_x1 + 2 * pm_phys2(16 * u), *d1 = s_x1 + 2 * pm_phys2(16 * (u + P.bpf_half));
#pragma unroll
			for (int q = 0; q < 8; q++) {
				*reinterpret_cast<float4 *>(d0 + 4 * q) = make_float4(f.lo(2 * q), f.lo(2 * q), f.lo(2 * q + 1), f.lo(2 * q + 1));
				*reinterpret_cast<float4 *>(d1 + 4 * q) = make_float4(f.hi(2 * q), f.hi(2 * q), f.hi(2 * q + 1), f.hi(2 * q + 1));
			}
		}
	} else
	for (int ub = tid - (tid & 31); ub < P.U_x; ub += PM_FRONT_THREADS) {        // warp-uniform control flow: uniform taps
		const int u = ub + (tid & 31);
		if (u >= P.U_x) continue;
		FirUnit<1> f;
		f.run(s_a, 16 * u, P.taps + P.bpf_off, nullptr, P.n_bpf);
		float *dst = s_x1 + 2 * pm_phys2(16 * u);
#pragma unroll
		for (int q = 0; q < 8; q++)
			*reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(f.a[2 * q], f.a[2 * q], f.a[2 * q + 1], f.a[2 * q + 1]);
	}
	__syncthreads();
	stage_done(1);

	// tone correlators and magnitudes (afsk.py:153-160): I and Q of one tone in the two halves of an FFMA2.
	// All threads of a warp work on the same tone(s), so that the tap / table operands are warp-uniform: an FFMA2 whose
	// tap comes from a uniform register runs at the full FP32 rate, one with three vector-register operands only at
	// ~78 % (tools/ubench/ffma2.cu) -- which is also why every table offset is indexed by the loop counter itself
	// (an offset fetched through a second constant load makes the compiler give up on the uniform datapath)
	for (int p = 0; p < P.n_pair; p++) {                                        // fused pairs: both tones in one pass
		if (!P.pair_fused[p]) continue;
		for (int ub = tid - (tid & 31); ub < P.U_m; ub += PM_FRONT_THREADS) {   // warp-uniform control flow
			const int ui = ub + (tid & 31);
			if (ui >= P.U_m) continue;
			if (MAGS) {
				EmitPieces em;
				const long long rows128 = mags.rows * 128;
#pragma unroll
				for (int q = 0; q < 3; q++) {
					em.ma[q] = mags.base + (long long)(P.pair_mark[p] * 3 + q) * rows128;
					em.mb[q] = mags.base + (long long)(P.pair_space[p] * 3 + q) * rows128;
				}
				em.off[0] = pm_mag_offset(n0 + 16 * ui);
				em.off[1] = pm_mag_offset(n0 + 16 * ui + 8);
				SlidePair::run(s_x1, 16 * ui, P.taps + P.pair_ea[p], P.taps + P.pair_eb[p], P.pair_fused[p], em);
			} else {
				EmitSmem em;
				em.dst = s_m + p * P.s_m_stride + 2 * pm_phys2(16 * ui);
				em.pa = em.pb = 0.f;
				SlidePair::run(s_x1, 16 * ui, P.taps + P.pair_ea[p], P.taps + P.pair_eb[p], P.pair_fused[p], em);
			}
		}
	}
	if (MAGS) {                       // the low-pass and the epilogue are csrc/lpf_tc.cu's
		if (trace) {
			__syncwarp();
			stage_done(2);
			atomicAdd(&guard.stage_clk[4], 1ull);
		}
		return;
	}
	for (int j = 0; j < P.n_mag; j++) {
	if (P.mag_dst_first[j] == P.mag_dst_first[j + 1]) continue;                 // both uses of the tone are fused pairs
	for (int ub = tid - (tid & 31); ub < P.U_m; ub += PM_FRONT_THREADS) {      // warp-uniform control flow
		const int ui = ub + (tid & 31);
		if (ui >= P.U_m) continue;
		float m[16];
		if (P.mag_slide[j]) {
			SlideUnit f;
			f.run(s_x1, 16 * ui, P.taps + P.mag_e_off[j], P.mag_slide[j]);
#pragma unroll
			for (int r = 0; r < 16; r++) m[r] = f.m[r];
		} else {
			FirUnitPair f;
			f.run(s_x1, 16 * ui, P.taps + P.mag_iq_off[j], P.mag_n[j]);
#pragma unroll
			for (int r = 0; r < 16; r++) {
				const float ci = f.lo(r), cq = f.hi(r);
				m[r] = fast_sqrt(fmaf(ci, ci, cq * cq));
			}
		}
		// the magnitude goes into the mark or space half of every pair stream this tone belongs to
		for (int di = P.mag_dst_first[j]; di < P.mag_dst_first[j + 1]; di++) {
			const int dd = P.mag_dst[di];
			float *dst = s_m + (dd >> 1) * P.s_m_stride + (dd & 1) + 2 * pm_phys2(16 * ui);
#pragma unroll
			for (int r = 0; r < 16; r++) dst[2 * r] = m[r];
		}
	}
	}
	__syncthreads();
	stage_done(2);

	// output low-pass of the mark and space magnitudes, per-chain combination,
	// sign packing (afsk.py:162-166 -> slicer.py:85,99)
	const int lane = tid & 31;
	const int total = P.n_pair * P.U_l;
	float tile_amax = s_wmax[0];
#pragma unroll
	for (int q = 1; q < PM_FRONT_THREADS / 32; q++) tile_amax = fmaxf(tile_amax, s_wmax[q]);
	for (int ub = tid - lane; ub < total; ub += PM_FRONT_THREADS) {
		const int u = ub + lane;
		const bool active = u < total;
		int p = 0, ui = 0;
		FirUnitPair fp;
		if (active) {
			p = u / P.U_l;
			ui = u - p * P.U_l;
			fp.run(s_m + p * P.s_m_stride, 16 * ui, P.taps + P.lpf2_off, P.n_lpf);
		}
		// chains are visited in lock-step by the whole warp (a warp may
		// straddle two pairs; lanes of the shorter pair idle)
		int cmax = 0;
		if (active) cmax = P.pair_first[p + 1] - P.pair_first[p];
		cmax = __reduce_max_sync(0xffffffffu, cmax);
		const long long nbase = n0 + 16 * ui;
		for (int ci = 0; ci < cmax; ci++) {
			unsigned int half = 0;
			int c = -1;
			if (active && ci < P.pair_first[p + 1] - P.pair_first[p]) {
				c = P.pair_first[p] + ci;
				const float g = P.chain_gain[c];
				const float neg_eps = -P.guard_eps;
				// guard: |y| < eps * (|L_mark| + g |L_space|) + abs_c, where abs_c = c_abs * 2^-24 * max|audio of the tile| *
				// sum|h_bpf| * N_corr * sum|h_lpf| * (1 + g) bounds what the band-pass's rounding at raw-input magnitude can
				// leave in y (host: build_groups; calibration: tools/guard_bound.py, tests/test_gpu_guard.py)
				const float abs_c = P.chain_guard_abs[c] * tile_amax;
				// six instructions per sample: y, the magnitude scale, |y| - abs_c - eps * scale, and one funnel shift each to
				// collect the sign bits of y and of the guard test (y is never -0: the accumulators start at +0)
				unsigned int neg = 0, near = 0;
#pragma unroll
				for (int r = 15; r >= 0; r--) {
					const float lm = fp.lo(r), ls = fp.hi(r);
					const float y = fmaf(-g, ls, lm);
					const float scale = fmaf(g, fabsf(ls), fabsf(lm));
					const float d = fmaf(neg_eps, scale, fabsf(y) - abs_c);
					neg = __funnelshift_l(__float_as_uint(y), neg, 1);
					near = __funnelshift_l(__float_as_uint(d), near, 1);
				}
				half = ~neg & 0xFFFFu;
				near &= 0xFFFFu;
				if (near || WRITE_SOFT) {
					const int gid = P.chain_gid[c];
					const long long nout = P.chain_nout[c];
					for (int r = 0; r < 16; r++) {
						if (nbase + r >= nout) break;
						if ((near >> r) & 1u) {
							unsigned int slot = atomicAdd(guard.count, 1u);
							if (slot < guard.cap)
								guard.entries[slot] = ((unsigned long long)gid << 48) | (unsigned long long)(nbase + r);
						}
					}
					if (WRITE_SOFT) {
#pragma unroll
						for (int r = 0; r < 16; r++)
							if (nbase + r < nout) soft[gid * soft_stride + nbase + r] = fmaf(-g, fp.hi(r), fp.lo(r));
					}
				}
			}
			// lanes 2k / 2k+1 hold the low / high half of one 32-sample word
			const unsigned int hi = __shfl_down_sync(0xffffffffu, half, 1);
			if (c >= 0 && !(lane & 1)) {
				const long long word = (nbase >> 5);
				sign[P.chain_gid[c] * sign_stride + word] = half | (hi << 16);
			}
		}
	}
	if (trace) {
		__syncwarp();
		stage_done(3);
		atomicAdd(&guard.stage_clk[4], 1ull);
	}
}

// Single-FIR front end: fsk.py:149-159 (y = audio (*) input_lpf, optional negate).
template <bool WRITE_SOFT>
__global__ void __launch_bounds__(PM_FRONT_THREADS, 2)
fir_front_kernel(const __grid_constant__ FirPlan P, const int16_t *__restrict__ audio, long long n_audio,
                 long long tile_first, uint32_t *__restrict__ sign, long long sign_stride,
                 float *__restrict__ soft, long long soft_stride, GuardList guard)
{
	extern __shared__ __align__(16) float smem[];
	float *s_a = smem;
	const long long n0 = (tile_first + blockIdx.x) * (long long)P.tile;
	const int tid = threadIdx.x;
	const int lane = tid & 31;

	(void)stage_audio(s_a, audio, n0, n_audio, P.a_len);
	__syncthreads();

	for (int ub = tid - lane; ub < P.U_y; ub += PM_FRONT_THREADS) {
		const int u = ub + lane;
		const bool active = u < P.U_y;
		FirUnit<1> f;
		float scale = 0.f;
		if (active) {
			f.run(s_a, 16 * u, P.taps + P.taps_off, nullptr, P.n_taps);
			// magnitude scale for the guard band: peak |x| of the unit's input
			// window (guard_eps already carries sum|h|); conservative, and the
			// FP64 re-evaluation of a single FIR output is cheap.
			for (int j = 0; j < P.n_taps + 16; j += 4) {
				const float4 v = lds4(s_a, 16 * u + j);
				scale = fmaxf(scale, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
			}
		}
		const long long nbase = n0 + 16 * u;
		for (int c = 0; c < P.n_chain; c++) {
			unsigned int half = 0;
			if (active) {
				const int gid = P.chain_gid[c];
				const long long nout = P.chain_nout[c];
#pragma unroll
				for (int r = 0; r < 16; r++) {
					const float y = P.chain_neg[c] ? -f.a[r] : f.a[r];
					if (y >= 0.f) half |= (1u << r);
					if (fabsf(y) < P.guard_eps * scale && nbase + r < nout) {
						unsigned int slot = atomicAdd(guard.count, 1u);
						if (slot < guard.cap)
							guard.entries[slot] = ((unsigned long long)gid << 48) | (unsigned long long)(nbase + r);
					}
					if (WRITE_SOFT && nbase + r < nout) soft[gid * soft_stride + nbase + r] = y;
				}
			}
			const unsigned int hi = __shfl_down_sync(0xffffffffu, half, 1);
			if (active && !(lane & 1))
				sign[P.chain_gid[c] * sign_stride + (nbase >> 5)] = half | (hi << 16);
		}
	}
}

// ---------------------------------------------------------------------------
// FP64 guard-band fix-up.  One CTA per queued (chain, sample): re-evaluate the
// reference's own float64 formula (its taps, its order of stages) for that one
// output and overwrite the sign bit (and the soft value when kept).
// ---------------------------------------------------------------------------

#define PM_FIX_THREADS 128       // four warps; every warp re-evaluates its own (chain, sample) entries
#define FIX_OB1 5                // band-pass outputs per lane and round (159 needed at 48 kHz: one round of 32 x 5)
#define FIX_OB2 7                // correlator outputs per lane and round (100 x {mark, space} -> 30 items: one round)

// One WARP per queued (chain, sample): no block-wide barriers, a handful of warps per SM keep the FP64 pipe busy.
// The audio window is staged once as doubles; each lane slides a small register window over the taps (one tap load
// and one sample load per FIX_OB multiply-adds), the low-pass is a warp reduction.
__global__ void __launch_bounds__(PM_FIX_THREADS)
guard_fixup_kernel(const Fp64Chain *__restrict__ chains, const int16_t *__restrict__ audio_all, long long /* n_audio: per chain */,
                   uint32_t *__restrict__ sign, long long sign_stride, float *__restrict__ soft,
                   long long soft_stride, GuardList guard, int warp_doubles)
{
	extern __shared__ __align__(16) double sm64[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	double *base = sm64 + (long long)wid * warp_doubles;
	const unsigned int n_entries = min(guard.to ? *guard.to : *guard.count, guard.cap);
	const unsigned int first = guard.from ? min(*guard.from, n_entries) : 0u;
	const unsigned int n_warps = gridDim.x * (PM_FIX_THREADS / 32);
	for (unsigned int e = first + blockIdx.x * (PM_FIX_THREADS / 32) + wid; e < n_entries; e += n_warps) {
		const unsigned long long ent = guard.entries[e];
		const int gid = (int)(ent >> 48);
		const long long n = (long long)(ent & 0xFFFFFFFFFFFFull);
		const Fp64Chain C = chains[gid];
		const int16_t *__restrict__ audio = audio_all + C.audio_off;      // this chain's recording
		const long long n_audio = C.n_audio;
		double part = 0.0;
		if (C.kind == 2) {
			// y[n] = sum_j hr[j] a[n+j]
			for (int j = lane; j < C.n_bpf; j += 32) part = fma(C.bpf[j], (double)audio[n + j], part);
		} else {
			const int nx = C.n_corr + C.n_lpf - 1;               // band-passed samples needed
			const int nx_pad = (nx + 32 * FIX_OB1 - 1) / (32 * FIX_OB1) * (32 * FIX_OB1);
			const int na = nx_pad + C.n_bpf;                     // audio samples staged (zero beyond the window)
			double *a = base;                                    // na
			double *x1 = a + na;                                 // nx_pad + n_corr + FIX_OB2 (zero padded)
			double *mag = x1 + nx_pad + C.n_corr + FIX_OB2 * 32; // 2 rows of n_lpf (+ padding)
			const int mrow = (C.n_lpf + 32 * FIX_OB2 - 1) / (32 * FIX_OB2) * (32 * FIX_OB2);
			for (int i = lane; i < na; i += 32)
				a[i] = (i < nx + C.n_bpf - 1 && n + i < n_audio) ? (double)audio[n + i] : 0.0;
			for (int i = nx + lane; i < nx_pad + C.n_corr + FIX_OB2 * 32; i += 32) x1[i] = 0.0;
			__syncwarp();
			for (int o0 = lane * FIX_OB1; o0 < nx_pad; o0 += 32 * FIX_OB1) {      // band-pass, FIX_OB1 outputs per lane
				double acc[FIX_OB1], w[FIX_OB1];
#pragma unroll
				for (int r = 0; r < FIX_OB1; r++) { acc[r] = 0.0; w[r] = a[o0 + r]; }
				for (int j = 0; j < C.n_bpf; j++) {
					const double h = C.bpf[j];
#pragma unroll
					for (int r = 0; r < FIX_OB1; r++) acc[r] = fma(h, w[r], acc[r]);
#pragma unroll
					for (int r = 0; r < FIX_OB1 - 1; r++) w[r] = w[r + 1];
					w[FIX_OB1 - 1] = a[o0 + j + FIX_OB1];
				}
#pragma unroll
				for (int r = 0; r < FIX_OB1; r++)
					if (o0 + r < nx) x1[o0 + r] = acc[r];
			}
			__syncwarp();
			const int n_ob = mrow / FIX_OB2;                     // output blocks per magnitude row
			for (int it = lane; it < 2 * n_ob; it += 32) {        // correlators: (block of outputs, mark | space)
				const int which = it >= n_ob, o0 = (it - which * n_ob) * FIX_OB2;
				const double *ti = which ? C.space_i : C.mark_i, *tq = which ? C.space_q : C.mark_q;
				double ai[FIX_OB2], aq[FIX_OB2], w[FIX_OB2];
#pragma unroll
				for (int r = 0; r < FIX_OB2; r++) { ai[r] = 0.0; aq[r] = 0.0; w[r] = x1[o0 + r]; }
				for (int j = 0; j < C.n_corr; j++) {
					const double hi = ti[j], hq = tq[j];
#pragma unroll
					for (int r = 0; r < FIX_OB2; r++) { ai[r] = fma(hi, w[r], ai[r]); aq[r] = fma(hq, w[r], aq[r]); }
#pragma unroll
					for (int r = 0; r < FIX_OB2 - 1; r++) w[r] = w[r + 1];
					w[FIX_OB2 - 1] = x1[o0 + j + FIX_OB2];
				}
#pragma unroll
				for (int r = 0; r < FIX_OB2; r++) mag[which * mrow + o0 + r] = sqrt(ai[r] * ai[r] + aq[r] * aq[r]);
			}
			__syncwarp();
			for (int i = lane; i < C.n_lpf; i += 32) part = fma(C.lpf[i], mag[i] - mag[mrow + i], part);
		}
		for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
		double y = part;
		if (C.kind == 2 && C.neg) y = -y;
		if (lane == 0) {
			uint32_t *w = sign + gid * sign_stride + (n >> 5);
			const uint32_t bit = 1u << (n & 31);
			if (y >= 0.0) atomicOr(w, bit); else atomicAnd(w, ~bit);
			if (soft) soft[gid * soft_stride + n] = (float)y;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------
// FP32 FFMA peak microbenchmark (roofline denominator for the kernels above).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float x)
{
	float a[16];
#pragma unroll
	for (int r = 0; r < 16; r++) a[r] = threadIdx.x * 0.001f + r;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 8; k++)
#pragma unroll
			for (int r = 0; r < 16; r++) a[r] = fmaf(a[r], x, 0.5f);
	}
	float s = 0.f;
#pragma unroll
	for (int r = 0; r < 16; r++) s += a[r];
	if (s == 12345.678f) out[0] = s;
}

// the guard count as of this point of the stream (a 4-byte cudaMemcpyAsync would queue on the copy engine behind the
// chunks of the recording that are already submitted, and stall the launch stream until they are through)
__global__ void guard_snapshot_kernel(const unsigned int *__restrict__ count, unsigned int *__restrict__ out)
{
	*out = *count;
}

// ---------------------------------------------------------------------------
// host-side launchers (called from engine.cu)
// ---------------------------------------------------------------------------
extern "C" cudaError_t pm_launch_afsk_front(const AfskPlan *plan, size_t smem_bytes, const int16_t *audio,
	long long n_audio, long long tile_first, int n_tiles, uint32_t *sign, long long sign_stride,
	float *soft, long long soft_stride, GuardList guard, MagOut mags, cudaStream_t st)
{
	if (n_tiles <= 0) return cudaSuccess;
	static bool attr_done = false;
	if (!attr_done) {
		// 227 KB per CTA in all, static shared memory (the per-warp raw-sample maxima) included: the engine plans
		// tiles of at most 226 KB of dynamic shared memory
		cudaError_t e1 = cudaFuncSetAttribute(afsk_front_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
		cudaError_t e2 = cudaFuncSetAttribute(afsk_front_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
		cudaError_t e3 = cudaFuncSetAttribute(afsk_front_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
		if (e1 != cudaSuccess) return e1;
		if (e2 != cudaSuccess) return e2;
		if (e3 != cudaSuccess) return e3;
		attr_done = true;
	}
	if (plan->tensor_lpf) {
		pm_kt_mark("afsk_front_kernel (magnitudes)", st);
		afsk_front_kernel<false, true><<<n_tiles, PM_FRONT_THREADS, smem_bytes, st>>>(*plan, audio, n_audio, tile_first,
			sign, sign_stride, soft, soft_stride, guard, mags);
		return cudaGetLastError();
	}
	pm_kt_mark("afsk_front_kernel", st);
	if (soft)
		afsk_front_kernel<true, false><<<n_tiles, PM_FRONT_THREADS, smem_bytes, st>>>(*plan, audio, n_audio, tile_first,
			sign, sign_stride, soft, soft_stride, guard, mags);
	else
		afsk_front_kernel<false, false><<<n_tiles, PM_FRONT_THREADS, smem_bytes, st>>>(*plan, audio, n_audio, tile_first,
			sign, sign_stride, soft, soft_stride, guard, mags);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_fir_front(const FirPlan *plan, size_t smem_bytes, const int16_t *audio,
	long long n_audio, long long tile_first, int n_tiles, uint32_t *sign, long long sign_stride,
	float *soft, long long soft_stride, GuardList guard, cudaStream_t st)
{
	if (n_tiles <= 0) return cudaSuccess;
	static bool attr_done = false;
	if (!attr_done) {
		cudaFuncSetAttribute(fir_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
		cudaFuncSetAttribute(fir_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
		attr_done = true;
	}
	pm_kt_mark("fir_front_kernel", st);
	if (soft)
		fir_front_kernel<true><<<n_tiles, PM_FRONT_THREADS, smem_bytes, st>>>(*plan, audio, n_audio, tile_first,
			sign, sign_stride, soft, soft_stride, guard);
	else
		fir_front_kernel<false><<<n_tiles, PM_FRONT_THREADS, smem_bytes, st>>>(*plan, audio, n_audio, tile_first,
			sign, sign_stride, soft, soft_stride, guard);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_guard_fixup(const Fp64Chain *chains, int warp_doubles, const int16_t *audio,
	long long n_audio, uint32_t *sign, long long sign_stride, float *soft, long long soft_stride,
	GuardList guard, int grid, cudaStream_t st)
{
	const size_t smem = sizeof(double) * (size_t)warp_doubles * (PM_FIX_THREADS / 32);
	static size_t attr = 0;
	if (smem > 48 * 1024 && smem > attr) {
		cudaFuncSetAttribute(guard_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		attr = smem;
	}
	pm_kt_mark("guard_fixup_kernel", st);
	guard_fixup_kernel<<<grid, PM_FIX_THREADS, smem, st>>>(chains, audio, n_audio, sign, sign_stride, soft,
		soft_stride, guard, warp_doubles);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_guard_snapshot(const unsigned int *count, unsigned int *out, cudaStream_t st)
{
	guard_snapshot_kernel<<<1, 1, 0, st>>>(count, out);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_ffma_peak(float *out, int blocks, int iters, cudaStream_t st)
{
	ffma_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, 0.999f);
	return cudaGetLastError();
}

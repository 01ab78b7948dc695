// Microbenchmark: FP32 FMA issue on sm_100a -- scalar FFMA vs packed FFMA2 (fma.rn.f32x2), alone and with the
// shared-memory window loads a register-blocked FIR needs (1 LDS.128 per 64 FMA).  Prints TFLOP/s per variant.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(unsigned long long &c, unsigned long long a, unsigned long long b)
{
	asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}


// cleaner variants: explicit windows
__global__ void __launch_bounds__(256, 2) ffma_only(float *out, const float *in, int iters)
{
	float a[16], w[16];
	for (int r = 0; r < 16; r++) { a[r] = 0.f; w[r] = in[(threadIdx.x + r) & 255]; }
	const float h0 = in[0], h1 = in[1];
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int k = 0; k < 16; k++)
#pragma unroll
			for (int r = 0; r < 16; r++) a[r] = fmaf((k & 1) ? h1 : h0, w[(r + k) & 15], a[r]);
	}
	float t = 0;
	for (int r = 0; r < 16; r++) t += a[r];
	out[blockIdx.x * 256 + threadIdx.x] = t;
}

__global__ void __launch_bounds__(256, 2) ffma2_only(float *out, const float *in, int iters)
{
	unsigned long long a[16], w[16];
	for (int r = 0; r < 16; r++) {
		a[r] = 0ull;
		float2 v = make_float2(in[(threadIdx.x + r) & 255], in[(threadIdx.x + r + 7) & 255]);
		w[r] = *reinterpret_cast<unsigned long long *>(&v);
	}
	float2 hv0 = make_float2(in[0], in[0]), hv1 = make_float2(in[1], in[1]);
	const unsigned long long h0 = *reinterpret_cast<unsigned long long *>(&hv0), h1 = *reinterpret_cast<unsigned long long *>(&hv1);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int k = 0; k < 16; k++)
#pragma unroll
			for (int r = 0; r < 16; r++) fma2(a[r], (k & 1) ? h1 : h0, w[(r + k) & 15]);
	}
	float t = 0;
	for (int r = 0; r < 16; r++) { float2 v = *reinterpret_cast<float2 *>(&a[r]); t += v.x + v.y; }
	out[blockIdx.x * 256 + threadIdx.x] = t;
}

// FFMA2 with the tap operand in a uniform register (kernel parameters): fewer vector register reads per instruction
struct TapParams { float2 h[16]; };
__global__ void __launch_bounds__(256, 2) ffma2_uniform(float *out, const float *in, int iters, const __grid_constant__ TapParams P)
{
	unsigned long long a[16], w[16];
	for (int r = 0; r < 16; r++) {
		a[r] = 0ull;
		float2 v = make_float2(in[(threadIdx.x + r) & 255], in[(threadIdx.x + r + 7) & 255]);
		w[r] = *reinterpret_cast<unsigned long long *>(&v);
	}
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int k = 0; k < 16; k++) {
			const unsigned long long h = *reinterpret_cast<const unsigned long long *>(&P.h[k]);
#pragma unroll
			for (int r = 0; r < 16; r++) fma2(a[r], h, w[(r + k) & 15]);
		}
	}
	float t = 0;
	for (int r = 0; r < 16; r++) { float2 v = *reinterpret_cast<float2 *>(&a[r]); t += v.x + v.y; }
	out[blockIdx.x * 256 + threadIdx.x] = t;
}

// FIR-like: 16 outputs x 16 taps per chunk, window refilled from shared memory (4 LDS.128 per 256 FMA)
__global__ void __launch_bounds__(256, 2) ffma_fir(float *out, const float *in, int iters)
{
	__shared__ __align__(16) float s[8192];
	for (int i = threadIdx.x; i < 8192; i += 256) s[i] = in[i & 255];
	__syncthreads();
	float a[16], w[32];
	for (int r = 0; r < 16; r++) { a[r] = 0.f; w[r] = s[threadIdx.x * 20 + r]; }
	const float4 *sv = reinterpret_cast<const float4 *>(s + threadIdx.x * 20);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int q = 0; q < 4; q++) {
			float4 v = sv[(it & 3) + q];
			w[16 + 4 * q] = v.x; w[17 + 4 * q] = v.y; w[18 + 4 * q] = v.z; w[19 + 4 * q] = v.w;
		}
#pragma unroll
		for (int k = 0; k < 16; k++) {
			const float h = in[(it * 16 + k) & 255];
#pragma unroll
			for (int r = 0; r < 16; r++) a[r] = fmaf(h, w[r + k], a[r]);
		}
#pragma unroll
		for (int r = 0; r < 16; r++) w[r] = w[r + 16];
	}
	float t = 0;
	for (int r = 0; r < 16; r++) t += a[r];
	out[blockIdx.x * 256 + threadIdx.x] = t;
}

// same arithmetic on float2 lanes: 16 outputs x 2 half-tiles, window of float2 (8 LDS.128 per 512 FMA)
__global__ void __launch_bounds__(256, 2) ffma2_fir(float *out, const float *in, int iters)
{
	__shared__ __align__(16) float s[11264];
	for (int i = threadIdx.x; i < 11264; i += 256) s[i] = in[i & 255];
	__syncthreads();
	unsigned long long a[16], w[32];
	for (int r = 0; r < 16; r++) { a[r] = 0ull; w[r] = reinterpret_cast<unsigned long long *>(s)[threadIdx.x * 20 + r]; }
	const ulonglong2 *sv = reinterpret_cast<const ulonglong2 *>(s + threadIdx.x * 40);
	const unsigned long long *hp = reinterpret_cast<const unsigned long long *>(in);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int q = 0; q < 8; q++) {
			ulonglong2 v = sv[(it & 3) + q];
			w[16 + 2 * q] = v.x; w[17 + 2 * q] = v.y;
		}
#pragma unroll
		for (int k = 0; k < 16; k++) {
			const unsigned long long h = hp[(it * 16 + k) & 127];
#pragma unroll
			for (int r = 0; r < 16; r++) fma2(a[r], h, w[r + k]);
		}
#pragma unroll
		for (int r = 0; r < 16; r++) w[r] = w[r + 16];
	}
	float t = 0;
	for (int r = 0; r < 16; r++) { float2 v = *reinterpret_cast<float2 *>(&a[r]); t += v.x + v.y; }
	out[blockIdx.x * 256 + threadIdx.x] = t;
}

template <typename F>
static double timeit(F launch, double flops)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	launch();
	cudaDeviceSynchronize();
	double best = 0;
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(e0);
		launch();
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		best = best > flops / (ms * 1e-3) / 1e12 ? best : flops / (ms * 1e-3) / 1e12;
	}
	return best;
}

int main()
{
	float *out, *in;
	const int blocks = 148 * 2 * 8, iters = 4096;
	cudaMalloc(&out, blocks * 256 * sizeof(float));
	cudaMalloc(&in, 4096 * sizeof(float));
	cudaMemset(in, 0, 4096 * sizeof(float));
	const double f1 = 2.0 * blocks * 256.0 * iters * 256;
	printf("ffma_only   %.1f TFLOP/s\n", timeit([&] { ffma_only<<<blocks, 256>>>(out, in, iters); }, f1));
	printf("ffma2_only  %.1f TFLOP/s\n", timeit([&] { ffma2_only<<<blocks, 256>>>(out, in, iters); }, 2 * f1));
	TapParams tp;
	for (int i = 0; i < 16; i++) tp.h[i] = make_float2(0.f, 0.f);
	printf("ffma2_unif  %.1f TFLOP/s\n", timeit([&] { ffma2_uniform<<<blocks, 256>>>(out, in, iters, tp); }, 2 * f1));
	printf("ffma_fir    %.1f TFLOP/s\n", timeit([&] { ffma_fir<<<blocks, 256>>>(out, in, iters); }, f1));
	printf("ffma2_fir   %.1f TFLOP/s\n", timeit([&] { ffma2_fir<<<blocks, 256>>>(out, in, iters); }, 2 * f1));
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}

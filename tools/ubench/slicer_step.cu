// Microbenchmark: the slicer's exact step (csrc/slicer.cu slicer_step, FAST form) on sm_100a -- cycles per warp-sample and
// scheduler for (0) the operand-select form the kernel uses, (1) the same with the lock multiply skipped when no lane of
// the warp has a zero crossing at that sample (warp-uniform branch on the OR of the lanes' crossing words), (2) two
// independent chains per thread, (3) the bare float64 chain (add + multiply, no selects, no mask), (4) add only.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o slicer_step slicer_step.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void step_full(double &c, uint32_t &m, uint32_t z, uint32_t bit, double cs, double a_roll, double a_keep, double lam)
{
	const bool roll = c >= cs;
	const double add = __hiloint2double(roll ? __double2hiint(a_roll) : __double2hiint(a_keep), 0);
	const double t = __dadd_rn(c, add);
	const bool cross = (z & bit) != 0;
	const double f = cross ? lam : 1.0;
	c = __dmul_rn(t, f);
	m = __funnelshift_l((uint32_t)__double2hiint(add), m, 1);
}
__device__ __forceinline__ void step_nomul(double &c, uint32_t &m, double cs, double a_roll, double a_keep)
{
	const bool roll = c >= cs;
	const double add = __hiloint2double(roll ? __double2hiint(a_roll) : __double2hiint(a_keep), 0);
	c = __dadd_rn(c, add);
	m = __funnelshift_l((uint32_t)__double2hiint(add), m, 1);
}

template <int V>
__global__ void __launch_bounds__(128) k(double *cp, const uint32_t *zp, uint32_t *mp, double cs, double ar, double ak, double lam, int n)
{
	const size_t tid = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
	double c = cp[tid], c2 = c + 3.0;
	const uint32_t *zz = zp + tid * (size_t)n;
	uint32_t acc = 0;
	uint32_t zn = zz[0];
	for (int w = 0; w < n; w++) {
		const uint32_t z = zn;
		if (w + 1 < n) zn = zz[w + 1];
		uint32_t m = 0, m2 = 0;
		if (V == 0) {
#pragma unroll
			for (int i = 0; i < 32; i++) step_full(c, m, z, 1u << i, cs, ar, ak, lam);
		} else if (V == 1) {
			const uint32_t zany = __reduce_or_sync(0xffffffffu, z);
#pragma unroll
			for (int i = 0; i < 32; i++) {
				if (zany & (1u << i)) step_full(c, m, z, 1u << i, cs, ar, ak, lam);
				else step_nomul(c, m, cs, ar, ak);
			}
		} else if (V == 2) {
#pragma unroll
			for (int i = 0; i < 32; i++) {
				step_full(c, m, z, 1u << i, cs, ar, ak, lam);
				step_full(c2, m2, ~z, 1u << i, cs, ar, ak, lam);
			}
		} else if (V == 3) {
#pragma unroll
			for (int i = 0; i < 32; i++) { c = __dadd_rn(c, ak); c = __dmul_rn(c, lam); }
		} else {
#pragma unroll
			for (int i = 0; i < 32; i++) c = __dadd_rn(c, ak);
		}
		acc ^= m ^ m2;
	}
	cp[tid] = c + c2;
	mp[tid] = acc;
}

int main()
{
	const int n = 1024;
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	const double sps = 40.0, thr = sps / 2 - 0.5, lam = 0.77;
	for (int blocks_per_sm = 3; blocks_per_sm <= 8; blocks_per_sm += (blocks_per_sm == 3 ? 3 : 2)) {
		const int blocks = prop.multiProcessorCount * blocks_per_sm;
		const size_t threads = (size_t)blocks * 128;
		std::vector<uint32_t> z(threads * n);
		srand(7);
		for (auto &v : z) {
			v = 0;
			for (int b = 0; b < 32; b++) if (rand() % 80 == 0) v |= 1u << b;
		}
		std::vector<double> c0(threads);
		for (auto &v : c0) v = -20.0 + (rand() % 4000) / 100.0;
		double *dc; uint32_t *dz, *dm;
		cudaMalloc(&dc, threads * 8); cudaMalloc(&dz, z.size() * 4); cudaMalloc(&dm, threads * 4);
		cudaMemcpy(dz, z.data(), z.size() * 4, cudaMemcpyHostToDevice);
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		for (int v = 0; v < 5; v++) {
			float best = 1e9f;
			for (int rep = 0; rep < 4; rep++) {
				cudaMemcpy(dc, c0.data(), threads * 8, cudaMemcpyHostToDevice);
				cudaEventRecord(e0);
				switch (v) {
				case 0: k<0><<<blocks, 128>>>(dc, dz, dm, thr - 1.0, -(sps - 1.0), 1.0, lam, n); break;
				case 1: k<1><<<blocks, 128>>>(dc, dz, dm, thr - 1.0, -(sps - 1.0), 1.0, lam, n); break;
				case 2: k<2><<<blocks, 128>>>(dc, dz, dm, thr - 1.0, -(sps - 1.0), 1.0, lam, n); break;
				case 3: k<3><<<blocks, 128>>>(dc, dz, dm, thr - 1.0, -(sps - 1.0), 1.0, lam, n); break;
				default: k<4><<<blocks, 128>>>(dc, dz, dm, thr - 1.0, -(sps - 1.0), 1.0, lam, n); break;
				}
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
				float ms; cudaEventElapsedTime(&ms, e0, e1);
				if (rep && ms < best) best = ms;
			}
			const double chains = (v == 2) ? 2.0 : 1.0;
			const double warp_samples = chains * (double)threads / 32.0 * n * 32.0;
			const double clk = prop.clockRate * 1e3;         // nominal; the real clock is sampled by the caller
			const double cyc = best * 1e-3 * clk * prop.multiProcessorCount * 4 / warp_samples;
			printf("warps/SM %2d variant %d: %.3f ms, %.2f scheduler cycles per warp-sample (at %.0f MHz), %.1f G thread-samples/s\n",
				blocks_per_sm * 4, v, best, cyc, clk / 1e6, warp_samples * 32 / best / 1e6);
		}
		cudaFree(dc); cudaFree(dz); cudaFree(dm);
	}
	cudaError_t err = cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(err));
	return err != cudaSuccess;
}

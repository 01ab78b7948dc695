// fir_umma.cu -- standalone experiment for DESIGN.md section 8.1: a 100-tap FIR as a tcgen05 GEMM whose data operand
// is the signal buffer itself.  NOT part of the product library; written without access to a GPU (it compiles for
// sm_100a; every run-time assumption is checked against a CPU reference when it is first run):
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fir_umma fir_umma.cu
//   timeout 60 ./fir_umma 0      # base_offset field of the shifted A descriptors left 0
//   timeout 60 ./fir_umma 1      # base_offset = (start address >> 7) & 7
//
// One CTA, one tile: D[128 x 64] (FP32, TMEM) = A[128 x 192] * B[192 x 64], bf16 operands.
//   A[i][d] = x[64 i + d]        K-major, SWIZZLE_128B: row i of the tile is 128 bytes = samples [64 i, 64 i + 64) of the
//                                 signal, which is how the signal lies in shared memory anyway (131 rows); the K blocks
//                                 64..127 and 128..191 are the same buffer seen through a descriptor that starts one and
//                                 two rows later.  Open question the two modes answer: how the swizzle phase of such a
//                                 start address (not 1024-byte aligned) has to be declared.
//   B[d][n] = h[d - n]           the banded Toeplitz matrix of the taps (0 outside [0, NTAPS)), N x K K-major SWIZZLE_128B,
//                                 three 64-wide K blocks of 64 rows.
// so D[i][n] = sum_t h[t] x[64 i + n + t] = output 64 i + n of the FIR (correlation form, taps already reversed).
// The run prints the largest deviation from a float64 reference on the same bf16-rounded operands and the cycles per
// 12-MMA tile when many tiles are issued back to back.
//
// Second part (fir_umma_split_kernel): the same tile on FP32 data and FP32 taps, each split into three bf16 pieces
// (x = x1 + x2 + x3 exactly), the six largest piece products accumulated into one TMEM tile -- the scheme DESIGN.md 8.1
// proposes for the front end's low-pass.  Prints its error and that of a sequential FP32 FMA loop (what the front end
// does today), both against float64 and relative to sum |h||x| (the guard scale): the emulation
// tools/lpf_split_error.py expects them to be equal (about 1e-7 RMS); what it cannot know is how the tensor core rounds
// inside and between the MMAs.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define NTAPS 100
#define ROWS 128          // M
#define NOUT 64           // N: outputs per row
#define KBLK 3            // K = 192 = 3 x 64
#define A_ROWS (ROWS + KBLK - 1 + 1)
#define A_BYTES (A_ROWS * 128)
#define B_BYTES (KBLK * NOUT * 128)
#define TMEM_COLS 64

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset inside a K-major SWIZZLE_128B region whose base is 1024-byte aligned: rows of 128 bytes, the 16-byte
// chunk index XORed with the row index mod 8 (Swizzle<3,4,3> on the address bits)
__host__ __device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte_in_row)
{
	const uint32_t o = row * 128u + byte_in_row;
	return o ^ (((o >> 7) & 7u) << 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int base_offset)
{
	// cute::UMMA::SmemDescriptor: start address >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version [46,48) = 1,
	// base_offset [49,52), layout type [61,64) = 2 (SWIZZLE_128B).  K-major SW128: LBO = 1 (unused), SBO = 1024 bytes.
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
	d |= (uint64_t)1u << 16;
	d |= (uint64_t)(1024u >> 4) << 32;
	d |= (uint64_t)1u << 46;
	d |= (uint64_t)(base_offset & 7) << 49;
	d |= (uint64_t)2u << 61;
	return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, long long max_spins)
{
	for (long long i = 0; i < max_spins; i++) {
		uint32_t ok;
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
		if (ok) return true;
	}
	return false;
}

// x: bf16 signal, at least 64 * A_ROWS samples; bt: the B region already laid out (swizzled) by the host
__global__ void __launch_bounds__(128, 1)
fir_umma_kernel(const __nv_bfloat16 *__restrict__ x, const uint8_t *__restrict__ bt, float *__restrict__ out,
                int base_offset_mode, int reps, long long *__restrict__ cycles, int *__restrict__ status)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	// the swizzle is a function of the address bits: make sure of the 1024-byte alignment (1 KB of slack is allocated)
	uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	uint8_t *sA = smem;                              // A_BYTES rounded up to 1024
	uint8_t *sB = smem + ((A_BYTES + 1023) / 1024) * 1024;
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmem_slot;
	const int tid = threadIdx.x, warp = tid >> 5;

	// stage the signal as rows of 128 bytes, swizzled by absolute position, and the prepared tap matrix
	for (int i = tid; i < A_ROWS * 8; i += blockDim.x) {             // 16-byte chunks
		const int row = i >> 3, chunk = i & 7;
		const uint4 v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(x) + row * 128 + chunk * 16);
		*reinterpret_cast<uint4 *>(sA + sw128(row, chunk * 16)) = v;
	}
	for (int i = tid; i < B_BYTES / 16; i += blockDim.x)
		*reinterpret_cast<uint4 *>(sB + i * 16) = *reinterpret_cast<const uint4 *>(bt + i * 16);
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the MMA's async proxy
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_d = tmem_slot;

	// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = BF16 [7,10) = 1, B = BF16 [10,13) = 1,
	// both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
	const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

	long long t0 = 0;
	bool ok = true;
	if (tid == 0) {
		t0 = clock64();
		for (int rep = 0; rep < reps; rep++) {
			for (int kb = 0; kb < KBLK; kb++) {
				const uint32_t a0 = smem_u32(sA) + 128u * kb;             // K block kb of row i = row i + kb of the buffer
				const uint32_t b0 = smem_u32(sB) + (uint32_t)(NOUT * 128) * kb;
				const int bo = base_offset_mode ? (int)((a0 >> 7) & 7u) : 0;
				for (int ks = 0; ks < 4; ks++)                             // 16 bf16 = 32 bytes per K step inside the 128-byte row
					umma_bf16(tmem_d, make_desc(a0 + 32u * ks, bo), make_desc(b0 + 32u * ks, 0), idesc,
						(rep | kb | ks) ? 1u : 0u);
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" :: "l"((unsigned long long)smem_u32(&bar)) : "memory");
	}
	ok = mbar_wait(smem_u32(&bar), 0, 1ll << 26);
	if (tid == 0) {
		cycles[0] = clock64() - t0;
		status[0] = ok ? 0 : 1;
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	if (ok) {
		// D row i = TMEM lane i; warp w reads lanes [32 w, 32 w + 32): one row of 64 columns per thread
		uint32_t v[64];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16) + 16u * j;
			asm volatile(
				"tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
				: "=r"(v[16 * j + 0]), "=r"(v[16 * j + 1]), "=r"(v[16 * j + 2]), "=r"(v[16 * j + 3]), "=r"(v[16 * j + 4]),
				  "=r"(v[16 * j + 5]), "=r"(v[16 * j + 6]), "=r"(v[16 * j + 7]), "=r"(v[16 * j + 8]), "=r"(v[16 * j + 9]),
				  "=r"(v[16 * j + 10]), "=r"(v[16 * j + 11]), "=r"(v[16 * j + 12]), "=r"(v[16 * j + 13]), "=r"(v[16 * j + 14]),
				  "=r"(v[16 * j + 15]) : "r"(taddr) : "memory");
		}
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
		for (int n = 0; n < 64; n++) out[tid * 64 + n] = __uint_as_float(v[n]) / (float)reps;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}

// x: FP32 signal; bt: three B regions (tap pieces 0, 1, 2), each laid out like the single one above
__global__ void __launch_bounds__(128, 1)
fir_umma_split_kernel(const float *__restrict__ x, const uint8_t *__restrict__ bt, float *__restrict__ out,
                      int base_offset_mode, long long *__restrict__ cycles, int *__restrict__ status)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	const uint32_t a_stride = ((A_BYTES + 1023) / 1024) * 1024;
	uint8_t *sA = smem;                              // 3 pieces
	uint8_t *sB = smem + 3 * a_stride;               // 3 pieces
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmem_slot;
	const int tid = threadIdx.x, warp = tid >> 5;

	for (int i = tid; i < A_ROWS * 8; i += blockDim.x) {             // 16-byte chunks = 8 samples
		const int row = i >> 3, chunk = i & 7;
		__nv_bfloat16 p[3][8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const float v = x[row * 64 + chunk * 8 + k];
			const __nv_bfloat16 p1 = __float2bfloat16_rn(v);
			const float r1 = v - __bfloat162float(p1);                   // exact
			const __nv_bfloat16 p2 = __float2bfloat16_rn(r1);
			const float r2 = r1 - __bfloat162float(p2);                  // exact
			p[0][k] = p1; p[1][k] = p2; p[2][k] = __float2bfloat16_rn(r2);
		}
#pragma unroll
		for (int q = 0; q < 3; q++)
			*reinterpret_cast<uint4 *>(sA + q * a_stride + sw128(row, chunk * 16)) = *reinterpret_cast<const uint4 *>(p[q]);
	}
	for (int i = tid; i < 3 * B_BYTES / 16; i += blockDim.x)
		*reinterpret_cast<uint4 *>(sB + i * 16) = *reinterpret_cast<const uint4 *>(bt + i * 16);
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_d = tmem_slot;
	const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

	long long t0 = 0;
	if (tid == 0) {
		t0 = clock64();
		// (tap piece, data piece): the small products first, the leading one last
		const int order[6][2] = {{2, 0}, {1, 1}, {0, 2}, {1, 0}, {0, 1}, {0, 0}};
		uint32_t first = 1;
		for (int o = 0; o < 6; o++)
			for (int kb = 0; kb < KBLK; kb++) {
				const uint32_t a0 = smem_u32(sA) + order[o][1] * a_stride + 128u * kb;
				const uint32_t b0 = smem_u32(sB) + order[o][0] * B_BYTES + (uint32_t)(NOUT * 128) * kb;
				const int bo = base_offset_mode ? (int)((a0 >> 7) & 7u) : 0;
				for (int ks = 0; ks < 4; ks++) {
					umma_bf16(tmem_d, make_desc(a0 + 32u * ks, bo), make_desc(b0 + 32u * ks, 0), idesc, first ? 0u : 1u);
					first = 0;
				}
			}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" :: "l"((unsigned long long)smem_u32(&bar)) : "memory");
	}
	const bool ok = mbar_wait(smem_u32(&bar), 0, 1ll << 26);
	if (tid == 0) {
		cycles[0] = clock64() - t0;
		status[0] = ok ? 0 : 1;
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (ok) {
		uint32_t v[64];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16) + 16u * j;
			asm volatile(
				"tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
				: "=r"(v[16 * j + 0]), "=r"(v[16 * j + 1]), "=r"(v[16 * j + 2]), "=r"(v[16 * j + 3]), "=r"(v[16 * j + 4]),
				  "=r"(v[16 * j + 5]), "=r"(v[16 * j + 6]), "=r"(v[16 * j + 7]), "=r"(v[16 * j + 8]), "=r"(v[16 * j + 9]),
				  "=r"(v[16 * j + 10]), "=r"(v[16 * j + 11]), "=r"(v[16 * j + 12]), "=r"(v[16 * j + 13]), "=r"(v[16 * j + 14]),
				  "=r"(v[16 * j + 15]) : "r"(taddr) : "memory");
		}
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
		for (int n = 0; n < 64; n++) out[tid * 64 + n] = __uint_as_float(v[n]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}

static float bf16_round(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
	memcpy(&f, &u, 4);
	return f;
}

int main(int argc, char **argv)
{
	const int mode = argc > 1 ? atoi(argv[1]) : 0;
	const int n_samples = 64 * A_ROWS;
	std::vector<float> xs(n_samples), h(NTAPS);
	srand(7);
	for (auto &v : xs) v = bf16_round((float)(rand() % 2001 - 1000) / 8.0f);
	double hs = 0;
	for (int t = 0; t < NTAPS; t++) { h[t] = 0.54f - 0.46f * cosf(2.0f * 3.14159265f * t / (NTAPS - 1)); hs += h[t]; }
	for (auto &v : h) v = bf16_round((float)(v / hs));
	std::vector<__nv_bfloat16> xb(n_samples);
	for (int i = 0; i < n_samples; i++) xb[i] = __float2bfloat16(xs[i]);
	// B region: K block kb, row n, column kk: element d = 64 kb + kk, value h[d - n]
	std::vector<uint8_t> bt(B_BYTES, 0);
	for (int kb = 0; kb < KBLK; kb++)
		for (int n = 0; n < NOUT; n++)
			for (int kk = 0; kk < 64; kk++) {
				const int t = 64 * kb + kk - n;
				const __nv_bfloat16 v = __float2bfloat16((t >= 0 && t < NTAPS) ? h[t] : 0.0f);
				memcpy(&bt[kb * NOUT * 128 + sw128(n, kk * 2)], &v, 2);
			}
	__nv_bfloat16 *dx; uint8_t *dbt; float *dout; long long *dcy; int *dst;
	cudaMalloc(&dx, n_samples * 2); cudaMalloc(&dbt, B_BYTES); cudaMalloc(&dout, ROWS * NOUT * 4);
	cudaMalloc(&dcy, 8); cudaMalloc(&dst, 4);
	cudaMemcpy(dx, xb.data(), n_samples * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(dbt, bt.data(), B_BYTES, cudaMemcpyHostToDevice);
	const size_t smem = ((A_BYTES + 1023) / 1024) * 1024 + B_BYTES + 1024;
	cudaFuncSetAttribute(fir_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	for (int reps : {1, 1000}) {
		cudaMemset(dout, 0, ROWS * NOUT * 4);
		fir_umma_kernel<<<1, 128, smem>>>(dx, dbt, dout, mode, reps, dcy, dst);
		cudaError_t ce = cudaDeviceSynchronize();
		if (ce != cudaSuccess) { printf("mode %d reps %d: CUDA error %s\n", mode, reps, cudaGetErrorString(ce)); return 2; }
		long long cy; int st;
		std::vector<float> out(ROWS * NOUT);
		cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(out.data(), dout, ROWS * NOUT * 4, cudaMemcpyDeviceToHost);
		if (st) { printf("mode %d reps %d: the MMAs never committed (mbarrier wait timed out)\n", mode, reps); return 3; }
		double worst = 0, ref_max = 0;
		for (int i = 0; i < ROWS; i++)
			for (int n = 0; n < NOUT; n++) {
				double r = 0;
				for (int t = 0; t < NTAPS; t++) r += (double)h[t] * xs[64 * i + n + t];
				worst = fmax(worst, fabs(r - out[i * NOUT + n]));
				ref_max = fmax(ref_max, fabs(r));
			}
		printf("mode %d reps %d: max |error| %.3e of max |y| %.3e (%s); %lld cycles = %.1f per 12-MMA tile (8192 outputs, one piece product)\n",
			mode, reps, worst, ref_max, worst <= 1e-4 * ref_max ? "MATCH" : "MISMATCH", cy, (double)cy / reps);
	}
	// ---- FP32 operands as three bf16 pieces each, six products -------------------------------------------------------
	{
		std::vector<float> xf(n_samples), hf(NTAPS);
		double sum = 0;
		for (int t = 0; t < NTAPS; t++) { hf[t] = (0.54f - 0.46f * cosf(2.0f * 3.14159265f * t / (NTAPS - 1))) * (1.0f + 0.013f * t); sum += hf[t]; }
		for (auto &v : hf) v = (float)(v / sum);
		for (int i = 0; i < n_samples; i++)      // magnitudes: large, positive, slowly varying + noise (like the front end's)
			xf[i] = 4.0e5f * (1.0f + 0.6f * sinf(i / 46.0f)) + 3.0e4f * (float)(rand() % 10007) / 10007.0f;
		std::vector<uint8_t> bt3(3 * B_BYTES, 0);
		for (int kb = 0; kb < KBLK; kb++)
			for (int n = 0; n < NOUT; n++)
				for (int kk = 0; kk < 64; kk++) {
					const int t = 64 * kb + kk - n;
					float rest = (t >= 0 && t < NTAPS) ? hf[t] : 0.0f;
					for (int q = 0; q < 3; q++) {
						const float pq = bf16_round(rest);
						const __nv_bfloat16 v = __float2bfloat16(pq);
						memcpy(&bt3[q * B_BYTES + kb * NOUT * 128 + sw128(n, kk * 2)], &v, 2);
						rest -= pq;                               // exact
					}
				}
		float *dxf; uint8_t *dbt3;
		cudaMalloc(&dxf, n_samples * 4); cudaMalloc(&dbt3, 3 * B_BYTES);
		cudaMemcpy(dxf, xf.data(), n_samples * 4, cudaMemcpyHostToDevice);
		cudaMemcpy(dbt3, bt3.data(), 3 * B_BYTES, cudaMemcpyHostToDevice);
		const size_t smem3 = 3 * (((A_BYTES + 1023) / 1024) * 1024) + 3 * B_BYTES + 1024;
		cudaFuncSetAttribute(fir_umma_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
		cudaMemset(dout, 0, ROWS * NOUT * 4);
		fir_umma_split_kernel<<<1, 128, smem3>>>(dxf, dbt3, dout, mode, dcy, dst);
		cudaError_t ce = cudaDeviceSynchronize();
		if (ce != cudaSuccess) { printf("split: CUDA error %s\n", cudaGetErrorString(ce)); return 2; }
		long long cy; int st;
		std::vector<float> out(ROWS * NOUT);
		cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(out.data(), dout, ROWS * NOUT * 4, cudaMemcpyDeviceToHost);
		if (st) { printf("split: the MMAs never committed\n"); return 3; }
		double e_tc = 0, e_f32 = 0, s_tc = 0, s_f32 = 0;
		for (int i = 0; i < ROWS; i++)
			for (int n = 0; n < NOUT; n++) {
				double r = 0, scale = 0;
				float acc = 0.f;
				for (int t = 0; t < NTAPS; t++) {
					r += (double)hf[t] * xf[64 * i + n + t];
					scale += fabs((double)hf[t] * xf[64 * i + n + t]);
					acc = fmaf(hf[t], xf[64 * i + n + t], acc);
				}
				const double a = fabs(r - out[i * NOUT + n]) / scale, b = fabs(r - acc) / scale;
				e_tc = fmax(e_tc, a); e_f32 = fmax(e_f32, b); s_tc += a * a; s_f32 += b * b;
			}
		printf("split 3 x bf16, 6 products (mode %d): error / sum|h||x|  rms %.2e max %.2e   | sequential FP32 FMA: rms %.2e max %.2e   | %lld cycles for 72 MMAs\n",
			mode, sqrt(s_tc / (ROWS * NOUT)), e_tc, sqrt(s_f32 / (ROWS * NOUT)), e_f32, cy);
	}
	return 0;
}

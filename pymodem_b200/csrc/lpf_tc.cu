// lpf_tc.cu -- the AFSK front end's output low-pass (afsk.py:162-166) on the 5th-generation tensor cores, and the
// epilogue that turns it into sign bits.
//
// After the sliding-window correlators the 100-tap low-pass of the four tone magnitudes was two thirds of the front
// end's multiply-adds (400 of 588 per sample) and half of its time on the FP32 pipe.  A FIR with constant taps is a
// banded-Toeplitz GEMM (pm_common.cuh, LpfTcPlan): with the magnitude stream laid out in rows of 64 samples the data
// operand IS the stream -- K blocks 1 and 2 are the same buffer seen through descriptors that start one and two rows
// later -- and the tap matrix is a 72 KB constant.  FP32 accuracy comes from three bf16 pieces per operand and the six
// largest piece products (measured on B200, tools/ubench/fir_umma.cu: 576 cycles per 12-MMA product, error 6.5e-8 rms
// of sum|h||m| against 1.5e-7 for the sequential FP32 FMAs it replaces).
//
// One persistent CTA per SM, six warps:
//   warp 5, one lane   producer: bulk-copies (cp.async.bulk, mbarrier complete_tx) the three piece buffers of one
//                      (tile, tone) item into a shared-memory slot.  Three slots: the copy for item i + 2 is issued as
//                      soon as item i - 1 has left its slot, so it has two items' time to land.
//   warp 4             MMA issuer: the whole warp runs the control loop (barrier waits, operand descriptors: warp-uniform
//                      values in uniform registers), one elected lane issues the 6 x NK tcgen05.mma of an item into that
//                      tone's 64 TMEM columns; tcgen05.commit releases the slot and, after a tile's last tone, hands the
//                      accumulators to the epilogue.
//   warps 0-3          epilogue: thread = TMEM lane = one row of 64 consecutive outputs.  Per tone pair and half row
//                      the mark and the space accumulators come out with tcgen05.ld (32 columns each), every chain of
//                      the pair forms y = L_mark - g L_space, the sign word (32 outputs = exactly one word of the
//                      slicer's sign stream) and the guard test (front.cu: |y| < eps (|L_mark| + g |L_space|) +
//                      c_abs-term) -- flagged samples are queued for the float64 fix-up as before.
// TMEM: 2 sets x 4 tones x 64 columns, so the MMAs of tile k + 1 run while the epilogue reads tile k.
#include "pm_common.cuh"

#define TC_THREADS 192
#define TC_SLOTS 3                        // operand slots: the copy for item i + 2 is in flight while item i + 1 waits and item i runs
#define TC_SMEM_BYTES (3 * TC_B_BYTES + TC_SLOTS * 3 * TC_A_STRIDE + 1024 + 128)   // + alignment slack + barriers: 231552 of the 232448 a CTA may have
#define TC_SPIN_LIMIT (1ll << 24)

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cute::UMMA::SmemDescriptor, K-major SWIZZLE_128B: start address >> 4 [0,14), LBO >> 4 [16,30) = 1 (unused), SBO >> 4
// [32,46) = 1024 bytes, version [46,48) = 1, base_offset [49,52) = 0 (also for starts that are whole rows into the
// swizzle atom: measured, fir_umma.cu mode 0), layout type [61,64) = 2
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr)
{
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
	d |= (uint64_t)1u << 16;
	d |= (uint64_t)(1024u >> 4) << 32;
	d |= (uint64_t)1u << 46;
	d |= (uint64_t)2u << 61;
	return d;
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// one lane of the (converged) warp, the way CUTLASS's elect_one_sync() asks for it: ptxas knows that the region this
// predicate guards has a single active thread
__device__ __forceinline__ bool tc_elect()
{
	uint32_t pred = 0;
	asm volatile(
		"{\n\t"
		".reg .pred P;\n\t"
		"elect.sync _|P, 0xFFFFFFFF;\n\t"
		"@P mov.s32 %0, 1;\n\t"
		"}\n" : "+r"(pred));
	return pred != 0;
}

__device__ __forceinline__ void tc_commit(uint32_t bar)
{
	// (with the state space spelled out the operand is the 32-bit shared address; without it the instruction takes a generic one)
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// bounded wait: a barrier that never completes (a bug, not a state the protocol knows) must not hang the GPU
__device__ __forceinline__ bool tc_wait(uint32_t bar, uint32_t parity, volatile int *abort_flag)
{
	for (long long i = 0; i < TC_SPIN_LIMIT; i++) {
		uint32_t ok;
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
		if (ok) return true;
		if ((i & 1023) == 1023 && *abort_flag) return false;
	}
	*abort_flag = 1;
	return false;
}

// 32 consecutive columns of this thread's TMEM lane.  The load is asynchronous: its destination registers are written
// some time after the instruction issues, and the compiler does not know -- left to itself it may move or copy them
// between the load and the wait (observed: the last registers of the last load came back stale, depending on timing).
// Load and wait therefore live in ONE asm statement, so the registers only become visible to the compiler once the
// data is there.
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t *v)
{
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
		"tcgen05.wait::ld.sync.aligned;"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
}

// NK: K steps of 16 per piece product, ceil((n_lpf + 63) / 16) -- the steps past the last tap hold only zeros and are not
// issued; a template parameter so that the issue loop is a straight run of tcgen05.mma with immediate descriptor offsets
template <bool WRITE_SOFT, int NK>
__global__ void __launch_bounds__(TC_THREADS, 1)
lpf_tc_kernel(const __grid_constant__ LpfTcPlan P, const unsigned char *__restrict__ mag, long long mag_rows,
              const unsigned char *__restrict__ btaps, const float *__restrict__ tile_amax, long long tile_first,
              long long n_tiles, uint32_t *__restrict__ sign, long long sign_stride, float *__restrict__ soft,
              long long soft_stride, GuardList guard, int *__restrict__ status)
{
	extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
	// the swizzle is a function of the address bits: 1024-byte alignment (1 KB of slack is allocated)
	unsigned char *smem = tc_smem_raw + ((1024u - (tc_smem_u32(tc_smem_raw) & 1023u)) & 1023u);
	unsigned char *sB = smem;                                   // 3 tap pieces x 3 K blocks x 64 rows x 128 bytes
	unsigned char *sA = smem + 3 * TC_B_BYTES;                  // TC_SLOTS slots x 3 data pieces x TC_A_STRIDE
	// barriers and two words of bookkeeping behind the operand buffers (no static shared memory: the kernel uses all but
	// 900 bytes of what a CTA can have): full_a[3], free_a[3], d_full[2], d_free[2]
	unsigned long long *bars = reinterpret_cast<unsigned long long *>(sA + TC_SLOTS * 3 * TC_A_STRIDE);
	uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(bars + 12);
	int &s_abort = *reinterpret_cast<int *>(bars + 13);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t bar_full_a = tc_smem_u32(&bars[0]), bar_free_a = tc_smem_u32(&bars[3]);
	const uint32_t bar_d_full = tc_smem_u32(&bars[6]), bar_d_free = tc_smem_u32(&bars[8]);

	// tiles of this CTA: tile_first + blockIdx.x, + gridDim.x, ...
	const long long my_tiles = (n_tiles > (long long)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

	for (int i = tid; i < 3 * TC_B_BYTES / 16; i += TC_THREADS)
		*reinterpret_cast<uint4 *>(sB + i * 16) = *reinterpret_cast<const uint4 *>(btaps + i * 16);
	if (tid == 0) {
		s_abort = 0;
		for (int b = 0; b < 8; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(&bars[b])) : "memory");
		for (int b = 8; b < 10; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(tc_smem_u32(&bars[b])) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 4) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&tmem_slot)), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the tap matrix was written with generic stores
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = tmem_slot;
	volatile int *abort_flag = &s_abort;

	if (warp == 4) {
		if (my_tiles > 0) {
			// The whole warp runs the control loop -- loop counters, barrier waits and operand descriptors are then
			// warp-uniform values that stay in uniform registers -- and one lane issues the copies, MMAs and commits.
			// (Issued from inside an `if (lane == 0)` region every tcgen05.mma had its operands moved from vector to uniform
			// registers through an elect/broadcast loop: 85 cycles of issue per MMA against the 48 the tensor core needs.)
			const bool leader = lane == 0;
			// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = BF16 [7,10) = 1, B = BF16 [10,13) = 1,
			// both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
			const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
			const long long n_items = my_tiles * P.n_mag;
			const uint64_t descB0 = tc_desc(tc_smem_u32(sB));
			// tracing (option "stage_clocks"): cycles spent waiting for data / for a slot / for the epilogue, and issuing
			const bool trace = guard.stage_clk != nullptr && leader;
			long long c_full = 0, c_free = 0, c_dfree = 0, c_issue = 0, t_a = 0;
			auto tick = [&]() { if (trace) t_a = clock64(); };
			auto tock = [&](long long &acc_c) { if (trace) acc_c += clock64() - t_a; };
			const long long t_loop0 = trace ? clock64() : 0;
			bool ok = true;
			for (long long it = 0; ok && it < n_items; it++) {
				const int slot = (int)(it % TC_SLOTS);
				const long long k = it / P.n_mag;
				const int tone = (int)(it % P.n_mag), set = (int)(k & 1);
				tick();
				if (!tc_wait(bar_full_a + 8u * slot, (uint32_t)((it / TC_SLOTS) & 1), abort_flag)) break;
				tock(c_full);
				// the epilogue has read this TMEM set (tile k - 2); the first two tiles pass at once
				tick();
				if (tone == 0 && !(P.debug_mask & 512) && !tc_wait(bar_d_free + 8u * set, (uint32_t)(((k >> 1) & 1) ^ 1), abort_flag)) break;
				tock(c_dfree);
				tick();
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t d_tmem = tmem_base + (uint32_t)((set * TC_MAX_TONES + tone) * TC_N);
				const uint64_t descA0 = tc_desc(tc_smem_u32(sA + slot * 3 * TC_A_STRIDE));
				// (tap piece, data piece): the small products first, the leading one last.  Plain loops over the products, the K
				// steps of a product unrolled with immediate descriptor offsets; the accumulate flag is off only for the very
				// first MMA of the item.
				if (!(P.debug_mask & 4)) {
#pragma unroll 1
					for (int o = 0; o < 6; o++) {
						const int tq = (0x210100 >> (4 * (5 - o))) & 3;          // tap piece of product o:  2 1 0 1 0 0
						const int dq = (0x012010 >> (4 * (5 - o))) & 3;          // data piece of product o: 0 1 2 0 1 0
						const uint64_t da0 = descA0 + (uint64_t)((dq * TC_A_STRIDE) >> 4);
						const uint64_t db0 = descB0 + (uint64_t)((tq * TC_B_BYTES) >> 4);
						if (tc_elect()) {
#pragma unroll
							for (int kk = 0; kk < NK; kk++) {
								const int kb = kk >> 2, ks = kk & 3;
								tc_mma(d_tmem, da0 + (uint64_t)((128 * kb + 32 * ks) >> 4), db0 + (uint64_t)((TC_N * 128 * kb + 32 * ks) >> 4),
									idesc, (kk == 0 && o == 0) ? 0u : 1u);
							}
						}
						__syncwarp();
					}
				}
				if (leader) {
					tc_commit(bar_free_a + 8u * slot);                  // the slot is free once these MMAs have read it
					if (tone == P.n_mag - 1) tc_commit(bar_d_full + 8u * set);
				}
				__syncwarp();
				tock(c_issue);
			}
			if (trace) {
				atomicAdd(&guard.stage_clk[0], (unsigned long long)c_full);
				atomicAdd(&guard.stage_clk[1], (unsigned long long)c_free);
				atomicAdd(&guard.stage_clk[2], (unsigned long long)c_dfree);
				atomicAdd(&guard.stage_clk[3], (unsigned long long)c_issue);
				atomicAdd(&guard.stage_clk[4], (unsigned long long)n_items);
				atomicAdd(&guard.stage_clk[7], (unsigned long long)(clock64() - t_loop0));     // the whole control loop of this CTA
			}
		}
	} else if (warp == 5) {
		// ---- producer warp: keeps the operand slots full, TC_SLOTS - 1 items ahead of the MMAs ----
		if (my_tiles > 0) {
			const bool leader = lane == 0;
			const long long n_items = my_tiles * P.n_mag;
			for (long long it = 0; it < n_items; it++) {
				const int slot = (int)(it % TC_SLOTS);
				// the slot's previous tenant (item it - TC_SLOTS) has been consumed; the first uses pass at once
				if (!tc_wait(bar_free_a + 8u * slot, (uint32_t)(((it / TC_SLOTS) & 1) ^ 1), abort_flag)) break;
				const long long tile = tile_first + blockIdx.x + (it / P.n_mag) * gridDim.x;
				const int tone = (int)(it % P.n_mag);
				const uint32_t bar = bar_full_a + 8u * slot;
				if (leader) {
					if (P.debug_mask & 2) {
						asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
					} else {
						asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(3u * TC_A_BYTES) : "memory");
#pragma unroll
						for (int q = 0; q < 3; q++) {
							const unsigned char *src = mag + ((long long)(tone * 3 + q) * mag_rows + tile * TC_ROWS) * 128;
							const uint32_t dst = tc_smem_u32(sA + (slot * 3 + q) * TC_A_STRIDE);
							asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
								:: "r"(dst), "l"(src), "r"((uint32_t)TC_A_BYTES), "r"(bar) : "memory");
						}
					}
				}
				__syncwarp();
			}
		}
	} else {
		// ---- epilogue warps: thread = TMEM lane = row of 64 outputs ----
		const int row = tid;                                         // 0 .. 127 (warp w reads lanes [32 w, 32 w + 32))
		for (long long k = 0; k < ((P.debug_mask & 512) ? 0 : my_tiles); k++) {      // (512: the epilogue warps leave at once, timing experiments)
			const int set = (int)(k & 1);
			const long long tile = tile_first + blockIdx.x + k * gridDim.x;
			const long long t_e0 = (guard.stage_clk && tid == 0) ? clock64() : 0;
			if (!tc_wait(bar_d_full + 8u * set, (uint32_t)((k >> 1) & 1), abort_flag)) break;
			const long long t_e1 = (guard.stage_clk && tid == 0) ? clock64() : 0;
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const long long n_row = tile * TC_TILE + (long long)row * TC_N;      // first output of this row
			// raw-input term of the guard: largest |sample| of the front tiles this row's outputs depend on
			float amax = 0.f;
			{
				long long t0 = n_row / P.tile_a, t1 = (n_row + TC_N - 1 + P.reach) / P.tile_a;
				if (t1 >= P.n_tile_a) t1 = P.n_tile_a - 1;
				if (!(P.debug_mask & 8))
					for (long long t = t0; t <= t1; t++) amax = fmaxf(amax, tile_amax[t]);
			}
			for (int p = 0; p < ((P.debug_mask & 64) ? 0 : P.n_pair); p++) {      // (64: no epilogue at all, timing experiments)
				const uint32_t t_mark = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)((set * TC_MAX_TONES + P.pair_mark[p]) * TC_N);
				const uint32_t t_space = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)((set * TC_MAX_TONES + P.pair_space[p]) * TC_N);
#pragma unroll 1
				for (int h = 0; h < 2; h++) {
					uint32_t lm[32], ls[32];
					tc_ld32(t_mark + 32 * h, lm);
					tc_ld32(t_space + 32 * h, ls);
					const long long nbase = n_row + 32 * h;
					for (int c = P.pair_first[p]; c < P.pair_first[p + 1]; c++) {
						const float g = P.chain_gain[c];
						const float neg_eps = -P.guard_eps;
						const float abs_c = P.chain_guard_abs[c] * amax;
						// most recent sample first, so that sample i ends up in bit i (front.cu epilogue)
						unsigned int neg = 0, near = 0;
#pragma unroll
						for (int r = 31; r >= 0; r--) {
							const float m = __uint_as_float(lm[r]), s = __uint_as_float(ls[r]);
							const float y = fmaf(-g, s, m);
							const float scale = fmaf(g, fabsf(s), fabsf(m));
							const float d = fmaf(neg_eps, scale, fabsf(y) - abs_c);
							neg = __funnelshift_l(__float_as_uint(y), neg, 1);
							near = __funnelshift_l(__float_as_uint(d), near, 1);
						}
						const int gid = P.chain_gid[c];
						const long long nout = P.chain_nout[c];
						if (nbase < nout && !(P.debug_mask & 1)) {
							sign[gid * sign_stride + (nbase >> 5)] = ~neg;
							if (near) {
								for (int r = 0; r < 32; r++) {
									if (nbase + r >= nout) break;
									if ((near >> r) & 1u) {
										unsigned int slot = atomicAdd(guard.count, 1u);
										if (slot < guard.cap)
											guard.entries[slot] = ((unsigned long long)gid << 48) | (unsigned long long)(nbase + r);
									}
								}
							}
							if (WRITE_SOFT) {
#pragma unroll
								for (int r = 0; r < 32; r++)
									if (nbase + r < nout)
										soft[gid * soft_stride + nbase + r] = fmaf(-g, __uint_as_float(ls[r]), __uint_as_float(lm[r]));
							}
						}
					}
				}
			}
			// this thread's loads of the set are complete (wait::ld above): hand the accumulators back
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar_d_free + 8u * set) : "memory");
			if (guard.stage_clk && tid == 0) {
				atomicAdd(&guard.stage_clk[5], (unsigned long long)(t_e1 - t_e0));          // waiting for the accumulators
				atomicAdd(&guard.stage_clk[6], (unsigned long long)(clock64() - t_e1));    // epilogue work of one tile
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 4)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
	if (tid == 0 && s_abort) atomicExch(status, 1);
}

// ---------------------------------------------------------------------------------------------------------
template <int NK>
static cudaError_t launch_lpf_tc_nk(const LpfTcPlan *plan, const unsigned char *mag, long long mag_rows, const unsigned char *btaps,
	const float *tile_amax, long long tile_first, long long n_tiles, uint32_t *sign, long long sign_stride, float *soft,
	long long soft_stride, GuardList guard, int *status, int grid, cudaStream_t st)
{
	static bool attr_done = false;
	if (!attr_done) {
		cudaError_t e1 = cudaFuncSetAttribute(lpf_tc_kernel<false, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
		cudaError_t e2 = cudaFuncSetAttribute(lpf_tc_kernel<true, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
		if (e1 != cudaSuccess) return e1;
		if (e2 != cudaSuccess) return e2;
		attr_done = true;
	}
	if (soft)
		lpf_tc_kernel<true, NK><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(*plan, mag, mag_rows, btaps, tile_amax, tile_first,
			n_tiles, sign, sign_stride, soft, soft_stride, guard, status);
	else
		lpf_tc_kernel<false, NK><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(*plan, mag, mag_rows, btaps, tile_amax, tile_first,
			n_tiles, sign, sign_stride, soft, soft_stride, guard, status);
	return cudaGetLastError();
}

extern "C" cudaError_t pm_launch_lpf_tc(const LpfTcPlan *plan, const unsigned char *mag, long long mag_rows,
	const unsigned char *btaps, const float *tile_amax, long long tile_first, long long n_tiles, uint32_t *sign,
	long long sign_stride, float *soft, long long soft_stride, GuardList guard, int *status, int sm_count, cudaStream_t st)
{
	if (n_tiles <= 0) return cudaSuccess;
	const int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
	const int nk = (plan->n_lpf + 63 + 15) >> 4;          // 5 (8 taps) .. 11 (TC_MAX_LPF taps)
	pm_kt_mark("lpf_tc_kernel", st);
#define TC_CASE(N) case N: return launch_lpf_tc_nk<N>(plan, mag, mag_rows, btaps, tile_amax, tile_first, n_tiles, sign, sign_stride, soft, soft_stride, guard, status, grid, st)
	switch (nk) {
		TC_CASE(5); TC_CASE(6); TC_CASE(7); TC_CASE(8); TC_CASE(9); TC_CASE(10); TC_CASE(11);
	}
#undef TC_CASE
	return cudaErrorInvalidValue;
}

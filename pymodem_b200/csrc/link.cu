// link.cu -- the shard hand-off of a multi-GPU run done by the GPUs themselves, over NVLink peer memory.
//
// One recording is split on the sample axis over the ranks of one NVSwitch box (pymodem_b200/sharded.py).  What the
// ranks owe each other per run is tiny -- slicer end states + symbol counts (40 B per chain), the last bits of each
// rank's stream for frames that straddle a boundary (2 KB per chain), and the decoded packet records -- but every
// host-driven collective costs ~0.25 ms of launch/sync latency, three of them per 14 ms step.  Here every rank owns a
// "link" buffer that all peers map (cudaIpc*), producers store straight into the consumers' buffers and raise an
// epoch flag, consumers spin on the flag inside a one-block kernel on their own stream.  The whole step is then ONE
// stream of kernels per rank with no host round trip between the slicer and the final record copy:
//
//   push states -> wait all states (verify every hand-off, prefix-sum the symbol counts, place the own bits)
//   -> gather -> tail to rank+1 -> wait tail from rank-1 -> descramble/decode -> push records to everyone
//   -> wait all records -> merge (chain-major block interleave, arena offsets rebased) -> one D2H.
//
// Every rank evaluates "all hand-offs verified" on the same data, so all of them agree on whether the fast path
// holds; when a speculated start state was wrong (rare: ~1 % of segment boundaries), the host falls back to the
// repair protocol of sharded.py.  Flags carry the run's epoch and the buffers are double-buffered on its parity: a
// rank can be at most one run ahead of the slowest one (its own run only completes when everybody's records arrived).
// A waiter gives up after a few seconds and reports an error instead of hanging the GPU.
#include "pm_common.cuh"
#include "../../include/pymodem_b200.h"

#define LINK_SPIN_LIMIT_NS 8000000000ll

__device__ __forceinline__ unsigned int ld_flag(const unsigned int *p)
{
	unsigned int v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_flag(unsigned int *p, unsigned int v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ long long now_ns()
{
	long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// spin until *flag == epoch; false on timeout
__device__ bool link_spin(const unsigned int *flag, unsigned int epoch)
{
	const long long t0 = now_ns();
	while (ld_flag(flag) != epoch) {
		__nanosleep(200);
		if (now_ns() - t0 > LINK_SPIN_LIMIT_NS) return false;
	}
	return true;
}

__device__ __forceinline__ unsigned char *slot_of(unsigned char *base, const LinkGeom &G, int parity)
{
	return base + (long long)parity * G.slot_bytes;
}

// ---- 1. states ------------------------------------------------------------------------------------------
// own (start state, end state, symbols in the own range) of every chain -> slot.states[rank] of EVERY rank
__global__ void __launch_bounds__(256)
link_push_states_kernel(LinkGeom G, LinkPeers peers, int parity, unsigned int epoch,
                        const SegState *__restrict__ S, const SegState *__restrict__ E, int n_seg, int k_end,
                        const unsigned long long *__restrict__ symcount, const unsigned int *__restrict__ counters,
                        int fast_passes, unsigned int guard_cap)
{
	// The run was enqueued without a host round trip: whether this rank's own verify passes converged (some pass
	// repaired nothing) and whether its guard list held is only known here, on the device.  If not, the states below are
	// not final: n_symbols = -1 tells every rank to take the host-driven protocol (link_wait_states_kernel).
	bool unsettled = false;
	if (counters) {
		bool conv = false;
		for (int p = 0; p < fast_passes; p++) conv = conv || counters[2 + p] == 0u;
		unsettled = !conv || counters[0] > guard_cap;
	}
	for (int c = threadIdx.x; c < G.nc; c += blockDim.x) {
		pm_shard_state st;
		const SegState s0 = S[(long long)c * n_seg], e1 = E[(long long)c * n_seg + (k_end - 1)];
		st.start_clock = s0.clock; st.start_last = s0.last; st.start_last_q = s0.last_q;
		st.end_clock = e1.clock; st.end_last = e1.last; st.end_last_q = e1.last_q;
		st.n_symbols = unsettled ? -1ll : (long long)symcount[c];
		for (int q = 0; q < G.world; q++) {
			pm_shard_state *dst = reinterpret_cast<pm_shard_state *>(slot_of(peers.base[q], G, parity) + G.off_states);
			dst[(long long)G.rank * G.nc + c] = st;
		}
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x < G.world)
		st_flag(reinterpret_cast<unsigned int *>(slot_of(peers.base[threadIdx.x], G, parity) + G.off_sflag) + G.rank, epoch);
}

// wait for the states of all ranks; status[0] = every hand-off verified; place the own bits (ShardBits)
__global__ void __launch_bounds__(256)
link_wait_states_kernel(LinkGeom G, unsigned char *own, int parity, unsigned int epoch,
                        const BitChain *__restrict__ chains, int first, int last, int tail_bits,
                        ShardBits *__restrict__ sb, int *__restrict__ status, long long *__restrict__ A0_out)
{
	__shared__ int s_ok, s_bad, s_err;
	unsigned char *slot = slot_of(own, G, parity);
	if (threadIdx.x == 0) { s_ok = 1; s_bad = 0; s_err = 0; }
	__syncthreads();
	if (threadIdx.x < G.world)
		if (!link_spin(reinterpret_cast<unsigned int *>(slot + G.off_sflag) + threadIdx.x, epoch)) atomicExch(&s_ok, 0);
	__syncthreads();
	if (!s_ok) {
		// a peer never showed up: leave an empty, bounded placement behind so that the kernels already queued
		// after this one do nothing harmful, and report
		for (int c = threadIdx.x; c < G.nc; c += blockDim.x) {
			ShardBits b;
			b.bit_off = 0; b.own_lo = 0; b.own_hi = 0; b.valid_from = 0; b.first = 1; b.pad = 0;
			sb[c] = b;
			A0_out[c] = 0;
		}
		if (threadIdx.x == 0) { status[0] = 0; status[1] = PM_ERR_STATE; status[2] = 1; }
		return;
	}
	__threadfence_system();
	const pm_shard_state *st = reinterpret_cast<const pm_shard_state *>(slot + G.off_states);
	for (int i = threadIdx.x; i < (G.world - 1) * G.nc; i += blockDim.x) {
		const int q = 1 + i / G.nc, c = i - (q - 1) * G.nc;
		const pm_shard_state a = st[(long long)(q - 1) * G.nc + c], b = st[(long long)q * G.nc + c];
		if (__double_as_longlong(a.end_clock) != __double_as_longlong(b.start_clock) || a.end_last != b.start_last ||
		    a.end_last_q != b.start_last_q)
			atomicExch(&s_bad, 1);
	}
	for (int i = threadIdx.x; i < G.world * G.nc; i += blockDim.x)
		if (st[i].n_symbols < 0) atomicExch(&s_bad, 1);          // some rank's own slicer pass is not settled yet
	__syncthreads();                                             // s_bad is final: the placement below reads it
	for (int c = threadIdx.x; c < G.nc; c += blockDim.x) {
		long long P = 0;
		for (int q = 0; q < G.rank; q++) {
			const long long v = st[(long long)q * G.nc + c].n_symbols;
			if (v > 0) P += v;
		}
		long long n_own = st[(long long)G.rank * G.nc + c].n_symbols;
		if (n_own < 0) n_own = 0;
		ShardBits b;
		b.first = first; b.pad = 0; b.valid_from = 0;
		if (first) { b.bit_off = 0; b.own_lo = 0; A0_out[c] = 0; }
		else {
			if (P < tail_bits && !s_bad) atomicExch(&s_err, 1);
			const long long A0 = ((P - tail_bits) >> 3) << 3;          // global bit index of local bit 0
			A0_out[c] = A0;
			b.bit_off = P - A0;
			b.own_lo = b.bit_off;
			int deg = 0;
			for (unsigned long long q = chains[c].lfsr_poly; q > 1; q >>= 1) deg++;
			b.valid_from = b.bit_off - tail_bits + deg;
		}
		if (last) b.own_hi = 0x7fffffffffffffffll;
		else {
			if (n_own < tail_bits && !s_bad) atomicExch(&s_err, 1);
			b.own_hi = b.bit_off + n_own;
		}
		sb[c] = b;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		status[0] = s_bad ? 0 : 1;
		status[1] = s_err ? PM_ERR_CAPACITY : 0;
		status[2] = 0;
	}
}

// ---- flags ----------------------------------------------------------------------------------------------
__global__ void link_set_flag_kernel(unsigned int *flag, unsigned int epoch)
{
	__threadfence_system();
	st_flag(flag, epoch);
}

__global__ void link_wait_flag_kernel(const unsigned int *flag, unsigned int epoch, int *status)
{
	if (!link_spin(flag, epoch)) { status[1] = PM_ERR_STATE; status[2] = 2; }
	__threadfence_system();
}

// ---- IL2P decoder state ---------------------------------------------------------------------------------
// The IL2P walk of a chain (il2p.cu: sync search, frame by frame) crosses the shard boundary with three numbers: where
// the search resumes (a GLOBAL stream bit), in which mode, and the corrected-byte count of failed frames not yet
// attributed (il2p.py:200-211).  Rank r - 1 stores them into rank r's slot when its own walk is done; rank r waits,
// rebases the position onto its local stream (A0 = global index of its local bit 0) and checks that everything the
// walk will look at lies in the hand-off tail -- the same test the host protocol makes (engine.cu shard_finish_impl);
// a frame that reaches back further marks the chain (tail_short) and the run recovers from the gathered bitstream.
// The walks of the ranks are serial by nature; what the link removes is the host round trip between them.
__global__ void link_il2p_in_kernel(LinkGeom G, unsigned char *own, int parity, unsigned int epoch, int first,
                                    const BitChain *__restrict__ chains, const ShardBits *__restrict__ sb,
                                    const long long *__restrict__ A0, Il2pHand *__restrict__ hand,
                                    ChainCounters *__restrict__ cc, int *__restrict__ status)
{
	__shared__ int s_ok;
	unsigned char *slot = slot_of(own, G, parity);
	if (threadIdx.x == 0) {
		s_ok = 1;
		if (!first && !link_spin(reinterpret_cast<const unsigned int *>(slot + G.off_iflag), epoch)) {
			s_ok = 0;
			status[1] = PM_ERR_STATE; status[2] = 2;
		}
		__threadfence_system();
	}
	__syncthreads();
	const pm_il2p_state *prev = reinterpret_cast<const pm_il2p_state *>(slot + G.off_il2p);
	for (int c = threadIdx.x; c < G.nc; c += blockDim.x) {
		Il2pHand h;
		h.pos = 0; h.mode = 0; h.leak = 0;
		if (!first && chains[c].codec == PM_CODEC_IL2P) {
			if (s_ok) {
				h.pos = prev[c].pos - A0[c];
				h.mode = prev[c].mode;
				h.leak = prev[c].leak;
			}
			const long long need = h.mode == 1 ? h.pos - 8 : h.pos - 63;
			if (!s_ok || h.mode > 2 || h.mode == 0 || need < sb[c].valid_from) {
				// the walk would restart outside what this shard holds: plain search from the first own bit, and say so
				h.pos = sb[c].own_lo; h.mode = 2; h.leak = 0;
				cc[c].tail_short = 1;
			}
		}
		hand[c] = h;
	}
}

__global__ void link_il2p_out_kernel(LinkGeom G, LinkPeers peers, int parity, unsigned int epoch,
                                     const Il2pHand *__restrict__ hand_out, const long long *__restrict__ A0)
{
	unsigned char *slot = slot_of(peers.base[G.rank + 1], G, parity);
	pm_il2p_state *dst = reinterpret_cast<pm_il2p_state *>(slot + G.off_il2p);
	for (int c = threadIdx.x; c < G.nc; c += blockDim.x) {
		pm_il2p_state o;
		o.pos = hand_out[c].pos + A0[c];
		o.mode = hand_out[c].mode;
		o.leak = hand_out[c].leak;
		dst[c] = o;
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) st_flag(reinterpret_cast<unsigned int *>(slot + G.off_iflag), epoch);
}

// ---- 2. records -----------------------------------------------------------------------------------------
// own records + arena -> slot.rdata[rank] of every rank (blockIdx.y = destination rank)
__global__ void __launch_bounds__(256)
link_push_records_kernel(LinkGeom G, LinkPeers peers, int parity, const PacketRecDev *__restrict__ recs,
                         const uint8_t *__restrict__ arena, const PacketTotals *__restrict__ totals, int *status)
{
	if (status[3] != 0) return;                    // nothing to trust in this shard's records
	const unsigned long long n = totals->n_packets, nb = totals->n_bytes;
	const unsigned long long rec_bytes = n * sizeof(PacketRecDev);
	if (rec_bytes + nb > (unsigned long long)G.rec_region) {
		if (blockIdx.x == 0 && threadIdx.x == 0) { status[1] = PM_ERR_CAPACITY; status[2] = 3; }
		return;
	}
	unsigned char *dst = slot_of(peers.base[blockIdx.y], G, parity) + G.off_rdata + (long long)G.rank * G.rec_region;
	const unsigned long long *src8 = reinterpret_cast<const unsigned long long *>(recs);
	unsigned long long *dst8 = reinterpret_cast<unsigned long long *>(dst);
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	for (unsigned long long i = t; i < rec_bytes / 8; i += stride) dst8[i] = src8[i];
	unsigned char *da = dst + rec_bytes;
	for (unsigned long long i = t; i < nb; i += stride) da[i] = arena[i];
}

// A shard whose decode could not be completed from what it holds -- a frame reaching back past the hand-off tail, or a
// gap long enough to overflow max_packet_length (ax25.py:46-51: the sequential replay needs a known state) -- says so
// in status[3]; every rank learns it with the records and all of them recover together from the gathered bitstream
// (sharded.py: pm_engine_shard_export -> pm_engine_decode_stream).
#define LINK_HDR_FAILED  (~0ull)
#define LINK_HDR_RECOVER (~0ull - 1ull)
__global__ void link_check_decode_kernel(const ChainCounters *__restrict__ cc, int nc, int *status)
{
	int bad = 0;
	for (int c = threadIdx.x; c < nc; c += blockDim.x)
		if (cc[c].tail_short || cc[c].seq_needed) bad = 1;
	if (__any_sync(0xffffffffu, bad) && threadIdx.x == 0) status[3] = 1;
}

__global__ void link_publish_records_kernel(LinkGeom G, LinkPeers peers, int parity, unsigned int epoch,
                                            const PacketTotals *__restrict__ totals, const int *status)
{
	const int q = threadIdx.x;
	if (q >= G.world) return;
	unsigned char *slot = slot_of(peers.base[q], G, parity);
	unsigned long long *hdr = reinterpret_cast<unsigned long long *>(slot + G.off_rhdr) + 2 * G.rank;
	const bool ok = status[1] == 0 && status[3] == 0;
	hdr[0] = ok ? totals->n_packets : (status[1] != 0 ? LINK_HDR_FAILED : LINK_HDR_RECOVER);   // not ok: do not trust the region
	hdr[1] = ok ? totals->n_bytes : 0ull;
	__threadfence_system();
	st_flag(reinterpret_cast<unsigned int *>(slot + G.off_rflag) + G.rank, epoch);
}

// wait for every rank's records, then plan the merge: lb[q][c] = first record of rank q with chain >= c,
// obase[q][c] = position of that block in the merged table, abase[q] = offset of rank q's arena
__global__ void __launch_bounds__(256)
link_merge_plan_kernel(LinkGeom G, unsigned char *own, int parity, unsigned int epoch,
                       unsigned int *__restrict__ lb, unsigned long long *__restrict__ obase,
                       unsigned long long *__restrict__ abase, PacketTotals *__restrict__ merged_totals, int *status)
{
	__shared__ int s_ok;
	unsigned char *slot = slot_of(own, G, parity);
	if (threadIdx.x == 0) s_ok = 1;
	__syncthreads();
	if (threadIdx.x < G.world)
		if (!link_spin(reinterpret_cast<unsigned int *>(slot + G.off_rflag) + threadIdx.x, epoch)) atomicExch(&s_ok, 0);
	__syncthreads();
	__threadfence_system();
	const unsigned long long *hdr = reinterpret_cast<const unsigned long long *>(slot + G.off_rhdr);
	if (threadIdx.x == 0) {
		if (!s_ok) { status[1] = PM_ERR_STATE; status[2] = 4; }
		for (int q = 0; q < G.world; q++) {
			if (s_ok && hdr[2 * q] == LINK_HDR_FAILED && status[1] == 0) { status[1] = PM_ERR_STATE; status[2] = 5; }
			if (s_ok && hdr[2 * q] == LINK_HDR_RECOVER) status[3] = 1;
		}
	}
	__syncthreads();
	if (status[1] != 0 || status[3] != 0) {
		if (threadIdx.x == 0) { merged_totals->n_packets = 0; merged_totals->n_bytes = 0; }
		return;
	}
	const int nc1 = G.nc + 1;
	for (int i = threadIdx.x; i < G.world * nc1; i += blockDim.x) {
		const int q = i / nc1, c = i - q * nc1;
		const PacketRecDev *r = reinterpret_cast<const PacketRecDev *>(slot + G.off_rdata + (long long)q * G.rec_region);
		unsigned int lo = 0, hi = (unsigned int)hdr[2 * q];
		while (lo < hi) {
			const unsigned int mid = (lo + hi) >> 1;
			if (r[mid].chain < (unsigned int)c) lo = mid + 1; else hi = mid;
		}
		lb[i] = lo;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long pos = 0, ab = 0;
		for (int c = 0; c < G.nc; c++)
			for (int q = 0; q < G.world; q++) {
				obase[q * nc1 + c] = pos;
				pos += lb[q * nc1 + c + 1] - lb[q * nc1 + c];
			}
		for (int q = 0; q < G.world; q++) { abase[q] = ab; ab += hdr[2 * q + 1]; }
		merged_totals->n_packets = pos;
		merged_totals->n_bytes = ab;
	}
}

// blockIdx.y = source rank
__global__ void __launch_bounds__(256)
link_merge_write_kernel(LinkGeom G, unsigned char *own, int parity, const unsigned int *__restrict__ lb,
                        const unsigned long long *__restrict__ obase, const unsigned long long *__restrict__ abase,
                        PacketRecDev *__restrict__ out_recs, unsigned long long rec_cap,
                        uint8_t *__restrict__ out_arena, unsigned long long arena_cap, const int *status)
{
	if (status[1] != 0 || status[3] != 0) return;
	unsigned char *slot = slot_of(own, G, parity);
	const int q = blockIdx.y, nc1 = G.nc + 1;
	const unsigned long long *hdr = reinterpret_cast<const unsigned long long *>(slot + G.off_rhdr);
	const unsigned long long n = hdr[2 * q], nb = hdr[2 * q + 1];
	const unsigned char *region = slot + G.off_rdata + (long long)q * G.rec_region;
	const PacketRecDev *r = reinterpret_cast<const PacketRecDev *>(region);
	const unsigned char *a = region + n * sizeof(PacketRecDev);
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long ab = abase[q];
	for (unsigned long long i = t; i < n; i += stride) {
		PacketRecDev p = r[i];
		const unsigned long long dst = obase[q * nc1 + p.chain] + (i - lb[q * nc1 + p.chain]);
		p.offset += ab;
		if (dst < rec_cap) out_recs[dst] = p;
	}
	for (unsigned long long i = t; i < nb; i += stride)
		if (ab + i < arena_cap) out_arena[ab + i] = a[i];
}

// ---------------------------------------------------------------------------------------------------------
extern "C" {

// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, and loading can wait for
// kernels that are already running -- such as this rank's own spinning wait kernel.  Load everything up front.
cudaError_t pm_link_preload(void)
{
	cudaFuncAttributes a;
	cudaError_t e;
	if ((e = cudaFuncGetAttributes(&a, link_push_states_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_wait_states_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_set_flag_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_wait_flag_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_push_records_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_check_decode_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_publish_records_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_il2p_in_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_il2p_out_kernel)) != cudaSuccess) return e;
	if ((e = cudaFuncGetAttributes(&a, link_merge_plan_kernel)) != cudaSuccess) return e;
	return cudaFuncGetAttributes(&a, link_merge_write_kernel);
}

cudaError_t pm_link_push_states(LinkGeom G, LinkPeers peers, int parity, unsigned int epoch, const SegState *S,
	const SegState *E, int n_seg, int k_end, const unsigned long long *symcount, const unsigned int *counters,
	int fast_passes, unsigned int guard_cap, cudaStream_t st)
{
	pm_kt_mark("link_push_states_kernel", st);
	link_push_states_kernel<<<1, 256, 0, st>>>(G, peers, parity, epoch, S, E, n_seg, k_end, symcount, counters, fast_passes,
		guard_cap);
	return cudaGetLastError();
}

cudaError_t pm_link_wait_states(LinkGeom G, unsigned char *own, int parity, unsigned int epoch, const BitChain *chains,
	int first, int last, int tail_bits, ShardBits *sb, int *status, long long *A0_out, cudaStream_t st)
{
	pm_kt_mark("link_wait_states_kernel", st);
	link_wait_states_kernel<<<1, 256, 0, st>>>(G, own, parity, epoch, chains, first, last, tail_bits, sb, status, A0_out);
	return cudaGetLastError();
}

cudaError_t pm_link_il2p_in(LinkGeom G, unsigned char *own, int parity, unsigned int epoch, int first, const BitChain *chains,
	const ShardBits *sb, const long long *A0, Il2pHand *hand, ChainCounters *cc, int *status, cudaStream_t st)
{
	pm_kt_mark("link_il2p_in_kernel", st);
	link_il2p_in_kernel<<<1, 128, 0, st>>>(G, own, parity, epoch, first, chains, sb, A0, hand, cc, status);
	return cudaGetLastError();
}

cudaError_t pm_link_il2p_out(LinkGeom G, LinkPeers peers, int parity, unsigned int epoch, const Il2pHand *hand_out,
	const long long *A0, cudaStream_t st)
{
	pm_kt_mark("link_il2p_out_kernel", st);
	link_il2p_out_kernel<<<1, 128, 0, st>>>(G, peers, parity, epoch, hand_out, A0);
	return cudaGetLastError();
}

cudaError_t pm_link_set_flag(unsigned int *flag, unsigned int epoch, cudaStream_t st)
{
	pm_kt_mark("link_set_flag_kernel", st);
	link_set_flag_kernel<<<1, 1, 0, st>>>(flag, epoch);
	return cudaGetLastError();
}

cudaError_t pm_link_wait_flag(const unsigned int *flag, unsigned int epoch, int *status, cudaStream_t st)
{
	pm_kt_mark("link_wait_flag_kernel", st);
	link_wait_flag_kernel<<<1, 1, 0, st>>>(flag, epoch, status);
	return cudaGetLastError();
}

cudaError_t pm_link_push_records(LinkGeom G, LinkPeers peers, int parity, unsigned int epoch, const PacketRecDev *recs,
	const uint8_t *arena, const PacketTotals *totals, const ChainCounters *cc, int *status, cudaStream_t st)
{
	pm_kt_mark("link_check_decode_kernel", st);
	link_check_decode_kernel<<<1, 32, 0, st>>>(cc, G.nc, status);
	pm_kt_mark("link_push_records_kernel", st);
	link_push_records_kernel<<<dim3(32, G.world), 256, 0, st>>>(G, peers, parity, recs, arena, totals, status);
	pm_kt_mark("link_publish_records_kernel", st);
	link_publish_records_kernel<<<1, 32, 0, st>>>(G, peers, parity, epoch, totals, status);
	return cudaGetLastError();
}

cudaError_t pm_link_merge(LinkGeom G, unsigned char *own, int parity, unsigned int epoch, unsigned int *lb,
	unsigned long long *obase, unsigned long long *abase, PacketTotals *merged_totals, PacketRecDev *out_recs,
	unsigned long long rec_cap, uint8_t *out_arena, unsigned long long arena_cap, int *status, cudaStream_t st)
{
	pm_kt_mark("link_merge_plan_kernel", st);
	link_merge_plan_kernel<<<1, 256, 0, st>>>(G, own, parity, epoch, lb, obase, abase, merged_totals, status);
	pm_kt_mark("link_merge_write_kernel", st);
	link_merge_write_kernel<<<dim3(64, G.world), 256, 0, st>>>(G, own, parity, lb, obase, abase, out_recs, rec_cap,
		out_arena, arena_cap, status);
	return cudaGetLastError();
}

}  // extern "C"

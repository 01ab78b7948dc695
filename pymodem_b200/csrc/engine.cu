// engine.cu -- host side of libpymodem_b200.so: chain table -> launch plans,
// device memory, the run pipeline, and the C ABI of include/pymodem_b200.h.
#include <algorithm>
#include <limits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pm_common.cuh"
#include "../../include/pymodem_b200.h"

static_assert(sizeof(pm_packet_rec) == 40 && sizeof(PacketRecDev) == 40, "pm_packet_rec layout");

// ---- per-kernel timing pass (option "kernel_times") --------------------------------------------------
struct KernelTimer {
	struct Mark { const char *name; cudaEvent_t ev; };
	std::vector<Mark> marks;
	std::vector<cudaEvent_t> pool;
	size_t used = 0;
	void reset() { marks.clear(); used = 0; }
	~KernelTimer() { for (auto ev : pool) cudaEventDestroy(ev); }
};
static thread_local KernelTimer *g_kt = nullptr;      // set while a timing pass runs on this host thread

extern "C" void pm_kt_mark(const char *name, cudaStream_t st)
{
	KernelTimer *t = g_kt;
	if (!t) return;
	if (t->used == t->pool.size()) {
		cudaEvent_t ev;
		if (cudaEventCreate(&ev) != cudaSuccess) return;
		t->pool.push_back(ev);
	}
	cudaEvent_t ev = t->pool[t->used++];
	cudaEventRecord(ev, st);
	t->marks.push_back({name, ev});
}

extern "C" {
cudaError_t pm_launch_afsk_front(const AfskPlan *, size_t, const int16_t *, long long, long long, int, uint32_t *,
	long long, float *, long long, GuardList, MagOut, cudaStream_t);
cudaError_t pm_launch_lpf_tc(const LpfTcPlan *, const unsigned char *, long long, const unsigned char *, const float *,
	long long, long long, uint32_t *, long long, float *, long long, GuardList, int *, int, cudaStream_t);
cudaError_t pm_launch_fir_front(const FirPlan *, size_t, const int16_t *, long long, long long, int, uint32_t *,
	long long, float *, long long, GuardList, cudaStream_t);
cudaError_t pm_launch_guard_fixup(const Fp64Chain *, int, const int16_t *, long long, uint32_t *, long long,
	float *, long long, GuardList, int, cudaStream_t);
cudaError_t pm_launch_ffma_peak(float *, int, int, cudaStream_t);
cudaError_t pm_launch_guard_snapshot(const unsigned int *, unsigned int *, cudaStream_t);
cudaError_t pm_launch_slicer_segments(const SlicerChain *, int, const uint32_t *, long long, uint32_t *, long long,
	SegState *, SegState *, SegState *, const SegState *, SlicerGeom, cudaStream_t);
cudaError_t pm_launch_slicer_verify(const SlicerChain *, int, const uint32_t *, long long, uint32_t *, long long,
	SegState *, const SegState *, SegState *, SegState *, const SegState *, SlicerGeom, unsigned int *,
	const unsigned int *, unsigned int *, cudaStream_t);
cudaError_t pm_launch_slicer_sweep(const SlicerChain *, int, const uint32_t *, long long, uint32_t *, long long,
	SegState *, SegState *, SegState *, const SegState *, SlicerGeom, unsigned int *, cudaStream_t);
cudaError_t pm_launch_slicer_count(const SlicerChain *, int, const uint32_t *, long long, long long, long long,
	unsigned long long *, cudaStream_t);
cudaError_t pm_launch_gather(const BitChain *, int, ChainCounters *, const uint32_t *, long long, const uint32_t *,
	long long, long long, long long, unsigned int *, unsigned int *, unsigned int *, uint32_t *, long long, uint32_t *,
	long long, const unsigned int *, const ShardBits *, cudaStream_t);
cudaError_t pm_launch_tail_extract(const uint32_t *, long long, const ShardBits *, int, int, uint32_t *, cudaStream_t);
cudaError_t pm_launch_tail_inject(uint32_t *, long long, const ShardBits *, int, int, const uint32_t *, cudaStream_t);
cudaError_t pm_launch_lfsr(const BitChain *, int, const ChainCounters *, const uint32_t *, uint32_t *, long long,
	cudaStream_t);
cudaError_t pm_launch_ax25(const BitChain *, int, ChainCounters *, const uint32_t *, long long, unsigned int *,
	unsigned int *, unsigned int *, unsigned int *, long long, const uint32_t *, long long, uint8_t *, long long,
	GapRec *, long long, const ShardBits *, int, unsigned int *, unsigned int *, cudaStream_t);
cudaError_t pm_launch_p64(const P64Chain *, const P64Chain *, int, const int16_t *, uint32_t *, long long, float *,
	long long, unsigned long long *, cudaStream_t);
cudaError_t pm_link_preload(void);
cudaError_t pm_link_push_states(LinkGeom, LinkPeers, int, unsigned int, const SegState *, const SegState *, int, int,
	const unsigned long long *, const unsigned int *, int, unsigned int, cudaStream_t);
cudaError_t pm_link_wait_states(LinkGeom, unsigned char *, int, unsigned int, const BitChain *, int, int, int, ShardBits *,
	int *, long long *, cudaStream_t);
cudaError_t pm_link_il2p_in(LinkGeom, unsigned char *, int, unsigned int, int, const BitChain *, const ShardBits *, const long long *,
	Il2pHand *, ChainCounters *, int *, cudaStream_t);
cudaError_t pm_link_il2p_out(LinkGeom, LinkPeers, int, unsigned int, const Il2pHand *, const long long *, cudaStream_t);
cudaError_t pm_link_set_flag(unsigned int *, unsigned int, cudaStream_t);
cudaError_t pm_link_wait_flag(const unsigned int *, unsigned int, int *, cudaStream_t);
cudaError_t pm_link_push_records(LinkGeom, LinkPeers, int, unsigned int, const PacketRecDev *, const uint8_t *,
	const PacketTotals *, const ChainCounters *, int *, cudaStream_t);
cudaError_t pm_link_merge(LinkGeom, unsigned char *, int, unsigned int, unsigned int *, unsigned long long *,
	unsigned long long *, PacketTotals *, PacketRecDev *, unsigned long long, uint8_t *, unsigned long long, int *,
	cudaStream_t);
cudaError_t pm_il2p_init_tables(void);
cudaError_t pm_launch_il2p(const BitChain *, int, ChainCounters *, const uint32_t *, long long, const unsigned int *,
	long long, const unsigned int *, int, unsigned char *, long long, Il2pRes *, const uint32_t *, long long,
	uint8_t *, long long, GapRec *, long long, const ShardBits *, const Il2pHand *, Il2pHand *, cudaStream_t);
cudaError_t pm_launch_packets(int, ChainCounters *, const GapRec *, long long,
	pm_packet_rec *, unsigned int *, unsigned long long, PacketTotals *, const uint8_t *, long long, uint8_t *,
	unsigned long long, long long, cudaStream_t);
cudaError_t pm_launch_stream_export(const ChainCounters *, int, const uint32_t *, long long, const uint32_t *,
	long long, uint8_t *, long long *, long long, cudaStream_t);
}

// ---------------------------------------------------------------------------
template <typename T>
struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	cudaError_t ensure(size_t want)
	{
		if (want <= n) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; n = 0;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e == cudaSuccess) n = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// pinned host buffer that only grows (results come back with plain async copies, no staging through pageable memory)
template <typename T>
struct HostBuf {
	T *p = nullptr;
	size_t cap = 0, n = 0;
	cudaError_t resize(size_t want)
	{
		if (want > cap) {
			if (p) cudaFreeHost(p);
			p = nullptr; cap = 0; n = 0;
			const size_t c = want + want / 4 + 1024;
			cudaError_t e = cudaHostAlloc((void **)&p, c * sizeof(T), cudaHostAllocDefault);
			if (e != cudaSuccess) return e;
			cap = c;
		}
		n = want;
		return cudaSuccess;
	}
	size_t size() const { return n; }
	bool empty() const { return n == 0; }
	T *data() { return p; }
	const T *data() const { return p; }
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = n = 0; }
};

struct HostChain {
	pm_chain_desc d;
	std::vector<double> bpf, mark_i, mark_q, space_i, space_q, space_ui, space_uq, lpf;
	int trim = 0;             // samples lost to 'valid' convolutions
	int group = -1;           // front-end launch group (-1: float64 pipeline)
	bool p64 = false;         // runs through csrc/loops.cu
	pm_loop_desc loop;
	std::vector<double> wavetable, hilbert;
	std::vector<int32_t> pd_table;
	int sign_q_row = 0;
};

struct FrontGroup {
	int kind = 0;             // PM_MODEM_AFSK / PM_MODEM_FSK
	std::vector<int> chains;  // engine chain ids
	AfskPlan afsk;
	FirPlan fir;
	size_t smem = 0;
	int tile = 0;
	int trim_max = 0;
	double macs_per_sample = 0;   // executed FP32 MACs per input sample (after sharing)
	// tensor-core low-pass (csrc/lpf_tc.cu): the front kernel stops at the magnitudes, lpf_tc_kernel does the rest
	bool tensor = false;
	LpfTcPlan tc;
	std::vector<unsigned char> btaps;   // the banded-Toeplitz tap matrix, three bf16 pieces, operand layout
	long long n_tile_a = 0, n_tile_b = 0, a_done = 0, b_done = 0;
	long long mag_rows = 0, mag_off = 0, amax_off = 0;     // this group's part of d_mag (bytes) / d_amax (floats)
	double tc_macs_per_sample = 0;      // executed bf16 multiply-adds per input sample on the tensor cores
	double lpf_macs_per_sample = 0;     // the FP32 multiply-adds they replace
};

struct pm_engine {
	int device = 0;
	cudaStream_t st = nullptr, st_copy = nullptr;
	cudaStream_t st_front[2] = {nullptr, nullptr};   // chunk launches of the host-buffer path alternate between these, so
	                                                 // that the last partial wave of one launch overlaps the next launch
	cudaEvent_t ev_front[2] = {nullptr, nullptr};
	// host-buffer runs start the tail while the copy is still going: guard fix-up (st_tail[0]) and slicer segments
	// (st_tail[1], st_tail[2] in turn) of the part of the recording whose sign words are final
	cudaStream_t st_tail[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t ev_tail[3] = {nullptr, nullptr, nullptr};
	std::vector<cudaEvent_t> ev_steps;               // front end of a chunk done / its fix-up done
	DevBuf<unsigned int> d_repair_list;              // segments a verify pass found to re-run (slicer_repair_kernel)
	DevBuf<unsigned int> d_snap;                     // guard count after the front-end launches of chunk i: d_snap[i + 1]
	int opt_early_tail = 1;                          // host-buffer runs: 1 guard fix-up chunk by chunk beside the copy, 2 slicer segments too, 0 neither
	int opt_trace = 0;                               // option "trace": timing events at the steps of a host-buffer run (pm_engine_trace)
	std::vector<cudaEvent_t> ev_trace;
	std::vector<std::string> trace_label;
	size_t trace_used = 0;
	int opt_early_batches = 12;                      // segments are launched in about this many batches
	bool early_fix = false, early_seg = false;       // this run's fix-up / segments were launched chunk by chunk
	std::string err;
	std::vector<HostChain> chains;
	std::vector<FrontGroup> groups;
	// options
	int opt_seg_words = 768;      // 24576 samples  (sweeps: profiles/r01_slicer_sweep.txt, tools/slicer_sweep2.py)
	int opt_warm_words = 1536;    // 49152 samples, of which the last 4096 sample by sample (profiles/r02aa_slicer_sweep3.txt)
	int opt_chk_words = 32;       // checkpoint every 1024 samples
	int opt_verify_passes = 6;    // parallel verify passes before the sequential sweep
	int opt_warm_exact_words = 128; // exact sample-by-sample tail of a warm-up (4096 samples; 16384 before the float64 far part); the part before it runs crossing by crossing (0: all exact)
	int opt_quiet_skip = 1;       // repairs copy the exactly repeating clock/mask of stretches without zero crossings (SlicerChain::quiet_words)
	int opt_warm_far_f64 = 1;     // the crossing-by-crossing part in float64 (0: FP32, round 1's form, needs a 16384-sample tail)
	double opt_guard_eps = 3.814697265625e-06;  // 2^-18 of the in-band magnitude scale: 4x the largest error seen (tools/guard_sweep.py, guard_bound.py)
	double opt_guard_abs = 0.25;                // c_abs of the raw-input term: 4x the largest error seen in units of
	                                            // 2^-24 max|audio| sum|h_bpf| N_corr sum|h_lpf| (1 + g) (tools/guard_bound.py)
	int opt_tile = 0;             // 0 = auto
	int opt_keep_soft = 0;
	int opt_fuse_pairs = 1;       // mark and space sliding windows of a pair in one pass (0: tone by tone)
	int opt_slide = 1;            // rotation-tap correlators as sliding window sums (0: always the direct FIR)
	int opt_slicer_fast = 1;      // shortened slicer clock update where it is exact (0: always the plain form)
	int opt_tensor_lpf = 1;       // low-pass of the AFSK front end on the tensor cores where the group qualifies
	int sm_count = 148;
	int opt_precise = 0;          // all AFSK chains through the float64 pipeline
	double opt_precise_ratio = 0.2;
	long long opt_h2d_chunk = 8 << 20;
	unsigned int guard_cap = 1u << 20;
	// device tables
	DevBuf<double> d_taps64;
	DevBuf<Fp64Chain> d_fp64;
	DevBuf<SlicerChain> d_slicer;
	DevBuf<BitChain> d_bitchain;
	DevBuf<ChainCounters> d_cc;
	DevBuf<SegState> d_init;
	// run buffers
	DevBuf<int16_t> d_audio;
	DevBuf<uint32_t> d_sign, d_mask, d_bits_raw, d_bits_lfsr, d_byte_addr;
	DevBuf<float> d_soft;
	DevBuf<unsigned long long> d_guard_entries;
	DevBuf<unsigned int> d_counters;      // [0] guard count, [1] repairs
	DevBuf<unsigned long long> d_stage_clk;   // option "stage_clocks": GuardList::stage_clk
	DevBuf<unsigned char> d_mag;          // tone magnitudes as bf16 pieces (MagOut), tensor-core low-pass only
	DevBuf<float> d_amax;
	std::vector<DevBuf<unsigned char>> d_btaps;   // per front group
	int opt_stage_clocks = 0;
	DevBuf<SegState> d_S, d_E0, d_E1, d_chk;
	DevBuf<ShardBits> d_shardbits;
	DevBuf<unsigned long long> d_symcount;
	DevBuf<uint32_t> d_tail;
	SegState *E_cur = nullptr, *E_alt = nullptr;
	SlicerGeom geom;
	pm_shard_plan plan;
	bool sharded = false;
	int phase = 0;                        // 0 idle, 1 begun (slicer converged locally), 2 gathered
	const int16_t *run_audio = nullptr;   // device pointer of the current run's audio
	long long own_w0 = 0, own_w1 = 0, end_w = 0;   // own range / processed range in words
	int k0 = 0;                           // first own segment (= plan.pre_segments)
	int k_end = 0;                        // segment after the one whose end state is the shard's end state
	std::vector<pm_shard_state> shard_out;
	std::vector<SegState> h_init;
	DevBuf<unsigned int> d_blk_count, d_blk_base, d_sym_totals, d_flag_totals, d_flag_pos, d_rec_src;
	DevBuf<uint8_t> d_scratch, d_arena;
	DevBuf<GapRec> d_gaps;
	DevBuf<unsigned int> d_gap_cand, d_gap_ncand;   // gaps that can emit (filtered by bit count), per chain
	DevBuf<pm_packet_rec> d_recs;
	DevBuf<PacketTotals> d_totals;
	DevBuf<unsigned char> d_il2p_slots;   // speculative IL2P decodes: (cand_cap + 1) slots per chain
	DevBuf<Il2pRes> d_il2p_res;
	DevBuf<Il2pHand> d_il2p_hand;         // [0, nc): where each chain's IL2P walk starts; [nc, 2 nc): where it stands at the end
	std::vector<long long> h_A0;          // global index of local stream bit 0, per chain (sharded runs)
	std::vector<long long> h_valid_from;  // first local stream bit whose descrambled value is right, per chain
	DevBuf<double> d_p64_work, d_p64_tabs;
	DevBuf<int> d_p64_pd;
	DevBuf<P64Chain> d_p64;
	DevBuf<unsigned long long> d_p64_max;
	std::vector<P64Chain> h_p64;
	std::vector<size_t> p64_tab_off;      // per p64 chain: offsets into d_p64_tabs (wavetable, hilbert) and d_p64_pd
	// shard link (csrc/link.cu)
	bool link_on = false;
	LinkGeom lg;
	LinkPeers lp;
	DevBuf<unsigned char> d_link;
	std::vector<void *> link_opened;      // peer mappings opened with cudaIpcOpenMemHandle
	unsigned int link_epoch = 0;
	int link_tail_bits = 0;
	DevBuf<int> d_link_status;
	DevBuf<long long> d_link_A0;          // global index of local stream bit 0, per chain (written by the link's wait kernel)
	DevBuf<unsigned int> d_link_lb;
	DevBuf<unsigned long long> d_link_obase;
	DevBuf<pm_packet_rec> d_mrecs;
	DevBuf<uint8_t> d_marena;
	DevBuf<PacketTotals> d_mtotals;
	int *h_link_status = nullptr;         // pinned
	PacketTotals *h_mtotals = nullptr;    // pinned
	int il2p_cand_cap = 0;
	double rec_scale = 1.0, il2p_cand_scale = 1.0;   // grown (and the run repeated) when packet buffers / IL2P candidate lists overflow
	int grow_hint = 0;                    // what the last PM_ERR_CAPACITY asked for: 1 packet buffers, 2 IL2P candidates
	bool skip_lfsr = false;               // pm_engine_decode_stream: the stream loaded is already descrambled
	// batched runs (pm_engine_run_batch): chain c decodes recording chains[c].d.recording = row of the audio buffer
	long long batch_stride = 0;           // samples per row; 0: a single recording
	std::vector<long long> batch_n;       // valid samples of every recording
	std::vector<Fp64Chain> h_fp64, up_fp64;
	int fast_passes = 0;                  // verify passes enqueued without a host round trip (slicer_enqueue_fast)
	bool fast_pending = false;            // ... whose counters have not been looked at yet
	size_t spec_recs = 4096, spec_arena = 1 << 18;   // records / packet bytes copied back before their count is known
	std::vector<ShardBits> h_sb;
	std::vector<SlicerChain> up_sl;       // what the device tables hold (prepare_run uploads only what changed)
	std::vector<BitChain> up_bc;
	std::vector<SegState> up_init;
	std::vector<P64Chain> up_p64;
	std::vector<ShardBits> up_sb;
	int opt_tc_debug = 0;
	int opt_debug_sync = 0;               // synchronise after every front-end launch and name the kernel that failed
	int opt_kernel_times = 0;             // record an event before every kernel of a run (pm_engine_kernel_times)
	KernelTimer kt;
	std::string kt_report;
	bool has_il2p = false, il2p_tables = false;
	unsigned int *h_counters = nullptr;   // pinned
	PacketTotals *h_totals = nullptr;     // pinned
	// geometry of the last run
	long long n_samples = 0, sign_stride = 0, bits_stride = 0, addr_stride = 0, flag_stride = 0,
	          scratch_stride = 0, soft_stride = 0;
	int n_seg = 0;
	int sign_rows = 0;
	long long sample_base = 0;
	std::vector<ChainCounters> h_cc;
	HostBuf<pm_packet_rec> h_recs;
	HostBuf<uint8_t> h_arena;
	pm_stats stats;
	cudaEvent_t ev[8] = {};
	std::vector<cudaEvent_t> ev_chunks;
	// pageable input: chunks are staged through a ring of pinned buffers (memcpy on host threads, then async H2D)
	static const int RING = 3;
	int16_t *h_ring[RING] = {nullptr, nullptr, nullptr};
	size_t ring_samples = 0;
	cudaEvent_t ev_ring[RING] = {nullptr, nullptr, nullptr};
	int opt_copy_threads = 4;
	int64_t staged_bytes = 0;             // bytes of the last run that went through the ring (0: the caller's buffer was pinned)
	bool have_run = false;
};

// audio of chain c: length and offset in the run's audio buffer (a single recording unless the run is batched)
static inline long long chain_n(const pm_engine *e, int c, long long n)
{
	return e->batch_stride ? e->batch_n[e->chains[c].d.recording] : n;
}
static inline long long chain_aoff(const pm_engine *e, int c)
{
	return e->batch_stride ? (long long)e->chains[c].d.recording * e->batch_stride : 0;
}

static int fail(pm_engine *e, int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (e) e->err = buf;
	return code;
}

#define CK(call)                                                                                        \
	do {                                                                                                \
		cudaError_t _e = (call);                                                                        \
		if (_e != cudaSuccess)                                                                          \
			return fail(e, PM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
	} while (0)

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <typename T>
static bool same_bytes(const std::vector<T> &a, const std::vector<T> &b)
{
	return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0);
}

static bool same_taps(const std::vector<double> &a, const std::vector<double> &b)
{
	return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
}

// reversed float taps, zero padded to a multiple of 4
static int push_taps(float *dst, int &used, const std::vector<double> &h, int &n_padded)
{
	n_padded = round_up((int)h.size(), 4);
	if (used + n_padded > PM_MAX_TAPS) return -1;
	const int off = used;
	const int M = (int)h.size();
	for (int j = 0; j < n_padded; j++) dst[off + j] = (j < M) ? (float)h[M - 1 - j] : 0.f;
	used += n_padded;
	return off;
}

// True when (ti[k], tq[k]) = a e^{i(phi + w k)}: a rectangular-window tone correlator (afsk.py:134-144).
static bool rotation_taps(const std::vector<double> &ti, const std::vector<double> &tq, double &w)
{
	const int N = (int)ti.size();
	if (N < 16 || (int)tq.size() != N) return false;
	const double a = std::hypot(ti[0], tq[0]);
	if (!(a > 0)) return false;
	const double phi = std::atan2(tq[0], ti[0]);
	// step angle from the first pair, refined over the whole window (the phase advance stays below pi per tap)
	w = std::atan2(tq[1] * ti[0] - ti[1] * tq[0], ti[1] * ti[0] + tq[1] * tq[0]);
	const double turns = std::round((w * (N - 1) + phi - std::atan2(tq[N - 1], ti[N - 1])) / (2 * M_PI));
	w = (std::atan2(tq[N - 1], ti[N - 1]) - phi + 2 * M_PI * turns) / (N - 1);
	for (int k = 0; k < N; k++)
		if (std::hypot(ti[k] - a * std::cos(phi + w * k), tq[k] - a * std::sin(phi + w * k)) > 1e-10 * a) return false;
	return true;
}

// ---- AFSK group geometry for a candidate tile ---------------------------------
struct AfskGeom {
	int U_x, U_m, U_l, a_len;
	int s_x1_off, s_m_off, s_m_stride;
	size_t smem;
	double cost;
};

static AfskGeom afsk_geom(const AfskPlan &p, int tile)
{
	AfskGeom g;
	int cmax = 0;
	double corr_sum = 0;
	for (int j = 0; j < p.n_mag; j++) { cmax = std::max(cmax, p.mag_n[j]); corr_sum += p.mag_n[j]; }
	auto warps = [](int units) { return (units + 31) / 32 * 32; };
	if (p.tensor_lpf) {
		// the kernel stops at the magnitudes (low-pass on the tensor cores): no low-pass halo, no pair streams
		g.U_l = 0;
		g.U_m = tile / 16;
		g.U_x = round_up((16 * g.U_m + cmax + 15) / 16, 2);      // two half tiles side by side in the band-pass
		g.a_len = round_up(16 * g.U_x + p.n_bpf + 16, 8);
		// the staged audio: (a[i], a[i + 8 U_x]) pairs for i < 8 U_x + n_bpf + 16
		const int a_phys = round_up(2 * pm_phys2(8 * g.U_x + p.n_bpf + 16) + 8, 4);
		const int x_phys = round_up(2 * pm_phys2(16 * g.U_x) + 8, 4);
		g.s_m_stride = 0;
		g.s_m_off = 0;
		g.s_x1_off = a_phys;
		g.smem = sizeof(float) * (size_t)(g.s_x1_off + x_phys);
		g.cost = ((double)warps(g.U_x) * p.n_bpf + (double)warps(g.U_m) * 1.15 * corr_sum) / tile;
		return g;
	}
	g.U_l = tile / 16;
	g.U_m = (tile + p.n_lpf + 15) / 16;
	g.U_x = (16 * g.U_m + cmax + 15) / 16;
	g.a_len = round_up(16 * g.U_x + p.n_bpf, 8);
	const int a_phys = round_up(pm_phys(g.a_len) + 4, 4);
	const int x_phys = round_up(2 * pm_phys2(16 * g.U_x) + 8, 4);   // band-passed samples stored as (x, x) pairs
	g.s_m_stride = round_up(2 * pm_phys2(16 * g.U_m) + 8, 4);      // one (mark, space) pair stream of float2
	// the staged audio is dead once the band-pass is done and the magnitudes are born after it: they share a region
	g.s_m_off = 0;
	g.s_x1_off = std::max(a_phys, p.n_pair * g.s_m_stride);
	g.smem = sizeof(float) * (size_t)(g.s_x1_off + x_phys);
	// issue-slot cost per output sample (warp granular)
	double c = (double)warps(g.U_x) * p.n_bpf;
	// magnitude units are laid out stream after stream
	// the correlator (I, Q) and the low-pass (mark, space) stages run packed FFMA2: one issue slot per tap and pair
	c += (double)warps(g.U_m) * 1.15 * corr_sum;          // tone after tone (uniform taps)
	c += (double)warps(p.n_pair * g.U_l) * 1.15 * p.n_lpf;
	g.cost = c / tile;
	return g;
}

// AFSK chains whose mark and space tones nearly coincide over the correlator window lose the soft value
// to cancellation in FP32 (|mark| - |space| is a small difference of large magnitudes): float64 pipeline.
static bool afsk_needs_fp64(const pm_engine *e, const HostChain &hc)
{
	if (e->opt_precise) return true;
	double diff = 0, ref = 0;
	for (size_t k = 0; k < hc.mark_i.size(); k++) {
		const double di = hc.mark_i[k] - hc.space_i[k], dq = hc.mark_q[k] - hc.space_q[k];
		diff += di * di + dq * dq;
		ref += hc.mark_i[k] * hc.mark_i[k] + hc.mark_q[k] * hc.mark_q[k];
	}
	return ref > 0 && std::sqrt(diff / ref) < e->opt_precise_ratio;
}

static int build_groups(pm_engine *e)
{
	e->groups.clear();
	const int nc = (int)e->chains.size();
	for (int c = 0; c < nc; c++) {
		HostChain &hc = e->chains[c];
		hc.group = -1;
		if (hc.d.modem_kind == PM_MODEM_AFSK) hc.p64 = afsk_needs_fp64(e, hc);
	}
	for (int c = 0; c < nc; c++) {
		HostChain &hc = e->chains[c];
		if (hc.group >= 0 || hc.p64 || hc.d.modem_kind == PM_MODEM_NONE) continue;
		FrontGroup g;
		g.kind = hc.d.modem_kind;
		if (g.kind != PM_MODEM_AFSK && g.kind != PM_MODEM_FSK)
			return fail(e, PM_ERR_UNSUPPORTED, "chain %d: modem kind %d not supported by this build", c, g.kind);
		for (int k = c; k < nc; k++) {
			HostChain &o = e->chains[k];
			if (o.group >= 0 || o.p64 || o.d.modem_kind != g.kind || o.d.recording != hc.d.recording) continue;
			if (!same_taps(o.bpf, hc.bpf)) continue;
			if (g.kind == PM_MODEM_AFSK && !same_taps(o.lpf, hc.lpf)) continue;
			if ((int)g.chains.size() >= PM_MAX_GCH) break;
			g.chains.push_back(k);
		}
		const int gi = (int)e->groups.size();
		if (g.kind == PM_MODEM_AFSK) {
			AfskPlan &p = g.afsk;
			memset(&p, 0, sizeof(p));
			int used = 0;
			p.bpf_off = push_taps(p.taps, used, hc.bpf, p.n_bpf);
			p.lpf_off = push_taps(p.taps, used, hc.lpf, p.n_lpf);
			if (p.bpf_off < 0 || p.lpf_off < 0) return fail(e, PM_ERR_CAPACITY, "too many FIR taps");
			// tone sets (unit gain) -> magnitude streams; chains -> (mark, space) pairs
			struct Tone { const std::vector<double> *i, *q; };
			std::vector<Tone> tones;
			auto tone_id = [&](const std::vector<double> &ti, const std::vector<double> &tq) -> int {
				for (size_t t = 0; t < tones.size(); t++)
					if (same_taps(*tones[t].i, ti) && same_taps(*tones[t].q, tq)) return (int)t;
				if ((int)tones.size() >= PM_MAX_MAG) return -1;
				tones.push_back({&ti, &tq});
				return (int)tones.size() - 1;
			};
			std::vector<int> cm, cs;
			std::vector<int> kept;
			for (int k : g.chains) {
				HostChain &o = e->chains[k];
				const int m = tone_id(o.mark_i, o.mark_q);
				const int s = tone_id(o.space_ui, o.space_uq);
				if (m < 0 || s < 0) continue;      // left for another group
				cm.push_back(m); cs.push_back(s); kept.push_back(k);
			}
			g.chains = kept;
			p.n_mag = (int)tones.size();
			for (int t = 0; t < p.n_mag; t++) {
				const std::vector<double> &ti = *tones[t].i, &tq = *tones[t].q;
				const int N = (int)ti.size(), npad = round_up(N, 4);
				p.mag_n[t] = npad;
				p.mag_slide[t] = 0;
				p.mag_iq_off[t] = p.mag_e_off[t] = 0;
				double w = 0;
				used = round_up(used, 4);
				if (e->opt_slide && rotation_taps(ti, tq, w)) {
					// rotation taps (afsk.py:134-144: cos/sin of w k, rectangular window): sliding-window path
					if (used + 2 * (N + 32) > PM_MAX_TAPS) return fail(e, PM_ERR_CAPACITY, "too many FIR taps");
					p.mag_slide[t] = N;
					p.mag_e_off[t] = used;
					const double a = std::hypot(ti[0], tq[0]);
					for (int k = 0; k < N + 16; k++) {
						p.taps[used + 2 * k] = (float)(a * std::cos(w * k));
						p.taps[used + 2 * k + 1] = (float)(a * std::sin(w * k));
					}
					used += 2 * (N + 16);
					for (int k = 0; k < 16; k++) {
						p.taps[used + 2 * k] = -p.taps[p.mag_e_off[t] + 2 * k];
						p.taps[used + 2 * k + 1] = -p.taps[p.mag_e_off[t] + 2 * k + 1];
					}
					used += 32;
				} else {
					// arbitrary taps: reversed, interleaved (i[0], q[0], i[1], q[1], ...), zero padded to a multiple of 4
					if (used + 2 * npad > PM_MAX_TAPS) return fail(e, PM_ERR_CAPACITY, "too many FIR taps");
					p.mag_iq_off[t] = used;
					for (int j = 0; j < npad; j++) {
						p.taps[used + 2 * j] = j < N ? (float)ti[N - 1 - j] : 0.f;
						p.taps[used + 2 * j + 1] = j < N ? (float)tq[N - 1 - j] : 0.f;
					}
					used += 2 * npad;
				}
			}
			// pairs, chains sorted by pair
			std::vector<std::pair<int, int>> pairs;
			std::vector<int> chain_pair(kept.size());
			for (size_t i = 0; i < kept.size(); i++) {
				std::pair<int, int> pr(cm[i], cs[i]);
				auto it = std::find(pairs.begin(), pairs.end(), pr);
				if (it == pairs.end()) { pairs.push_back(pr); chain_pair[i] = (int)pairs.size() - 1; }
				else chain_pair[i] = (int)(it - pairs.begin());
			}
			if ((int)pairs.size() > PM_MAX_PAIR) return fail(e, PM_ERR_CAPACITY, "too many tone pairs in a group");
			p.n_pair = (int)pairs.size();
			int ci = 0;
			for (int pi = 0; pi < p.n_pair; pi++) {
				p.pair_mark[pi] = pairs[pi].first;
				p.pair_space[pi] = pairs[pi].second;
				p.pair_first[pi] = ci;
				for (size_t i = 0; i < kept.size(); i++)
					if (chain_pair[i] == pi) {
						p.chain_gid[ci] = kept[i];
						p.chain_gain[ci] = (float)e->chains[kept[i]].d.space_gain;
						ci++;
					}
			}
			p.pair_first[p.n_pair] = ci;
			p.n_chain = ci;
			p.guard_eps = (float)e->opt_guard_eps;
			{
				// raw-input term of the guard (front.cu epilogue): the band-pass accumulates at the magnitude of the raw
				// samples, so its rounding error -- carried through the window sum (N_corr terms) and the low-pass
				// (sum|h_lpf|) into y = L_mark - g L_space -- scales with max|audio|, whatever the in-band level is
				double hb = 0, hl = 0;
				for (double v : hc.bpf) hb += std::fabs(v);
				for (double v : hc.lpf) hl += std::fabs(v);
				for (int i = 0; i < p.n_chain; i++) {
					const HostChain &o = e->chains[p.chain_gid[i]];
					p.chain_guard_abs[i] = (float)(e->opt_guard_abs * std::ldexp(1.0, -24) * hb * (double)o.mark_i.size() * hl *
						(1.0 + std::fabs(o.d.space_gain)));
				}
			}
			// pairs whose two tones are sliding-window tones of one length, each used by this pair only, are computed
			// together (front.cu SlidePair)
			for (int pi = 0; pi < p.n_pair; pi++) {
				const int m = p.pair_mark[pi], sp = p.pair_space[pi];
				int uses_m = 0, uses_s = 0;
				for (int q = 0; q < p.n_pair; q++) {
					uses_m += (p.pair_mark[q] == m) + (p.pair_space[q] == m);
					uses_s += (p.pair_mark[q] == sp) + (p.pair_space[q] == sp);
				}
				const bool fuse = e->opt_fuse_pairs && m != sp && p.mag_slide[m] > 0 && p.mag_slide[m] == p.mag_slide[sp] &&
					uses_m == 1 && uses_s == 1;
				p.pair_fused[pi] = fuse ? p.mag_slide[m] : 0;
				p.pair_ea[pi] = p.mag_e_off[m];
				p.pair_eb[pi] = p.mag_e_off[sp];
			}
			// where each remaining tone's magnitude goes: the (mark, space) slots of the pair streams the low-pass reads
			{
				int k = 0;
				for (int t = 0; t < p.n_mag; t++) {
					p.mag_dst_first[t] = k;
					for (int pi = 0; pi < p.n_pair; pi++) {
						if (p.pair_fused[pi]) continue;
						if (p.pair_mark[pi] == t) p.mag_dst[k++] = pi * 2;
						if (p.pair_space[pi] == t) p.mag_dst[k++] = pi * 2 + 1;
					}
				}
				p.mag_dst_first[p.n_mag] = k;
			}
			// band-pass taps twice in a row each, for the route that runs two half tiles on FFMA2
			if (used + 2 * p.n_bpf > PM_MAX_TAPS) return fail(e, PM_ERR_CAPACITY, "too many FIR taps");
			p.bpf2_off = used;
			for (int j = 0; j < p.n_bpf; j++) p.taps[used + 2 * j] = p.taps[used + 2 * j + 1] = p.taps[p.bpf_off + j];
			used += 2 * p.n_bpf;
			// low-pass taps once more, each twice in a row: the (h, h) operand of the packed FFMA2
			if (used + 2 * p.n_lpf > PM_MAX_TAPS) return fail(e, PM_ERR_CAPACITY, "too many FIR taps");
			p.lpf2_off = used;
			for (int j = 0; j < p.n_lpf; j++) p.taps[used + 2 * j] = p.taps[used + 2 * j + 1] = p.taps[p.lpf_off + j];
			used += 2 * p.n_lpf;
			// Tensor-core low-pass (csrc/lpf_tc.cu) where the group qualifies: every pair a fused sliding-window pair (the
			// magnitudes then leave the correlator stage pair by pair), at most TC_MAX_TONES tones, taps within K = 192
			{
				bool all_fused = p.n_pair > 0;
				for (int pi = 0; pi < p.n_pair; pi++) all_fused = all_fused && p.pair_fused[pi] > 0;
				g.tensor = e->opt_tensor_lpf && all_fused && p.n_mag <= TC_MAX_TONES && (int)hc.lpf.size() <= TC_MAX_LPF &&
					(int)hc.lpf.size() >= 8;
				p.tensor_lpf = g.tensor ? 1 : 0;
			}
			// tile: cheapest issue-slot cost that still fits two CTAs per SM
			int best = 0;
			double best_cost = 1e300;
			const size_t smem_2cta = 112 * 1024, smem_max = 226 * 1024;
			for (int pass = 0; pass < 2 && !best; pass++)
				for (int tile = 256; tile <= 16384; tile += 32) {
					AfskGeom gg = afsk_geom(p, tile);
					if (gg.smem > (pass == 0 ? smem_2cta : smem_max)) break;
					if (gg.cost < best_cost) { best_cost = gg.cost; best = tile; }
				}
			if (e->opt_tile > 0) best = round_up(e->opt_tile, 32);
			if (!best) return fail(e, PM_ERR_CAPACITY, "AFSK front end does not fit in shared memory");
			AfskGeom gg = afsk_geom(p, best);
			if (gg.smem > smem_max) return fail(e, PM_ERR_CAPACITY, "tile %d needs %zu B shared memory", best, gg.smem);
			p.tile = best; p.U_x = gg.U_x; p.U_m = gg.U_m; p.U_l = gg.U_l; p.a_len = gg.a_len;
			p.bpf_half = gg.U_x / 2;
			p.s_x1_off = gg.s_x1_off; p.s_m_off = gg.s_m_off; p.s_m_stride = gg.s_m_stride;
			g.smem = gg.smem;
			g.tile = best;
			double macs = hc.bpf.size() + (g.tensor ? 0.0 : 2.0 * p.n_pair * hc.lpf.size());
			for (int t = 0; t < p.n_mag; t++)
				macs += p.mag_slide[t] ? 2.0 * (p.mag_slide[t] + 30) / 16.0 : 2.0 * tones[t].i->size();
			g.macs_per_sample = macs;
			g.lpf_macs_per_sample = 2.0 * p.n_pair * hc.lpf.size();
			if (g.tensor) {
				LpfTcPlan &T = g.tc;
				memset(&T, 0, sizeof(T));
				T.n_mag = p.n_mag; T.n_pair = p.n_pair; T.n_chain = p.n_chain;
				for (int pi = 0; pi < p.n_pair; pi++) { T.pair_mark[pi] = p.pair_mark[pi]; T.pair_space[pi] = p.pair_space[pi]; }
				for (int pi = 0; pi <= p.n_pair; pi++) T.pair_first[pi] = p.pair_first[pi];
				for (int i = 0; i < p.n_chain; i++) {
					T.chain_gid[i] = p.chain_gid[i]; T.chain_gain[i] = p.chain_gain[i]; T.chain_guard_abs[i] = p.chain_guard_abs[i];
				}
				T.guard_eps = p.guard_eps;
				T.n_lpf = (int)hc.lpf.size();
				T.debug_mask = e->opt_tc_debug;
				T.tile_a = p.tile;
				int cmax = 0;
				for (int t = 0; t < p.n_mag; t++) cmax = std::max(cmax, (int)tones[t].i->size());
				T.reach = ((int)hc.bpf.size() - 1) + (cmax - 1) + ((int)hc.lpf.size() - 1);
				// B[d][n] = hr[d - n] (hr = the taps in correlation order), three bf16 pieces, per piece three K blocks of
				// 64 rows (n) x 64 columns (d within the block), K-major rows of 128 bytes, SWIZZLE_128B
				const int M = (int)hc.lpf.size();
				g.btaps.assign(3 * TC_B_BYTES, 0);
				auto bf16_rn = [](float f) -> uint16_t {
					uint32_t u;
					memcpy(&u, &f, 4);
					u = u + 0x7FFFu + ((u >> 16) & 1u);
					return (uint16_t)(u >> 16);
				};
				for (int kb = 0; kb < TC_KBLK; kb++)
					for (int nn = 0; nn < TC_N; nn++)
						for (int kk = 0; kk < 64; kk++) {
							const int t = 64 * kb + kk - nn;
							float rest = (t >= 0 && t < M) ? (float)hc.lpf[M - 1 - t] : 0.0f;
							for (int q = 0; q < 3; q++) {
								const uint16_t h16 = bf16_rn(rest);
								uint32_t back = (uint32_t)h16 << 16;
								float pq;
								memcpy(&pq, &back, 4);
								const uint32_t o = (uint32_t)nn * 128u + (uint32_t)kk * 2u;
								const uint32_t sw = o ^ (((o >> 7) & 7u) << 4);
								memcpy(&g.btaps[(size_t)q * TC_B_BYTES + (size_t)kb * TC_N * 128 + sw], &h16, 2);
								rest -= pq;                      // exact
							}
						}
				// executed on the tensor cores per input sample: six piece products x the K steps that hold taps x 64 columns
				int ksteps = 0;
				for (int k0 = 0; k0 < 64 * TC_KBLK; k0 += 16) if (k0 < M + 63) ksteps++;
				g.tc_macs_per_sample = 6.0 * ksteps * 16.0 * p.n_mag;
			}
		} else {
			FirPlan &p = g.fir;
			memset(&p, 0, sizeof(p));
			int used = 0;
			p.taps_off = push_taps(p.taps, used, hc.bpf, p.n_taps);
			if (p.taps_off < 0) return fail(e, PM_ERR_CAPACITY, "too many FIR taps");
			p.n_chain = (int)g.chains.size();
			double abs_sum = 0;
			for (double v : hc.bpf) abs_sum += std::fabs(v);
			for (int i = 0; i < p.n_chain; i++) {
				p.chain_gid[i] = g.chains[i];
				p.chain_neg[i] = e->chains[g.chains[i]].d.invert_soft;
			}
			// |FP32 FIR - exact| <= (n + 1) 2^-24 sum|h| max|x| (one rounding per tap plus the tap's own conversion): the
			// guard is never narrower than that worst case, so the single-FIR front end's signs are exact by bound
			p.guard_eps = (float)(std::max(e->opt_guard_eps, 1.02 * (hc.bpf.size() + 2) * std::ldexp(1.0, -24)) * abs_sum);
			int tile = e->opt_tile > 0 ? round_up(e->opt_tile, 32) : 4096;
			p.tile = tile;
			p.U_y = tile / 16;
			p.a_len = round_up(tile + p.n_taps + 16, 8);
			g.smem = sizeof(float) * (size_t)(pm_phys(p.a_len) + 8);
			if (g.smem > 226 * 1024) return fail(e, PM_ERR_CAPACITY, "FIR front end does not fit in shared memory");
			g.tile = tile;
			g.macs_per_sample = (double)hc.bpf.size();
		}
		g.trim_max = 0;
		for (int k : g.chains) {
			e->chains[k].group = gi;
			g.trim_max = std::max(g.trim_max, e->chains[k].trim);
		}
		e->groups.push_back(g);
	}
	for (auto &b : e->d_btaps) b.release();
	e->d_btaps.assign(e->groups.size(), DevBuf<unsigned char>());
	for (size_t gi = 0; gi < e->groups.size(); gi++) {
		FrontGroup &g = e->groups[gi];
		if (!g.tensor) continue;
		CK(e->d_btaps[gi].ensure(g.btaps.size()));
		CK(cudaMemcpy(e->d_btaps[gi].p, g.btaps.data(), g.btaps.size(), cudaMemcpyHostToDevice));
	}
	return PM_OK;
}

// ---------------------------------------------------------------------------
extern "C" const char *pm_version(void) { return "pymodem_b200 0.1 (sm_100a)"; }

extern "C" int pm_engine_create(int device, pm_engine **out)
{
	if (!out) return PM_ERR_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
		return PM_ERR_CUDA;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PM_ERR_CUDA;
	if (prop.major != 10) return PM_ERR_CUDA;          // sm_100a code only: no fallback path
	pm_engine *e = new pm_engine();
	e->device = device;
	e->sm_count = prop.multiProcessorCount;
	memset(&e->stats, 0, sizeof(e->stats));
	int prio_lo = 0, prio_hi = 0;
	if (cudaSetDevice(device) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&e->st_front[0], cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&e->st_front[1], cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreateWithFlags(&e->ev_front[0], cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&e->ev_front[1], cudaEventDisableTiming) != cudaSuccess ||
	    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess ||
	    // (highest priority: their few blocks are placed as soon as a front-end CTA retires instead of after the whole grid)
	    cudaStreamCreateWithPriority(&e->st_tail[0], cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
	    cudaStreamCreateWithPriority(&e->st_tail[1], cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
	    cudaStreamCreateWithPriority(&e->st_tail[2], cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
	    cudaEventCreateWithFlags(&e->ev_tail[0], cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&e->ev_tail[1], cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&e->ev_tail[2], cudaEventDisableTiming) != cudaSuccess ||
	    cudaHostAlloc((void **)&e->h_counters, 64, cudaHostAllocDefault) != cudaSuccess ||
	    cudaHostAlloc((void **)&e->h_totals, sizeof(PacketTotals), cudaHostAllocDefault) != cudaSuccess) {
		delete e;
		return PM_ERR_CUDA;
	}
	for (auto &ev : e->ev) cudaEventCreate(&ev);
	*out = e;
	return PM_OK;
}

extern "C" void pm_engine_destroy(pm_engine *e)
{
	if (!e) return;
	cudaSetDevice(e->device);
	cudaStreamSynchronize(e->st);
	cudaStreamSynchronize(e->st_copy);
	e->d_taps64.release(); e->d_fp64.release(); e->d_slicer.release(); e->d_bitchain.release();
	e->d_cc.release(); e->d_init.release(); e->d_audio.release(); e->d_sign.release(); e->d_mask.release();
	e->d_bits_raw.release(); e->d_bits_lfsr.release(); e->d_byte_addr.release(); e->d_soft.release();
	e->d_chk.release(); e->d_shardbits.release(); e->d_symcount.release(); e->d_tail.release();
	e->d_mag.release(); e->d_amax.release();
	for (auto &b : e->d_btaps) b.release();
	e->d_guard_entries.release(); e->d_repair_list.release(); e->d_counters.release(); e->d_stage_clk.release(); e->d_S.release(); e->d_E0.release();
	e->d_E1.release(); e->d_blk_count.release(); e->d_blk_base.release(); e->d_sym_totals.release();
	e->d_flag_totals.release(); e->d_flag_pos.release(); e->d_rec_src.release(); e->d_scratch.release();
	e->d_gap_cand.release(); e->d_gap_ncand.release();
	e->d_arena.release(); e->d_gaps.release(); e->d_recs.release(); e->d_totals.release();
	e->d_il2p_slots.release(); e->d_il2p_res.release(); e->d_il2p_hand.release();
	e->d_p64_work.release(); e->d_p64_tabs.release(); e->d_p64_pd.release(); e->d_p64.release(); e->d_p64_max.release();
	for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
	for (auto &ev : e->ev_chunks) cudaEventDestroy(ev);
	for (auto &ev : e->ev_trace) cudaEventDestroy(ev);
	for (void *p : e->link_opened) cudaIpcCloseMemHandle(p);
	e->d_link.release(); e->d_link_status.release(); e->d_link_A0.release(); e->d_link_lb.release(); e->d_link_obase.release();
	e->d_mrecs.release(); e->d_marena.release(); e->d_mtotals.release();
	if (e->h_link_status) cudaFreeHost(e->h_link_status);
	if (e->h_mtotals) cudaFreeHost(e->h_mtotals);
	e->h_recs.release(); e->h_arena.release();
	if (e->h_counters) cudaFreeHost(e->h_counters);
	if (e->h_totals) cudaFreeHost(e->h_totals);
	for (int i = 0; i < pm_engine::RING; i++) {
		if (e->h_ring[i]) cudaFreeHost(e->h_ring[i]);
		if (e->ev_ring[i]) cudaEventDestroy(e->ev_ring[i]);
	}
	cudaStreamDestroy(e->st);
	cudaStreamDestroy(e->st_copy);
	for (int i = 0; i < 2; i++) {
		if (e->st_front[i]) cudaStreamDestroy(e->st_front[i]);
		if (e->ev_front[i]) cudaEventDestroy(e->ev_front[i]);
	}
	for (int i = 0; i < 3; i++) {
		if (e->st_tail[i]) cudaStreamDestroy(e->st_tail[i]);
		if (e->ev_tail[i]) cudaEventDestroy(e->ev_tail[i]);
	}
	for (auto ev : e->ev_steps) cudaEventDestroy(ev);
	delete e;
}

extern "C" const char *pm_last_error(const pm_engine *e) { return e ? e->err.c_str() : "no engine"; }

static void copy_vec(std::vector<double> &dst, const double *src, int n)
{
	dst.clear();
	if (src && n > 0) dst.assign(src, src + n);
}

extern "C" int pm_engine_load_chains(pm_engine *e, const pm_chain_desc *descs, int32_t n)
{
	if (!e || !descs || n <= 0) return fail(e, PM_ERR_ARG, "load_chains: bad arguments");
	if (n >= 65536) return fail(e, PM_ERR_ARG, "too many chains");
	for (int c = 0; c < n; c++)
		if (descs[c].recording < 0 || descs[c].recording >= 65536) return fail(e, PM_ERR_ARG, "chain %d: recording index out of range", c);
	cudaSetDevice(e->device);
	e->chains.clear();
	e->have_run = false;
	e->up_sl.clear(); e->up_bc.clear(); e->up_init.clear(); e->up_p64.clear(); e->up_sb.clear();
	for (int c = 0; c < n; c++) {
		HostChain hc;
		hc.d = descs[c];
		const pm_chain_desc &d = descs[c];
		if (d.modem_kind == PM_MODEM_AFSK) {
			if (!d.bpf || !d.mark_i || !d.mark_q || !d.space_i || !d.space_q || !d.space_unit_i ||
			    !d.space_unit_q || !d.lpf || d.n_bpf <= 0 || d.n_corr <= 0 || d.n_lpf <= 0)
				return fail(e, PM_ERR_ARG, "chain %d: AFSK needs bpf/mark/space/lpf taps", c);
			copy_vec(hc.bpf, d.bpf, d.n_bpf);
			copy_vec(hc.mark_i, d.mark_i, d.n_corr); copy_vec(hc.mark_q, d.mark_q, d.n_corr);
			copy_vec(hc.space_i, d.space_i, d.n_corr); copy_vec(hc.space_q, d.space_q, d.n_corr);
			copy_vec(hc.space_ui, d.space_unit_i, d.n_corr); copy_vec(hc.space_uq, d.space_unit_q, d.n_corr);
			copy_vec(hc.lpf, d.lpf, d.n_lpf);
			hc.trim = (d.n_bpf - 1) + (d.n_corr - 1) + (d.n_lpf - 1);
		} else if (d.modem_kind == PM_MODEM_FSK) {
			if (!d.bpf || d.n_bpf <= 0) return fail(e, PM_ERR_ARG, "chain %d: FSK needs input filter taps", c);
			copy_vec(hc.bpf, d.bpf, d.n_bpf);
			hc.trim = d.n_bpf - 1;
		} else if (d.modem_kind == PM_MODEM_BPSK || d.modem_kind == PM_MODEM_MPSK || d.modem_kind == PM_MODEM_AFSK_PLL) {
			const pm_loop_desc *lp = d.loop;
			if (!d.bpf || d.n_bpf <= 0 || !d.lpf || d.n_lpf <= 0 || !lp || !lp->nco_wavetable || lp->nco_size <= 0)
				return fail(e, PM_ERR_ARG, "chain %d: PSK/PLL modem needs input/output taps and loop constants", c);
			if (lp->nco_size > 1024) return fail(e, PM_ERR_CAPACITY, "chain %d: NCO wavetable larger than 1024", c);
			if (d.n_bpf > P64_MAX_TAPS || d.n_lpf > P64_MAX_TAPS)
				return fail(e, PM_ERR_CAPACITY, "chain %d: more than %d taps in a float64 FIR", c, P64_MAX_TAPS);
			copy_vec(hc.bpf, d.bpf, d.n_bpf);
			copy_vec(hc.lpf, d.lpf, d.n_lpf);
			hc.loop = *lp;
			hc.wavetable.assign(lp->nco_wavetable, lp->nco_wavetable + lp->nco_size);
			hc.trim = (d.n_bpf - 1) + (d.n_lpf - 1);
			if (d.modem_kind == PM_MODEM_MPSK) {
				if (!lp->hilbert || lp->n_hilbert <= 0 || lp->n_hilbert > P64_MAX_TAPS || !lp->pd_table ||
				    lp->pd_granularity <= 0 || lp->pd_granularity > 64 || (lp->pd_granularity & (lp->pd_granularity - 1)) ||
				    lp->hilbert_delay != lp->n_hilbert / 2)
					return fail(e, PM_ERR_ARG, "chain %d: MPSK needs Hilbert taps and the phase-error table", c);
				hc.hilbert.assign(lp->hilbert, lp->hilbert + lp->n_hilbert);
				hc.pd_table.assign(lp->pd_table, lp->pd_table + lp->pd_granularity * lp->pd_granularity);
				hc.trim += lp->n_hilbert - 1;
			}
			hc.loop.nco_wavetable = nullptr; hc.loop.pd_table = nullptr; hc.loop.hilbert = nullptr;
			hc.p64 = true;
		} else if (d.modem_kind == PM_MODEM_NONE) {
			hc.trim = 0;                     // per-stage calls only (pm_engine_slice_soft / unscramble_stream / decode_stream)
		} else {
			return fail(e, PM_ERR_UNSUPPORTED, "chain %d: modem kind %d not supported by this build", c, d.modem_kind);
		}
		if (d.slicer_kind != PM_SLICER_BINARY && d.slicer_kind != PM_SLICER_QUADRATURE)
			return fail(e, PM_ERR_UNSUPPORTED, "chain %d: slicer kind %d not supported by this build", c, d.slicer_kind);
		// the reference's duck typing: MPSKModem.demod returns IQData (psk.py:748), which only
		// QuadratureSlicer.slice can iterate (slicer.py:198); every other modem returns an ndarray
		if (d.modem_kind != PM_MODEM_NONE && (d.slicer_kind == PM_SLICER_QUADRATURE) != (d.modem_kind == PM_MODEM_MPSK))
			return fail(e, PM_ERR_ARG, "chain %d: the quadrature slicer goes with the mpsk modem (and only with it)", c);
		if (d.slicer_kind == PM_SLICER_QUADRATURE &&
		    (d.bits_per_symbol < 1 || d.bits_per_symbol > 2 || d.state_mask > 15 || 8 % d.bits_per_symbol))
			return fail(e, PM_ERR_ARG, "chain %d: quadrature slicer needs 1 or 2 bits per symbol and a 4-bit state", c);
		if (d.codec_kind != PM_CODEC_AX25 && d.codec_kind != PM_CODEC_IL2P)
			return fail(e, PM_ERR_UNSUPPORTED, "chain %d: codec kind %d not supported by this build", c, d.codec_kind);
		if (d.codec_kind == PM_CODEC_IL2P && (d.il2p_sync_tol < 0 || d.il2p_sync_tol > 8))
			return fail(e, PM_ERR_ARG, "chain %d: il2p sync_tol must be 0..8", c);
		if (!(d.symbol_rate > 0) || !(d.slicer_sample_rate / d.symbol_rate >= 3.0))
			return fail(e, PM_ERR_ARG, "chain %d: needs >= 3 samples per symbol", c);
		// pointers in the copy are not valid after this call
		hc.d.bpf = hc.d.mark_i = hc.d.mark_q = hc.d.space_i = hc.d.space_q = nullptr;
		hc.d.space_unit_i = hc.d.space_unit_q = hc.d.lpf = nullptr;
		hc.d.loop = nullptr;
		e->chains.push_back(std::move(hc));
	}
	e->has_il2p = false;
	for (auto &hc : e->chains) if (hc.d.codec_kind == PM_CODEC_IL2P) e->has_il2p = true;
	if (e->has_il2p && !e->il2p_tables) {
		CK(pm_il2p_init_tables());
		e->il2p_tables = true;
	}
	int rc = build_groups(e);
	if (rc != PM_OK) return rc;

	// FP64 taps (reversed = correlation order) for the guard-band fix-up and the float64 pipeline
	std::vector<double> flat;
	std::vector<Fp64Chain> f64(n);
	memset(f64.data(), 0, n * sizeof(Fp64Chain));
	auto push = [&](const std::vector<double> &h, bool reverse = true) {
		size_t off = flat.size();
		for (size_t j = 0; j < h.size(); j++) flat.push_back(reverse ? h[h.size() - 1 - j] : h[j]);
		return off;
	};
	struct Offs { size_t bpf, mi, mq, si, sq, lpf, hil, wt; };
	std::vector<Offs> o(n);
	std::vector<int> pd_flat;
	std::vector<size_t> pd_off(n, 0);
	for (int c = 0; c < n; c++) {
		HostChain &hc = e->chains[c];
		o[c].bpf = push(hc.bpf); o[c].mi = push(hc.mark_i); o[c].mq = push(hc.mark_q);
		o[c].si = push(hc.space_i); o[c].sq = push(hc.space_q); o[c].lpf = push(hc.lpf);
		o[c].hil = push(hc.hilbert); o[c].wt = push(hc.wavetable, false);
		pd_off[c] = pd_flat.size();
		pd_flat.insert(pd_flat.end(), hc.pd_table.begin(), hc.pd_table.end());
	}
	CK(e->d_taps64.ensure(flat.size() + 1));
	CK(cudaMemcpy(e->d_taps64.p, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
	CK(e->d_p64_pd.ensure(pd_flat.size() + 1));
	if (!pd_flat.empty())
		CK(cudaMemcpy(e->d_p64_pd.p, pd_flat.data(), pd_flat.size() * sizeof(int), cudaMemcpyHostToDevice));
	for (int c = 0; c < n; c++) {
		HostChain &hc = e->chains[c];
		Fp64Chain &f = f64[c];
		f.kind = hc.d.modem_kind;
		f.n_bpf = (int)hc.bpf.size(); f.n_corr = (int)hc.mark_i.size(); f.n_lpf = (int)hc.lpf.size();
		f.neg = hc.d.invert_soft;
		f.bpf = e->d_taps64.p + o[c].bpf;
		f.mark_i = e->d_taps64.p + o[c].mi; f.mark_q = e->d_taps64.p + o[c].mq;
		f.space_i = e->d_taps64.p + o[c].si; f.space_q = e->d_taps64.p + o[c].sq;
		f.lpf = e->d_taps64.p + o[c].lpf;
	}
	CK(e->d_fp64.ensure(n));
	e->h_fp64 = f64;                    // audio_off / n_audio are filled in per run (prepare_run)
	e->up_fp64.clear();
	// float64 pipeline table (buffers and lengths are filled in per run)
	e->h_p64.clear();
	for (int c = 0; c < n; c++) {
		HostChain &hc = e->chains[c];
		if (!hc.p64) continue;
		if (hc.bpf.size() > P64_MAX_TAPS || hc.lpf.size() > P64_MAX_TAPS || hc.mark_i.size() > P64_MAX_TAPS)
			return fail(e, PM_ERR_CAPACITY, "chain %d: more than %d taps in a float64 FIR", c, P64_MAX_TAPS);
		P64Chain P;
		memset(&P, 0, sizeof(P));
		P.kind = hc.d.modem_kind;
		P.gid = c;
		P.n_bpf = (int)hc.bpf.size();
		P.n_out = (int)hc.lpf.size();
		P.bpf = e->d_taps64.p + o[c].bpf;
		P.out_taps = e->d_taps64.p + o[c].lpf;
		if (P.kind == PM_MODEM_AFSK) {
			P.n_mid = (int)hc.mark_i.size();
			P.mid0 = e->d_taps64.p + o[c].mi; P.mid1 = e->d_taps64.p + o[c].mq;
			P.mid2 = e->d_taps64.p + o[c].si; P.mid3 = e->d_taps64.p + o[c].sq;
		} else {
			const pm_loop_desc &l = hc.loop;
			LoopConst &L = P.lc;
			L.agc_scaled_attack = l.agc_scaled_attack; L.agc_scaled_decay = l.agc_scaled_decay;
			L.agc_sustain_time = l.agc_sustain_time; L.agc_sustain_increment = l.agc_sustain_increment;
			L.agc_target = l.agc_target;
			L.nco_phase_scale = l.nco_phase_scale; L.nco_index_scale = l.nco_index_scale;
			L.nco_set_frequency = l.nco_set_frequency; L.nco_two_pi = l.nco_two_pi; L.nco_quarter = l.nco_quarter;
			L.iir_b0 = l.iir_b0; L.iir_b1 = l.iir_b1; L.iir_a1 = l.iir_a1;
			L.pi_gain = l.pi_gain; L.pi_p = l.pi_p; L.pi_i = l.pi_i; L.pi_limit = l.pi_limit;
			L.pi_integral0 = l.pi_integral0;
			P.wavetable = e->d_taps64.p + o[c].wt;
			P.wt_size = (int)hc.wavetable.size();
			if (P.kind == PM_MODEM_MPSK) {
				P.n_mid = (int)hc.hilbert.size();
				P.mid_delay = l.hilbert_delay;
				P.mid0 = e->d_taps64.p + o[c].hil;
				P.pd_table = e->d_p64_pd.p + pd_off[c];
				P.pd_g = (int)l.pd_granularity;
			}
		}
		e->h_p64.push_back(P);
	}
	CK(e->d_p64.ensure(e->h_p64.size() + 1));
	CK(e->d_p64_max.ensure(e->h_p64.size() + 1));
	CK(e->d_slicer.ensure(n));
	CK(e->d_bitchain.ensure(n));
	CK(e->d_cc.ensure(n));
	CK(e->d_init.ensure(n));
	CK(e->d_sym_totals.ensure(n));
	CK(e->d_flag_totals.ensure(n));
	CK(e->d_counters.ensure(16));
	CK(e->d_totals.ensure(1));
	return PM_OK;
}

extern "C" int pm_engine_set_option(pm_engine *e, const char *key, double value)
{
	if (!e || !key) return PM_ERR_ARG;
	std::string k(key);
	bool replan = false;
	if (k == "segment_len") e->opt_seg_words = std::max(1, (int)(value / 32));
	else if (k == "warmup_len") e->opt_warm_words = std::max(0, (int)((value + 31) / 32));
	else if (k == "checkpoint_len") e->opt_chk_words = std::max(1, (int)(value / 32));
	else if (k == "verify_passes") e->opt_verify_passes = std::max(0, (int)value);
	else if (k == "warmup_exact_len") e->opt_warm_exact_words = std::max(0, (int)((value + 31) / 32));
	else if (k == "warmup_far_f64") e->opt_warm_far_f64 = value != 0;
	else if (k == "quiet_skip") e->opt_quiet_skip = value != 0;
	else if (k == "guard_eps") { e->opt_guard_eps = value; replan = true; }
	else if (k == "guard_abs") { e->opt_guard_abs = value; replan = true; }
	else if (k == "tile") { e->opt_tile = (int)value; replan = true; }
	else if (k == "keep_soft") e->opt_keep_soft = value != 0;
	else if (k == "slicer_fast") e->opt_slicer_fast = value != 0;
	else if (k == "slide_correlator") { e->opt_slide = value != 0; replan = true; }
	else if (k == "fuse_pairs") { e->opt_fuse_pairs = value != 0; replan = true; }
	else if (k == "tensor_lpf") { e->opt_tensor_lpf = value != 0; replan = true; }
	else if (k == "precise") {
		if (!e->chains.empty()) return fail(e, PM_ERR_STATE, "set 'precise' before loading chains");
		e->opt_precise = value != 0;
	}
	else if (k == "stage_clocks") {
		e->opt_stage_clocks = value != 0;
		if (e->opt_stage_clocks) {
			cudaSetDevice(e->device);
			CK(e->d_stage_clk.ensure(8));
			CK(cudaMemset(e->d_stage_clk.p, 0, 8 * sizeof(unsigned long long)));
		}
	}
	else if (k == "early_tail") e->opt_early_tail = std::max(0, std::min(2, (int)value));
	else if (k == "trace") e->opt_trace = (int)value;
	else if (k == "early_batches") e->opt_early_batches = std::max(1, (int)value);
	else if (k == "h2d_chunk") e->opt_h2d_chunk = std::max<long long>(1 << 16, (long long)value);
	else if (k == "kernel_times") e->opt_kernel_times = value != 0;
	else if (k == "debug_sync") e->opt_debug_sync = value != 0;
	else if (k == "tc_debug") { e->opt_tc_debug = (int)value; replan = true; }
	else if (k == "tc_grid") e->sm_count = std::max(1, (int)value);          // CTAs of the persistent low-pass kernel (default: one per SM)
	else if (k == "copy_threads") e->opt_copy_threads = std::max(1, std::min(16, (int)value));
	else if (k == "guard_cap") e->guard_cap = (unsigned int)std::max(1024.0, value);
	else return fail(e, PM_ERR_ARG, "unknown option '%s'", key);
	if (replan && !e->chains.empty()) return build_groups(e);
	return PM_OK;
}

// ---------------------------------------------------------------------------
// One run = begin (front end + slicer) -> [handoff]* -> gather -> finish.
// An unsharded run is the same pipeline with "everything is mine".
// ---------------------------------------------------------------------------
// only_chain >= 0: a per-stage call (pm_engine_slice_soft / unscramble_stream / decode_stream) works on one chain, the
// others see an empty recording; min_bits: stream bits the bit-level buffers have to hold whatever n says
static int prepare_run(pm_engine *e, long long n, const pm_shard_plan &plan, bool sharded, int only_chain = -1,
                       long long min_bits = 0)
{
	const int nc = (int)e->chains.size();
	if (nc == 0) return fail(e, PM_ERR_STATE, "no chains loaded");
	if (n <= 0 || n >= (1ll << 32) - (1 << 20)) return fail(e, PM_ERR_ARG, "n_samples out of range");
	const int seg_words = e->opt_seg_words;
	int chk_words = std::min(e->opt_chk_words, seg_words);
	while (seg_words % chk_words) chk_words--;
	const long long seg_len = 32ll * seg_words;
	if (plan.own_begin < 0 || plan.own_begin % seg_len)
		return fail(e, PM_ERR_ARG, "own_begin must be a multiple of segment_len (%lld)", seg_len);
	if (!plan.last && (plan.own_len <= 0 || plan.own_len % seg_len))
		return fail(e, PM_ERR_ARG, "own_len must be a multiple of segment_len (%lld) on all but the last shard", seg_len);
	if (plan.first && (plan.own_begin != 0 || plan.sample_base != 0))
		return fail(e, PM_ERR_ARG, "the first shard starts at sample 0");
	if (plan.tail_bits < 0 || plan.tail_bits % 32) return fail(e, PM_ERR_ARG, "tail_bits must be a multiple of 32");
	if (plan.pre_segments < 0 || (plan.first && plan.pre_segments) || plan.own_begin < (long long)plan.pre_segments * seg_len)
		return fail(e, PM_ERR_ARG, "pre_segments needs own_begin >= pre_segments * segment_len (and 0 on the first shard)");
	if (sharded && !(plan.first && plan.last) && plan.tail_bits < 128)
		return fail(e, PM_ERR_ARG, "tail_bits must be at least 128");
	e->n_samples = n;
	e->plan = plan;
	e->sharded = sharded;
	e->sample_base = plan.sample_base;
	int n_quad = 0;
	for (int c = 0; c < nc; c++)
		if (e->chains[c].d.slicer_kind == PM_SLICER_QUADRATURE) e->chains[c].sign_q_row = nc + n_quad++;
	e->sign_rows = nc + n_quad;
	long long max_words = 0, max_bits = 0, max_nout = 0;
	std::vector<SlicerChain> sl(nc);
	std::vector<BitChain> bc(nc);
	memset(sl.data(), 0, nc * sizeof(SlicerChain));
	e->h_init.assign(nc, SegState());
	long long rec_cap = 16, arena_cap = 64;
	for (int c = 0; c < nc; c++) {
		HostChain &hc = e->chains[c];
		if (sharded && (hc.d.slicer_kind != PM_SLICER_BINARY || (hc.p64 && hc.d.modem_kind != PM_MODEM_AFSK)))
			return fail(e, PM_ERR_UNSUPPORTED, "chain %d: chains with a carrier loop (AGC takes max() of the whole "
				"recording, agc.py:67) cannot be sharded on the sample axis", c);
		const long long nout = (only_chain >= 0 && c != only_chain) ? 0 : std::max<long long>(0, chain_n(e, c, n) - hc.trim);
		max_nout = std::max(max_nout, nout);
		const int tile = (hc.p64 || hc.group < 0) ? P64_TILE : e->groups[hc.group].tile;
		const long long tiles = (nout + tile - 1) / tile;
		max_words = std::max(max_words, tiles * tile / 32);
		SlicerChain &s = sl[c];
		s.sps = hc.d.slicer_sample_rate / hc.d.symbol_rate;       // slicer.py:51
		s.thr = (s.sps / 2.0) - 0.5;                              // slicer.py:52
		s.lock = hc.d.lock_rate;
		s.nout = nout;
		const bool quad = hc.d.slicer_kind == PM_SLICER_QUADRATURE;
		s.quadrature = quad ? 1 : 0; s.sign_q_row = quad ? hc.sign_q_row : 0; s.sign_row = c;
		{
			// shortened clock update (SlicerChain): c_star = smallest double with fl(c_star + 1.0) >= thr; usable when
			// [c_star, thr + 1) lies inside one binade
			volatile double probe;
			double cs = s.thr - 1.0;
			bool fast = cs >= 1.0 && e->opt_slicer_fast;
			if (fast) {
				for (int k = 0; k < 4; k++) cs = std::nextafter(cs, -INFINITY);
				for (int guard = 0; guard < 16; guard++) {
					probe = cs + 1.0;
					if (probe >= s.thr) break;
					cs = std::nextafter(cs, INFINITY);
				}
				probe = cs + 1.0;
				int ex = 0;
				std::frexp(cs, &ex);                       // cs in [2^(ex-1), 2^ex)
				fast = probe >= s.thr && cs >= 1.0 && s.thr + 1.0 <= std::ldexp(1.0, ex);
				probe = std::nextafter(cs, -INFINITY) + 1.0;
				fast = fast && !(probe >= s.thr);          // cs really is the smallest
			}
			s.fast = fast ? 1 : 0;
			memcpy(&s.c_star_bits, &cs, sizeof(double));
			s.sps_m1 = s.sps - 1.0;
			// exact repetition of the clock in stretches without crossings (SlicerChain::quiet_words): sps a dyadic
			// rational, P = its smallest integer multiple, the repetition lcm(32, P) / 32 words long
			s.quiet_words = 0; s.quiet_lead = 0;
			if (e->opt_quiet_skip && s.sps >= 2.0 && s.sps * 65536.0 == std::floor(s.sps * 65536.0)) {
				long long P = 0;
				for (int k = 1; k <= 64 && !P; k++)
					if (s.sps * k == std::floor(s.sps * k)) P = (long long)(s.sps * k);
				if (P > 0 && P <= 4096) {
					long long a = 32, b = P;
					while (b) { const long long t = a % b; a = b; b = t; }       // a = gcd(32, P)
					const long long lw = P / a;
					if (lw <= PM_QUIET_MAX) { s.quiet_words = (int)lw; s.quiet_lead = (int)((P + 31) / 32) + 1; }
				}
			}
		}
		BitChain &b = bc[c];
		memset(&b, 0, sizeof(b));
		b.nout = nout; b.sign_row = c; b.bps = 1; b.lfsr_poly = hc.d.lfsr_poly; b.lfsr_invert = hc.d.lfsr_invert;
		if (quad) {
			b.quadrature = 1; b.sign_q_row = hc.sign_q_row; b.bps = (int)hc.d.bits_per_symbol;
			b.state_mask = hc.d.state_mask;
			for (int i = 0; i < 16; i++) b.demap[i] = hc.d.demap[i];
		}
		b.codec = hc.d.codec_kind;
		b.il2p_crc = hc.d.il2p_crc; b.il2p_disable_rs = hc.d.il2p_disable_rs;
		b.il2p_min_dist = hc.d.il2p_min_dist; b.il2p_sync_tol = hc.d.il2p_sync_tol;
		e->h_init[c].clock = 0.0; e->h_init[c].last = 1; e->h_init[c].last_q = 1;    // slicer.py:50,55
		const long long min_gap = std::max<long long>(1, (long long)std::ceil(s.thr));
		const long long mb = std::max((nout / min_gap + 8) * b.bps + 64 + plan.tail_bits, min_bits);
		max_bits = std::max(max_bits, mb);
		// shortest frame on the air: AX.25 18 bytes + a flag = 152 stream bits; IL2P a header-only frame without trailing
		// CRC = 24 sync bits + 15 bytes = 144 (back to back, il2p.py:367-409); buffers grow and the run repeats when a
		// recording still overflows them (rec_scale)
		rec_cap += (long long)((mb / (hc.d.codec_kind == PM_CODEC_IL2P ? 144 : 152) + 4) * e->rec_scale);
		arena_cap += (long long)((mb / 8 + 64) * e->rec_scale);
	}
	max_words = round_up((int)max_words, 4) + 4;
	e->sign_stride = max_words;
	// slicer geometry: segment 0 starts at own_begin
	e->own_w0 = plan.own_begin / 32;
	e->end_w = (max_nout + 31) / 32;
	if (plan.last || plan.own_begin + plan.own_len >= max_nout) e->own_w1 = e->end_w;
	else e->own_w1 = (plan.own_begin + plan.own_len) / 32;
	if (e->own_w0 > e->end_w) e->own_w0 = e->end_w;
	SlicerGeom &G = e->geom;
	e->k0 = plan.pre_segments;
	G.origin_w = e->own_w0 - (long long)e->k0 * seg_words;
	G.k_init = 0;
	G.seg_words = seg_words;
	G.warm_words = e->opt_warm_words;
	G.warm_f32_words = (e->opt_warm_exact_words > 0 && e->opt_warm_exact_words < G.warm_words) ?
		G.warm_words - e->opt_warm_exact_words : 0;
	G.warm_far_f64 = e->opt_warm_far_f64;
	G.chk_words = chk_words;
	G.n_chk = seg_words / chk_words;
	G.n_seg = e->k0 + (int)std::max<long long>(1, (e->end_w - e->own_w0 + seg_words - 1) / seg_words);
	G.true_start = plan.first ? 1 : 0;
	G.k_first = 0;
	G.k_count = G.n_seg;
	e->k_end = e->k0 + (int)std::min<long long>(G.n_seg - e->k0,
		std::max<long long>(1, (e->own_w1 - e->own_w0 + seg_words - 1) / seg_words));
	e->n_seg = G.n_seg;
	e->bits_stride = round_up((int)((max_bits + 31) / 32) + 8, 4);
	e->addr_stride = max_bits / 8 + 16;
	e->flag_stride = max_bits / 7 + 16;
	e->scratch_stride = max_bits / 8 + 64;
	CK(e->d_sign.ensure((size_t)e->sign_rows * e->sign_stride));
	CK(e->d_mask.ensure((size_t)nc * e->sign_stride));
	CK(e->d_S.ensure((size_t)nc * G.n_seg));
	CK(e->d_E0.ensure((size_t)nc * G.n_seg));
	CK(e->d_E1.ensure((size_t)nc * G.n_seg));
	CK(e->d_chk.ensure((size_t)nc * G.n_seg * G.n_chk));
	CK(e->d_repair_list.ensure((size_t)nc * G.n_seg));
	CK(e->d_shardbits.ensure(nc));
	CK(e->d_symcount.ensure(nc));
	CK(e->d_tail.ensure((size_t)nc * std::max(1, plan.tail_bits / 32)));
	const long long n_gblk = std::max((max_words + 1023) / 1024, (e->bits_stride + 1023) / 1024) + 1;
	CK(e->d_blk_count.ensure((size_t)nc * n_gblk));
	CK(e->d_blk_base.ensure((size_t)nc * n_gblk));
	CK(e->d_bits_raw.ensure((size_t)nc * e->bits_stride));
	CK(e->d_bits_lfsr.ensure((size_t)nc * e->bits_stride));
	CK(e->d_byte_addr.ensure((size_t)nc * e->addr_stride));
	CK(e->d_flag_pos.ensure((size_t)nc * e->flag_stride));
	CK(e->d_gaps.ensure((size_t)nc * e->flag_stride));
	CK(e->d_gap_cand.ensure((size_t)nc * e->flag_stride));
	CK(e->d_gap_ncand.ensure(nc));
	CK(e->d_scratch.ensure((size_t)nc * e->scratch_stride));
	CK(e->d_recs.ensure((size_t)rec_cap));
	CK(e->d_rec_src.ensure((size_t)rec_cap));
	CK(e->d_arena.ensure((size_t)arena_cap));
	CK(e->d_guard_entries.ensure(e->guard_cap));
	if (e->has_il2p) {
		e->il2p_cand_cap = (int)std::min<long long>(e->flag_stride, (long long)((max_bits / 256 + 64) * e->il2p_cand_scale));
		CK(e->d_il2p_slots.ensure((size_t)nc * (e->il2p_cand_cap + 1) * IL2P_SLOT));
		CK(e->d_il2p_res.ensure((size_t)nc * e->il2p_cand_cap));
		CK(e->d_il2p_hand.ensure((size_t)2 * nc));
	}
	if (e->opt_keep_soft) {
		e->soft_stride = n;
		CK(e->d_soft.ensure((size_t)e->sign_rows * n));
	}
	if (!e->h_p64.empty()) {
		// work buffers of the float64 pipeline: A (+B, +C, +D), n doubles each
		size_t need = 0;
		for (auto &P : e->h_p64) need += (size_t)(P.kind == PM_MODEM_MPSK ? 4 : 2) * (size_t)n;
		CK(e->d_p64_work.ensure(need + 8));
		double *w = e->d_p64_work.p;
		for (size_t i = 0; i < e->h_p64.size(); i++) {
			P64Chain &P = e->h_p64[i];
			const HostChain &hc = e->chains[P.gid];
			P.sign_row = P.gid;
			P.sign_q_row = hc.sign_q_row;
			P.n_audio = chain_n(e, P.gid, n);
			P.audio_off = chain_aoff(e, P.gid);
			P.L1 = std::max<long long>(0, P.n_audio - (P.n_bpf - 1));
			if (P.kind == PM_MODEM_AFSK || P.kind == PM_MODEM_MPSK) P.L2 = std::max<long long>(0, P.L1 - (P.n_mid - 1));
			else P.L2 = P.L1;
			P.L3 = std::max<long long>(0, P.L2 - (P.n_out - 1));
			if (P.L3 == 0) P.L1 = P.L2 = 0;     // numpy.convolve 'valid' on a signal shorter than the taps: nothing usable
			P.A = w; w += n;
			P.B = w; w += n;
			if (P.kind == PM_MODEM_MPSK) { P.C = w; w += n; P.D = w; w += n; }
			P.max_slot = e->d_p64_max.p + i;
		}
		if (!same_bytes(e->up_p64, e->h_p64)) {
			CK(cudaMemcpyAsync(e->d_p64.p, e->h_p64.data(), e->h_p64.size() * sizeof(P64Chain), cudaMemcpyHostToDevice, e->st));
			e->up_p64 = e->h_p64;
		}
	}
	// the per-run tables only travel when they changed (same recording length and plan as last time: nothing to upload;
	// a copy from pageable memory returns once the source has been staged, so the vectors may go out of scope)
	if (!same_bytes(e->up_sl, sl)) {
		CK(cudaMemcpyAsync(e->d_slicer.p, sl.data(), nc * sizeof(SlicerChain), cudaMemcpyHostToDevice, e->st));
		e->up_sl = sl;
	}
	if (!same_bytes(e->up_bc, bc)) {
		CK(cudaMemcpyAsync(e->d_bitchain.p, bc.data(), nc * sizeof(BitChain), cudaMemcpyHostToDevice, e->st));
		e->up_bc = bc;
	}
	if (!same_bytes(e->up_init, e->h_init)) {
		CK(cudaMemcpyAsync(e->d_init.p, e->h_init.data(), nc * sizeof(SegState), cudaMemcpyHostToDevice, e->st));
		e->up_init = e->h_init;
	}
	CK(cudaMemsetAsync(e->d_counters.p, 0, 16 * sizeof(unsigned int), e->st));
	CK(cudaMemsetAsync(e->d_totals.p, 0, sizeof(PacketTotals), e->st));
	for (int c = 0; c < nc; c++) {
		e->h_fp64[c].audio_off = chain_aoff(e, c);
		e->h_fp64[c].n_audio = chain_n(e, c, n);
	}
	if (!same_bytes(e->up_fp64, e->h_fp64)) {
		CK(cudaMemcpyAsync(e->d_fp64.p, e->h_fp64.data(), nc * sizeof(Fp64Chain), cudaMemcpyHostToDevice, e->st));
		e->up_fp64 = e->h_fp64;
	}
	long long mag_bytes = 0, amax_floats = 0;
	for (auto &g : e->groups) {
		if (g.kind == PM_MODEM_AFSK)
			for (int i = 0; i < g.afsk.n_chain; i++)
				g.afsk.chain_nout[i] = std::max<long long>(0, chain_n(e, g.afsk.chain_gid[i], n) - e->chains[g.afsk.chain_gid[i]].trim);
		else
			for (int i = 0; i < g.fir.n_chain; i++)
				g.fir.chain_nout[i] = std::max<long long>(0, chain_n(e, g.fir.chain_gid[i], n) - e->chains[g.fir.chain_gid[i]].trim);
		g.a_done = g.b_done = 0;
		if (g.tensor) {
			long long nout_max = 0;
			for (int i = 0; i < g.afsk.n_chain; i++) {
				g.tc.chain_nout[i] = (only_chain >= 0) ? 0 : g.afsk.chain_nout[i];
				nout_max = std::max(nout_max, g.tc.chain_nout[i]);
			}
			// low-pass tiles of TC_TILE outputs; the front tiles have to cover their inputs: TC_KBLK more rows of 64
			g.n_tile_b = (nout_max + TC_TILE - 1) / TC_TILE;
			g.n_tile_a = g.n_tile_b ? (g.n_tile_b * TC_TILE + 64 * TC_KBLK + g.tile - 1) / g.tile : 0;
			g.tc.n_tile_a = g.n_tile_a;
			g.mag_rows = ((g.n_tile_a * g.tile + 63) / 64 + 8 + 7) / 8 * 8;
			g.mag_off = mag_bytes;
			g.amax_off = amax_floats;
			mag_bytes += (long long)g.afsk.n_mag * 3 * g.mag_rows * 128;
			amax_floats += g.n_tile_a + 1;
		}
	}
	if (mag_bytes) {
		CK(e->d_mag.ensure((size_t)mag_bytes));
		CK(e->d_amax.ensure((size_t)amax_floats));
	}
	return PM_OK;
}

static GuardList guard_of(pm_engine *e)
{
	GuardList g;
	g.entries = e->d_guard_entries.p;
	g.count = e->d_counters.p;
	g.cap = e->guard_cap;
	g.stage_clk = e->opt_stage_clocks ? e->d_stage_clk.p : nullptr;
	g.from = g.to = nullptr;
	return g;
}

// launch the front-end tiles of every group that became complete with the
// samples in [0, avail_to) and were not launched for [0, avail_from)
static int launch_front(pm_engine *e, const int16_t *d_audio, long long n, long long avail_from, long long avail_to,
                        bool last, cudaStream_t stream = nullptr)
{
	if (!stream) stream = e->st;
	const int16_t *audio_all = d_audio;
	const long long n_all = n;
	for (auto &g : e->groups) {
		const int a_len = (g.kind == PM_MODEM_AFSK) ? g.afsk.a_len : g.fir.a_len;
		// the group's recording (all its chains share it)
		const long long n = chain_n(e, g.chains[0], n_all);
		const int16_t *d_audio = audio_all + chain_aoff(e, g.chains[0]);
		long long nout_max = 0;
		for (int k : g.chains) nout_max = std::max(nout_max, std::max<long long>(0, n - e->chains[k].trim));
		const long long tiles_total = g.tensor ? g.n_tile_a : (nout_max + g.tile - 1) / g.tile;
		auto ready = [&](long long upto, bool fin) -> long long {
			if (fin) return tiles_total;
			long long t = (upto - a_len) / g.tile + 1;      // tiles with t*tile + a_len <= upto
			if (upto < a_len) t = 0;
			return std::min(std::max<long long>(t, 0), tiles_total);
		};
		const long long t0 = ready(avail_from, false), t1 = ready(avail_to, last);
		long long t = t0;
		while (t < t1) {
			const int cnt = (int)std::min<long long>(t1 - t, 1 << 30);
			cudaError_t ce;
			MagOut mo;
			mo.base = g.tensor ? e->d_mag.p + g.mag_off : nullptr;
			mo.rows = g.mag_rows;
			mo.tile_amax = g.tensor ? e->d_amax.p + g.amax_off : nullptr;
			if (g.kind == PM_MODEM_AFSK)
				ce = pm_launch_afsk_front(&g.afsk, g.smem, d_audio, n, t, cnt, e->d_sign.p, e->sign_stride,
					e->opt_keep_soft ? e->d_soft.p : nullptr, e->soft_stride, guard_of(e), mo, stream);
			else
				ce = pm_launch_fir_front(&g.fir, g.smem, d_audio, n, t, cnt, e->d_sign.p, e->sign_stride,
					e->opt_keep_soft ? e->d_soft.p : nullptr, e->soft_stride, guard_of(e), stream);
			if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "front-end launch failed: %s", cudaGetErrorString(ce));
			if (e->opt_debug_sync && (ce = cudaStreamSynchronize(stream)) != cudaSuccess)
				return fail(e, PM_ERR_CUDA, "front-end kernel (tiles %lld..%lld of %lld) failed: %s", t, t + cnt, tiles_total, cudaGetErrorString(ce));
			e->stats.kernel_launches++;
			e->stats.front_launches++;
			t += cnt;
		}
		g.a_done = std::max(g.a_done, t1);
		if (g.tensor) {
			// low-pass tiles whose input rows are all written now: tile k reads the magnitudes [TC_TILE k, TC_TILE (k + 1) + 64 TC_KBLK)
			long long b_ready = (g.a_done >= g.n_tile_a) ? g.n_tile_b : (g.a_done * g.tile - 64 * TC_KBLK) / TC_TILE;
			b_ready = std::min(std::max<long long>(b_ready, 0), g.n_tile_b);
			if (b_ready > g.b_done) {
				const size_t gi = (size_t)(&g - e->groups.data());
				cudaError_t ce = pm_launch_lpf_tc(&g.tc, e->d_mag.p + g.mag_off, g.mag_rows, e->d_btaps[gi].p,
					e->d_amax.p + g.amax_off, g.b_done, b_ready - g.b_done, e->d_sign.p, e->sign_stride,
					e->opt_keep_soft ? e->d_soft.p : nullptr, e->soft_stride, guard_of(e), (int *)(e->d_counters.p + 15),
					e->sm_count, stream);
				if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "tensor-core low-pass launch failed: %s", cudaGetErrorString(ce));
				if (e->opt_debug_sync && (ce = cudaStreamSynchronize(stream)) != cudaSuccess)
					return fail(e, PM_ERR_CUDA, "tensor-core low-pass kernel (tiles %lld..%lld of %lld, %lld rows, %lld front tiles of %d) failed: %s",
						g.b_done, b_ready, g.n_tile_b, g.mag_rows, g.n_tile_a, g.tile, cudaGetErrorString(ce));
				e->stats.kernel_launches++;
				g.b_done = b_ready;
			}
		}
	}
	return PM_OK;
}

// Verify/repair until every hand-off inside the local buffer is bit-exact.
static int slicer_converge(pm_engine *e)
{
	const int nc = (int)e->chains.size();
	cudaError_t ce;
	for (int pass = 0;; pass++) {
		CK(cudaMemsetAsync(e->d_counters.p + 1, 0, sizeof(unsigned int), e->st));
		if (pass < e->opt_verify_passes) {
			ce = pm_launch_slicer_verify(e->d_slicer.p, nc, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride,
				e->d_S.p, e->E_cur, e->E_alt, e->d_chk.p, e->d_init.p, e->geom, e->d_counters.p + 1, nullptr, e->d_repair_list.p, e->st);
			std::swap(e->E_cur, e->E_alt);
		} else {
			ce = pm_launch_slicer_sweep(e->d_slicer.p, nc, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride,
				e->d_S.p, e->E_cur, e->d_chk.p, e->d_init.p, e->geom, e->d_counters.p + 1, e->st);
		}
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "slicer verify launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches += 2;
		pm_kt_mark("d2h counters + host sync", e->st);
		CK(cudaMemcpyAsync(e->h_counters, e->d_counters.p, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, e->st));
		CK(cudaStreamSynchronize(e->st));
		e->stats.slicer_repairs += e->h_counters[1];
		if (pass >= e->opt_verify_passes) break;          // the sweep is exact by construction
		if (e->h_counters[1] == 0) break;
	}
	return PM_OK;
}

// The same passes without asking the host in between: up to FAST_PASSES verify passes are enqueued back to back, each
// one skipping itself when its predecessor repaired nothing (d_counters[2 + p] = repairs of pass p).  Whether that was
// enough is read from the counters with the run's results (slicer_fast_converged); almost always it is -- pass 0
// repairs the ~0.15 % of hand-offs whose warm-up did not become bit-identical, pass 1 finds nothing; the later passes
// are there for short segments, where a repaired segment more often ends in a different state than the first run did.
#define FAST_PASSES 8      // d_counters[2 .. 10): a repair whose segment does not merge with a checkpoint moves the next hand-off, one segment per pass
static int slicer_enqueue_fast(pm_engine *e)
{
	const int nc = (int)e->chains.size();
	e->fast_passes = std::min(e->opt_verify_passes, FAST_PASSES);
	for (int p = 0; p < e->fast_passes; p++) {
		cudaError_t ce = pm_launch_slicer_verify(e->d_slicer.p, nc, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride,
			e->d_S.p, e->E_cur, e->E_alt, e->d_chk.p, e->d_init.p, e->geom, e->d_counters.p + 2 + p,
			p ? e->d_counters.p + 2 + p - 1 : nullptr, e->d_repair_list.p, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "slicer verify launch failed: %s", cudaGetErrorString(ce));
		std::swap(e->E_cur, e->E_alt);
		e->stats.kernel_launches += 2;
	}
	e->fast_pending = true;
	return PM_OK;
}

// after the counters came back: did some enqueued pass find nothing to repair?
static bool slicer_fast_converged(pm_engine *e)
{
	bool ok = false;
	for (int p = 0; p < e->fast_passes; p++) {
		e->stats.slicer_repairs += e->h_counters[2 + p];
		if (e->h_counters[2 + p] == 0) ok = true;
	}
	e->fast_pending = false;
	return ok;
}

// start state (S of the first own segment k0), end state (E of segment k_end-1) and own symbol count of every chain
static int read_shard_states(pm_engine *e)
{
	const int nc = (int)e->chains.size();
	const int n_seg = e->geom.n_seg;
	std::vector<SegState> s0(nc), e1(nc);
	std::vector<unsigned long long> cnt(nc);
	cudaError_t ce = pm_launch_slicer_count(e->d_slicer.p, nc, e->d_mask.p, e->sign_stride, e->own_w0, e->own_w1,
		e->d_symcount.p, e->st);
	if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "symbol count launch failed: %s", cudaGetErrorString(ce));
	e->stats.kernel_launches++;
	CK(cudaMemcpy2DAsync(s0.data(), sizeof(SegState), e->d_S.p + e->k0, (size_t)n_seg * sizeof(SegState), sizeof(SegState), nc,
		cudaMemcpyDeviceToHost, e->st));
	CK(cudaMemcpy2DAsync(e1.data(), sizeof(SegState), e->E_cur + (e->k_end - 1), (size_t)n_seg * sizeof(SegState),
		sizeof(SegState), nc, cudaMemcpyDeviceToHost, e->st));
	CK(cudaMemcpyAsync(cnt.data(), e->d_symcount.p, nc * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->st));
	CK(cudaStreamSynchronize(e->st));
	e->shard_out.resize(nc);
	for (int c = 0; c < nc; c++) {
		pm_shard_state &o = e->shard_out[c];
		o.start_clock = s0[c].clock; o.start_last = s0[c].last; o.start_last_q = s0[c].last_q;
		o.end_clock = e1[c].clock; o.end_last = e1[c].last; o.end_last_q = e1[c].last_q;
		o.n_symbols = (int64_t)cnt[c];
	}
	return PM_OK;
}

// shard states to the host; on a later shard the device-side init[] holds the speculated start states
static int fetch_shard_states(pm_engine *e)
{
	int rc = read_shard_states(e);
	if (rc != PM_OK) return rc;
	if (!e->plan.first)
		for (size_t c = 0; c < e->chains.size(); c++) {
			e->h_init[c].clock = e->shard_out[c].start_clock;
			e->h_init[c].last = e->shard_out[c].start_last;
			e->h_init[c].last_q = e->shard_out[c].start_last_q;
		}
	return PM_OK;
}

// true when the driver can DMA straight out of [p, ...): cudaHostAlloc'ed, cudaHostRegister'ed or managed memory
static bool host_pointer_is_pinned(const void *p)
{
	cudaPointerAttributes pa;
	if (cudaPointerGetAttributes(&pa, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged;
}

// memcpy on a few host threads: one thread copies ~10 GB/s, the PCIe link takes 50
static void parallel_copy(void *dst, const void *src, size_t bytes, int threads)
{
	if (threads <= 1 || bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	const size_t per = ((bytes + threads - 1) / threads + 4095) & ~(size_t)4095;
	for (int t = 1; t < threads; t++) {
		const size_t off = per * t;
		if (off >= bytes) break;
		th.emplace_back([=]() { memcpy((char *)dst + off, (const char *)src + off, std::min(per, bytes - off)); });
	}
	memcpy(dst, src, std::min(per, bytes));
	for (auto &t : th) t.join();
}

// option "trace": a timing event on a stream, reported by pm_engine_trace relative to the start of the run
static void trace_mark(pm_engine *e, const char *what, int i, cudaStream_t st)
{
	if (!e->opt_trace) return;
	if (e->trace_used == e->ev_trace.size()) {
		cudaEvent_t ev;
		if (cudaEventCreate(&ev) != cudaSuccess) return;
		e->ev_trace.push_back(ev);
		e->trace_label.emplace_back();
	}
	char buf[64];
	snprintf(buf, sizeof(buf), "%s %d", what, i);
	e->trace_label[e->trace_used] = buf;
	cudaEventRecord(e->ev_trace[e->trace_used++], st);
}

// guard_fixup_kernel, doubles per warp: audio window, band-passed window, 2 magnitude rows
static int fixup_doubles(const pm_engine *e)
{
	int max_sum = 8;
	for (auto &hc : e->chains) {
		const int nx = (int)(hc.mark_i.size() + hc.lpf.size());
		const int nx_pad = (nx + 159) / 160 * 160, mrow = ((int)hc.lpf.size() + 223) / 224 * 224;
		max_sum = std::max(max_sum, nx_pad + (int)hc.bpf.size() + nx_pad + (int)hc.mark_i.size() + 7 * 32 + 2 * mrow + 8);
	}
	return max_sum;
}

static int begin_once(pm_engine *e, const int16_t *audio, long long n, bool on_host, const pm_shard_plan &plan,
                      bool sharded, bool defer)
{
	const int nc = (int)e->chains.size();
	for (int c = 0; c < nc; c++)
		if (e->chains[c].d.modem_kind == PM_MODEM_NONE)
			return fail(e, PM_ERR_STATE, "chain %d has no modem: it only serves the per-stage calls (slice_soft / unscramble_stream / decode_stream)", c);
	memset(&e->stats, 0, sizeof(e->stats));
	int rc = prepare_run(e, n, plan, sharded);
	if (rc != PM_OK) return rc;
	const int16_t *d_audio = audio;
	e->early_fix = e->early_seg = false;
	e->trace_used = 0;
	CK(cudaEventRecord(e->ev[0], e->st));
	pm_kt_mark("(launch gap)", e->st);
	if (on_host) {
		CK(e->d_audio.ensure((size_t)n + 64));
		d_audio = e->d_audio.p;
		// chunk sizes ramp up (512 Ki samples, doubling to "h2d_chunk"): the front end cannot start before the first
		// chunk has landed, so that one is small; afterwards the copy runs ahead of the (slower) front-end kernels
		std::vector<long long> cuts;
		{
			long long done = 0, len = std::min<long long>(e->opt_h2d_chunk, 512 << 10);
			while (done < n) {
				const long long take = std::min(len, n - done);
				done += take;
				cuts.push_back(done);
				len = std::min(e->opt_h2d_chunk, len * 2);
			}
		}
		const int n_chunks = (int)cuts.size();
		while ((int)e->ev_chunks.size() < n_chunks) {
			cudaEvent_t ev;
			CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
			e->ev_chunks.push_back(ev);
		}
		// Early tail (option early_tail).  1: the guard fix-up does not wait for the last chunk.  After the front-end launches
		// of a chunk the guard count is snapshot on the launch stream and the fix-up of the entries since the previous
		// snapshot runs on st_tail[0], beside the front end of the following chunks: what follows the last byte of the
		// recording is one chunk's front end, a small fix-up, then slicer and bit level as on the device-resident path.
		// 2: the slicer's speculative segments are launched early too, in batches, on st_tail[1] / st_tail[2], as soon as
		// all sign words below them are final.  Measured and NOT the default: a batch of segments takes as long as ALL of
		// them (a thread's chain of ~74000 dependent steps is 1.1 ms, the whole kernel 1.1 ms), so the batch after the
		// last chunk costs what the single launch costs, and the earlier ones hold the front end back while they share its
		// SMs (profiles/r02ad_e2e_trace.txt: 9.4-11.9 ms against 8.9).
		const int early = (defer && e->h_p64.empty() && n_chunks > 2) ? e->opt_early_tail : 0;
		int k_done = 0, n_slc = 0, snap_prev = 0;
		const int k_batch = std::max(1, e->geom.n_seg / std::max(1, e->opt_early_batches));
		if (early) {
			CK(e->d_snap.ensure((size_t)n_chunks + 2));
			CK(cudaMemsetAsync(e->d_snap.p, 0, sizeof(unsigned int), e->st));
			while ((int)e->ev_steps.size() < 2 * n_chunks) {
				cudaEvent_t ev;
				CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
				e->ev_steps.push_back(ev);
			}
			e->E_cur = e->d_E0.p; e->E_alt = e->d_E1.p;
		}
		// the copy stream must not start before the engine stream reached this point
		CK(cudaEventRecord(e->ev[6], e->st));
		CK(cudaStreamWaitEvent(e->st_copy, e->ev[6], 0));
		for (int q = 0; q < 2; q++) CK(cudaStreamWaitEvent(e->st_front[q], e->ev[6], 0));
		if (early)
			for (int q = 0; q < 3; q++) CK(cudaStreamWaitEvent(e->st_tail[q], e->ev[6], 0));
		// pageable caller memory: the driver would stage it through its own small bounce buffers, synchronously.  Stage
		// it ourselves through a ring of pinned buffers: the memcpy of chunk i+1 (host threads) overlaps the DMA of chunk i
		const bool staged = !host_pointer_is_pinned(audio);
		e->staged_bytes = staged ? (int64_t)n * 2 : 0;
		if (staged) {
			const size_t want = (size_t)std::min<long long>(e->opt_h2d_chunk, n);
			if (want > e->ring_samples) {
				for (int i = 0; i < pm_engine::RING; i++) {
					if (e->h_ring[i]) cudaFreeHost(e->h_ring[i]);
					e->h_ring[i] = nullptr;
					CK(cudaHostAlloc((void **)&e->h_ring[i], want * sizeof(int16_t), cudaHostAllocDefault));
					if (!e->ev_ring[i]) CK(cudaEventCreateWithFlags(&e->ev_ring[i], cudaEventDisableTiming));
				}
				e->ring_samples = want;
			}
		}
		bool any_tensor = false;
		for (auto &g : e->groups) any_tensor = any_tensor || g.tensor;
		long long done = 0;
		for (int i = 0; i < n_chunks; i++) {
			const long long len = cuts[i] - done;
			// (a low-pass tile reads magnitudes of front tiles from earlier chunks: one launch stream keeps them ordered)
			cudaStream_t fs = e->st_front[(any_tensor || early) ? 0 : (i & 1)];
			const int16_t *src = audio + done;
			if (staged) {
				const int slot = i % pm_engine::RING;
				if (i >= pm_engine::RING) CK(cudaEventSynchronize(e->ev_ring[slot]));      // its previous DMA has read the buffer
				parallel_copy(e->h_ring[slot], audio + done, (size_t)len * sizeof(int16_t), e->opt_copy_threads);
				src = e->h_ring[slot];
			}
			CK(cudaMemcpyAsync(e->d_audio.p + done, src, (size_t)len * sizeof(int16_t),
				cudaMemcpyHostToDevice, e->st_copy));
			if (staged) CK(cudaEventRecord(e->ev_ring[i % pm_engine::RING], e->st_copy));
			CK(cudaEventRecord(e->ev_chunks[i], e->st_copy));
			trace_mark(e, "copied", i, e->st_copy);
			CK(cudaStreamWaitEvent(fs, e->ev_chunks[i], 0));
			rc = launch_front(e, d_audio, n, done, done + len, i == n_chunks - 1, fs);
			if (rc != PM_OK) return rc;
			done += len;
			trace_mark(e, "front", i, fs);
			if (early) {
				const bool last = i == n_chunks - 1;
				const SlicerGeom &G = e->geom;
				long long upto = std::numeric_limits<long long>::max();     // sign words below this sample are written
				for (auto &g : e->groups) upto = std::min(upto, g.tensor ? g.b_done * (long long)TC_TILE : g.a_done * (long long)g.tile);
				const int k_ready = last ? G.n_seg
					: (int)std::min<long long>(G.n_seg, std::max<long long>(0, (upto / 32 - G.origin_w) / G.seg_words));
				if (last || early == 1 || k_ready - k_done >= k_batch) {
					if (pm_launch_guard_snapshot(e->d_counters.p, e->d_snap.p + i + 1, fs) != cudaSuccess)
						return fail(e, PM_ERR_CUDA, "guard snapshot launch failed");
					CK(cudaEventRecord(e->ev_steps[2 * i], fs));
					CK(cudaStreamWaitEvent(e->st_tail[0], e->ev_steps[2 * i], 0));
					GuardList gl = guard_of(e);
					gl.from = e->d_snap.p + snap_prev;
					gl.to = e->d_snap.p + i + 1;
					snap_prev = i + 1;
					cudaError_t ce = pm_launch_guard_fixup(e->d_fp64.p, fixup_doubles(e), d_audio, n, e->d_sign.p, e->sign_stride,
						e->opt_keep_soft ? e->d_soft.p : nullptr, e->soft_stride, gl, last ? 148 * 5 : 148, e->st_tail[0]);
					if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "fixup launch failed: %s", cudaGetErrorString(ce));
					e->stats.kernel_launches++;
					CK(cudaEventRecord(e->ev_steps[2 * i + 1], e->st_tail[0]));
					trace_mark(e, "fixup", i, e->st_tail[0]);
					if (early == 2 && k_ready > k_done) {
						cudaStream_t ss = e->st_tail[1 + (n_slc++ & 1)];
						CK(cudaStreamWaitEvent(ss, e->ev_steps[2 * i + 1], 0));
						SlicerGeom G2 = G;
						G2.k_first = k_done;
						G2.k_count = k_ready - k_done;
						ce = pm_launch_slicer_segments(e->d_slicer.p, nc, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride,
							e->d_S.p, e->E_cur, e->d_chk.p, e->d_init.p, G2, ss);
						if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "slicer launch failed: %s", cudaGetErrorString(ce));
						e->stats.kernel_launches++;
						trace_mark(e, "segments", i, ss);
						k_done = k_ready;
					}
				}
			}
		}
		if (early) {
			for (int q = 0; q < 3; q++) {
				CK(cudaEventRecord(e->ev_tail[q], e->st_tail[q]));
				CK(cudaStreamWaitEvent(e->st, e->ev_tail[q], 0));
			}
			e->early_fix = true;
			e->early_seg = early == 2;
		}
		for (int q = 0; q < 2; q++) {          // the engine stream continues when both launch streams are done
			CK(cudaEventRecord(e->ev_front[q], e->st_front[q]));
			CK(cudaStreamWaitEvent(e->st, e->ev_front[q], 0));
		}
		e->stats.h2d_bytes = (int64_t)n * 2;
	} else {
		rc = launch_front(e, d_audio, n, 0, n, true);
		if (rc != PM_OK) return rc;
	}
	if (!e->h_p64.empty()) {
		// float64 pipeline (needs the whole recording: AGC.normal = max over the band-passed buffer)
		cudaError_t pe = pm_launch_p64(e->d_p64.p, e->h_p64.data(), (int)e->h_p64.size(), d_audio, e->d_sign.p,
			e->sign_stride, e->opt_keep_soft ? e->d_soft.p : nullptr, e->soft_stride, e->d_p64_max.p, e->st);
		if (pe != cudaSuccess) return fail(e, PM_ERR_CUDA, "float64 pipeline launch failed: %s", cudaGetErrorString(pe));
		e->stats.kernel_launches += 6;
	}
	e->run_audio = d_audio;
	CK(cudaEventRecord(e->ev[1], e->st));

	cudaError_t ce;
	if (!e->early_fix) {
		// FP64 guard-band fix-up
		ce = pm_launch_guard_fixup(e->d_fp64.p, fixup_doubles(e), d_audio, n, e->d_sign.p, e->sign_stride,
			e->opt_keep_soft ? e->d_soft.p : nullptr, e->soft_stride, guard_of(e), 148 * 5, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "fixup launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches++;
	}
	CK(cudaEventRecord(e->ev[2], e->st));

	// slicer: speculative segments, then verify/repair
	if (!e->early_seg) {
		e->E_cur = e->d_E0.p; e->E_alt = e->d_E1.p;
		ce = pm_launch_slicer_segments(e->d_slicer.p, nc, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride,
			e->d_S.p, e->E_cur, e->d_chk.p, e->d_init.p, e->geom, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "slicer launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches++;
	}
	e->stats.slicer_segments = (int64_t)nc * e->geom.n_seg;
	if (!plan.first) {
		// the state at own_begin is only speculated so far: take it as given until the hand-off
		CK(cudaMemcpy2DAsync(e->d_init.p, sizeof(SegState), e->d_S.p, (size_t)e->geom.n_seg * sizeof(SegState),
			sizeof(SegState), nc, cudaMemcpyDeviceToDevice, e->st));
		e->up_init.clear();
	}
	if (defer) {
		// no host round trip: the verify passes are enqueued blind, the guard and repair counters come back with the
		// results (run_impl / run_linked_end look at them)
		rc = slicer_enqueue_fast(e);
		if (rc != PM_OK) return rc;
		CK(cudaEventRecord(e->ev[3], e->st));
		return PM_OK;
	}
	rc = slicer_converge(e);
	if (rc != PM_OK) return rc;
	if (e->h_counters[0] > e->guard_cap) return PM_ERR_CAPACITY;    // caller grows the list and re-runs
	e->stats.guard_flagged = e->h_counters[0];
	CK(cudaEventRecord(e->ev[3], e->st));
	return PM_OK;
}

static int shard_begin_impl(pm_engine *e, const int16_t *audio, long long n, bool on_host, const pm_shard_plan &plan,
                            bool sharded, bool read_states = true)
{
	if (!e) return PM_ERR_ARG;
	if (!audio) return fail(e, PM_ERR_ARG, "audio is NULL");
	cudaSetDevice(e->device);
	e->phase = 0;
	e->have_run = false;
	e->fast_pending = false;
	const bool defer = !sharded || !read_states;        // unsharded and linked runs: one host synchronisation, at the end
	for (int attempt = 0; attempt < 4; attempt++) {
		int rc = begin_once(e, audio, n, on_host, plan, sharded, defer);
		if (rc == PM_ERR_CAPACITY && e->h_counters[0] > e->guard_cap) {
			// guard list overflowed: grow it and run again
			e->guard_cap = e->h_counters[0] + e->h_counters[0] / 4 + 1024;
			cudaStreamSynchronize(e->st);
			continue;
		}
		if (rc != PM_OK) return rc;
		if (sharded && read_states) {
			rc = fetch_shard_states(e);
			if (rc != PM_OK) return rc;
		}
		e->phase = 1;
		return PM_OK;
	}
	return fail(e, PM_ERR_CAPACITY, "guard list kept overflowing");
}

static int shard_gather_impl(pm_engine *e, const int64_t *symbols_before, uint32_t *tail_out)
{
	if (!e || e->phase != 1) return fail(e, PM_ERR_STATE, "shard_gather: call shard_begin first");
	const int nc = (int)e->chains.size();
	const pm_shard_plan &plan = e->plan;
	std::vector<ShardBits> &sb = e->h_sb;
	sb.assign(nc, ShardBits());
	memset(sb.data(), 0, nc * sizeof(ShardBits));
	e->h_A0.assign(nc, 0);
	e->h_valid_from.assign(nc, 0);
	for (int c = 0; c < nc; c++) {
		ShardBits &b = sb[c];
		b.first = plan.first ? 1 : 0; b.pad = 0;
		b.valid_from = 0;
		if (plan.first) {
			b.bit_off = 0; b.own_lo = 0;
		} else {
			const long long P = symbols_before ? symbols_before[c] : 0;      // bps == 1: global bit index
			if (P < plan.tail_bits)
				return fail(e, PM_ERR_CAPACITY, "chain %d: earlier shards hold fewer bits (%lld) than the hand-off tail", c, P);
			const long long A0 = ((P - plan.tail_bits) >> 3) << 3;              // global bit index of local bit 0
			e->h_A0[c] = A0;
			b.bit_off = P - A0;
			b.own_lo = b.bit_off;
			int deg = 0;                                                        // LFSR history (lfsr.py:38-44)
			for (unsigned long long q = e->chains[c].d.lfsr_poly; q > 1; q >>= 1) deg++;
			b.valid_from = b.bit_off - plan.tail_bits + deg;
			e->h_valid_from[c] = b.valid_from;
		}
		if (plan.last) b.own_hi = 0x7fffffffffffffffll;
		else {
			const long long n_own = e->sharded ? e->shard_out[c].n_symbols : 0;
			if (n_own < plan.tail_bits)
				return fail(e, PM_ERR_CAPACITY, "chain %d: shard holds fewer bits (%lld) than the hand-off tail", c, n_own);
			b.own_hi = b.bit_off + n_own;
		}
	}
	pm_kt_mark("(launch gap)", e->st);
	if (!same_bytes(e->up_sb, sb)) {
		CK(cudaMemcpyAsync(e->d_shardbits.p, sb.data(), nc * sizeof(ShardBits), cudaMemcpyHostToDevice, e->st));
		e->up_sb = sb;
	}
	cudaError_t ce = pm_launch_gather(e->d_bitchain.p, nc, e->d_cc.p, e->d_sign.p, e->sign_stride, e->d_mask.p,
		e->sign_stride, e->own_w0, std::max<long long>(1, e->end_w - e->own_w0), e->d_blk_count.p, e->d_blk_base.p,
		e->d_sym_totals.p, e->d_bits_raw.p, e->bits_stride, e->d_byte_addr.p, e->addr_stride, nullptr,
		e->d_shardbits.p, e->st);
	if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "gather launch failed: %s", cudaGetErrorString(ce));
	e->stats.kernel_launches += 4;
	if (e->sharded && !plan.last && plan.tail_bits > 0 && tail_out) {
		ce = pm_launch_tail_extract(e->d_bits_raw.p, e->bits_stride, e->d_shardbits.p, nc, plan.tail_bits / 32,
			e->d_tail.p, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "tail extract launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches++;
		CK(cudaMemcpyAsync(tail_out, e->d_tail.p, (size_t)nc * (plan.tail_bits / 32) * sizeof(uint32_t),
			cudaMemcpyDeviceToHost, e->st));
	}
	if (e->sharded && tail_out) CK(cudaStreamSynchronize(e->st));       // tail_out is ready
	e->phase = 2;
	return PM_OK;
}

static int shard_finish_impl(pm_engine *e, const uint32_t *tail_in, const pm_il2p_state *il2p_prev = nullptr,
                             pm_il2p_state *il2p_out = nullptr)
{
	if (!e || e->phase != 2) return fail(e, PM_ERR_STATE, "shard_finish: call shard_gather first");
	const int nc = (int)e->chains.size();
	const pm_shard_plan &plan = e->plan;
	cudaError_t ce;
	std::vector<Il2pHand> hand(nc);
	if (e->has_il2p) {
		if (e->sharded && !(plan.first && plan.last) && (!il2p_out || (!plan.first && !il2p_prev)))
			return fail(e, PM_ERR_ARG, "a sharded run with IL2P chains finishes through pm_engine_shard_finish_il2p, rank after rank");
		if (e->h_A0.size() != (size_t)nc) e->h_A0.assign(nc, 0);
		for (int c = 0; c < nc; c++) {
			Il2pHand &h = hand[c];
			h.pos = 0; h.mode = 0; h.leak = 0;
			if (e->sharded && !plan.first && e->chains[c].d.codec_kind == PM_CODEC_IL2P) {
				h.pos = il2p_prev[c].pos - e->h_A0[c];
				h.mode = il2p_prev[c].mode;
				h.leak = il2p_prev[c].leak;
				// the walk restarts inside the previous shard's bits: everything it looks at has to lie in the hand-off tail
				const long long vf = e->h_valid_from.size() == (size_t)nc ? e->h_valid_from[c] : 0;
				const long long need = h.mode == 1 ? h.pos - 8 : h.pos - 63;
				if (h.mode > 2 || h.mode == 0 || need < vf)
					return fail(e, PM_ERR_STATE, "chain %d: an IL2P frame reaches back past the %d-bit hand-off tail", c, plan.tail_bits);
			}
		}
		CK(cudaMemcpyAsync(e->d_il2p_hand.p, hand.data(), nc * sizeof(Il2pHand), cudaMemcpyHostToDevice, e->st));
	}
	if (e->sharded && !plan.first && plan.tail_bits > 0) {
		if (!tail_in) return fail(e, PM_ERR_ARG, "shard_finish: the previous shard's tail is required");
		CK(cudaMemcpyAsync(e->d_tail.p, tail_in, (size_t)nc * (plan.tail_bits / 32) * sizeof(uint32_t),
			cudaMemcpyHostToDevice, e->st));
		ce = pm_launch_tail_inject(e->d_bits_raw.p, e->bits_stride, e->d_shardbits.p, nc, plan.tail_bits / 32,
			e->d_tail.p, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "tail inject launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches++;
	}
	if (!e->skip_lfsr) {
		ce = pm_launch_lfsr(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_raw.p, e->d_bits_lfsr.p, e->bits_stride, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "lfsr launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches++;
	}
	ce = pm_launch_ax25(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_lfsr.p, e->bits_stride, e->d_blk_count.p,
		e->d_blk_base.p, e->d_flag_totals.p, e->d_flag_pos.p, e->flag_stride, e->d_byte_addr.p, e->addr_stride,
		e->d_scratch.p, e->scratch_stride, e->d_gaps.p, e->flag_stride, e->d_shardbits.p, e->sharded ? 0 : 1,
		e->d_gap_cand.p, e->d_gap_ncand.p, e->st);
	if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "ax25 launch failed: %s", cudaGetErrorString(ce));
	e->stats.kernel_launches += e->sharded ? 4 : 5;
	if (e->has_il2p) {
		ce = pm_launch_il2p(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_lfsr.p, e->bits_stride, e->d_flag_pos.p,
			e->flag_stride, e->d_flag_totals.p, e->il2p_cand_cap, e->d_il2p_slots.p,
			(long long)(e->il2p_cand_cap + 1) * IL2P_SLOT, e->d_il2p_res.p, e->d_byte_addr.p, e->addr_stride,
			e->d_scratch.p, e->scratch_stride, e->d_gaps.p, e->flag_stride, e->d_shardbits.p, e->d_il2p_hand.p,
			e->d_il2p_hand.p + nc, e->st);
		if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "il2p launch failed: %s", cudaGetErrorString(ce));
		e->stats.kernel_launches += 2;
		CK(cudaMemcpyAsync(hand.data(), e->d_il2p_hand.p + nc, nc * sizeof(Il2pHand), cudaMemcpyDeviceToHost, e->st));
	}
	ce = pm_launch_packets(nc, e->d_cc.p, e->d_gaps.p, e->flag_stride, e->d_recs.p, e->d_rec_src.p,
		e->d_recs.n, e->d_totals.p, e->d_scratch.p, e->scratch_stride, e->d_arena.p, e->d_arena.n, e->sample_base,
		e->st);
	if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "packet launch failed: %s", cudaGetErrorString(ce));
	e->stats.kernel_launches += 3;
	CK(cudaEventRecord(e->ev[4], e->st));

	// results to host: ONE synchronisation.  The record count is not known yet, so as many records / packet bytes as the
	// last run produced (plus a quarter) are copied back blind with the counters; only a run that produced more
	// pays a second copy.
	pm_kt_mark("d2h results + host sync", e->st);
	CK(cudaMemcpyAsync(e->h_totals, e->d_totals.p, sizeof(PacketTotals), cudaMemcpyDeviceToHost, e->st));
	CK(cudaMemcpyAsync(e->h_counters, e->d_counters.p, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, e->st));
	e->h_cc.resize(nc);
	CK(cudaMemcpyAsync(e->h_cc.data(), e->d_cc.p, nc * sizeof(ChainCounters), cudaMemcpyDeviceToHost, e->st));
	const size_t spec_np = std::min(e->spec_recs, e->d_recs.n), spec_nb = std::min(e->spec_arena, e->d_arena.n);
	CK(e->h_recs.resize(spec_np));
	CK(e->h_arena.resize(spec_nb));
	if (spec_np) CK(cudaMemcpyAsync(e->h_recs.data(), e->d_recs.p, spec_np * sizeof(pm_packet_rec), cudaMemcpyDeviceToHost, e->st));
	if (spec_nb) CK(cudaMemcpyAsync(e->h_arena.data(), e->d_arena.p, spec_nb, cudaMemcpyDeviceToHost, e->st));
	CK(cudaEventRecord(e->ev[5], e->st));
	CK(cudaStreamSynchronize(e->st));
	if (e->has_il2p && il2p_out)
		for (int c = 0; c < nc; c++) {
			il2p_out[c].pos = hand[c].pos + e->h_A0[c];
			il2p_out[c].mode = hand[c].mode;
			il2p_out[c].leak = hand[c].leak;
		}
	if (e->h_counters[15] != 0)
		return fail(e, PM_ERR_CUDA, "tensor-core low-pass: a pipeline barrier timed out (internal error)");
	for (int c = 0; c < nc; c++) {
		if (e->h_cc[c].tail_short)
			return fail(e, PM_ERR_STATE, "chain %d: a frame closing in this shard reaches back past the %d-bit hand-off tail",
				c, plan.tail_bits);
		if (e->h_cc[c].seq_needed == 2) {
			e->grow_hint = 2;
			return fail(e, PM_ERR_CAPACITY, "chain %d: IL2P sync candidates overflowed the candidate list", c);
		}
		// a gap long enough to overflow max_packet_length (ax25.py:46-51) needs the sequential replay from a known
		// state: the shards recover through pm_engine_shard_export + pm_engine_decode_stream (sharded.py)
		if (e->sharded && e->h_cc[c].seq_needed)
			return fail(e, PM_ERR_STATE, "chain %d: needs the sequential AX.25 replay (recover from the gathered bitstream)", c);
	}
	const unsigned long long np = e->h_totals->n_packets, nb = e->h_totals->n_bytes;
	if (np > e->d_recs.n || nb > e->d_arena.n) {
		e->grow_hint = 1;
		return fail(e, PM_ERR_CAPACITY, "packet buffers too small (%llu records, %llu bytes)", np, nb);
	}
	if (np > spec_np || nb > spec_nb) {
		// more than was copied blind: fetch everything again (the pinned buffers may have to grow)
		CK(e->h_recs.resize(np));
		CK(e->h_arena.resize(nb));
		pm_kt_mark("d2h records (second copy)", e->st);
		if (np) CK(cudaMemcpyAsync(e->h_recs.data(), e->d_recs.p, np * sizeof(pm_packet_rec), cudaMemcpyDeviceToHost, e->st));
		if (nb) CK(cudaMemcpyAsync(e->h_arena.data(), e->d_arena.p, nb, cudaMemcpyDeviceToHost, e->st));
		CK(cudaEventRecord(e->ev[5], e->st));
		CK(cudaStreamSynchronize(e->st));
	} else {
		CK(e->h_recs.resize(np));           // shrinks the logical size only: the data stays
		CK(e->h_arena.resize(nb));
	}
	e->spec_recs = np + np / 4 + 256;
	e->spec_arena = nb + nb / 4 + 16384;
	e->stats.d2h_bytes = (int64_t)(np * sizeof(pm_packet_rec) + nb + sizeof(PacketTotals) + nc * sizeof(ChainCounters) + 8);
	e->stats.n_packets = (int64_t)np;
	e->stats.n_stream_bits = 0;
	for (auto &c : e->h_cc) e->stats.n_stream_bits += c.nbits;
	float ms = 0;
	auto el = [&](int a, int b) { ms = 0; cudaEventElapsedTime(&ms, e->ev[a], e->ev[b]); return (double)ms; };
	e->stats.total_ms = el(0, 5);
	e->stats.front_ms = el(0, 1);
	e->stats.fixup_ms = el(1, 2);
	e->stats.slicer_ms = el(2, 3);
	e->stats.bits_ms = el(3, 4);
	e->stats.d2h_ms = el(4, 5);
	e->phase = 0;
	e->have_run = true;
	return PM_OK;
}

// elapsed time between consecutive marks of a timing pass, summed per name in order of first appearance
static void kt_finish(pm_engine *e)
{
	cudaStreamSynchronize(e->st);
	std::vector<std::string> names;
	std::vector<double> ms;
	std::vector<int> cnt;
	auto &m = e->kt.marks;
	for (size_t i = 0; i + 1 < m.size(); i++) {
		float dt = 0;
		if (cudaEventElapsedTime(&dt, m[i].ev, m[i + 1].ev) != cudaSuccess) { cudaGetLastError(); continue; }
		size_t k = 0;
		while (k < names.size() && names[k] != m[i].name) k++;
		if (k == names.size()) { names.push_back(m[i].name); ms.push_back(0); cnt.push_back(0); }
		ms[k] += dt; cnt[k]++;
	}
	e->kt_report.clear();
	char line[160];
	for (size_t k = 0; k < names.size(); k++) {
		snprintf(line, sizeof(line), "%s\t%d\t%.6f\n", names[k].c_str(), cnt[k], ms[k]);
		e->kt_report += line;
	}
}

// a finish that ran out of room says what to grow (grow_hint); true when the run should be repeated
static bool grow_after_capacity(pm_engine *e, int rc)
{
	if (rc != PM_ERR_CAPACITY || !e->grow_hint) return false;
	if (e->grow_hint == 1) e->rec_scale *= 2.0;
	else e->il2p_cand_scale *= 4.0;
	e->grow_hint = 0;
	return e->rec_scale <= 64.0 && e->il2p_cand_scale <= 4096.0;
}

static int run_impl(pm_engine *e, const int16_t *audio, long long n, bool on_host, bool batched = false)
{
	if (e && !batched) {
		e->batch_stride = 0;
		for (size_t c = 0; c < e->chains.size(); c++)
			if (e->chains[c].d.recording != 0)
				return fail(e, PM_ERR_STATE, "chain %zu decodes recording %d of a batch: use pm_engine_run_batch", c, e->chains[c].d.recording);
	}
	pm_shard_plan plan;
	memset(&plan, 0, sizeof(plan));
	plan.first = plan.last = 1;
	for (;;) {
		if (e) e->grow_hint = 0;
		const bool timing = e && e->opt_kernel_times;
		if (timing) { e->kt.reset(); g_kt = &e->kt; }
		int rc = shard_begin_impl(e, audio, n, on_host, plan, false);
		if (rc == PM_OK) rc = shard_gather_impl(e, nullptr, nullptr);
		if (rc == PM_OK) rc = shard_finish_impl(e, nullptr);
		if (e && e->fast_pending && (rc == PM_OK || rc == PM_ERR_CAPACITY)) {
			// the one synchronisation of the run is behind us: now look at what the blind part assumed
			const unsigned int flagged = e->h_counters[0];
			const bool converged = slicer_fast_converged(e);
			e->stats.guard_flagged = flagged;
			if (flagged > e->guard_cap) {
				// guard list overflowed (samples went without their float64 re-evaluation): grow it, run again
				e->guard_cap = flagged + flagged / 4 + 1024;
				if (timing) g_kt = nullptr;
				continue;
			}
			if (!converged && rc == PM_OK) {
				// rare (stretches without zero crossings): finish the verify/repair passes with the host in the loop,
				// then redo the bit-level stages on the now exact roll-over mask
				rc = slicer_converge(e);
				if (rc == PM_OK) { e->phase = 1; rc = shard_gather_impl(e, nullptr, nullptr); }
				if (rc == PM_OK) rc = shard_finish_impl(e, nullptr);
				e->stats.guard_flagged = flagged;
			}
		}
		if (timing) {
			pm_kt_mark("end", e->st);
			g_kt = nullptr;
			kt_finish(e);
		}
		if (!grow_after_capacity(e, rc)) return rc;      // packet buffers / IL2P candidate list too small: grow, run again
	}
}

extern "C" int pm_engine_run(pm_engine *e, const int16_t *audio_host, int64_t n_samples)
{
	return run_impl(e, audio_host, n_samples, true);
}

extern "C" int pm_engine_run_device(pm_engine *e, const int16_t *audio_dev, int64_t n_samples)
{
	return run_impl(e, audio_dev, n_samples, false);
}

// R recordings x their chains in one run: the rows travel in one copy, then every stage runs over all chains at once
extern "C" int pm_engine_run_batch(pm_engine *e, const int16_t *audio_host, int64_t row_stride, const int64_t *n_samples,
                                   int32_t n_recordings)
{
	if (!e || !audio_host || !n_samples || n_recordings <= 0 || row_stride <= 0)
		return fail(e, PM_ERR_ARG, "run_batch: bad arguments");
	cudaSetDevice(e->device);
	long long n_max = 0;
	e->batch_n.assign(n_samples, n_samples + n_recordings);
	for (int r = 0; r < n_recordings; r++) {
		if (n_samples[r] <= 0 || n_samples[r] > row_stride) return fail(e, PM_ERR_ARG, "run_batch: recording %d has %lld samples (row stride %lld)", r, (long long)n_samples[r], (long long)row_stride);
		n_max = std::max<long long>(n_max, n_samples[r]);
	}
	for (size_t c = 0; c < e->chains.size(); c++)
		if (e->chains[c].d.recording >= n_recordings)
			return fail(e, PM_ERR_ARG, "run_batch: chain %zu names recording %d of %d", c, e->chains[c].d.recording, n_recordings);
	const size_t total = (size_t)row_stride * n_recordings;
	if (total >= (1ull << 32)) return fail(e, PM_ERR_ARG, "run_batch: more than 2^32 samples in a batch");
	CK(e->d_audio.ensure(total + 64));
	CK(cudaMemcpyAsync(e->d_audio.p, audio_host, total * sizeof(int16_t), cudaMemcpyHostToDevice, e->st));
	e->batch_stride = row_stride;
	const int rc = run_impl(e, e->d_audio.p, n_max, false, true);
	e->stats.h2d_bytes = (int64_t)total * 2;
	return rc;
}

extern "C" int pm_engine_shard_begin(pm_engine *e, const int16_t *audio, int64_t n_samples, int32_t audio_on_device,
                                     const pm_shard_plan *plan, pm_shard_state *out)
{
	if (!e || !plan || !out) return fail(e, PM_ERR_ARG, "shard_begin: bad arguments");
	int rc = shard_begin_impl(e, audio, n_samples, !audio_on_device, *plan, true);
	if (rc != PM_OK) return rc;
	memcpy(out, e->shard_out.data(), e->shard_out.size() * sizeof(pm_shard_state));
	return PM_OK;
}

extern "C" int pm_engine_shard_handoff(pm_engine *e, const pm_shard_state *prev, pm_shard_state *out, int32_t *changed)
{
	if (!e || !out || !changed) return fail(e, PM_ERR_ARG, "shard_handoff: bad arguments");
	if (e->phase != 1 || !e->sharded) return fail(e, PM_ERR_STATE, "shard_handoff: call shard_begin first");
	cudaSetDevice(e->device);
	const int nc = (int)e->chains.size();
	*changed = 0;
	if (!e->plan.first) {
		if (!prev) return fail(e, PM_ERR_ARG, "shard_handoff: the previous shard's state is required");
		bool mismatch = false;
		for (int c = 0; c < nc; c++) {
			SegState want;
			want.clock = prev[c].end_clock; want.last = prev[c].end_last; want.last_q = prev[c].end_last_q;
			if (memcmp(&want.clock, &e->h_init[c].clock, sizeof(double)) || want.last != e->h_init[c].last ||
			    want.last_q != e->h_init[c].last_q)
				mismatch = true;
			e->h_init[c] = want;
		}
		if (mismatch) {
			const std::vector<pm_shard_state> before = e->shard_out;
			e->geom.k_init = e->k0;      // the true state enters at the first own segment; the history before it is final
			CK(cudaMemcpyAsync(e->d_init.p, e->h_init.data(), nc * sizeof(SegState), cudaMemcpyHostToDevice, e->st));
			e->up_init.clear();
			int rc = slicer_converge(e);
			if (rc != PM_OK) return rc;
			rc = read_shard_states(e);
			if (rc != PM_OK) return rc;
			for (int c = 0; c < nc; c++)
				if (memcmp(&before[c].end_clock, &e->shard_out[c].end_clock, sizeof(double)) ||
				    before[c].end_last != e->shard_out[c].end_last || before[c].end_last_q != e->shard_out[c].end_last_q ||
				    before[c].n_symbols != e->shard_out[c].n_symbols)
					*changed = 1;
		}
	}
	memcpy(out, e->shard_out.data(), e->shard_out.size() * sizeof(pm_shard_state));
	return PM_OK;
}

extern "C" int pm_engine_shard_gather(pm_engine *e, const int64_t *symbols_before, uint32_t *tail_out)
{
	if (!e) return PM_ERR_ARG;
	cudaSetDevice(e->device);
	return shard_gather_impl(e, symbols_before, tail_out);
}

extern "C" int pm_engine_shard_finish(pm_engine *e, const uint32_t *tail_in)
{
	if (!e) return PM_ERR_ARG;
	cudaSetDevice(e->device);
	return shard_finish_impl(e, tail_in);
}


// ---------------------------------------------------------------------------
// Per-stage entry points: the reference's duck-typed blocks one at a time (chain_execute.py:32-47) --
// slicer.slice(soft), stream.stream_unscramble_8bit(AddressedData[]), codec.decode(AddressedData[]) -- on the same
// kernels as the whole-chain run.  They work on ONE chain of the loaded table; the other chains see nothing.
// ---------------------------------------------------------------------------
extern "C" cudaError_t pm_launch_soft_signs(const double *, long long, uint32_t *, cudaStream_t);

static int stage_chain_ok(pm_engine *e, int32_t chain)
{
	if (!e) return PM_ERR_ARG;
	if (chain < 0 || chain >= (int)e->chains.size()) return fail(e, PM_ERR_ARG, "chain %d out of range", chain);
	cudaSetDevice(e->device);
	e->phase = 0;
	e->have_run = false;
	memset(&e->stats, 0, sizeof(e->stats));
	return PM_OK;
}

static int fetch_counters(pm_engine *e)
{
	const int nc = (int)e->chains.size();
	e->h_cc.resize(nc);
	CK(cudaMemcpyAsync(e->h_cc.data(), e->d_cc.p, nc * sizeof(ChainCounters), cudaMemcpyDeviceToHost, e->st));
	CK(cudaStreamSynchronize(e->st));
	return PM_OK;
}

extern "C" int pm_engine_slice_soft(pm_engine *e, int32_t chain, const double *soft_i, const double *soft_q, int64_t n)
{
	int rc = stage_chain_ok(e, chain);
	if (rc != PM_OK) return rc;
	HostChain &hc = e->chains[chain];
	const bool quad = hc.d.slicer_kind == PM_SLICER_QUADRATURE;
	if (n < 0 || (n > 0 && !soft_i) || (quad && n > 0 && !soft_q))
		return fail(e, PM_ERR_ARG, "slice_soft: soft values missing (the quadrature slicer needs I and Q, slicer.py:198-199)");
	if (n == 0) {                         // slicing nothing gives nothing (slicer.py:75: the loop body never runs)
		ChainCounters zero;
		memset(&zero, 0, sizeof(zero));
		e->h_cc.assign(e->chains.size(), zero);
		CK(e->h_recs.resize(0));
		CK(e->h_arena.resize(0));
		e->have_run = true;
		return PM_OK;
	}
	pm_shard_plan plan;
	memset(&plan, 0, sizeof(plan));
	plan.first = plan.last = 1;
	rc = prepare_run(e, (long long)n + hc.trim, plan, false, chain);     // this chain's valid soft samples = n
	if (rc != PM_OK) return rc;
	const int nc = (int)e->chains.size();
	DevBuf<double> d_x;
	CK(d_x.ensure((size_t)std::max<int64_t>(n, 1)));
	for (int comp = 0; comp < (quad ? 2 : 1); comp++) {
		const int row = comp ? hc.sign_q_row : chain;
		if (n) CK(cudaMemcpyAsync(d_x.p, comp ? soft_q : soft_i, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->st));
		cudaError_t ce = pm_launch_soft_signs(d_x.p, n, e->d_sign.p + (size_t)row * e->sign_stride, e->st);
		if (ce != cudaSuccess) { d_x.release(); return fail(e, PM_ERR_CUDA, "sign launch failed: %s", cudaGetErrorString(ce)); }
		CK(cudaStreamSynchronize(e->st));       // the caller's buffer (pageable) has been read
	}
	d_x.release();
	e->E_cur = e->d_E0.p; e->E_alt = e->d_E1.p;
	cudaError_t ce = pm_launch_slicer_segments(e->d_slicer.p, nc, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride,
		e->d_S.p, e->E_cur, e->d_chk.p, e->d_init.p, e->geom, e->st);
	if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "slicer launch failed: %s", cudaGetErrorString(ce));
	e->stats.kernel_launches += 2;
	rc = slicer_converge(e);
	if (rc != PM_OK) return rc;
	e->phase = 1;
	rc = shard_gather_impl(e, nullptr, nullptr);
	if (rc != PM_OK) return rc;
	rc = fetch_counters(e);
	if (rc != PM_OK) return rc;
	e->phase = 0;
	e->have_run = true;
	return PM_OK;
}

// AddressedData list of one chain -> the device's packed stream (first bit of a byte = its MSB, bits.cu) + byte addresses
static int load_stream(pm_engine *e, int32_t chain, const uint8_t *bytes, const int64_t *addresses, int64_t n, bool descrambled)
{
	int rc = stage_chain_ok(e, chain);
	if (rc != PM_OK) return rc;
	if (n < 0 || (n > 0 && (!bytes || !addresses))) return fail(e, PM_ERR_ARG, "stream: bytes/addresses missing");
	if (n >= (1ll << 28)) return fail(e, PM_ERR_ARG, "stream too long");
	pm_shard_plan plan;
	memset(&plan, 0, sizeof(plan));
	plan.first = plan.last = 1;
	rc = prepare_run(e, (long long)e->chains[chain].trim + 64, plan, false, chain, 8 * (long long)n + 64);
	if (rc != PM_OK) return rc;
	const int nc = (int)e->chains.size();
	std::vector<uint32_t> words((size_t)(n + 3) / 4 + 1, 0u), addr((size_t)n + 1, 0u);
	for (int64_t b = 0; b < n; b++) {
		uint32_t v = bytes[b];
		v = ((v & 0xF0u) >> 4) | ((v & 0x0Fu) << 4);
		v = ((v & 0xCCu) >> 2) | ((v & 0x33u) << 2);
		v = ((v & 0xAAu) >> 1) | ((v & 0x55u) << 1);
		words[(size_t)b >> 2] |= v << ((b & 3) * 8);
		if (addresses[b] < 0 || addresses[b] > 0xFFFFFFFFll) return fail(e, PM_ERR_ARG, "stream: address %lld out of range", (long long)addresses[b]);
		addr[(size_t)b] = (uint32_t)addresses[b];
	}
	uint32_t *row = (descrambled ? e->d_bits_lfsr.p : e->d_bits_raw.p) + (size_t)chain * e->bits_stride;
	CK(cudaMemsetAsync(row, 0, (size_t)e->bits_stride * sizeof(uint32_t), e->st));
	CK(cudaMemcpyAsync(row, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->st));
	CK(cudaMemcpyAsync(e->d_byte_addr.p + (size_t)chain * e->addr_stride, addr.data(), (size_t)n * sizeof(uint32_t) + 4,
		cudaMemcpyHostToDevice, e->st));
	std::vector<ChainCounters> cc(nc);
	std::vector<ShardBits> sb(nc);
	memset(cc.data(), 0, nc * sizeof(ChainCounters));
	memset(sb.data(), 0, nc * sizeof(ShardBits));
	for (int c = 0; c < nc; c++) { sb[c].first = 1; sb[c].own_hi = 0x7fffffffffffffffll; }
	cc[chain].nbits = 8 * (long long)n;
	cc[chain].nbytes = n;
	CK(cudaMemcpyAsync(e->d_cc.p, cc.data(), nc * sizeof(ChainCounters), cudaMemcpyHostToDevice, e->st));
	CK(cudaMemcpyAsync(e->d_shardbits.p, sb.data(), nc * sizeof(ShardBits), cudaMemcpyHostToDevice, e->st));
	e->up_sb.clear();
	CK(cudaEventRecord(e->ev[0], e->st));
	for (int i = 1; i <= 3; i++) CK(cudaEventRecord(e->ev[i], e->st));
	CK(cudaStreamSynchronize(e->st));           // the staging vectors go out of scope
	e->h_A0.assign(nc, 0);
	e->h_valid_from.assign(nc, 0);
	e->phase = 2;
	return PM_OK;
}

extern "C" int pm_engine_unscramble_stream(pm_engine *e, int32_t chain, const uint8_t *bytes, const int64_t *addresses, int64_t n)
{
	int rc = load_stream(e, chain, bytes, addresses, n, false);
	if (rc != PM_OK) return rc;
	const int nc = (int)e->chains.size();
	cudaError_t ce = pm_launch_lfsr(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_raw.p, e->d_bits_lfsr.p, e->bits_stride, e->st);
	if (ce != cudaSuccess) return fail(e, PM_ERR_CUDA, "lfsr launch failed: %s", cudaGetErrorString(ce));
	e->stats.kernel_launches++;
	rc = fetch_counters(e);
	if (rc != PM_OK) return rc;
	e->phase = 0;
	e->have_run = true;
	e->h_recs.resize(0);
	e->h_arena.resize(0);
	return PM_OK;
}

extern "C" int pm_engine_decode_stream(pm_engine *e, int32_t chain, const uint8_t *bytes, const int64_t *addresses, int64_t n)
{
	for (;;) {
		if (e) e->grow_hint = 0;
		int rc = load_stream(e, chain, bytes, addresses, n, true);
		if (rc != PM_OK) return rc;
		e->skip_lfsr = true;
		rc = shard_finish_impl(e, nullptr);
		e->skip_lfsr = false;
		if (!grow_after_capacity(e, rc)) return rc;
	}
}

// What a shard holds of one chain's sliced stream after shard_gather (also after a finish that failed): the local
// packed bits and byte addresses, for the recovery path of sharded.py.  info4 = {local stream bits, local position of
// the first own bit, own bits, sample_base}.  With bits == NULL only info4 is filled.
extern "C" int pm_engine_shard_export(pm_engine *e, int32_t chain, uint32_t *bits, int64_t cap_words, uint32_t *byte_addr,
                                      int64_t cap_bytes, int64_t *info4)
{
	if (!e || !info4) return PM_ERR_ARG;
	if (chain < 0 || chain >= (int)e->chains.size()) return fail(e, PM_ERR_ARG, "chain %d out of range", chain);
	if (e->phase != 2 && !(e->have_run && e->sharded)) return fail(e, PM_ERR_STATE, "shard_export: call shard_gather first");
	cudaSetDevice(e->device);
	ChainCounters cc;
	ShardBits sb;
	CK(cudaStreamSynchronize(e->st));
	CK(cudaMemcpy(&cc, e->d_cc.p + chain, sizeof(cc), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(&sb, e->d_shardbits.p + chain, sizeof(sb), cudaMemcpyDeviceToHost));
	const long long nbits = cc.nbits;
	info4[0] = nbits;
	info4[1] = sb.bit_off;
	info4[2] = (sb.own_hi == 0x7fffffffffffffffll) ? nbits - sb.bit_off : sb.own_hi - sb.bit_off;
	info4[3] = e->sample_base;
	if (!bits && !byte_addr) return PM_OK;
	const long long nw = (nbits + 31) / 32, nby = (nbits + 7) / 8;
	if (!bits || !byte_addr || cap_words < nw || cap_bytes < nby) return fail(e, PM_ERR_CAPACITY, "shard_export: buffers too small");
	if (nw) CK(cudaMemcpy(bits, e->d_bits_raw.p + (size_t)chain * e->bits_stride, (size_t)nw * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	if (nby) CK(cudaMemcpy(byte_addr, e->d_byte_addr.p + (size_t)chain * e->addr_stride, (size_t)nby * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return PM_OK;
}

// ---------------------------------------------------------------------------
// Shard link: the hand-off done on the devices over peer memory (csrc/link.cu)
// ---------------------------------------------------------------------------
static long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

extern "C" int pm_engine_link_create(pm_engine *e, int32_t rank, int32_t world, int32_t tail_bits, int64_t max_samples,
                                     void *ipc_handle_out, void **base_out)
{
	if (!e || rank < 0 || world < 1 || rank >= world || world > LINK_MAX_WORLD || tail_bits < 0 || tail_bits % 32 || max_samples <= 0)
		return fail(e, PM_ERR_ARG, "link_create: bad arguments");
	if (world > 1 && tail_bits < 128) return fail(e, PM_ERR_ARG, "tail_bits must be at least 128");
	const int nc = (int)e->chains.size();
	if (nc == 0) return fail(e, PM_ERR_STATE, "load chains before creating the link");
	if ((long long)nc * world > 8192) return fail(e, PM_ERR_CAPACITY, "too many chains x ranks for the link");
	cudaSetDevice(e->device);
	// worst-case records + packet bytes of one rank (the bounds prepare_run uses)
	long long rec_cap = 16, arena_cap = 64;
	for (auto &hc : e->chains) {
		if (hc.d.slicer_kind != PM_SLICER_BINARY || (hc.p64 && hc.d.modem_kind != PM_MODEM_AFSK) ||
		    (hc.d.codec_kind != PM_CODEC_AX25 && hc.d.codec_kind != PM_CODEC_IL2P))
			return fail(e, PM_ERR_UNSUPPORTED, "only binary-slicer AX.25 / IL2P chains without a carrier loop can be sharded");
		const double thr = hc.d.slicer_sample_rate / hc.d.symbol_rate / 2.0 - 0.5;
		const long long min_gap = std::max<long long>(1, (long long)std::ceil(thr));
		const long long mb = (max_samples / min_gap + 8) + 64 + tail_bits;
		rec_cap += mb / (hc.d.codec_kind == PM_CODEC_IL2P ? 144 : 152) + 4;
		arena_cap += mb / 8 + 64;
	}
	LinkGeom &G = e->lg;
	memset(&G, 0, sizeof(G));
	G.rank = rank; G.world = world; G.nc = nc; G.tail_words = tail_bits / 32;
	G.rec_region = align_up(rec_cap * (long long)sizeof(pm_packet_rec) + arena_cap, 256);
	long long off = 0;
	G.off_states = off; off = align_up(off + (long long)world * nc * sizeof(pm_shard_state), 256);
	G.off_sflag = off; off = align_up(off + 4ll * world, 256);
	G.off_tail = off; off = align_up(off + 4ll * nc * std::max(1, G.tail_words), 256);
	G.off_tflag = off; off += 256;
	G.off_rflag = off; off = align_up(off + 4ll * world, 256);
	G.off_rhdr = off; off = align_up(off + 16ll * world, 256);
	G.off_il2p = off; off = align_up(off + (long long)nc * sizeof(pm_il2p_state), 256);
	G.off_iflag = off; off += 256;
	G.off_rdata = off; off += (long long)world * G.rec_region;
	G.slot_bytes = align_up(off, 4096);
	CK(e->d_link.ensure((size_t)(2 * G.slot_bytes)));
	CK(cudaMemset(e->d_link.p, 0, (size_t)(2 * G.slot_bytes)));
	CK(e->d_link_status.ensure(8));
	CK(e->d_link_A0.ensure((size_t)nc));
	CK(e->d_link_lb.ensure((size_t)world * (nc + 1)));
	CK(e->d_link_obase.ensure((size_t)world * (nc + 1) + world));
	CK(e->d_mrecs.ensure((size_t)(rec_cap * world)));
	CK(e->d_marena.ensure((size_t)(arena_cap * world)));
	CK(e->d_mtotals.ensure(1));
	if (!e->h_link_status) CK(cudaHostAlloc((void **)&e->h_link_status, 64, cudaHostAllocDefault));
	if (!e->h_mtotals) CK(cudaHostAlloc((void **)&e->h_mtotals, sizeof(PacketTotals), cudaHostAllocDefault));
	e->link_tail_bits = tail_bits;
	e->link_epoch = 0;
	e->link_on = false;
	if (ipc_handle_out) {
		cudaIpcMemHandle_t h;
		CK(cudaIpcGetMemHandle(&h, e->d_link.p));
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
		memcpy(ipc_handle_out, &h, sizeof(h));
	}
	if (base_out) *base_out = e->d_link.p;
	return PM_OK;
}

extern "C" int pm_engine_link_connect(pm_engine *e, const void *handles, int32_t use_ipc)
{
	if (!e || !handles || !e->d_link.p) return fail(e, PM_ERR_STATE, "link_connect: create the link first");
	cudaSetDevice(e->device);
	for (void *p : e->link_opened) cudaIpcCloseMemHandle(p);
	e->link_opened.clear();
	memset(&e->lp, 0, sizeof(e->lp));
	for (int q = 0; q < e->lg.world; q++) {
		if (q == e->lg.rank) { e->lp.base[q] = e->d_link.p; continue; }
		if (use_ipc) {
			cudaIpcMemHandle_t h;
			memcpy(&h, (const char *)handles + 64 * q, 64);
			void *p = nullptr;
			CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
			e->link_opened.push_back(p);
			e->lp.base[q] = (unsigned char *)p;
		} else {
			e->lp.base[q] = (unsigned char *)((void *const *)handles)[q];
		}
	}
	CK(pm_link_preload());
	e->link_on = true;
	return PM_OK;
}

#define CKL(call)                                                                                      \
	do {                                                                                               \
		cudaError_t _e = (call);                                                                       \
		if (_e != cudaSuccess) return fail(e, PM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
	} while (0)

extern "C" int pm_engine_run_linked_begin(pm_engine *e, const int16_t *audio, int64_t n_samples, int32_t audio_on_device,
                                          const pm_shard_plan *plan)
{
	if (!e || !plan) return fail(e, PM_ERR_ARG, "run_linked_begin: bad arguments");
	if (!e->link_on) return fail(e, PM_ERR_STATE, "run_linked_begin: connect the link first");
	const LinkGeom &G = e->lg;
	if (plan->tail_bits != e->link_tail_bits || (plan->first != 0) != (G.rank == 0) || (plan->last != 0) != (G.rank == G.world - 1))
		return fail(e, PM_ERR_ARG, "run_linked_begin: plan does not match the link (rank %d of %d, tail %d)", G.rank, G.world,
			e->link_tail_bits);
	int rc = shard_begin_impl(e, audio, n_samples, !audio_on_device, *plan, true, false);
	if (rc != PM_OK) return rc;
	const int nc = (int)e->chains.size();
	const unsigned int epoch = ++e->link_epoch;
	const int parity = (int)(epoch & 1u);
	cudaStream_t st = e->st;
	unsigned char *own = e->d_link.p;
	int *status = e->d_link_status.p;
	CK(cudaMemsetAsync(status, 0, 8 * sizeof(int), st));
	CKL(pm_launch_slicer_count(e->d_slicer.p, nc, e->d_mask.p, e->sign_stride, e->own_w0, e->own_w1, e->d_symcount.p, st));
	CKL(pm_link_push_states(G, e->lp, parity, epoch, e->d_S.p + e->k0, e->E_cur, e->geom.n_seg, e->k_end, e->d_symcount.p,
		e->d_counters.p, e->fast_passes, e->guard_cap, st));
	CKL(pm_link_wait_states(G, own, parity, epoch, e->d_bitchain.p, plan->first, plan->last, plan->tail_bits,
		e->d_shardbits.p, status, e->d_link_A0.p, st));
	e->up_sb.clear();                       // the placement was written on the device
	CKL(pm_launch_gather(e->d_bitchain.p, nc, e->d_cc.p, e->d_sign.p, e->sign_stride, e->d_mask.p, e->sign_stride, e->own_w0,
		std::max<long long>(1, e->end_w - e->own_w0), e->d_blk_count.p, e->d_blk_base.p, e->d_sym_totals.p, e->d_bits_raw.p,
		e->bits_stride, e->d_byte_addr.p, e->addr_stride, nullptr, e->d_shardbits.p, st));
	e->stats.kernel_launches += 7;
	if (!plan->last && G.tail_words > 0) {
		unsigned char *next = e->lp.base[G.rank + 1] + (long long)parity * G.slot_bytes;
		CKL(pm_launch_tail_extract(e->d_bits_raw.p, e->bits_stride, e->d_shardbits.p, nc, G.tail_words,
			(uint32_t *)(next + G.off_tail), st));
		CKL(pm_link_set_flag((unsigned int *)(next + G.off_tflag), epoch, st));
		e->stats.kernel_launches += 2;
	}
	if (!plan->first && G.tail_words > 0) {
		unsigned char *slot = own + (long long)parity * G.slot_bytes;
		CKL(pm_link_wait_flag((const unsigned int *)(slot + G.off_tflag), epoch, status, st));
		CKL(pm_launch_tail_inject(e->d_bits_raw.p, e->bits_stride, e->d_shardbits.p, nc, G.tail_words,
			(const uint32_t *)(slot + G.off_tail), st));
		e->stats.kernel_launches += 2;
	}
	CKL(pm_launch_lfsr(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_raw.p, e->d_bits_lfsr.p, e->bits_stride, st));
	CKL(pm_launch_ax25(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_lfsr.p, e->bits_stride, e->d_blk_count.p, e->d_blk_base.p,
		e->d_flag_totals.p, e->d_flag_pos.p, e->flag_stride, e->d_byte_addr.p, e->addr_stride, e->d_scratch.p,
		e->scratch_stride, e->d_gaps.p, e->flag_stride, e->d_shardbits.p, 0, e->d_gap_cand.p, e->d_gap_ncand.p, st));
	if (e->has_il2p) {
		// the IL2P walk crosses the boundary rank after rank: wait for where the previous rank's walk stands, decode,
		// tell the next rank (csrc/link.cu) -- no host in between
		CKL(pm_link_il2p_in(G, own, parity, epoch, plan->first, e->d_bitchain.p, e->d_shardbits.p, e->d_link_A0.p,
			e->d_il2p_hand.p, e->d_cc.p, status, st));
		CKL(pm_launch_il2p(e->d_bitchain.p, nc, e->d_cc.p, e->d_bits_lfsr.p, e->bits_stride, e->d_flag_pos.p,
			e->flag_stride, e->d_flag_totals.p, e->il2p_cand_cap, e->d_il2p_slots.p,
			(long long)(e->il2p_cand_cap + 1) * IL2P_SLOT, e->d_il2p_res.p, e->d_byte_addr.p, e->addr_stride,
			e->d_scratch.p, e->scratch_stride, e->d_gaps.p, e->flag_stride, e->d_shardbits.p, e->d_il2p_hand.p,
			e->d_il2p_hand.p + nc, st));
		if (!plan->last) CKL(pm_link_il2p_out(G, e->lp, parity, epoch, e->d_il2p_hand.p + nc, e->d_link_A0.p, st));
		e->stats.kernel_launches += plan->last ? 3 : 4;
	}
	CKL(pm_launch_packets(nc, e->d_cc.p, e->d_gaps.p, e->flag_stride, e->d_recs.p, e->d_rec_src.p, e->d_recs.n,
		e->d_totals.p, e->d_scratch.p, e->scratch_stride, e->d_arena.p, e->d_arena.n, e->sample_base, st));
	CK(cudaEventRecord(e->ev[4], st));
	CKL(pm_link_push_records(G, e->lp, parity, epoch, (const PacketRecDev *)e->d_recs.p, e->d_arena.p, e->d_totals.p, e->d_cc.p,
		status, st));
	CKL(pm_link_merge(G, own, parity, epoch, e->d_link_lb.p, e->d_link_obase.p, e->d_link_obase.p + (size_t)G.world * (nc + 1),
		e->d_mtotals.p, (PacketRecDev *)e->d_mrecs.p, e->d_mrecs.n, e->d_marena.p, e->d_marena.n, status, st));
	e->stats.kernel_launches += 13;
	CK(cudaMemcpyAsync(e->h_counters, e->d_counters.p, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(e->h_link_status, status, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(e->h_mtotals, e->d_mtotals.p, sizeof(PacketTotals), cudaMemcpyDeviceToHost, st));
	// (nothing here may block the host: a copy into pageable memory would wait for the stream, i.e. for the peers)
	e->phase = 3;
	return PM_OK;
}

extern "C" int pm_engine_run_linked_end(pm_engine *e, int32_t *verified)
{
	if (!e || !verified) return fail(e, PM_ERR_ARG, "run_linked_end: bad arguments");
	if (e->phase != 3) return fail(e, PM_ERR_STATE, "run_linked_end: call run_linked_begin first");
	cudaSetDevice(e->device);
	const int nc = (int)e->chains.size();
	CK(cudaStreamSynchronize(e->st));
	*verified = e->h_link_status[0];
	if (*verified && e->fast_pending) {
		// all hand-offs verified implies every rank's own passes converged and no guard list overflowed (a rank for
		// which that does not hold publishes n_symbols = -1, which un-verifies the run for everybody)
		e->stats.guard_flagged = e->h_counters[0];
		slicer_fast_converged(e);
	}
	if (e->h_link_status[1] != 0) {
		e->phase = 0;
		static const char *where[] = {"", "waiting for the slicer states of the other ranks", "waiting for the previous rank's bit tail",
			"pushing packet records (link region too small)", "waiting for the packet records of the other ranks",
			"another rank reported a failure"};
		const int w = e->h_link_status[2];
		return fail(e, e->h_link_status[1], "shard link failed while %s", (w >= 1 && w <= 5) ? where[w] : "running");
	}
	if (!*verified) {
		// some rank's speculated start state was wrong: every rank saw the same states and falls back to the
		// host-driven repair protocol (pm_engine_shard_states -> handoff -> gather -> finish)
		e->phase = 1;
		return PM_OK;
	}
	if (e->h_link_status[3] != 0) {
		// some rank could not finish its decode from what it holds (a frame longer than the hand-off tail, or the
		// max_packet_length overflow of ax25.py:46-51): every rank was told with the records, all of them export
		// their bits (pm_engine_shard_export) and decode the gathered stream (pm_engine_decode_stream)
		*verified = 2;
		e->phase = 2;
		return PM_OK;
	}
	const unsigned long long np = e->h_mtotals->n_packets, nb = e->h_mtotals->n_bytes;
	if (np > e->d_mrecs.n || nb > e->d_marena.n)
		return fail(e, PM_ERR_CAPACITY, "merged packet buffers too small (%llu records, %llu bytes)", np, nb);
	CK(e->h_recs.resize(np));
	CK(e->h_arena.resize(nb));
	e->h_cc.resize(nc);
	CK(cudaMemcpyAsync(e->h_cc.data(), e->d_cc.p, nc * sizeof(ChainCounters), cudaMemcpyDeviceToHost, e->st));
	if (np) CK(cudaMemcpyAsync(e->h_recs.data(), e->d_mrecs.p, np * sizeof(pm_packet_rec), cudaMemcpyDeviceToHost, e->st));
	if (nb) CK(cudaMemcpyAsync(e->h_arena.data(), e->d_marena.p, nb, cudaMemcpyDeviceToHost, e->st));
	CK(cudaEventRecord(e->ev[5], e->st));
	CK(cudaStreamSynchronize(e->st));
	e->stats.d2h_bytes = (int64_t)(np * sizeof(pm_packet_rec) + nb + sizeof(PacketTotals) + nc * sizeof(ChainCounters) + 16);
	e->stats.n_packets = (int64_t)np;
	e->stats.n_stream_bits = 0;
	for (auto &c : e->h_cc) e->stats.n_stream_bits += c.nbits;
	float ms = 0;
	auto el = [&](int a, int b) { ms = 0; cudaEventElapsedTime(&ms, e->ev[a], e->ev[b]); return (double)ms; };
	e->stats.total_ms = el(0, 5);
	e->stats.front_ms = el(0, 1);
	e->stats.fixup_ms = el(1, 2);
	e->stats.slicer_ms = el(2, 3);
	e->stats.bits_ms = el(3, 4);
	e->stats.d2h_ms = el(4, 5);
	e->phase = 0;
	e->have_run = true;
	return PM_OK;
}

// the states shard_begin reports, for a rank that began through run_linked_begin and has to fall back
extern "C" int pm_engine_shard_states(pm_engine *e, pm_shard_state *out)
{
	if (!e || !out) return fail(e, PM_ERR_ARG, "shard_states: bad arguments");
	if (e->phase != 1 || !e->sharded) return fail(e, PM_ERR_STATE, "shard_states: no sharded run in progress");
	cudaSetDevice(e->device);
	if (e->fast_pending) {
		// the linked run was enqueued blind (run_linked_begin): settle what it left open before the host-driven protocol
		// takes over -- a guard list that overflowed means starting over, verify passes that did not converge are finished
		const unsigned int flagged = e->h_counters[0];
		const bool converged = slicer_fast_converged(e);
		e->stats.guard_flagged = flagged;
		if (flagged > e->guard_cap) {
			e->guard_cap = flagged + flagged / 4 + 1024;
			const pm_shard_plan plan = e->plan;
			int rc = begin_once(e, e->run_audio, e->n_samples, false, plan, true, false);
			if (rc != PM_OK) return rc;
			e->phase = 1;
		} else if (!converged) {
			int rc = slicer_converge(e);
			if (rc != PM_OK) return rc;
		}
	}
	int rc = fetch_shard_states(e);
	if (rc != PM_OK) return rc;
	memcpy(out, e->shard_out.data(), e->shard_out.size() * sizeof(pm_shard_state));
	return PM_OK;
}

extern "C" int pm_engine_shard_finish_il2p(pm_engine *e, const uint32_t *tail_in, const pm_il2p_state *prev, pm_il2p_state *out)
{
	if (!e || !out) return fail(e, PM_ERR_ARG, "shard_finish_il2p: bad arguments");
	cudaSetDevice(e->device);
	return shard_finish_impl(e, tail_in, prev, out);
}

extern "C" int64_t pm_engine_num_packets(const pm_engine *e) { return (e && e->have_run) ? (int64_t)e->h_recs.size() : -1; }
extern "C" int64_t pm_engine_arena_bytes(const pm_engine *e) { return (e && e->have_run) ? (int64_t)e->h_arena.size() : -1; }

extern "C" int pm_engine_get_packets(const pm_engine *e, pm_packet_rec *recs, int64_t rec_cap, uint8_t *arena,
                                     int64_t arena_cap)
{
	if (!e || !e->have_run) return PM_ERR_STATE;
	if ((int64_t)e->h_recs.size() > rec_cap || (int64_t)e->h_arena.size() > arena_cap) return PM_ERR_CAPACITY;
	if (!e->h_recs.empty()) memcpy(recs, e->h_recs.data(), e->h_recs.size() * sizeof(pm_packet_rec));
	if (!e->h_arena.empty()) memcpy(arena, e->h_arena.data(), e->h_arena.size());
	return PM_OK;
}

extern "C" int64_t pm_engine_soft_len(const pm_engine *e, int32_t chain)
{
	if (!e || !e->have_run || chain < 0 || chain >= (int)e->chains.size()) return -1;
	return std::max<long long>(0, chain_n(e, chain, e->n_samples) - e->chains[chain].trim);
}

extern "C" int pm_engine_get_soft(const pm_engine *ce, int32_t chain, int32_t component, float *out, int64_t cap)
{
	pm_engine *e = const_cast<pm_engine *>(ce);
	if (!e || !e->have_run) return PM_ERR_STATE;
	if (!e->opt_keep_soft || !e->d_soft.p) return fail(e, PM_ERR_STATE, "soft values were not kept (option keep_soft)");
	const int64_t len = pm_engine_soft_len(e, chain);
	if (len < 0 || cap < len) return PM_ERR_CAPACITY;
	int row = chain;
	if (component == 1 && e->chains[chain].d.slicer_kind == PM_SLICER_QUADRATURE) row = e->chains[chain].sign_q_row;
	else if (component != 0) return fail(e, PM_ERR_ARG, "component %d not available", component);
	cudaSetDevice(e->device);
	CK(cudaMemcpy(out, e->d_soft.p + (size_t)row * e->soft_stride, (size_t)len * sizeof(float), cudaMemcpyDeviceToHost));
	return PM_OK;
}

// packed sign words of a chain's soft values as the slicer reads them (after the float64 fix-up)
extern "C" int pm_engine_get_signs(const pm_engine *ce, int32_t chain, int32_t component, uint32_t *out, int64_t cap_words)
{
	pm_engine *e = const_cast<pm_engine *>(ce);
	if (!e || !e->have_run) return PM_ERR_STATE;
	const int64_t len = pm_engine_soft_len(e, chain);
	if (len < 0 || !out) return PM_ERR_ARG;
	const int64_t words = (len + 31) / 32;
	if (cap_words < words) return PM_ERR_CAPACITY;
	int row = chain;
	if (component == 1 && e->chains[chain].d.slicer_kind == PM_SLICER_QUADRATURE) row = e->chains[chain].sign_q_row;
	else if (component != 0) return fail(e, PM_ERR_ARG, "component %d not available", component);
	cudaSetDevice(e->device);
	if (words) CK(cudaMemcpy(out, e->d_sign.p + (size_t)row * e->sign_stride, (size_t)words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	if (len & 31) out[words - 1] &= (1u << (len & 31)) - 1u;      // bits past the last soft sample are not defined
	return PM_OK;
}

extern "C" int64_t pm_engine_stream_len(const pm_engine *e, int32_t chain)
{
	if (!e || !e->have_run || chain < 0 || chain >= (int)e->chains.size()) return -1;
	return e->h_cc[chain].nbytes;
}

extern "C" int pm_engine_get_stream(const pm_engine *ce, int32_t chain, int32_t stage, uint8_t *bytes,
                                    int64_t *addresses, int64_t cap)
{
	pm_engine *e = const_cast<pm_engine *>(ce);
	if (!e || !e->have_run) return PM_ERR_STATE;
	const int64_t len = pm_engine_stream_len(e, chain);
	if (len < 0 || cap < len) return PM_ERR_CAPACITY;
	if (len == 0) return PM_OK;
	cudaSetDevice(e->device);
	uint8_t *d_b = nullptr;
	long long *d_a = nullptr;
	CK(cudaMalloc((void **)&d_b, (size_t)len));
	CK(cudaMalloc((void **)&d_a, (size_t)len * sizeof(long long)));
	const uint32_t *src = (stage == 0) ? e->d_bits_raw.p : e->d_bits_lfsr.p;
	cudaError_t c1 = pm_launch_stream_export(e->d_cc.p, chain, src, e->bits_stride, e->d_byte_addr.p, e->addr_stride,
		d_b, d_a, e->sample_base, e->st);
	cudaError_t c2 = cudaMemcpyAsync(bytes, d_b, (size_t)len, cudaMemcpyDeviceToHost, e->st);
	cudaError_t c3 = cudaMemcpyAsync(addresses, d_a, (size_t)len * sizeof(long long), cudaMemcpyDeviceToHost, e->st);
	cudaError_t c4 = cudaStreamSynchronize(e->st);
	cudaFree(d_b);
	cudaFree(d_a);
	if (c1 != cudaSuccess || c2 != cudaSuccess || c3 != cudaSuccess || c4 != cudaSuccess)
		return fail(e, PM_ERR_CUDA, "stream export failed");
	return PM_OK;
}

// The last run's per-kernel times (option "kernel_times" = 1 before the run): one line per kernel name,
// "name<TAB>launches<TAB>milliseconds", in order of first launch.  Returns the number of bytes written (without the
// terminating 0) or PM_ERR_CAPACITY.
extern "C" int64_t pm_engine_kernel_times(const pm_engine *e, char *buf, int64_t cap)
{
	if (!e || !buf) return PM_ERR_ARG;
	if ((int64_t)e->kt_report.size() + 1 > cap) return PM_ERR_CAPACITY;
	memcpy(buf, e->kt_report.c_str(), e->kt_report.size() + 1);
	return (int64_t)e->kt_report.size();
}

// option "trace": "label ms" lines of the last host-buffer run, times relative to its start
extern "C" int64_t pm_engine_trace(pm_engine *e, char *buf, int64_t cap)
{
	if (!e || !buf || cap <= 0) return -1;
	cudaSetDevice(e->device);
	cudaDeviceSynchronize();
	std::string out;
	for (size_t i = 0; i < e->trace_used; i++) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, e->ev[0], e->ev_trace[i]) != cudaSuccess) { cudaGetLastError(); continue; }
		char line[96];
		snprintf(line, sizeof(line), "%s %.3f\n", e->trace_label[i].c_str(), ms);
		out += line;
	}
	const int64_t n = std::min<int64_t>((int64_t)out.size(), cap - 1);
	memcpy(buf, out.data(), (size_t)n);
	buf[n] = 0;
	return (int64_t)out.size();
}

extern "C" int pm_engine_get_stats(const pm_engine *e, pm_stats *out)
{
	if (!e || !out) return PM_ERR_ARG;
	*out = e->stats;
	return PM_OK;
}

// executed FP32 MACs per input sample (all chains, after sharing) -- used by
// bench.py for the roofline
extern "C" double pm_engine_front_macs_per_sample(const pm_engine *e)
{
	double m = 0;
	if (e) for (auto &g : e->groups) m += g.macs_per_sample;
	return m;
}

// executed bf16 multiply-adds per input sample on the tensor cores (csrc/lpf_tc.cu: six piece products x the K steps
// issued x 16 per output and tone; about eleven times the 100 useful ones), 0 when no group takes that route
extern "C" double pm_engine_front_tensor_macs_per_sample(const pm_engine *e)
{
	double m = 0;
	if (e) for (auto &g : e->groups) if (g.tensor) m += g.tc_macs_per_sample;
	return m;
}

// the FP32 multiply-adds per input sample the low-pass takes on the FP32 pipe (what the tensor-core route replaces)
extern "C" double pm_engine_front_lpf_macs_per_sample(const pm_engine *e)
{
	double m = 0;
	if (e) for (auto &g : e->groups) m += g.lpf_macs_per_sample;
	return m;
}

// host-side check behind the sliding-window correlators (no device needed): are (ti[k], tq[k]) = a e^{i(phi + w k)}?
extern "C" int pm_taps_are_rotation(const double *ti, const double *tq, int32_t n, double *step)
{
	if (!ti || !tq || n <= 0) return 0;
	double w = 0;
	const bool ok = rotation_taps(std::vector<double>(ti, ti + n), std::vector<double>(tq, tq + n), w);
	if (ok && step) *step = w;
	return ok ? 1 : 0;
}

// cycles per stage of the AFSK front end summed over the CTAs launched since the last call (option "stage_clocks")
extern "C" int pm_engine_stage_clocks(pm_engine *e, uint64_t *out8)
{
	if (!e || !out8) return PM_ERR_ARG;
	if (!e->opt_stage_clocks || !e->d_stage_clk.p) return fail(e, PM_ERR_STATE, "set option 'stage_clocks' first");
	cudaSetDevice(e->device);
	CK(cudaStreamSynchronize(e->st));
	CK(cudaMemcpy(out8, e->d_stage_clk.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	CK(cudaMemset(e->d_stage_clk.p, 0, 8 * sizeof(unsigned long long)));
	return PM_OK;
}

extern "C" int pm_engine_front_tile(const pm_engine *e, int group)
{
	if (!e || group < 0 || group >= (int)e->groups.size()) return -1;
	return e->groups[group].tile;
}

extern "C" int pm_measure_fp32_peak(int device, double *tflops)
{
	if (!tflops) return PM_ERR_ARG;
	if (cudaSetDevice(device) != cudaSuccess) return PM_ERR_CUDA;
	float *d = nullptr;
	if (cudaMalloc((void **)&d, 16) != cudaSuccess) return PM_ERR_CUDA;
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	const int blocks = 148 * 8, iters = 4096;
	pm_launch_ffma_peak(d, blocks, 64, 0);
	cudaDeviceSynchronize();
	double best = 0;
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(a, 0);
		pm_launch_ffma_peak(d, blocks, iters, 0);
		cudaEventRecord(b, 0);
		if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return PM_ERR_CUDA; }
		float ms = 0;
		cudaEventElapsedTime(&ms, a, b);
		const double flops = 2.0 * blocks * 256.0 * iters * 8 * 16;
		best = std::max(best, flops / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(a); cudaEventDestroy(b);
	cudaFree(d);
	*tflops = best;
	return PM_OK;
}

// Pin a caller-owned buffer in place (page-locks it: ~20 ms per 100 MB, once), so that pm_engine_run can DMA straight
// out of it on every later call.  Read-only mappings (a memory-mapped WAV) are registered read-only.
extern "C" int pm_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) return PM_ERR_ARG;
	if (host_pointer_is_pinned(p)) return PM_OK;
	cudaError_t ce = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
	if (ce != cudaSuccess) {
		cudaGetLastError();
		ce = cudaHostRegister(p, bytes, cudaHostRegisterReadOnly);
	}
	if (ce != cudaSuccess) { cudaGetLastError(); return PM_ERR_CUDA; }
	return PM_OK;
}

extern "C" int pm_host_unregister(void *p)
{
	if (!p) return PM_ERR_ARG;
	if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return PM_ERR_CUDA; }
	return PM_OK;
}

// Dependent float64 operation latency in nanoseconds (and SM cycles): the floor under the sequential carrier loops,
// whose ~25 operations per sample form one dependency chain (csrc/loops.cu)
extern "C" cudaError_t pm_launch_fp64_chain(double *, long long *, int, cudaStream_t);
extern "C" int pm_measure_fp64_chain(int device, double *ns_per_op, double *cycles_per_op)
{
	if (!ns_per_op || !cycles_per_op) return PM_ERR_ARG;
	if (cudaSetDevice(device) != cudaSuccess) return PM_ERR_CUDA;
	double *d = nullptr;
	long long *c = nullptr;
	if (cudaMalloc((void **)&d, 8) != cudaSuccess || cudaMalloc((void **)&c, 8) != cudaSuccess) return PM_ERR_CUDA;
	cudaMemset(d, 0, 8);
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	const int iters = 20000;                       // 320 000 dependent operations
	pm_launch_fp64_chain(d, c, 100, 0);
	cudaDeviceSynchronize();
	cudaEventRecord(a, 0);
	pm_launch_fp64_chain(d, c, iters, 0);
	cudaEventRecord(b, 0);
	int rc = PM_OK;
	if (cudaEventSynchronize(b) != cudaSuccess) rc = PM_ERR_CUDA;
	float ms = 0;
	long long cyc = 0;
	cudaEventElapsedTime(&ms, a, b);
	cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
	*ns_per_op = (double)ms * 1e6 / (16.0 * iters);
	*cycles_per_op = (double)cyc / (16.0 * iters);
	cudaEventDestroy(a); cudaEventDestroy(b);
	cudaFree(d); cudaFree(c);
	return rc;
}

// pinned host memory for callers without torch
extern "C" void *pm_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
	return p;
}
extern "C" void pm_host_free(void *p) { if (p) cudaFreeHost(p); }

// pm_common.cuh -- shared definitions for the sm_100a kernels of libpymodem_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Per-kernel timing (option "kernel_times", pm_engine_kernel_times): every launcher announces the kernel it is about to
// launch; the engine records a CUDA event on that stream when a timing pass is active on the calling thread, and does
// nothing otherwise (csrc/engine.cu).
extern "C" void pm_kt_mark(const char *name, cudaStream_t st);

#define PM_MAX_MAG     8      // tone-magnitude streams per AFSK front-end group
#define PM_MAX_PAIR    8      // (mark stream, space stream) pairs per group
#define PM_MAX_GCH     16     // chains per front-end group
#define PM_MAX_TAPS    3072   // floats of FIR taps carried in the kernel parameter block
#define PM_FRONT_THREADS 256

// Shared-memory sample arrays use a padded layout: 4 floats of padding after
// every 16.  A thread owns 16 consecutive outputs ("unit"), so unit bases are
// 80 B apart instead of 64 B and the 128-bit window loads of the 8 lanes of a
// quarter-warp fall into 8 different 16-byte bank groups (conflict-free).
__host__ __device__ __forceinline__ int pm_phys(int i) { return i + ((i >> 4) << 2); }
// The same idea for arrays of float2 (a mark/space magnitude pair per sample): 2 float2 (16 B) of padding after every
// 16, so unit bases are 144 B apart and a quarter-warp's 128-bit loads hit 8 different 16-byte bank groups.
// Returns the index in float2 units.
__host__ __device__ __forceinline__ int pm_phys2(int i) { return i + ((i >> 4) << 1); }

// Plan of one fused AFSK front-end launch (all chains that share the input
// band-pass and output low-pass taps).  Passed as a __grid_constant__ kernel
// parameter, so taps are read through the constant bank (uniform loads).
struct AfskPlan {
	int tile;                 // chain output samples per CTA (multiple of 32)
	int U_x, U_m, U_l;        // 16-output units per CTA: BPF, magnitude, LPF
	int a_len;                // audio samples staged per CTA
	int n_bpf, bpf_off;       // input BPF (taps reversed, count multiple of 4)
	int n_mag;
	int mag_n[PM_MAX_MAG];    // correlator taps (multiple of 4)
	int mag_iq_off[PM_MAX_MAG];   // taps interleaved (i[0], q[0], i[1], q[1], ...): operands of the packed FFMA2
	int mag_slide[PM_MAX_MAG];    // window length N when the taps are a rotation e^{i(phi + w k)} (afsk.py:134-144): sliding
	                              // window sum instead of the FIR; 0 = arbitrary taps, direct FIR
	int mag_e_off[PM_MAX_MAG];    // (cos wk, sin wk) for k in [0, N + 16), then (-cos wk, -sin wk) for k in [0, 16)
	int n_lpf, lpf_off;
	int lpf2_off;             // the low-pass taps, each stored twice in a row: (h, h) operands of the packed FFMA2
	int bpf2_off;             // the band-pass taps likewise (tensor_lpf route: the band-pass runs two half tiles side by side on FFMA2)
	int bpf_half;             // that route: 16-output units per half tile (U_x = 2 * bpf_half); the staged audio is an array of
	                          // (a[i], a[i + 16 * bpf_half]) pairs
	int n_pair;
	int pair_mark[PM_MAX_PAIR];
	int pair_space[PM_MAX_PAIR];
	int pair_first[PM_MAX_PAIR + 1];   // chains of pair p: chain_*[pair_first[p] .. pair_first[p+1])
	int pair_fused[PM_MAX_PAIR];       // window length when both tones are sliding-window tones of one length used by this pair
	                                   // only: one pass computes both (front.cu SlidePair; their mag_dst lists are empty); else 0
	int pair_ea[PM_MAX_PAIR];          // mag_e_off of the mark and of the space tone of a fused pair (indexed by the pair,
	int pair_eb[PM_MAX_PAIR];          // so that the table addresses stay warp-uniform for the compiler)
	int n_chain;
	int chain_gid[PM_MAX_GCH];         // engine-wide chain index
	float chain_gain[PM_MAX_GCH];      // space_gain (afsk.py:143)
	long long chain_nout[PM_MAX_GCH];  // valid demod outputs of the chain (N - trim)
	int s_x1_off, s_m_off, s_m_stride; // shared-memory carve-up, in floats; s_m_stride = one PAIR stream (float2 per sample)
	int mag_dst_first[PM_MAX_MAG + 1]; // tone j feeds mag_dst[mag_dst_first[j] .. mag_dst_first[j+1])
	int mag_dst[2 * PM_MAX_PAIR];      // pair * 2 + slot (0: mark, 1: space)
	int tensor_lpf;           // 1: the kernel stops after the tone magnitudes and writes them, split into three bf16 pieces
	                          // each, to global memory (MagOut); the low-pass and the epilogue run on the tensor cores
	                          // (csrc/lpf_tc.cu).  The tile then has no low-pass halo: U_m = tile / 16.
	float guard_eps;
	float chain_guard_abs[PM_MAX_GCH];  // raw-input term of the sign guard per unit of max|audio| over the tile:
	                                    // c_abs * 2^-24 * sum|h_bpf| * N_corr * sum|h_lpf| * (1 + space_gain)
	alignas(16) float taps[PM_MAX_TAPS];   // every tap set starts at a multiple of 4 floats: read as float4
};

// ---- tensor-core low-pass (csrc/lpf_tc.cu) ------------------------------------------------------------
// The 100-tap low-pass of every tone magnitude as a tcgen05 GEMM: D[128 x 64] (FP32, TMEM) = A[128 x 192] * B[192 x 64],
// A[i][d] = m[64 i + d] (the magnitude stream itself, rows of 64 samples), B[d][n] = h[d - n] (banded Toeplitz matrix
// of the taps), so D[i][n] = sum_t h[t] m[64 i + n + t] = output 64 i + n.  FP32 accuracy from bf16 operands: every
// magnitude and every tap is split into three bf16 pieces (x = x1 + x2 + x3 exactly) and the six largest piece
// products are accumulated into one TMEM tile (tools/ubench/fir_umma.cu: error 6.5e-8 of sum|h||m| rms, below the
// sequential FP32 FMA's 1.5e-7).
//
// Magnitude streams in global memory (written by afsk_front_kernel<.., MAGS = true>, read by lpf_tc_kernel with bulk
// copies): per tone t and piece q an array of `rows` rows of 128 bytes; row r holds samples [64 r, 64 r + 64) as bf16,
// its 16-byte chunks XOR-swizzled with r & 7 (the shared-memory SWIZZLE_128B pattern, so a run of rows starting at a
// multiple of 8 can be copied verbatim into a 1024-byte aligned operand buffer).
#define TC_ROWS 128                      // M: rows (of 64 outputs) per tile
#define TC_N 64
#define TC_KBLK 3                        // K = 192 = 3 x 64 >= n_lpf + 63
#define TC_TILE (TC_ROWS * TC_N)         // 8192 outputs of every tone per tile
#define TC_A_ROWS (TC_ROWS + TC_KBLK)    // rows of one operand buffer: K block kb of row i is row i + kb
#define TC_A_BYTES (TC_A_ROWS * 128)     // 16768
#define TC_A_STRIDE 17408                // ... rounded up to 1024
#define TC_B_BYTES (TC_KBLK * TC_N * 128)    // one tap piece: 24576
#define TC_MAX_TONES 4                   // TMEM: 2 sets x 4 tones x 64 columns = 512
#define TC_MAX_LPF (TC_KBLK * 64 - 16 - 63)   // 113 taps: eleven K steps of 16 (see lpf_tc.cu on the twelfth)

struct MagOut {
	unsigned char *base;      // piece array (t, q) starts at base + (t * 3 + q) * rows * 128
	long long rows;
	float *tile_amax;         // largest raw |sample| staged by front tile i (the guard's raw-input term), one float per tile
};

__host__ __device__ __forceinline__ long long pm_mag_offset(long long sample)
{
	const long long r = sample >> 6;
	const int c = (int)((sample & 63) >> 3);
	return r * 128 + (long long)((c ^ (int)(r & 7)) << 4);        // byte offset of the 16-byte chunk holding `sample`
}

struct LpfTcPlan {
	int n_mag;                         // tones = TMEM accumulators per tile
	int n_pair, n_chain;
	int pair_mark[PM_MAX_PAIR], pair_space[PM_MAX_PAIR];
	int pair_first[PM_MAX_PAIR + 1];
	int chain_gid[PM_MAX_GCH];
	float chain_gain[PM_MAX_GCH];
	float chain_guard_abs[PM_MAX_GCH];
	long long chain_nout[PM_MAX_GCH];
	float guard_eps;
	int n_lpf;                         // taps of the low-pass (K steps beyond n_lpf + 63 hold only zeros and are skipped)
	int tile_a;                        // outputs per front tile (granularity of tile_amax)
	int reach;                         // raw samples an output depends on beyond its own index: sum of (taps - 1)
	long long n_tile_a;                // front tiles per run
	int debug_mask;                    // development aid (option "tc_debug", tools/tc_debug.py): 1 no epilogue stores, 2 no bulk copies, 4 no MMAs, 8 no tile_amax reads
};

// Plan of a single-FIR front end (FSK: fsk.py:149-159).
struct FirPlan {
	int tile;
	int U_y;
	int a_len;
	int n_taps, taps_off;
	int n_chain;
	int chain_gid[PM_MAX_GCH];
	int chain_neg[PM_MAX_GCH];         // fsk.py:153 invert
	long long chain_nout[PM_MAX_GCH];
	float guard_eps;
	alignas(16) float taps[PM_MAX_TAPS];
};

// Samples whose FP32 soft value is too close to zero to trust its sign are
// queued here and re-evaluated in FP64 by the fix-up kernel.
struct GuardList {
	unsigned long long *entries;       // (chain gid << 48) | sample index
	unsigned int *count;
	unsigned int cap;
	unsigned long long *stage_clk;     // optional tracing (option "stage_clocks"): per-CTA cycles of the AFSK front end's
	                                   // stages summed over CTAs [stage, band-pass, correlators, low-pass+epilogue], [4] = CTAs
	const unsigned int *from, *to;     // guard_fixup_kernel only: entries [*from, *to) (snapshots of count taken between the
	                                   // front-end launches of a host-buffer run); null: [0, *count)
};

// ---- tables shared by the kernels and the host engine -------------------------
struct Fp64Chain {            // float64 taps for the guard-band fix-up
	int kind;                 // PM_MODEM_AFSK (1) or PM_MODEM_FSK (2)
	int n_bpf, n_corr, n_lpf;
	int neg;
	long long audio_off;      // where this chain's recording starts in the audio buffer (batched runs), and its length
	long long n_audio;
	const double *bpf;        // reversed (correlation order)
	const double *mark_i, *mark_q, *space_i, *space_q;
	const double *lpf;
};

struct SlicerChain {
	double sps;              // samples_per_symbol   slicer.py:51
	double thr;              // rollover_threshold   slicer.py:52
	double lock;             // lock_rate
	long long nout;          // valid soft samples of the chain
	int quadrature;          // zero crossings on I or Q (slicer.py:226-233)
	int sign_q_row;          // row of the Q sign stream (quadrature)
	int sign_row;            // row of the (I) sign stream
	int fast;                // the shortened update below is exact for this chain
	// Shortened clock update.  A roll-over happens when fl(clock + 1.0) >= thr, i.e. (rounding is monotonic) when
	// clock >= c_star, the smallest double whose successor step reaches thr; clock < thr always, so for positive
	// c_star this is an integer comparison of the bit patterns.  When [c_star, thr + 1) lies inside one binade,
	// clock + 1.0 is exact in that range and fl(fl(clock + 1.0) - sps) == fl(clock - (sps - 1.0)): both candidates
	// of the next clock come straight from the old one, one double operation deep instead of two plus a compare.
	long long c_star_bits;
	double sps_m1;           // sps - 1.0 (exact for sps >= 1)
	// Stretches without zero crossings (digital silence: a squelched receiver, the gaps of clean synthetic audio).  The
	// clock then only counts and rolls over, and when sps is a dyadic rational it becomes EXACTLY periodic within one
	// period: the first pass through the top binade rounds it onto that binade's grid, after which every +1.0 and -sps
	// is exact (tools/slicer_quiet_check.py: every start state periodic from sample < P on, P = smallest integer
	// multiple of sps).  Mask words and word-end clocks then repeat every quiet_words = lcm(32, P) / 32 words: a repair
	// (slicer_verify / slicer_sweep) copies them instead of stepping.  0: not available for this chain.
	int quiet_words;
	int quiet_lead;          // quiet words to step exactly before the repetition may be trusted: ceil(P / 32) + 1
};

#define PM_QUIET_MAX 8       // longest repetition (in words) the repairs keep
struct QuietState {           // carried by a repairing thread across its run_words calls
	int run;                 // whole quiet words processed since the last word with a crossing
	double c[PM_QUIET_MAX];  // clock after the most recent quiet word w with w % quiet_words == slot
	uint32_t m[PM_QUIET_MAX];// its mask word
};

struct SegState {
	double clock;            // phase_clock
	unsigned int last;       // sign bit of last_sample (1: >= 0.0); slicer.py:55 starts at 0.0
	unsigned int last_q;
};

// Geometry of one slicer pass over the local sign stream.
struct SlicerGeom {
	long long origin_w;      // first word of segment 0 (own range start / 32)
	int n_seg;
	int seg_words;
	int warm_words;
	int chk_words;           // checkpoint spacing (divides seg_words)
	int n_chk;               // seg_words / chk_words
	int true_start;          // local sample 0 is the true start of the recording
	int warm_f32_words;      // the first warm_f32_words of every warm-up run in FP32 (3x shorter dependency chain per sample):
	                         // a warm-up only has to get NEAR the true state, the exact float64 tail after it contracts what
	                         // is left (every zero crossing multiplies the error by lock_rate) and the verify pass decides
	int k_init;              // segment whose predecessor state is init[] (0 during the local pass; the first own
	                         // segment when a hand-off repair starts there -- the segments before it are frozen)
	int warm_far_f64;        // the far part crossing by crossing in float64 instead of FP32: an ulp or two off instead of
	                         // 2^-24, so a 4096-sample exact tail suffices (tools/slicer_warm_sim.py)
	int k_first;             // slicer_segments_kernel handles segments [k_first, k_first + k_count): the host-buffer
	int k_count;             // path launches the segments of every chunk of audio as its sign words become final
};

// Per-chain placement of the local bitstream when the recording is sharded on
// the sample axis (all zero / "everything is mine" for an unsharded run).
struct ShardBits {
	long long bit_off;       // local stream position of the first own bit (tail bits + byte-alignment padding before it)
	long long own_lo;        // packets are emitted by the shard that holds their closing bit: own_lo <= pos < own_hi
	long long own_hi;
	long long valid_from;    // first local position whose descrambled bit is known to be right: the bits before it are
	                         // alignment padding or lack LFSR history (0 on the first shard)
	int first;               // first shard: the stream start is the true start
	int pad;
};

struct BitChain {
	long long nout;            // valid soft samples
	int sign_row, sign_q_row;
	int quadrature;
	int bps;                   // bits per symbol (1 binary / bpsk, 2 qpsk)
	unsigned int state_mask;
	unsigned int demap[16];
	unsigned long long lfsr_poly;
	int lfsr_invert;
	int codec;                 // PM_CODEC_*
	int il2p_crc, il2p_disable_rs, il2p_min_dist, il2p_sync_tol;   // il2p.py:140-145
};

#define IL2P_SLOT 1056         // bytes per speculative IL2P decode (15-byte AX.25 header + PID, 1023 payload, FCS)
#define IL2P_FAIL 0u
#define IL2P_OK 1u
#define IL2P_INCOMPLETE 2u
// Where the IL2P decoder of a chain stands at a shard boundary (local bit positions on the device)
struct Il2pHand {
	long long pos;             // the search (re)starts at this stream bit
	unsigned int mode;         // 0: start of the recording (register = 0xFFFFFF, il2p.py:119); 1: a frame ended at pos-1
	                           // (register = its last byte, zero above, il2p.py:147-153); 2: plain search, 32 real bits behind
	unsigned int leak;         // corrected-byte counts of failed frames not yet attributed (il2p.py:200-211)
};

struct Il2pRes {
	long long end_bit;         // last stream bit the frame consumed (search resumes at end_bit + 1)
	unsigned int status;
	unsigned int len;          // bytes of the emitted packet (status OK)
	unsigned int corrected;    // RS corrections of this frame (of the blocks decoded so far when it failed)
	unsigned int pad;
};

// ---- float64 pipeline (csrc/loops.cu): the modems with a recursive stage (BPSK Costas loop, MPSK
// decision-directed loop, AFSK PLL) and AFSK chains whose tone pair is too close for FP32 ----------
#define P64_TILE 1024          // outputs per CTA of a float64 FIR stage
#define P64_THREADS 256
#define P64_MAX_TAPS 4096

struct LoopConst {             // mirrors pm_loop_desc (include/pymodem_b200.h) without the pointers
	double agc_scaled_attack, agc_scaled_decay, agc_sustain_time, agc_sustain_increment, agc_target;
	double nco_phase_scale, nco_index_scale, nco_set_frequency, nco_two_pi, nco_quarter;
	double iir_b0, iir_b1, iir_a1;
	double pi_gain, pi_p, pi_i, pi_limit, pi_integral0;
};

struct P64Chain {
	int kind;                  // PM_MODEM_AFSK (1), PM_MODEM_BPSK (3), PM_MODEM_MPSK (4), PM_MODEM_AFSK_PLL (5)
	int gid;                   // engine chain index (row of the soft-value export)
	int sign_row, sign_q_row;
	int n_bpf, n_mid, n_out, mid_delay;
	long long n_audio;
	long long audio_off;       // where this chain's recording starts in the audio buffer (batched runs)
	long long L1, L2, L3;      // samples after the input FIR, after the middle stage, final soft samples
	const double *bpf;         // all taps reversed (correlation order)
	const double *mid0, *mid1, *mid2, *mid3;   // AFSK: mark_i, mark_q, space_i, space_q; MPSK: mid0 = Hilbert
	const double *out_taps;
	double *A, *B, *C, *D;
	LoopConst lc;
	const double *wavetable;
	const int *pd_table;
	int wt_size, pd_g;
	unsigned long long *max_slot;   // ordered-integer image of max(A) for AGC.normal (agc.py:67)
};

// ---- shard link over NVLink peer memory (csrc/link.cu) -------------------------------------------------
#define LINK_MAX_WORLD 16
struct LinkGeom {
	int rank, world, nc, tail_words;
	long long slot_bytes;      // one parity slot
	long long off_states;      // pm_shard_state[world][nc]
	long long off_sflag;       // unsigned[world]
	long long off_tail;        // uint32[nc][tail_words]   (written by rank - 1)
	long long off_tflag;       // unsigned
	long long off_rflag;       // unsigned[world]
	long long off_rhdr;        // u64[world][2]: records, arena bytes
	long long off_rdata;       // world regions of rec_region bytes: records then arena
	long long rec_region;
	long long off_il2p;        // pm_il2p_state[nc]: where the previous rank's IL2P walk stands (written by rank - 1)
	long long off_iflag;       // unsigned
};
struct LinkPeers {
	unsigned char *base[LINK_MAX_WORLD];
};

struct ChainCounters {
	long long nbits;           // local stream bits of the chain (incl. hand-off tail and padding when sharded)
	long long nbytes;          // nbits / 8 (trailing partial byte dropped, slicer.py:95)
	int nflags;
	int seq_needed;            // ax25 max_packet_length overflow seen: replay sequentially
	int tail_short;            // sharded: a frame closing in the own range reaches back past the hand-off tail
	int n_emit;                // packets emitted by this chain
	long long n_emit_bytes;
};

struct GapRec {
	unsigned int emit;        // 1: a packet was emitted at the closing flag
	unsigned int len;         // bytes appended since the previous flag event
	unsigned int scratch_off; // where they are in the chain's scratch bytes
	unsigned int addr;        // byte address of the closing flag's byte (ax25.py:82)
	unsigned int corrected;   // BytesCorrected (IL2P; 0 for AX.25)
};

struct PacketRecDev {          // mirrors pm_packet_rec (include/pymodem_b200.h)
	unsigned int chain;
	unsigned int len;
	unsigned long long offset;
	long long streamaddress;
	unsigned int bytes_corrected;
	unsigned short calculated_crc;
	unsigned short carried_crc;
	unsigned char valid_crc;
	unsigned char valid_header;
	unsigned char pad[6];
};

struct PacketTotals {
	unsigned long long n_packets;
	unsigned long long n_bytes;
};

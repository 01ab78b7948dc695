/*
 * pymodem_b200.h -- C ABI of libpymodem_b200.so, the B200 (sm_100a) engine for
 * pymodem's demod_chain hot path.
 *
 * The reference (ninocarrillo/pymodem) is plain Python and has no FFI; the seam
 * this library replaces is the per-chain process fan-out
 *     pymodem.py:140-166      Process(target=multiprocess_chain, args=[chain, audio, queue])
 *     chain_execute.py:6-52   modem.demod -> slicer.slice -> stream.stream_unscramble_8bit -> codec.decode
 * and the entry points below are what a ctypes binding on the reference side
 * calls instead (see INTEGRATION.md for the stub).  Conventions: plain pointers
 * and sizes only, int return code (0 = ok, <0 = error; pm_last_error() gives the
 * text), no exceptions cross the boundary, the caller owns every buffer it
 * passes, the engine owns all device memory, one host thread per engine.
 * There is NO CPU fallback: every entry point fails with PM_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef PYMODEM_B200_H
#define PYMODEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK              0
#define PM_ERR_ARG        -1
#define PM_ERR_CUDA       -2
#define PM_ERR_UNSUPPORTED -3
#define PM_ERR_CAPACITY   -4
#define PM_ERR_STATE      -5

/* modem kinds -- chain_builder.py:17-38 */
#define PM_MODEM_NONE      0   /* no modem: the chain only serves the per-stage calls (pm_engine_slice_soft, ...) */
#define PM_MODEM_AFSK      1   /* afsk.py:13   AFSKModem  */
#define PM_MODEM_FSK       2   /* fsk.py:15    FSKModem   */
#define PM_MODEM_BPSK      3   /* psk.py:20    BPSKModem  */
#define PM_MODEM_MPSK      4   /* psk.py:479   MPSKModem  */
#define PM_MODEM_AFSK_PLL  5   /* afsk_pll.py:16 AFSKPLLModem */
/* slicer kinds -- chain_builder.py:40-52 */
#define PM_SLICER_BINARY     1 /* slicer.py:9   */
#define PM_SLICER_QUADRATURE 2 /* slicer.py:109 */
/* codec kinds -- chain_builder.py:62-69 */
#define PM_CODEC_AX25      1   /* ax25.py:11  */
#define PM_CODEC_IL2P      2   /* il2p.py:109 */

/*
 * One demod_chain = {modem, slicer, stream, codec} (pymodem.py:68-114), flattened
 * to plain data.  All filter taps are the float64 vectors produced by the
 * reference's own tune() arithmetic on the host (afsk.py:102-146, fsk.py:115-147,
 * psk.py:111-160/640-703); the engine never re-derives them.
 */
/*
 * Constants of the recursive part of a PSK / PLL modem, computed on the host by the reference's own
 * expressions (math.sin, math.tan, atan2 of the host libm), so that the device works from bit-identical
 * numbers: AGC agc.py:8-24, NCO nco.py:14-32, IIR_1 iir.py:15-29, PI_control pi_control.py:8-13,
 * PhaseDetector.qpsk_error_table phase_detector.py:34-45, Hilbert hilbert.py:9-34.
 */
typedef struct pm_loop_desc {
	double agc_scaled_attack;       /* attack_rate / sample_rate            agc.py:15 */
	double agc_scaled_decay;        /* decay_rate / sample_rate             agc.py:16 */
	double agc_sustain_time;
	double agc_sustain_increment;   /* 1 / sample_rate                      agc.py:17 */
	double agc_target;              /* target_amplitude */
	double nco_phase_scale;         /* 2.0 * pi / sample_rate               nco.py:30 */
	double nco_index_scale;         /* wavetable_size / (2.0 * pi)          nco.py:28 */
	double nco_set_frequency;       /* carrier_freq */
	double nco_two_pi;              /* 2.0 * pi */
	double nco_quarter;             /* wavetable_size / 4.0                 nco.py:46 */
	const double *nco_wavetable;    /* amplitude * sin(i * 2.0 * pi / size) nco.py:24-26 */
	int64_t nco_size;               /* <= 1024 */
	double iir_b0, iir_b1, iir_a1;  /* gain*b0, gain*b1, a1                 iir.py:24-29 */
	double pi_gain, pi_p, pi_i, pi_limit;
	double pi_integral0;            /* PI integral at the first sample (psk.py:703 presets -max_freq_offset) */
	const int32_t *pd_table;        /* [granularity][granularity], MPSK only */
	int64_t pd_granularity;         /* <= 64 */
	const double *hilbert;          /* Hilbert taps, MPSK only (numpy.convolve order) */
	int32_t n_hilbert;
	int32_t hilbert_delay;          /* hilbert.py:14 tap_count // 2 */
} pm_loop_desc;

typedef struct pm_chain_desc {
	int32_t modem_kind;
	int32_t slicer_kind;
	int32_t codec_kind;
	int32_t invert_soft;            /* fsk.py:153-154  audio = -audio */

	/* --- modem FIR taps (numpy.convolve order, i.e. h[0] first) --- */
	const double *bpf;              /* afsk.py:112 input_bpf / fsk.py:133 input_lpf / psk input_bpf */
	int32_t n_bpf;
	int32_t n_corr;                 /* afsk.py:134 correlator length */
	const double *mark_i;           /* afsk.py:138 */
	const double *mark_q;           /* afsk.py:139 */
	const double *space_i;          /* afsk.py:143 (space_gain applied, as the reference convolves it) */
	const double *space_q;          /* afsk.py:144 */
	const double *space_unit_i;     /* cos/sin before space_gain: lets chains that differ only  */
	const double *space_unit_q;     /* in space_gain share one correlator pass                  */
	double space_gain;              /* afsk.py:143 */
	const double *lpf;              /* afsk.py:122 output_lpf / psk.py rrc taps */
	int32_t n_lpf;
	int32_t recording;              /* batch slot: which recording of pm_engine_run_batch this chain decodes (0 for the
	                                   single-recording calls).  Chains of different recordings never share a front-end pass. */

	/* --- slicer (slicer.py:49-56, 181-191) --- */
	double slicer_sample_rate;      /* pymodem.py:86-90 */
	double symbol_rate;
	double lock_rate;
	uint32_t state_mask;            /* quadrature */
	uint32_t bits_per_symbol;       /* quadrature */
	uint32_t demap[16];             /* quadrature */

	/* --- stream (lfsr.py:10-20) --- */
	uint64_t lfsr_poly;
	int32_t lfsr_invert;

	/* --- codec options (il2p.py:140-145) --- */
	int32_t il2p_crc;
	int32_t il2p_disable_rs;
	int32_t il2p_min_dist;
	int32_t il2p_sync_tol;
	int32_t reserved1;

	/* --- PSK / PLL loop constants (psk.py, afsk_pll.py); see pm_loop_desc --- */
	const pm_loop_desc *loop;       /* NULL unless modem_kind is BPSK/MPSK/AFSK_PLL */
} pm_chain_desc;

/* One decoded packet == one reference PacketMeta (packet_meta.py:178-208). */
typedef struct pm_packet_rec {
	uint32_t chain;                 /* index into the loaded chain table (SourceDecoder = its object_name) */
	uint32_t len;                   /* len(PacketMeta.data) */
	uint64_t offset;                /* offset of data[0] in the byte arena */
	int64_t  streamaddress;         /* ax25.py:82 / il2p.py:364 */
	uint32_t bytes_corrected;       /* il2p.py:200 */
	uint16_t calculated_crc;        /* crc_functions.py:45-54  (PacketMeta.CalcCRC) */
	uint16_t carried_crc;           /* crc_functions.py:44 */
	uint8_t  valid_crc;             /* crc_functions.py:55-61 */
	uint8_t  valid_header;          /* packet_meta.py:21-41 */
	uint8_t  pad[6];
} pm_packet_rec;

/* Timings are CUDA-event milliseconds on the engine's own stream. */
typedef struct pm_stats {
	double total_ms;                /* first enqueue .. results on host */
	double h2d_ms;                  /* sum of host->device audio copies (0 for run_device) */
	double front_ms;                /* FIR front-end kernels (all launches) */
	double fixup_ms;                /* FP64 guard-band re-evaluation */
	double slicer_ms;               /* symbol-timing recovery + hand-off verification */
	double bits_ms;                 /* gather + LFSR + codec kernels */
	double d2h_ms;
	int64_t kernel_launches;        /* kernels of this library launched by the last run */
	int64_t front_launches;
	int64_t guard_flagged;          /* samples re-evaluated in FP64 */
	int64_t slicer_repairs;         /* segments re-run because the warm-up hand-off did not verify */
	int64_t slicer_segments;
	int64_t h2d_bytes;
	int64_t d2h_bytes;
	int64_t n_packets;
	int64_t n_stream_bits;          /* sliced bits over all chains */
} pm_stats;

typedef struct pm_engine pm_engine;

/* Create an engine bound to one CUDA device (one process per GPU). */
int pm_engine_create(int device, pm_engine **out);
void pm_engine_destroy(pm_engine *e);
const char *pm_last_error(const pm_engine *e);

/* Replaces building demod_stack (pymodem.py:58-114): copies the chain table and
 * uploads taps.  Chains keep their index (= config order). */
int pm_engine_load_chains(pm_engine *e, const pm_chain_desc *chains, int32_t n_chains);

/* Tunables (results do not depend on any of them; they trade time against the amount of float64 re-evaluation):
 * "segment_len", "warmup_len", "checkpoint_len" (samples, multiples of 32; slicer geometry, defaults 24576 / 49152 / 1024),
 * "warmup_exact_len" (samples of exact, sample-by-sample tail of every slicer warm-up, default 4096; the part before it
 *   runs crossing by crossing in closed form -- a warm-up only has to get near the true state; 0 = all exact),
 *   "warmup_far_f64" (1, default: that part in float64; 0: in FP32, which needs an exact tail of 16384 samples),
 *   "verify_passes", "slicer_fast" (0: always the plain clock update),
 * "guard_eps" (relative width of the FP32 front end's sign guard band, default 2^-18; samples inside it are
 *   re-evaluated in float64), "guard_abs" (the guard's raw-input term: multiples, default 0.25, of 2^-24 max|audio of the tile|
 *   sum|h_bpf| N_corr sum|h_lpf| (1 + space_gain) added to the band -- the band-pass rounds at the magnitude of the RAW
 *   samples, DC / hum / out-of-band energy included), "guard_cap" (initial capacity of the guard list; grows on demand),
 * "slide_correlator" (0: tone correlators as plain FIRs even when their taps are a rotation), "fuse_pairs" (0: sliding
 *   windows tone by tone instead of mark and space of a pair together), "tensor_lpf" (0: the AFSK low-pass on the FP32
 *   pipe, fused into the front kernel, even where the tensor-core route applies), "tile" (front-end outputs
 *   per CTA, 0 = cost model), "keep_soft" (1: keep the soft values for pm_engine_get_soft), "h2d_chunk" (samples per
 *   host-to-device copy of pm_engine_run), "copy_threads" (host threads that stage pageable input, default 4), "stage_clocks" (1: trace the front end's stages, pm_engine_stage_clocks),
 * "early_tail" (host-buffer runs; 1, default: the float64 guard fix-up runs chunk by chunk beside the copy; 2: batches of
 *   slicer segments too -- measured slower; 0: both after the last chunk), "early_batches" (about how many batches, default 12),
 *   "trace" (1: timing events at the steps of a host-buffer run, pm_engine_trace), "kernel_times" (1: pm_engine_kernel_times),
 * "precise" (1: every AFSK chain takes the float64 pipeline; default: only chains whose tone pair is so
 *   close that |mark| - |space| cancels below FP32 resolution; set before pm_engine_load_chains). */
int pm_engine_set_option(pm_engine *e, const char *key, double value);

/*
 * Replaces the fan-out + gather of pymodem.py:140-166 for every loaded chain:
 * demod -> slice -> unscramble -> decode over the whole recording.
 * pm_engine_run:        audio in HOST memory (pinned preferred); the H2D copy is
 *                       chunked and overlapped with the front-end kernels.
 * pm_engine_run_device: audio already resident in device memory.
 * Results stay in the engine until the next run; fetch with pm_engine_get_packets.
 */
int pm_engine_run(pm_engine *e, const int16_t *audio_host, int64_t n_samples);
int pm_engine_run_device(pm_engine *e, const int16_t *audio_dev, int64_t n_samples);

/*
 * Batched form: R recordings x the chains that name them (pm_chain_desc.recording = 0 .. R-1) in one engine call.  Row r of
 * audio_host ([R][row_stride] int16, host memory) holds recording r, n_samples[r] <= row_stride of it valid.  Every
 * stage runs over all chains of all recordings at once -- this is what fills a B200 when one recording cannot: the
 * carrier-loop modems (psk.py:162-195, 705-773; afsk_pll.py:140-170) are sequential per chain, so 64 short recordings
 * x their chains are 64+ loops side by side.  Records come back ordered by chain index as usual.
 */
int pm_engine_run_batch(pm_engine *e, const int16_t *audio_host, int64_t row_stride, const int64_t *n_samples,
                        int32_t n_recordings);

/*
 * Sharded form used by the multi-GPU host (one engine per rank; ONE recording is
 * split on the sample axis; pymodem_b200/sharded.py drives it).  Positions are in
 * soft samples (the demod output the slicer addresses, slicer.py:75): global soft
 * sample = local + sample_base, and audio[i] is the recording's sample
 * sample_base + i.  A rank owns soft samples [own_begin, own_begin + own_len) of
 * its local buffer; the samples before own_begin are slicer history/warm-up (and
 * FIR history), the ones after the own range let the last stream byte complete.
 * own_begin and own_len must be multiples of the segment length (option
 * "segment_len"), except own_len on the last shard.
 *
 *   begin   : front end + slicer over the local buffer, speculating the state at
 *             own_begin from the warm-up (true start state on the first shard)
 *   handoff : given the previous shard's state, verify the speculation bit for
 *             bit and repair if it was wrong; *changed tells whether this shard's
 *             own end state / symbol count changed (then the next shard has to
 *             look again).  Repeat over all shards until nothing changes.
 *   gather  : place the own bits at their global byte alignment
 *             (symbols_before = symbols of all earlier shards, per chain) and
 *             return the last tail_bits own bits of every chain
 *   finish  : prepend the previous shard's tail, descramble, decode; a packet is
 *             emitted by the shard that holds its closing bit
 * PM_ERR_STATE from finish means a frame reached back past the hand-off tail or a
 * gap is long enough to overflow max_packet_length (ax25.py:46-51; the sequential
 * replay needs a known decoder state).  The shards then recover together: every
 * rank exports what it holds of the sliced stream (pm_engine_shard_export), the
 * exports are gathered, and the complete stream is decoded by
 * pm_engine_unscramble_stream / pm_engine_decode_stream -- exactly what an
 * unsharded run decodes (pymodem_b200/sharded.py: recover_from_bitstream).
 */
typedef struct pm_shard_plan {
	int64_t sample_base;
	int64_t own_begin;
	int64_t own_len;
	int32_t first;                  /* shard holds the start of the recording */
	int32_t last;                   /* shard holds the end of the recording */
	int32_t tail_bits;              /* hand-off tail per chain, multiple of 32 */
	int32_t pre_segments;           /* whole segments of slicer history processed (and verified) before own_begin, on
	                                   top of the warm-up of the first of them: the speculated state at own_begin then
	                                   rests on (pre_segments + 1) x warm-up of samples; needs own_begin >=
	                                   pre_segments * segment_len.  0 on the first shard. */
} pm_shard_plan;

typedef struct pm_shard_state {     /* one per chain */
	double   start_clock;           /* slicer.py:50 phase_clock reached at own_begin */
	double   end_clock;             /* ... at own_begin + own_len */
	uint32_t start_last, start_last_q;   /* sign (1: >= 0) of slicer.py:56 last_sample at own_begin */
	uint32_t end_last, end_last_q;
	int64_t  n_symbols;             /* symbols taken inside the own range */
} pm_shard_state;

int pm_engine_shard_begin(pm_engine *e, const int16_t *audio, int64_t n_samples, int32_t audio_on_device,
                          const pm_shard_plan *plan, pm_shard_state *out);
int pm_engine_shard_handoff(pm_engine *e, const pm_shard_state *prev, pm_shard_state *out, int32_t *changed);
int pm_engine_shard_gather(pm_engine *e, const int64_t *symbols_before, uint32_t *tail_out);
int pm_engine_shard_finish(pm_engine *e, const uint32_t *tail_in);

/*
 * IL2P chains keep decoder state across a shard boundary (is the decoder searching or inside a frame, il2p.py:360-519;
 * corrected-byte counts of failed frames, il2p.py:200-211), so their shards finish ONE AFTER THE OTHER: rank r calls
 * shard_finish_il2p with the state rank r-1 returned (NULL on the first shard) and passes its own on.  Positions are
 * global stream bits.  A frame is decoded by the shard that holds its last bit; the previous shard's tail_bits must
 * cover the longest frame (a 1023-byte payload is 10 552 bits on the air).  Engines without IL2P chains may use
 * either call.
 */
typedef struct pm_il2p_state {      /* one per chain (ignored for AX.25 chains) */
	int64_t  pos;                   /* the sync search (re)starts at this stream bit */
	uint32_t mode;                  /* 0 start of recording, 1 right after a frame, 2 plain search */
	uint32_t leak;                  /* corrected bytes of failed frames since the last emitted packet */
} pm_il2p_state;
int pm_engine_shard_finish_il2p(pm_engine *e, const uint32_t *tail_in, const pm_il2p_state *prev, pm_il2p_state *out);

/*
 * Shard link: the same hand-off carried out by the GPUs themselves over NVLink peer memory (csrc/link.cu); IL2P chains
 * included: the pm_il2p_state of the previous rank arrives through the link buffer instead of pm_engine_shard_finish_il2p.
 * Every rank creates a link buffer, the ranks exchange the 64-byte CUDA IPC handles once (any transport) and
 * map each other's buffers.  run_linked_begin then enqueues the whole sharded run -- front end, slicer, state
 * push, bit placement, tail push/wait, decode, record push, merge -- on the engine's stream with no host round
 * trip after the slicer; run_linked_end waits for it and leaves the MERGED records of all ranks (ordered like an
 * unsharded run: chain, then stream position; their `offset` fields point into an arena that holds rank 0's packet
 * bytes first, then rank 1's, ...) in the engine for pm_engine_get_packets.
 *   *verified == 2: some rank could not finish its decode from what it holds (see PM_ERR_STATE above).  Every rank
 *   learns it with the records, so all of them get 2 and recover through pm_engine_shard_export.
 *   *verified == 0: some rank's speculated slicer start state was wrong.  All ranks see the same states, so all
 *   of them get 0 and continue with the host-driven protocol: pm_engine_shard_states (what shard_begin would
 *   have returned) -> shard_handoff ... -> shard_gather -> shard_finish.
 * use_ipc = 1: `handles` is world x 64 bytes of cudaIpcMemHandle_t (one process per GPU);
 * use_ipc = 0: `handles` is world device pointers (several engines in one process, e.g. tests on one GPU).
 * A rank that waits more than a few seconds for a peer gives up with PM_ERR_STATE instead of hanging.
 */
int pm_engine_link_create(pm_engine *e, int32_t rank, int32_t world, int32_t tail_bits, int64_t max_samples,
                          void *ipc_handle_out /* 64 bytes */, void **base_out);
int pm_engine_link_connect(pm_engine *e, const void *handles, int32_t use_ipc);
int pm_engine_run_linked_begin(pm_engine *e, const int16_t *audio, int64_t n_samples, int32_t audio_on_device,
                               const pm_shard_plan *plan);
int pm_engine_run_linked_end(pm_engine *e, int32_t *verified);
int pm_engine_shard_states(pm_engine *e, pm_shard_state *out);
/* What this shard holds of one chain's sliced stream after shard_gather (also after a finish that failed, or a linked
 * run that ended with *verified == 2): bits[(info4[0] + 31) / 32] packed LSB-first in time, byte_addr[(info4[0] + 7) / 8]
 * local sample addresses of the stream bytes (valid for the bytes whose last bit is an own bit; add info4[3]).
 * info4 = {local stream bits, local position of the first own bit, own bits, sample_base}.  bits == byte_addr == NULL:
 * only info4 is filled. */
int pm_engine_shard_export(pm_engine *e, int32_t chain, uint32_t *bits, int64_t cap_words, uint32_t *byte_addr,
                           int64_t cap_bytes, int64_t *info4);

/*
 * Per-stage entry points: the reference's duck-typed blocks one at a time (chain_execute.py:32-47), on the same
 * kernels as the whole-chain run, for ONE chain of the loaded table.
 *   slice_soft         slicer.slice(soft)  slicer.py:59-107 / 193-242: float64 soft values (I, and Q for the quadrature
 *                      slicer) -> the AddressedData stream, read back with pm_engine_get_stream(stage 0)
 *   unscramble_stream  stream.stream_unscramble_8bit(list[AddressedData])  lfsr.py:22-52 -> pm_engine_get_stream(stage 1)
 *   decode_stream      codec.decode(list[AddressedData])  ax25.py:25-93 / il2p.py:360-519 (incl. the sequential replay
 *                      after a max_packet_length overflow) -> pm_engine_get_packets
 * bytes[i] / addresses[i] = AddressedData.data / .address; addresses must fit 32 bits.
 */
int pm_engine_slice_soft(pm_engine *e, int32_t chain, const double *soft_i, const double *soft_q, int64_t n);
int pm_engine_unscramble_stream(pm_engine *e, int32_t chain, const uint8_t *bytes, const int64_t *addresses, int64_t n);
int pm_engine_decode_stream(pm_engine *e, int32_t chain, const uint8_t *bytes, const int64_t *addresses, int64_t n);

int64_t pm_engine_num_packets(const pm_engine *e);
int64_t pm_engine_arena_bytes(const pm_engine *e);
/* Records are ordered by (chain, position in stream) == the order the
 * reference's per-chain decode() lists have. */
int pm_engine_get_packets(const pm_engine *e, pm_packet_rec *recs, int64_t rec_cap,
                          uint8_t *arena, int64_t arena_cap);

/* Intermediates for parity tests (need option keep_soft=1 for the soft values). */
int64_t pm_engine_soft_len(const pm_engine *e, int32_t chain);
int pm_engine_get_soft(const pm_engine *e, int32_t chain, int32_t component /*0=I/real,1=Q*/,
                       float *out, int64_t cap);
/* The packed signs of the soft values (bit i of word w = soft[32 w + i] >= 0; slicer.py:85, 99-102 read nothing else of
 * the demod output), after the float64 fix-up: (soft_len + 31) / 32 words.  Lets a test compare the FP32 front end's
 * signs with the float64 route's (option "precise") over whole recordings. */
int pm_engine_get_signs(const pm_engine *e, int32_t chain, int32_t component, uint32_t *out, int64_t cap_words);
/* AddressedData streams: stage 0 = slicer output (slicer.py:97), 1 = after LFSR (lfsr.py:47-51). */
int64_t pm_engine_stream_len(const pm_engine *e, int32_t chain);
int pm_engine_get_stream(const pm_engine *e, int32_t chain, int32_t stage,
                         uint8_t *bytes, int64_t *addresses, int64_t cap);

int pm_engine_get_stats(const pm_engine *e, pm_stats *out);
/* Per-kernel times of the last pm_engine_run / pm_engine_run_device when option "kernel_times" was 1: a CUDA event is
 * recorded before every kernel launch and copy of the run; buf receives one line per kernel name,
 * "name<TAB>launches<TAB>milliseconds" (the time up to the next launch on the stream), in order of first launch.  Returns
 * the bytes written or PM_ERR_CAPACITY.  The events cost a few microseconds each: a timing pass is not a benchmark run. */
int64_t pm_engine_kernel_times(const pm_engine *e, char *buf, int64_t cap);
/* Timeline of the last pm_engine_run on a host buffer when option "trace" was 1: "label milliseconds" lines (chunk i
 * copied / its front end done / its guard fix-up done / its batch of slicer segments done), relative to the start of the
 * run.  Synchronises the device.  Returns the length of the text (it is cut at cap - 1). */
int64_t pm_engine_trace(pm_engine *e, char *buf, int64_t cap);

/* FP32 FFMA peak microbenchmark (roofline denominator for the FIR kernels):
 * returns achieved TFLOP/s of a register-resident FFMA loop on the device. */
int pm_measure_fp32_peak(int device, double *tflops);

/* Latency of one dependent float64 operation on the device (alternating add / multiply on its own result, one thread):
 * the floor under the sequential carrier loops of psk.py:173-189, 727-747 and afsk_pll.py:147-166, whose operations per
 * sample form one dependency chain.  bench.py --config reports a loop's measured time per sample against it. */
int pm_measure_fp64_chain(int device, double *ns_per_op, double *cycles_per_op);

/* Executed FP32 multiply-adds per input sample over all loaded chains (after
 * the sharing of common filter passes) and the tile length of a front-end
 * launch group -- reported by bench.py next to the roofline. */
double pm_engine_front_macs_per_sample(const pm_engine *e);
/* Executed bf16 multiply-adds per input sample on the tensor cores when the AFSK low-pass runs there (option
 * "tensor_lpf", default on where a chain group qualifies: every tone pair a sliding-window pair, at most 4 tones, 8..113
 * low-pass taps): three bf16 pieces per operand, six piece products, K padded to the Toeplitz band -- about eleven
 * times the useful multiply-adds, on a pipe with thirty times the FP32 rate.  0 when no group takes that route. */
double pm_engine_front_tensor_macs_per_sample(const pm_engine *e);
/* The useful multiply-adds per input sample of the AFSK low-pass (2 x tone pairs x taps): what the FP32 route executes
 * for it and what the tensor-core figure above should be compared with. */
double pm_engine_front_lpf_macs_per_sample(const pm_engine *e);
int pm_engine_front_tile(const pm_engine *e, int group);

/* Host-side helper (no device): 1 when the correlator taps (i[k], q[k]), k < n, are a rotation a*e^{i(phi + step*k)} with
 * n >= 16 -- the form afsk.py:134-144 generates -- which is what lets the engine compute the tone magnitudes as sliding
 * window sums; *step receives the angle per tap.  0 otherwise (the engine then runs the taps as plain FIRs). */
int pm_taps_are_rotation(const double *i, const double *q, int32_t n, double *step);

/* Tracing (option "stage_clocks" = 1): SM cycles per stage of the AFSK front-end kernel, summed over the CTAs launched
 * since the last call: out8[0..3] = staging, band-pass, tone correlators, low-pass + epilogue (each including the
 * wait for the CTA's slowest warp), out8[4] = number of CTAs.  Resets the counters. */
int pm_engine_stage_clocks(pm_engine *e, uint64_t *out8);

/* Host memory the GPU can read directly.  pm_engine_run takes any host pointer: from pinned memory (pm_host_alloc,
 * pm_host_register, torch pin_memory ...) the chunks are DMA'ed in place; from pageable memory they are staged through
 * the engine's ring of pinned buffers by a few host threads (options "h2d_chunk", "copy_threads") -- correct but bounded
 * by the host's memcpy rate.  pm_host_alloc: read a recording straight into pinned memory (what python -m pymodem_b200
 * does with the WAV, pymodem.py:46).  pm_host_register: pin a buffer the caller already owns, once, for repeated runs
 * (unregister before freeing it). */
void *pm_host_alloc(size_t bytes);
void pm_host_free(void *p);
int pm_host_register(void *p, size_t bytes);
int pm_host_unregister(void *p);

const char *pm_version(void);

#ifdef __cplusplus
}
#endif
#endif

"""ctypes binding of libpymodem_b200.so (include/pymodem_b200.h).

There is no CPU fallback: importing this module without the built library, or
creating an engine without an sm_100 GPU, raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpymodem_b200.so")

PM_OK, PM_ERR_ARG, PM_ERR_CUDA, PM_ERR_UNSUPPORTED, PM_ERR_CAPACITY, PM_ERR_STATE = 0, -1, -2, -3, -4, -5
PM_MODEM_NONE, PM_MODEM_AFSK, PM_MODEM_FSK, PM_MODEM_BPSK, PM_MODEM_MPSK, PM_MODEM_AFSK_PLL = 0, 1, 2, 3, 4, 5
PM_SLICER_BINARY, PM_SLICER_QUADRATURE = 1, 2
PM_CODEC_AX25, PM_CODEC_IL2P = 1, 2

_dp = ctypes.POINTER(ctypes.c_double)


class LoopDesc(ctypes.Structure):
	"""pm_loop_desc"""
	_fields_ = [(n, ctypes.c_double) for n in (
		"agc_scaled_attack", "agc_scaled_decay", "agc_sustain_time", "agc_sustain_increment", "agc_target",
		"nco_phase_scale", "nco_index_scale", "nco_set_frequency", "nco_two_pi", "nco_quarter")] + [
		("nco_wavetable", _dp), ("nco_size", ctypes.c_int64),
		("iir_b0", ctypes.c_double), ("iir_b1", ctypes.c_double), ("iir_a1", ctypes.c_double),
		("pi_gain", ctypes.c_double), ("pi_p", ctypes.c_double), ("pi_i", ctypes.c_double),
		("pi_limit", ctypes.c_double), ("pi_integral0", ctypes.c_double),
		("pd_table", ctypes.POINTER(ctypes.c_int32)), ("pd_granularity", ctypes.c_int64),
		("hilbert", _dp), ("n_hilbert", ctypes.c_int32), ("hilbert_delay", ctypes.c_int32)]


class ChainDesc(ctypes.Structure):
	"""pm_chain_desc"""
	_fields_ = [
		("modem_kind", ctypes.c_int32), ("slicer_kind", ctypes.c_int32),
		("codec_kind", ctypes.c_int32), ("invert_soft", ctypes.c_int32),
		("bpf", _dp), ("n_bpf", ctypes.c_int32), ("n_corr", ctypes.c_int32),
		("mark_i", _dp), ("mark_q", _dp), ("space_i", _dp), ("space_q", _dp),
		("space_unit_i", _dp), ("space_unit_q", _dp), ("space_gain", ctypes.c_double),
		("lpf", _dp), ("n_lpf", ctypes.c_int32), ("recording", ctypes.c_int32),
		("slicer_sample_rate", ctypes.c_double), ("symbol_rate", ctypes.c_double),
		("lock_rate", ctypes.c_double), ("state_mask", ctypes.c_uint32),
		("bits_per_symbol", ctypes.c_uint32), ("demap", ctypes.c_uint32 * 16),
		("lfsr_poly", ctypes.c_uint64), ("lfsr_invert", ctypes.c_int32),
		("il2p_crc", ctypes.c_int32), ("il2p_disable_rs", ctypes.c_int32),
		("il2p_min_dist", ctypes.c_int32), ("il2p_sync_tol", ctypes.c_int32),
		("reserved1", ctypes.c_int32), ("loop", ctypes.POINTER(LoopDesc)),
	]


class PacketRec(ctypes.Structure):
	"""pm_packet_rec"""
	_fields_ = [
		("chain", ctypes.c_uint32), ("len", ctypes.c_uint32), ("offset", ctypes.c_uint64),
		("streamaddress", ctypes.c_int64), ("bytes_corrected", ctypes.c_uint32),
		("calculated_crc", ctypes.c_uint16), ("carried_crc", ctypes.c_uint16),
		("valid_crc", ctypes.c_uint8), ("valid_header", ctypes.c_uint8), ("pad", ctypes.c_uint8 * 6),
	]


class Stats(ctypes.Structure):
	"""pm_stats"""
	_fields_ = [(n, ctypes.c_double) for n in
		("total_ms", "h2d_ms", "front_ms", "fixup_ms", "slicer_ms", "bits_ms", "d2h_ms")] + \
		[(n, ctypes.c_int64) for n in
		("kernel_launches", "front_launches", "guard_flagged", "slicer_repairs", "slicer_segments",
		 "h2d_bytes", "d2h_bytes", "n_packets", "n_stream_bits")]

	def as_dict(self):
		return {n: getattr(self, n) for n, _ in self._fields_}


class ShardPlan(ctypes.Structure):
	"""pm_shard_plan"""
	_fields_ = [
		("sample_base", ctypes.c_int64), ("own_begin", ctypes.c_int64), ("own_len", ctypes.c_int64),
		("first", ctypes.c_int32), ("last", ctypes.c_int32), ("tail_bits", ctypes.c_int32), ("pre_segments", ctypes.c_int32),
	]


class Il2pState(ctypes.Structure):
	"""pm_il2p_state"""
	_fields_ = [("pos", ctypes.c_int64), ("mode", ctypes.c_uint32), ("leak", ctypes.c_uint32)]


class ShardState(ctypes.Structure):
	"""pm_shard_state"""
	_fields_ = [
		("start_clock", ctypes.c_double), ("end_clock", ctypes.c_double),
		("start_last", ctypes.c_uint32), ("start_last_q", ctypes.c_uint32),
		("end_last", ctypes.c_uint32), ("end_last_q", ctypes.c_uint32), ("n_symbols", ctypes.c_int64),
	]


# every symbol include/pymodem_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_char_p
PROTOTYPES = {
	"pm_engine_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
	"pm_engine_destroy": (None, [_vp]),
	"pm_last_error": (_cp, [_vp]),
	"pm_engine_load_chains": (ctypes.c_int, [_vp, ctypes.POINTER(ChainDesc), _i32]),
	"pm_engine_set_option": (ctypes.c_int, [_vp, _cp, ctypes.c_double]),
	"pm_engine_run": (ctypes.c_int, [_vp, _vp, _i64]),
	"pm_engine_run_device": (ctypes.c_int, [_vp, _vp, _i64]),
	"pm_engine_run_batch": (ctypes.c_int, [_vp, _vp, _i64, ctypes.POINTER(_i64), _i32]),
	"pm_engine_shard_begin": (ctypes.c_int, [_vp, _vp, _i64, _i32, ctypes.POINTER(ShardPlan), ctypes.POINTER(ShardState)]),
	"pm_engine_shard_handoff": (ctypes.c_int, [_vp, ctypes.POINTER(ShardState), ctypes.POINTER(ShardState), ctypes.POINTER(_i32)]),
	"pm_engine_shard_gather": (ctypes.c_int, [_vp, ctypes.POINTER(_i64), _vp]),
	"pm_engine_shard_finish": (ctypes.c_int, [_vp, _vp]),
	"pm_engine_shard_finish_il2p": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(Il2pState), ctypes.POINTER(Il2pState)]),
	"pm_engine_link_create": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i64, _vp, ctypes.POINTER(_vp)]),
	"pm_engine_link_connect": (ctypes.c_int, [_vp, _vp, _i32]),
	"pm_engine_run_linked_begin": (ctypes.c_int, [_vp, _vp, _i64, _i32, ctypes.POINTER(ShardPlan)]),
	"pm_engine_run_linked_end": (ctypes.c_int, [_vp, ctypes.POINTER(_i32)]),
	"pm_engine_shard_states": (ctypes.c_int, [_vp, ctypes.POINTER(ShardState)]),
	"pm_engine_shard_export": (ctypes.c_int, [_vp, _i32, _vp, _i64, _vp, _i64, ctypes.POINTER(_i64)]),
	"pm_engine_slice_soft": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i64]),
	"pm_engine_unscramble_stream": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i64]),
	"pm_engine_decode_stream": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i64]),
	"pm_engine_get_signs": (ctypes.c_int, [_vp, _i32, _i32, _vp, _i64]),
	"pm_engine_num_packets": (_i64, [_vp]),
	"pm_engine_arena_bytes": (_i64, [_vp]),
	"pm_engine_get_packets": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64]),
	"pm_engine_soft_len": (_i64, [_vp, _i32]),
	"pm_engine_get_soft": (ctypes.c_int, [_vp, _i32, _i32, _vp, _i64]),
	"pm_engine_stream_len": (_i64, [_vp, _i32]),
	"pm_engine_get_stream": (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _i64]),
	"pm_engine_get_stats": (ctypes.c_int, [_vp, ctypes.POINTER(Stats)]),
	"pm_engine_kernel_times": (_i64, [_vp, ctypes.c_char_p, _i64]),
	"pm_engine_trace": (_i64, [_vp, ctypes.c_char_p, _i64]),
	"pm_engine_front_macs_per_sample": (ctypes.c_double, [_vp]),
	"pm_engine_front_tensor_macs_per_sample": (ctypes.c_double, [_vp]),
	"pm_engine_front_lpf_macs_per_sample": (ctypes.c_double, [_vp]),
	"pm_engine_stage_clocks": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
	"pm_taps_are_rotation": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int32,
		ctypes.POINTER(ctypes.c_double)]),
	"pm_engine_front_tile": (ctypes.c_int, [_vp, ctypes.c_int]),
	"pm_measure_fp32_peak": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
	"pm_measure_fp64_chain": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
	"pm_host_alloc": (_vp, [ctypes.c_size_t]),
	"pm_host_free": (None, [_vp]),
	"pm_host_register": (ctypes.c_int, [_vp, ctypes.c_size_t]),
	"pm_host_unregister": (ctypes.c_int, [_vp]),
	"pm_version": (_cp, []),
}

_lib = None


def load():
	"""Load the CUDA library; raises (never falls back) when it is missing."""
	global _lib
	if _lib is None:
		if not os.path.exists(LIB_PATH):
			raise ImportError(
				f"{LIB_PATH} is missing: build it with `python -m pymodem_b200.build` "
				"(there is no CPU fallback for the demod_chain path)")
		lib = ctypes.CDLL(LIB_PATH)
		for name, (res, args) in PROTOTYPES.items():
			fn = getattr(lib, name)      # AttributeError if the library does not export it
			fn.restype = res
			fn.argtypes = args
		_lib = lib
	return _lib

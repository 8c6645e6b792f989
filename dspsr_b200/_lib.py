"""ctypes binding of libb200dsp.so (include/b200dsp.h).  No CPU fallback: a missing library or a
missing GPU is an error, never a silent detour."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libb200dsp.so")   # B200_LIB: developer override (A/B builds)

OK = 0
FMT_CASPSR8, FMT_GENERIC8, FMT_MEERKAT8, FMT_UWB16, FMT_FLOAT32, FMT_TWOBIT = range(6)
INTENSITY, PPQQ, COHERENCE, STOKES = range(4)
STATE = {"Intensity": INTENSITY, "PPQQ": PPQQ, "Coherence": COHERENCE, "Stokes": STOKES}


class TwoBitDesc(C.Structure):
    _fields_ = [
        ("table_type", C.c_int),
        ("npol", C.c_uint),
        ("ndat_per_weight", C.c_uint),
        ("nlow_min", C.c_uint),
        ("nlow_max", C.c_uint),
        ("lo", C.c_float * 513),
        ("hi", C.c_float * 513),
    ]


class TimeDivide(C.Structure):
    _fields_ = [("division_seconds", C.c_double), ("lower", C.c_double), ("upper", C.c_double),
                ("current_end", C.c_double), ("is_valid", C.c_int), ("division", C.c_uint64)]


class TimeBounds(C.Structure):
    _fields_ = [("is_valid", C.c_int), ("new_division", C.c_int), ("end_reached", C.c_int), ("in_next", C.c_int),
                ("idat_start", C.c_uint64), ("ndat", C.c_uint64), ("division", C.c_uint64)]


class UnpackDesc(C.Structure):
    _fields_ = [
        ("format", C.c_int),
        ("nchan", C.c_uint),
        ("npol", C.c_uint),
        ("ndim", C.c_uint),
        ("lut", C.c_float * 256),
        ("scale", C.c_float),
        ("sample_swap", C.c_uint),
        ("twobit", C.POINTER(TwoBitDesc)),
    ]


class FbDesc(C.Structure):
    _fields_ = [
        ("input_real", C.c_int),
        ("input_nchan", C.c_uint),
        ("npol", C.c_uint),
        ("nchan_subband", C.c_uint),
        ("freq_res", C.c_uint),
        ("nfilt_pos", C.c_uint),
        ("nfilt_neg", C.c_uint),
        ("h_response", C.c_void_p),
        ("max_npart", C.c_uint),
    ]


class FbInfo(C.Structure):
    _fields_ = [
        ("n_fft", C.c_uint),
        ("nsamp_fft", C.c_uint),
        ("nsamp_overlap", C.c_uint),
        ("nsamp_step", C.c_uint),
        ("nkeep", C.c_uint),
        ("fft_rows", C.c_uint),
        ("fft_cols", C.c_uint),
        ("batch_npart", C.c_uint),
        ("scratch_bytes", C.c_uint64),
    ]


class PipelineDesc(C.Structure):
    _fields_ = [
        ("unpack", UnpackDesc),
        ("fb", FbDesc),
        ("detect_state", C.c_int),
        ("detect_ndim", C.c_uint),
        ("nbin", C.c_uint),
    ]


class Dedispersion(C.Structure):
    _fields_ = [
        ("centre_frequency", C.c_double),
        ("bandwidth", C.c_double),
        ("dispersion_measure", C.c_double),
        ("input_nchan", C.c_uint),
        ("nchan", C.c_uint),
        ("input_dual_sideband", C.c_int),
        ("input_dc_centred", C.c_int),
        ("input_swap", C.c_int),
        ("frequency_resolution", C.c_uint),
        ("impulse_pos", C.c_uint),
        ("impulse_neg", C.c_uint),
        ("ndat", C.c_uint),
    ]


class Polyco(C.Structure):
    _fields_ = [
        ("tmid_day", C.c_int),
        ("tmid_sec", C.c_double),
        ("rphase_int", C.c_double),
        ("rphase_frac", C.c_double),
        ("f0", C.c_double),
        ("span_min", C.c_double),
        ("obsfreq", C.c_double),
        ("dm", C.c_double),
        ("ncoef", C.c_int),
        ("coef", C.c_double * 32),
    ]


class Mjd(C.Structure):
    _fields_ = [("day", C.c_int), ("sec", C.c_int), ("frac", C.c_double)]


class Observation(C.Structure):
    _fields_ = [
        ("telescope", C.c_char * 32), ("receiver", C.c_char * 32), ("source", C.c_char * 32),
        ("mode", C.c_char * 32), ("machine", C.c_char * 32), ("format", C.c_char * 32),
        ("centre_frequency", C.c_double), ("bandwidth", C.c_double), ("rate", C.c_double), ("scale", C.c_double),
        ("dispersion_measure", C.c_double), ("rotation_measure", C.c_double),
        ("nchan", C.c_uint), ("npol", C.c_uint), ("ndim", C.c_uint), ("nbit", C.c_uint),
        ("state", C.c_int), ("type", C.c_int), ("basis", C.c_int), ("swap", C.c_int), ("nsub_swap", C.c_int),
        ("dc_centred", C.c_int),
        ("start_time", Mjd),
        ("ndat", C.c_uint64),
    ]


class PhaseSeries(C.Structure):
    _fields_ = [
        ("obs", Observation),
        ("nbin", C.c_uint), ("hits_nchan", C.c_uint),
        ("integration_length", C.c_double),
        ("ndat_total", C.c_uint64), ("ndat_expected", C.c_uint64),
        ("end_time", Mjd),
        ("folding_period", C.c_double), ("reference_phase", C.c_double),
        ("data", C.c_void_p),
        ("hits", C.c_void_p),
    ]


class SigprocHeader(C.Structure):
    _fields_ = [
        ("rawdatafile", C.c_char * 80), ("source_name", C.c_char * 80),
        ("machine_id", C.c_int), ("telescope_id", C.c_int), ("nchans", C.c_int), ("nbits", C.c_int), ("nifs", C.c_int),
        ("nbeams", C.c_int), ("ibeam", C.c_int),
        ("src_raj", C.c_double), ("src_dej", C.c_double), ("az_start", C.c_double), ("za_start", C.c_double),
        ("fch1", C.c_double), ("foff", C.c_double), ("tstart", C.c_double), ("tsamp", C.c_double),
    ]


class PhaseSegment(C.Structure):
    _fields_ = [
        ("start", C.c_uint64),
        ("count", C.c_uint64),
        ("a0", C.c_uint64),
        ("step", C.c_uint64),
        ("scale_exp", C.c_int),
        ("pad", C.c_int),
    ]


_vp, _u64, _u, _i, _d = C.c_void_p, C.c_uint64, C.c_uint, C.c_int, C.c_double
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/b200dsp.h declares
SIGNATURES = {
    "b200_rescale_create": (_i, [_vp, C.c_uint, C.c_uint, C.c_uint64, _i, C.POINTER(C.c_void_p)]),
    "b200_rescale_destroy": (_i, [_vp]),
    "b200_rescale_transform": (_i, [_vp, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64]),
    "b200_rescale_get": (_i, [_vp, _vp, _vp]),
    "b200_sigproc_digitize8": (_i, [_vp, _vp, C.c_uint64, C.c_uint, C.c_uint, C.c_uint64, C.c_float, C.c_float,
                                    C.c_float, _i, _i, _vp]),
    "b200_time_divide_init": (_i, [C.POINTER(TimeDivide), C.c_double]),
    "b200_time_divide_set_bounds": (_i, [C.POINTER(TimeDivide), C.c_double, C.c_double, C.c_uint64,
                                         C.POINTER(TimeBounds)]),
    "b200_twobit_prepare": (_i, [C.c_double, C.c_float, _i, C.c_uint, C.c_uint, C.POINTER(TwoBitDesc)]),
    "b200_unpack_twobit": (_i, [_vp, C.POINTER(TwoBitDesc), _vp, C.c_uint64, _vp, C.c_uint64, _vp]),
    "b200_version": (_i, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_context_create": (_i, [_i, _vp, _pvp]),
    "b200_context_destroy": (_i, [_vp]),
    "b200_context_synchronize": (_i, [_vp]),
    "b200_context_launch_count": (C.c_ulonglong, [_vp]),
    "b200_context_stream": (_vp, [_vp]),
    "b200_context_set_timing": (_i, [_vp, _i]),
    "b200_context_read_timing": (_i, [_vp, C.POINTER(_d), C.POINTER(C.c_ulonglong)]),
    "b200_malloc": (_i, [_vp, _u64, _pvp]),
    "b200_free": (_i, [_vp, _vp]),
    "b200_malloc_host": (_i, [_vp, _u64, _pvp]),
    "b200_free_host": (_i, [_vp, _vp]),
    "b200_memset": (_i, [_vp, _vp, _i, _u64]),
    "b200_memcpy_h2d": (_i, [_vp, _vp, _vp, _u64]),
    "b200_memcpy_d2d": (_i, [_vp, _vp, _vp, _u64]),
    "b200_memcpy_d2h": (_i, [_vp, _vp, _vp, _u64]),
    "b200_unpack": (_i, [_vp, C.POINTER(UnpackDesc), _vp, _u64, _vp, _u64]),
    "b200_fb_plan_create": (_i, [_vp, C.POINTER(FbDesc), _pvp]),
    "b200_fb_plan_info": (_i, [_vp, C.POINTER(FbInfo)]),
    "b200_fb_plan_destroy": (_i, [_vp]),
    "b200_fb_perform": (_i, [_vp, _vp, _u64, _vp, _u64, _u64, _u64, _u64]),
    "b200_detect": (_i, [_vp, _i, _u, _vp, _u64, _u, _u, _u64, _vp, _u64]),
    "b200_fold_create": (_i, [_vp, _u, _u, _u, _u, _pvp]),
    "b200_fold_destroy": (_i, [_vp]),
    "b200_fold_set_bins": (_i, [_vp, _d, _d, _u64, _u64, C.POINTER(_u64)]),
    "b200_fold_set_bins_weighted": (_i, [_vp, _d, _d, _u64, _u64, _vp, _u64, _u, _u64]),
    "b200_fold_weighted": (_i, [_vp]),
    "b200_fold_set_deterministic": (_i, [_vp, C.c_float]),
    "b200_pipeline_set_deterministic": (_i, [_vp, C.c_float]),
    "b200_fold_get_bin_hits": (_i, [_vp, _vp]),
    "b200_fold_fold": (_i, [_vp, _vp, _u64]),
    "b200_fold_fold_into": (_i, [_vp, _vp, _u64, _vp, _u64]),
    "b200_fold_synch": (_i, [_vp, _vp]),
    "b200_fold_get_hits": (_i, [_vp, _vp, C.POINTER(_u64)]),
    "b200_fold_zero": (_i, [_vp]),
    "b200_fold_device_profile": (_vp, [_vp]),
    "b200_fold_device_hits": (_vp, [_vp]),
    "b200_pipeline_create": (_i, [_vp, C.POINTER(PipelineDesc), _pvp]),
    "b200_pipeline_destroy": (_i, [_vp]),
    "b200_pipeline_reserve": (_i, [_vp, _u64]),
    "b200_weights_convolve": (_i, [_vp, _vp, _u64, _u, _u64, _u64, _u, _u, _vp, _vp]),
    "b200_weights_scrunch": (_i, [_vp, _vp, C.POINTER(_u64), C.POINTER(_u), C.POINTER(_u64), _u, _vp]),
    "b200_pipeline_info": (_i, [_vp, C.POINTER(FbInfo)]),
    "b200_pipeline_execute": (_i, [_vp, _vp, _u64, _u64, _u64, _d, _d, _vp, _u64]),
    "b200_pipeline_execute_host": (_i, [_vp, _vp, _u64, _u64, _u64, _d, _d, _vp, _u64]),
    "b200_pipeline_input_consumed": (_i, [_vp]),
    "b200_pipeline_synch": (_i, [_vp, _vp, _vp, C.POINTER(_u64)]),
    "b200_pipeline_zero": (_i, [_vp]),
    "b200_pipeline_fold": (_vp, [_vp]),
    "b200_pipeline_set_observation": (_i, [_vp, C.POINTER(Observation)]),
    "b200_pipeline_set_predictor": (_i, [_vp, C.POINTER(Polyco), _d]),
    "b200_pipeline_set_folding_period": (_i, [_vp, _d, _d, C.POINTER(Mjd)]),
    "b200_pipeline_execute_obs": (_i, [_vp, _vp, _u64, _u64, _u64, _u64]),
    "b200_pipeline_execute_host_obs": (_i, [_vp, _vp, _u64, _u64, _u64, _u64]),
    "b200_pipeline_stream_begin": (_i, [_vp, _u64, _u64]),
    "b200_pipeline_feed_host": (_i, [_vp, _vp, _u64, _vp, _u64, C.POINTER(_u64)]),
    "b200_pipeline_feed": (_i, [_vp, _vp, _u64, _vp, _u64, C.POINTER(_u64)]),
    "b200_pipeline_get_phase_series": (_i, [_vp, C.POINTER(PhaseSeries)]),
    "b200_pipeline_reset": (_i, [_vp]),
    "b200_bittable8": (_i, [_i, _vp, C.POINTER(_d)]),
    "b200_dedispersion_prepare": (_i, [C.POINTER(Dedispersion)]),
    "b200_dedispersion_build": (_i, [C.POINTER(Dedispersion), _vp]),
    "b200_dedispersion_build_channels": (_i, [C.POINTER(Dedispersion), C.c_uint, C.c_uint, _vp]),
    "b200_optimal_fft_length": (C.c_int64, [_u64, _u64]),
    "b200_polyco_parse": (_i, [C.c_char_p, C.POINTER(Polyco)]),
    "b200_polyco_phase": (_d, [C.POINTER(Polyco), _i, _i, _d, C.POINTER(_d)]),
    "b200_polyco_frequency": (_d, [C.POINTER(Polyco), _i, _i, _d]),
    "b200_observation_combinable": (_i, [C.POINTER(Observation), C.POINTER(Observation), C.c_char_p, C.c_uint]),
    "b200_mjd_diff": (_d, [C.POINTER(Mjd), C.POINTER(Mjd)]),
    "b200_mjd_add": (Mjd, [C.POINTER(Mjd), _d]),
    "b200_phase_series_mixable": (_i, [C.POINTER(PhaseSeries), C.POINTER(Observation), C.c_uint, C.c_int64, C.c_int64]),
    "b200_phase_series_folded": (_i, [C.POINTER(PhaseSeries), _u64, _u64]),
    "b200_phase_series_combine": (_i, [C.POINTER(PhaseSeries), C.POINTER(PhaseSeries)]),
    "b200_phase_series_normalise": (_i, [C.POINTER(PhaseSeries), _vp, _vp, C.POINTER(C.c_uint)]),
    "b200_phase_series_unload": (_i, [C.POINTER(PhaseSeries), C.c_char_p]),
    "b200_phase_series_load": (_i, [C.c_char_p, C.POINTER(PhaseSeries), _vp, _vp, _vp, _vp]),
    "b200_sigproc_header_from_observation": (_i, [C.POINTER(Observation), _u, C.POINTER(SigprocHeader)]),
    "b200_sigproc_header_write": (C.c_int64, [C.POINTER(SigprocHeader), _vp, _u64]),
    "b200_sigproc_file_write": (_i, [C.c_char_p, C.POINTER(SigprocHeader), _vp, _u64, _i]),
    "b200_phase_segments": (C.c_int64, [_d, _d, _u64, C.POINTER(PhaseSegment), _u64, C.POINTER(_d)]),
    "b200_phase_bins_sequential": (None, [_d, _d, _u, _u64, _vp, C.POINTER(_d)]),
}

_lib = None


def load():
    """Load libb200dsp.so; raises if it has not been built (python __graft_entry__.py / make)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libb200dsp.so is not built (%s). Run `make -C dspsr_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class B200Error(RuntimeError):
    """Mirror of the reference's `Error` exceptions (status + message of b200_last_error)."""

    def __init__(self, status, message):
        super().__init__("b200 status %d: %s" % (status, message))
        self.status = status


def check(status):
    if status != OK:
        raise B200Error(status, load().b200_last_error().decode(errors="replace"))

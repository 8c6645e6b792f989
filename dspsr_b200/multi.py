"""ctypes handle on libb200multi.so (include/b200multi.h): one host PROCESS driving one pipeline per GPU with a
worker thread each, NCCL reduce of the PhaseSeries at sub-integration boundaries -- the library-side counterpart
of dspsr's MultiThread (Signal/General/MultiThread.C).  bench.py's torchrun mode (one process per GPU) is the other
multi-GPU host; both share sharding.py's partition arithmetic."""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from . import engine as E
from . import phaseseries as P

_HERE = os.path.dirname(os.path.abspath(__file__))
MULTI_PATH = os.path.join(_HERE, "libb200multi.so")
SHARD_TIME, SHARD_CHANNEL = 0, 1

_vp, _u64p = C.c_void_p, C.POINTER(C.c_uint64)
SIGNATURES = {
    "b200_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_uint, C.POINTER(_vp)]),
    "b200_multi_destroy": (C.c_int, [_vp]),
    "b200_multi_ndev": (C.c_uint, [_vp]),
    "b200_multi_context": (_vp, [_vp, C.c_uint]),
    "b200_multi_set_pipeline": (C.c_int, [_vp, C.c_uint, _vp]),
    "b200_multi_execute_host_obs": (C.c_int, [_vp, C.POINTER(_vp), _u64p, _u64p, _u64p, _u64p]),
    "b200_multi_combine": (C.c_int, [_vp, C.c_int, C.POINTER(L.PhaseSeries)]),
    "b200_multi_reset": (C.c_int, [_vp]),
    "b200_multi_synchronize": (C.c_int, [_vp]),
    "b200_multi_last_error": (C.c_char_p, []),
    "b200_multi_nccl_version": (C.c_int, []),
}
_mlib = None


def load():
    global _mlib
    if _mlib is None:
        L.load()
        if not os.path.exists(MULTI_PATH):
            raise RuntimeError("libb200multi.so is not built (%s): make -C dspsr_b200/host" % MULTI_PATH)
        lib = C.CDLL(MULTI_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _mlib = lib
    return _mlib


def _check(rc):
    if rc != 0:
        raise L.B200Error(rc, load().b200_multi_last_error().decode())


class _BorrowedContext(E.Context):
    """An engine.Context view of a context owned by the multi host (not destroyed by Python)."""

    def __init__(self, handle, device):
        self.lib = L.load()
        self.h = C.c_void_p(handle)
        self.device = device
        self.torch_stream = None

    def __del__(self):
        pass


class MultiHost:
    def __init__(self, devices):
        self.lib = load()
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = _vp()
        _check(self.lib.b200_multi_create(arr, len(self.devices), C.byref(h)))
        self.h = h
        self.pipes = [None] * len(self.devices)

    def context(self, i):
        return _BorrowedContext(self.lib.b200_multi_context(self.h, i), self.devices[i])

    def set_pipeline(self, i, pipe):
        self.pipes[i] = pipe
        _check(self.lib.b200_multi_set_pipeline(self.h, i, pipe.h))

    def execute_host_obs(self, inputs, nparts, obs_samples, first_samples=None):
        """inputs: per device a numpy uint8 array / pinned torch tensor (or None when nparts[i] == 0)."""
        n = len(self.devices)
        ptrs, nbytes = (_vp * n)(), (C.c_uint64 * n)()
        for i, x in enumerate(inputs):
            if x is None:
                continue
            if hasattr(x, "data_ptr"):
                ptrs[i], nbytes[i] = x.data_ptr(), x.numel() * x.element_size()
            else:
                ptrs[i], nbytes[i] = x.ctypes.data, x.nbytes
        fs = (C.c_uint64 * n)(*(first_samples or [0] * n))
        _check(self.lib.b200_multi_execute_host_obs(self.h, ptrs, nbytes, fs, (C.c_uint64 * n)(*nparts),
                                                    (C.c_uint64 * n)(*obs_samples)))

    def combine(self, mode, nchan_total, npol, ndim, nbin):
        out = P.PhaseSeries(nchan_total, npol, ndim, nbin)
        _check(self.lib.b200_multi_combine(self.h, mode, C.byref(out.ps)))
        out._bind()
        return out

    def reset(self):
        _check(self.lib.b200_multi_reset(self.h))

    def synchronize(self):
        _check(self.lib.b200_multi_synchronize(self.h))

    def nccl_version(self):
        return self.lib.b200_multi_nccl_version()

    def close(self):
        if self.h:
            self.pipes = [None] * len(self.devices)       # pipelines first (they live on the host's contexts)
            import gc
            gc.collect()
            self.lib.b200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

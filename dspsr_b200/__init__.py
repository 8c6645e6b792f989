"""dspsr_b200 -- B200-native (sm_100a) implementation of dspsr's baseband hot path
(unpack -> overlap-save coherent dedispersion / filterbank -> detect -> fold) behind the
reference's engine interfaces.  The product is libb200dsp.so (C ABI, include/b200dsp.h);
this package holds its sources (csrc/), the C++ engine shims (host/) and Python handles."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]

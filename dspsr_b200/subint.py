"""dsp::Subint<Fold> on the B200 fold engine: sub-integrations of fixed length in seconds (dspsr -L).

Mirrors Subint<Fold>::transformation (Signal/Pulsar/dsp/Subint.h:235-305): every input block is cut at the
division boundaries by TimeDivide::set_bounds (b200_time_divide_set_bounds), each slice is folded with
Fold::fold's own phase set-up for the slice's first sample (Fold.C:650-657), and when the end of a division
is reached the accumulated PhaseSeries is handed to `unload(division, profile, hits, ndat_total, partial)`
and zeroed.  As in the reference, the first division of a stream is flagged partial (unload_partial): the
data may have started part-way through it."""
import ctypes as C

from . import _lib as L
from . import hostmath as HM


class SubintFolder:
    def __init__(self, fold_engine, predictor, start_mjd, division_seconds, unload, reference_phase=0.0):
        self.fe = fold_engine
        self.pred = predictor
        self.start = start_mjd                   # (day, sec, frac) of the observation start
        self.unload = unload
        self.reference_phase = reference_phase
        self.td = L.TimeDivide()
        L.check(L.load().b200_time_divide_init(C.byref(self.td), division_seconds))
        self.first_division = True
        self.have_data = False

    def fold_block(self, d_block, block_start_seconds, rate):
        """d_block: detected CUDA tensor [nchan, npol, ndat*ndim] starting `block_start_seconds` after the
        observation start, sampled at `rate` Hz."""
        ndat = d_block.shape[2] // self.fe.ndim
        lib = L.load()
        b = L.TimeBounds()
        more = True
        first_in_block = True
        while more:
            L.check(lib.b200_time_divide_set_bounds(C.byref(self.td), block_start_seconds, rate, ndat, C.byref(b)))
            more = bool(b.in_next)
            if first_in_block and b.new_division and self.have_data:
                self._flush(self.td.division, partial=True)       # uncontiguous input (Subint.h:262-270)
            first_in_block = False
            if not b.is_valid:
                continue
            t_block = HM.mjd_add(self.start, block_start_seconds)
            phi, pps = HM.fold_phase(self.pred, t_block, b.idat_start, rate, self.reference_phase)
            self.fe.set_bins(phi, pps, b.ndat, b.idat_start)
            self.fe.fold(d_block)
            self.have_data = True
            if b.end_reached:
                self._flush(b.division, partial=self.first_division)
                self.first_division = False

    def finish(self):
        if self.have_data:
            self._flush(self.td.division, partial=True)

    def _flush(self, division, partial):
        prof = self.fe.synch()
        hits, ntot = self.fe.hits()
        self.unload(int(division), prof, hits, ntot, partial)
        self.fe.zero()
        self.have_data = False

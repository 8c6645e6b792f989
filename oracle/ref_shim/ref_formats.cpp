// oracle/ref_shim/ref_formats.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors into the reference's own format unpackers, compiled in place by oracle/ref.mk:
//   Kernel/Formats/caspsr/CASPSRUnpacker.C      unpack() -> unpack_single_thread() -> unpack(ndat, from, into, ...)
//   Kernel/Formats/kat/MeerKATUnpacker.C     (row a4)             unpack(), FPT branch
//   Kernel/Formats/uwb/UWBUnpacker.C         (row a5)             unpack()
// against ref_shim/dsp/HistUnpacker.h (data-holder stand-ins for BitSeries / TimeSeries / Unpacker) and the reference's
// real dsp/BitTable.h + BitTable.C.  This file only builds the inputs and calls the protected unpack() through a
// derived probe class.  (Separate library from ref_cxx.cpp: the two use different stand-ins for dsp::Observation.)
#include "dsp/CASPSRUnpacker.h"
#include "dsp/MeerKATUnpacker.h"
#include "dsp/UWBUnpacker.h"

bool dsp::Unpacker::verbose = false;

namespace {
template <class U>
class Probe : public U {
 public:
  void run(const dsp::BitSeries* in, dsp::TimeSeries* out) {
    this->input = in;
    this->output = out;
    this->unpack();
  }
  double table_scale() { return this->table->get_scale(); }
};
class UwbProbe : public dsp::UWBUnpacker {
 public:
  void run(const dsp::BitSeries* in, dsp::TimeSeries* out) {
    input = in;
    output = out;
    unpack();
  }
};
void fill(dsp::BitSeries& in, dsp::TimeSeries& out, const unsigned char* raw, uint64_t ndat, unsigned nchan,
          unsigned npol, unsigned ndim, unsigned nbit, const char* machine, float* dst, uint64_t span) {
  in.raw = raw;
  in.ndat = out.ndat = ndat;
  in.nchan = out.nchan = nchan;
  in.npol = out.npol = npol;
  in.ndim = out.ndim = ndim;
  in.nbit = nbit;
  in.machine = machine;
  out.base = dst;
  out.span = span;
}
}  // namespace

extern "C" {
int ref_unpack_caspsr(const unsigned char* raw, uint64_t ndat, float* out, uint64_t span) {
  try {
    dsp::BitSeries in;
    dsp::TimeSeries ts;
    fill(in, ts, raw, ndat, 1, 2, 1, 8, "CASPSR", out, span);
    Probe<dsp::CASPSRUnpacker> u;
    u.run(&in, &ts);
  } catch (Error& e) { return -1; }
  return 0;
}
// sample_swap 1: MKBF, 2: MKBFRo.  *scale receives float(table->get_scale()) as the unpacker uses it.
int ref_unpack_meerkat(const unsigned char* raw, uint64_t ndat, unsigned nchan, unsigned npol, int sample_swap,
                       float* out, uint64_t span, float* scale) {
  try {
    dsp::BitSeries in;
    dsp::TimeSeries ts;
    fill(in, ts, raw, ndat, nchan, npol, 2, 8, sample_swap == 2 ? "MKBFRo" : "MKBF", out, span);
    Probe<dsp::MeerKATUnpacker> u;
    u.run(&in, &ts);
    if (scale) *scale = float(u.table_scale());
  } catch (Error& e) { return -1; }
  return 0;
}
int ref_unpack_uwb(const unsigned char* raw, uint64_t ndat, unsigned npol, float* out, uint64_t span) {
  try {
    dsp::BitSeries in;
    dsp::TimeSeries ts;
    fill(in, ts, raw, ndat, 1, npol, 2, 16, "UWB", out, span);
    UwbProbe u;
    u.run(&in, &ts);
  } catch (Error& e) { return -1; }
  return 0;
}
}

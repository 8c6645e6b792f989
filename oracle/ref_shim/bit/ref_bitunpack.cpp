// oracle/ref_shim/bit/ref_bitunpack.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" door into the reference's generic 8-bit unpacker, compiled in place by oracle/ref.mk:
//   Kernel/Classes/BitUnpacker.C        unpack(): the (ichan, ipol, idim) walk over TFP bytes, one histogram per digitizer
//   Kernel/Classes/EightBitUnpacker.C   unpack(ndat, from, nskip, into, fskip, hist): hist[*from]++, *into = lookup[*from]
//   Kernel/Classes/BitTable.C           the 256-entry table
// with the reference's own dsp/BitUnpacker.h / dsp/EightBitUnpacker.h on top of the data-holder stand-ins of
// ref_shim/dsp/HistUnpacker.h.
#include "dsp/EightBitUnpacker.h"

bool dsp::Unpacker::verbose = false;

namespace {
class Probe : public dsp::EightBitUnpacker {
 public:
  void run(const dsp::BitSeries* in, dsp::TimeSeries* out) {
    input = in;
    output = out;
    dsp::BitUnpacker::unpack();
  }
  unsigned long* histogram(unsigned idig) { return get_histogram(idig); }
};
}  // namespace

// raw: TFP bytes [idat][ichan][ipol][idim]; out: FPT planes of `span` floats; hist (nullable): [nchan*npol*ndim][256]
extern "C" int ref_unpack_generic8(const unsigned char* raw, uint64_t ndat, unsigned nchan, unsigned npol, unsigned ndim,
                                   int twos_complement, float* out, uint64_t span, unsigned long* hist) {
  try {
    dsp::BitSeries in;
    dsp::TimeSeries ts;
    in.raw = raw;
    in.ndat = ts.ndat = ndat;
    in.nchan = ts.nchan = nchan;
    in.npol = ts.npol = npol;
    in.ndim = ts.ndim = ndim;
    in.nbit = 8;
    ts.base = out;
    ts.span = span;
    Probe u;
    u.set_table(new dsp::BitTable(8, twos_complement ? dsp::BitTable::TwosComplement : dsp::BitTable::OffsetBinary));
    u.run(&in, &ts);
    if (hist)
      for (unsigned d = 0; d < nchan * npol * ndim; d++)
        for (unsigned s = 0; s < 256; s++) hist[d * 256u + s] = u.histogram(d)[s];
  } catch (Error& e) { return -1; }
  return 0;
}

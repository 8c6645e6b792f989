// ref_shim/dsp/HistUnpacker.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for the Unpacker / HistUnpacker / BitSeries /
// TimeSeries / Memory tree (Kernel/Classes/dsp/*.h, which needs PSRCHIVE): plain data holders with exactly the
// members the format unpackers' unpack() bodies touch (CASPSRUnpacker.C, MeerKATUnpacker.C, UWBUnpacker.C are
// compiled in place against this).  FPT layout as DataSeries.C:246-259: plane (ichan, ipol) at base + (ichan*npol+ipol)*span.
#ifndef REF_SHIM_DSP_HISTUNPACKER_H
#define REF_SHIM_DSP_HISTUNPACKER_H
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>
#include <iostream>
#include <string>
#include <vector>
#include "Reference.h"
#include "Error.h"
#include "dsp/BitTable.h"       // the reference's own
namespace dsp {
class Memory : public Reference::Able {
 public:
  static Memory* get_manager() { static Memory m; return &m; }
};
class Observation : public Reference::Able {
 public:
  Observation() : ndat(0), nchan(1), npol(2), ndim(1), nbit(8) {}
  uint64_t get_ndat() const { return ndat; }
  unsigned get_nchan() const { return nchan; }
  unsigned get_npol() const { return npol; }
  unsigned get_ndim() const { return ndim; }
  unsigned get_nbit() const { return nbit; }
  std::string get_machine() const { return machine; }
  uint64_t ndat;
  unsigned nchan, npol, ndim, nbit;
  std::string machine;
};
class BitSeries : public Observation {
 public:
  BitSeries() : raw(0) {}
  const unsigned char* get_rawptr() const { return raw; }
  const unsigned char* raw;
};
class TimeSeries : public Observation {
 public:
  enum Order { OrderFPT, OrderTFP };
  TimeSeries() : base(0), span(0), order(OrderFPT) {}
  Order get_order() const { return order; }
  float* get_datptr(unsigned ichan, unsigned ipol) { return base + (uint64_t(ichan) * npol + ipol) * span; }
  float* get_dattfp() { return base; }
  float* base;
  uint64_t span;
  Order order;
};
class Unpacker : public Reference::Able {
 public:
  static bool verbose;
  Unpacker() : input(0), output(0), output_order(TimeSeries::OrderFPT) {}
  virtual ~Unpacker() {}
  virtual void set_device(Memory*) {}
  const BitSeries* input;
  TimeSeries* output;
  TimeSeries::Order output_order;
};
class HistUnpacker : public Unpacker {
 public:
  HistUnpacker(const char* = "HistUnpacker") : nstate(256), ndig(2) {}
  ~HistUnpacker() { clear(); }
  void set_nstate(unsigned n) { nstate = n; clear(); }
  void set_ndig(unsigned n) { ndig = n; }
  unsigned get_ndig() const { return ndig; }
  // one heap block per digitizer: pointers handed out earlier stay valid when more digitizers are asked for
  unsigned long* get_histogram(unsigned idig) {
    if (histograms.size() <= idig) histograms.resize(idig + 1, (unsigned long*)0);
    if (!histograms[idig]) histograms[idig] = new unsigned long[nstate]();
    return histograms[idig];
  }
 protected:
  void clear() {
    for (size_t i = 0; i < histograms.size(); i++) delete[] histograms[i];
    histograms.clear();
  }
  unsigned nstate, ndig;
  std::vector<unsigned long*> histograms;
};
}  // namespace dsp
#endif

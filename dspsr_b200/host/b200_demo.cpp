// b200_demo.cpp -- drives the four engine shims through the (stand-in) reference operators the
// way dspsr's LoadToFold does (Signal/Pulsar/LoadToFold1.C:117-599, SingleThread.C:405-431):
//   raw CASPSR bytes -> Unpacker device hook -> Filterbank(engine) -> Detection(engine) -> Fold(engine)
// Usage: b200_demo raw.bin response.c64 nchan freq_res nfilt_pos nfilt_neg nbin phi pps out.bin
// Output: nchan*4*nbin float32 profile (Coherence, ndim 4) followed by nbin uint32 hits.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "B200Engines.h"

struct HostMemory : public dsp::Memory {
  void* do_allocate(size_t n) { return malloc(n); }
  void do_free(void* p) { free(p); }
  bool on_host() const { return true; }
};

static std::vector<char> slurp(const char* fn) {
  FILE* f = fopen(fn, "rb");
  if (!f) { perror(fn); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> b(n);
  if (fread(b.data(), 1, n, f) != size_t(n)) { perror("fread"); exit(2); }
  fclose(f);
  return b;
}

int main(int argc, char** argv) try {
  if (argc != 11) {
    fprintf(stderr, "usage: %s raw.bin response.c64 nchan freq_res nfilt_pos nfilt_neg nbin phi pps out.bin\n", argv[0]);
    return 2;
  }
  std::vector<char> raw = slurp(argv[1]);
  std::vector<char> resp = slurp(argv[2]);
  const unsigned nchan = atoi(argv[3]), freq_res = atoi(argv[4]), npos = atoi(argv[5]), nneg = atoi(argv[6]);
  const unsigned nbin = atoi(argv[7]);
  const double phi = atof(argv[8]), pps = atof(argv[9]);

  b200_context* ctx = 0;
  B200::check(b200_context_create(0, NULL, &ctx), "b200_context_create");
  Reference::To<dsp::Memory> device = new B200::DeviceMemory(ctx);
  Reference::To<dsp::Memory> host = new HostMemory;

  // ---- IOManager: load + unpack on the device (SingleThread.C:249-275, File.C:213-272) ----
  const uint64_t ndat = raw.size() / 2 / 4 * 4;
  void* d_raw = 0;
  B200::check(b200_malloc(ctx, raw.size(), &d_raw), "b200_malloc");
  B200::check(b200_memcpy_h2d(ctx, d_raw, raw.data(), raw.size()), "b200_memcpy_h2d");
  Reference::To<dsp::TimeSeries> unpacked = new dsp::TimeSeries;
  unpacked->set_memory(device);
  unpacked->set_state(Signal::Nyquist);
  unpacked->set_nchan(1); unpacked->set_npol(2); unpacked->set_ndim(1);
  unpacked->set_rate(800e6);
  unpacked->resize(ndat);
  b200_unpack_desc ud;
  ud.format = B200_FMT_CASPSR8; ud.nchan = 1; ud.npol = 2; ud.ndim = 1; ud.scale = 0; ud.sample_swap = 1;
  B200::check(b200_bittable8(1, ud.lut, NULL), "b200_bittable8");
  B200::check(b200_unpack(ctx, &ud, d_raw, ndat, unpacked->get_datptr(0, 0), unpacked->get_nfloat_span()), "b200_unpack");

  // ---- Filterbank with the dedispersion response (LoadToFold1.C:295-328) ----
  Reference::To<dsp::Response> kernel = new dsp::Response;
  kernel->configure(nchan, freq_res, npos, nneg);
  if (resp.size() != size_t(nchan) * freq_res * 8) { fprintf(stderr, "response size mismatch\n"); return 2; }
  memcpy(kernel->writable(), resp.data(), resp.size());
  Reference::To<dsp::TimeSeries> filtered = new dsp::TimeSeries;
  filtered->set_memory(device);
  Reference::To<dsp::Filterbank> filterbank = new dsp::Filterbank;
  filterbank->set_input(unpacked);
  filterbank->set_output(filtered);
  filterbank->set_nchan(nchan);
  filterbank->set_response(kernel);
  filterbank->set_engine(new B200::FilterbankEngine(ctx));
  filterbank->prepare();
  filterbank->operate();

  // ---- Detection (LoadToFold1.C:541-556,1098-1153) ----
  Reference::To<dsp::TimeSeries> detected = new dsp::TimeSeries;
  detected->set_memory(device);
  Reference::To<dsp::Detection> detect = new dsp::Detection;
  detect->set_input(filtered);
  detect->set_output(detected);
  detect->set_output_state(Signal::Coherence);
  detect->set_output_ndim(4);
  detect->set_engine(new B200::DetectionEngine(ctx));
  detect->operate();

  // ---- Fold (LoadToFold1.C:927-969,1155-1242) ----
  Reference::To<dsp::PhaseSeries> profiles = new dsp::PhaseSeries;
  profiles->set_memory(host);
  profiles->copy_configuration(detected);
  profiles->resize_bins(nbin);
  Reference::To<dsp::Fold> fold = new dsp::Fold;
  fold->set_input(detected);
  fold->set_output(profiles);
  fold->set_nbin(nbin);
  fold->set_phase(phi, pps);
  fold->set_engine(new B200::FoldEngine(ctx));
  fold->operate();
  dsp::PhaseSeries* result = fold->get_result();

  FILE* f = fopen(argv[10], "wb");
  for (unsigned c = 0; c < result->get_nchan(); c++)
    for (unsigned p = 0; p < result->get_npol(); p++)
      fwrite(result->get_datptr(c, p), sizeof(float), size_t(nbin) * result->get_ndim(), f);
  fwrite(result->get_hits(), sizeof(unsigned), nbin, f);
  fclose(f);
  printf("b200_demo: nchan=%u npol'=%u ndim'=%u nbin=%u ndat_out=%llu ndat_total=%llu integration_length=%g s\n",
         result->get_nchan(), result->get_npol(), result->get_ndim(), nbin,
         (unsigned long long)detected->get_ndat(), (unsigned long long)result->ndat_total, result->integration_length);
  b200_free(ctx, d_raw);
  return 0;
} catch (Error& e) {
  fprintf(stderr, "Error in %s: %s\n", e.function.c_str(), e.message.c_str());
  return 1;
}

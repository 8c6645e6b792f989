// b200_demo.cpp -- drives the four engine shims through the (stand-in) reference operators the
// way dspsr's LoadToFold does (Signal/Pulsar/LoadToFold1.C:117-599, SingleThread.C:405-431):
//   raw CASPSR bytes -> Unpacker device hook -> Filterbank(engine) -> Detection(engine) -> Fold(engine)
// Usage: b200_demo raw.bin response.c64 nchan freq_res nfilt_pos nfilt_neg nbin phi pps out.bin
//        b200_demo --meerkat raw.bin response.c64 nchan freq_res nfilt_pos nfilt_neg nbin phi pps out.bin
// The second form is BASELINE configs[2]'s wiring: MeerKATUnpacker(engine) -> Convolution(engine) -> Detection(engine)
// -> Fold(engine) on `nchan` input channels, the blocks folded in TWO calls so that the engine-owned PhaseSeries
// accumulates across Fold::transformation calls, with a Fold::reset() + refold in between to exercise zero().
// Output: nchan*4*nbin float32 profile (Coherence, ndim 4) followed by nbin uint32 hits, then (double) integration
// length and (uint64) ndat_total.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "B200Engines.h"

struct HostMemory : public dsp::Memory {
  void* do_allocate(size_t n) { return malloc(n); }
  void do_free(void* p) { free(p); }
  void do_zero(void* p, size_t n) { memset(p, 0, n); }
  void do_copy(void* to, const void* from, size_t n) { memcpy(to, from, n); }
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

static int write_result(const char* fn, dsp::PhaseSeries* result, unsigned nbin) {
  FILE* f = fopen(fn, "wb");
  if (!f) { perror(fn); return 2; }
  for (unsigned c = 0; c < result->get_nchan(); c++)
    for (unsigned p = 0; p < result->get_npol(); p++)
      fwrite(result->get_datptr(c, p), sizeof(float), size_t(nbin) * result->get_ndim(), f);
  fwrite(result->get_hits(), sizeof(unsigned), nbin, f);
  fwrite(&result->integration_length, sizeof(double), 1, f);
  fwrite(&result->ndat_total, sizeof(uint64_t), 1, f);
  fclose(f);
  return 0;
}

// BASELINE configs[2] wiring through the stand-in operators; see the usage comment
static int meerkat_main(char** argv) {
  std::vector<char> raw = slurp(argv[0]);
  std::vector<char> resp = slurp(argv[1]);
  const unsigned nchan = atoi(argv[2]), freq_res = atoi(argv[3]), npos = atoi(argv[4]), nneg = atoi(argv[5]);
  const unsigned nbin = atoi(argv[6]);
  const double phi = atof(argv[7]), pps = atof(argv[8]);
  b200_context* ctx = 0;
  B200::check(b200_context_create(0, NULL, &ctx), "b200_context_create");
  Reference::To<dsp::Memory> device = new B200::DeviceMemory(ctx);
  Reference::To<dsp::Memory> host = new HostMemory;

  // the block's raw bytes on the device (File::load_bytes_device) as a BitSeries
  const uint64_t ndat = raw.size() / (uint64_t(nchan) * 2 * 2) / 256 * 256;
  void* d_raw = 0;
  B200::check(b200_malloc(ctx, raw.size(), &d_raw), "b200_malloc");
  B200::check(b200_memcpy_h2d(ctx, d_raw, raw.data(), raw.size()), "b200_memcpy_h2d");
  Reference::To<dsp::BitSeries> bits = new dsp::BitSeries;
  bits->set_machine("MKBF");
  bits->set_state(Signal::Analytic);
  bits->set_nchan(nchan); bits->set_npol(2); bits->set_ndim(2); bits->set_nbit(8);
  bits->set_rate(856e6 / 1024);
  bits->set_start_time(MJD(55299, 7545, 0.0));
  bits->set_rawptr(static_cast<const unsigned char*>(d_raw), ndat);

  Reference::To<dsp::TimeSeries> unpacked = new dsp::TimeSeries;
  unpacked->set_memory(device);
  Reference::To<dsp::MeerKATUnpacker> unpacker = new dsp::MeerKATUnpacker;
  double table_scale = 0;
  float lut[256];
  B200::check(b200_bittable8(1, lut, &table_scale), "b200_bittable8");
  unpacker->set_table_scale(table_scale);
  unpacker->set_engine(new B200::MeerKATUnpackerEngine(ctx));
  if (!unpacker->get_device_supported(device)) { fprintf(stderr, "device not supported\n"); return 1; }
  unpacker->set_device(device);
  unpacker->set_input(bits);
  unpacker->set_output(unpacked);
  unpacker->operate();

  Reference::To<dsp::Response> kernel = new dsp::Response;
  kernel->configure(nchan, freq_res, npos, nneg);
  if (resp.size() != size_t(nchan) * freq_res * 8) { fprintf(stderr, "response size mismatch\n"); return 2; }
  memcpy(kernel->writable(), resp.data(), resp.size());
  Reference::To<dsp::TimeSeries> convolved = new dsp::TimeSeries;
  convolved->set_memory(device);
  Reference::To<dsp::Convolution> convolution = new dsp::Convolution;
  convolution->set_input(unpacked);
  convolution->set_output(convolved);
  convolution->set_response(kernel);
  convolution->set_engine(new B200::ConvolutionEngine(ctx));
  convolution->prepare();
  convolution->operate();

  Reference::To<dsp::TimeSeries> detected = new dsp::TimeSeries;
  detected->set_memory(device);
  Reference::To<dsp::Detection> detect = new dsp::Detection;
  detect->set_input(convolved);
  detect->set_output(detected);
  detect->set_output_state(Signal::Coherence);
  detect->set_output_ndim(4);
  detect->set_engine(new B200::DetectionEngine(ctx));
  detect->operate();

  // Fold: the host PhaseSeries is only the destination of synch(); the accumulator is the engine's
  Reference::To<dsp::PhaseSeries> profiles = new dsp::PhaseSeries;
  profiles->set_memory(host);
  Reference::To<dsp::Fold> fold = new dsp::Fold;
  fold->set_input(detected);
  fold->set_output(profiles);
  fold->set_nbin(nbin);
  fold->set_phase(phi, pps);
  fold->set_engine(new B200::FoldEngine(ctx));
  fold->operate();
  fold->get_result();                 // synch #1
  fold->reset();                      // Fold::reset -> engine->zero() (Fold.C:137-148)
  fold->operate();                    // the same block again into the zeroed accumulator ...
  fold->operate();                    // ... and once more: the engine-owned PhaseSeries accumulates across calls
  dsp::PhaseSeries* result = fold->get_result();
  if (fold->get_output() == profiles.get()) { fprintf(stderr, "Fold::get_output() must be the engine's PhaseSeries\n"); return 1; }
  int rc = write_result(argv[9], result, nbin);
  printf("b200_demo --meerkat: nchan=%u nbin=%u ndat_out=%llu ndat_total=%llu integration_length=%g s start=%d+%d end-start=%g s\n",
         result->get_nchan(), nbin, (unsigned long long)detected->get_ndat(), (unsigned long long)result->ndat_total,
         result->integration_length, result->get_start_time().intday(), result->get_start_time().get_secs(),
         result->get_end_time() - result->get_start_time());
  b200_free(ctx, d_raw);
  return rc;
}

int main(int argc, char** argv) try {
  if (argc == 12 && !strcmp(argv[1], "--meerkat")) return meerkat_main(argv + 2);
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
  Reference::To<dsp::Fold> fold = new dsp::Fold;
  fold->set_input(detected);
  fold->set_output(profiles);
  fold->set_nbin(nbin);
  fold->set_phase(phi, pps);
  fold->set_engine(new B200::FoldEngine(ctx));
  fold->operate();
  dsp::PhaseSeries* result = fold->get_result();

  if (write_result(argv[10], result, nbin)) return 2;
  printf("b200_demo: nchan=%u npol'=%u ndim'=%u nbin=%u ndat_out=%llu ndat_total=%llu integration_length=%g s\n",
         result->get_nchan(), result->get_npol(), result->get_ndim(), nbin,
         (unsigned long long)detected->get_ndat(), (unsigned long long)result->ndat_total, result->integration_length);
  b200_free(ctx, d_raw);
  return 0;
} catch (Error& e) {
  fprintf(stderr, "Error in %s: %s\n", e.function.c_str(), e.message.c_str());
  return 1;
}

// sigproc.cpp -- SIGPROC filterbank (.fil) header and file writer: the last stage of digifil.
//
// Replaces: dsp::SigProcOutputFile::write_header (Kernel/Formats/sigproc/SigProcOutputFile.C:35-60) ->
// SigProcObservation::unload_global (SigProcObservation.C:228-275) -> filterbank_header (filterbank_header.c:44-100,
// send_stuff.c), for the BitSeries that SigProcDigitizer::pack emits (SigProcDigitizer.C:84-86: bandwidth =
// -|bandwidth|, no swap -- ChannelSort has put the channels in descending frequency order).
// Keyword strings and binary values are written exactly as send_string / send_int / send_double do
// (native-endian int32 length prefix, no terminator).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/b200dsp.h"

namespace {

struct Out {
  unsigned char* p;
  uint64_t cap, n;
  bool ok;
  void bytes(const void* src, uint64_t len) {
    if (n + len > cap) { ok = false; return; }
    memcpy(p + n, src, len);
    n += len;
  }
  void str(const char* s) {
    const int len = (int)strlen(s);
    bytes(&len, sizeof len);
    bytes(s, (uint64_t)len);
  }
  void i32(const char* name, int v) { str(name); bytes(&v, sizeof v); }
  void f64(const char* name, double v) { str(name); bytes(&v, sizeof v); }
};

std::string upper(const char* s) {
  std::string r(s);
  for (auto& c : r) c = (char)toupper((unsigned char)c);
  return r;
}

// get_sigproc_telescope_id (SigProcObservation.C:103-140): PSRCHIVE's Tempo::itoa_code resolves aliases to the ITOA
// two-letter code; restated here for the telescopes sigproc knows (aliases.c)
int telescope_id(const char* name) {
  const std::string n = upper(name);
  if (n == "AO" || n == "ARECIBO") return 1;
  if (n == "NC" || n == "NANCAY") return 3;
  if (n == "PK" || n == "PKS" || n == "PARKES") return 4;
  if (n == "JB" || n == "JODRELL" || n == "JBO" || n == "LOVELL") return 5;
  if (n == "GB" || n == "GBT") return 6;
  if (n == "GM" || n == "GMRT") return 7;
  if (n == "EF" || n == "EFFELSBERG") return 8;
  if (n == "LF" || n == "LOFAR") return 11;
  if (n == "VL" || n == "VLA") return 12;
  return 0;
}

}  // namespace

extern "C" {

int b200_sigproc_header_from_observation(const b200_observation* det, unsigned nbit, b200_sigproc_header* h) {
  if (!det || !h || det->nchan == 0 || det->rate <= 0) return B200_ERR_INVALID;
  memset(h, 0, sizeof *h);
  strcpy(h->rawdatafile, "unknown");                              // inpfile (:251-252)
  strncpy(h->source_name, det->source, sizeof h->source_name - 1);
  h->telescope_id = telescope_id(det->telescope);
  const std::string m = det->machine;
  h->machine_id = m == "BPSR" ? 10 : m == "SCAMP" ? 6 : m == "COBALT" ? 11 : 0;
  // the digitizer's output: bandwidth = -|bandwidth|, swap = false, nsub_swap = 0 (SigProcDigitizer.C:84-86)
  const double bw = -std::fabs(det->bandwidth);
  const double base = det->dc_centred ? det->centre_frequency - 0.5 * bw
                                      : det->centre_frequency - 0.5 * bw + 0.5 * bw / double(det->nchan);   // Observation.C:445-451
  h->fch1 = base;                                                  // get_centre_frequency(0), :438-442
  h->foff = bw / double(det->nchan);
  h->nchans = (int)det->nchan;
  h->nifs = (int)det->npol;
  h->nbits = (int)nbit;
  h->tsamp = 1.0 / det->rate;
  h->tstart = double(det->start_time.day) + (double(det->start_time.sec) + det->start_time.frac) / 86400.0;   // MJD::in_days
  h->src_raj = h->src_dej = h->az_start = h->za_start = 0.0;
  return B200_OK;
}

// filterbank_header (filterbank_header.c:44-100) in filterbank mode (zerolagdump = 0, sumifs = 0, every IF selected)
int64_t b200_sigproc_header_write(const b200_sigproc_header* h, unsigned char* buf, uint64_t buflen) {
  if (!h || !buf) return -1;
  Out o{buf, buflen, 0, true};
  o.str("HEADER_START");
  if (h->rawdatafile[0]) { o.str("rawdatafile"); o.str(h->rawdatafile); }
  if (h->source_name[0]) { o.str("source_name"); o.str(h->source_name); }
  o.i32("machine_id", h->machine_id);
  o.i32("telescope_id", h->telescope_id);
  // send_coords: its tests `(x != 0.0) || (x != -1.0)` are always true -- all four are written
  o.f64("src_raj", h->src_raj);
  o.f64("src_dej", h->src_dej);
  o.f64("az_start", h->az_start);
  o.f64("za_start", h->za_start);
  o.i32("data_type", 1);
  o.f64("fch1", h->fch1);
  o.f64("foff", h->foff);
  o.i32("nchans", h->nchans);
  o.i32("nbeams", h->nbeams);
  o.i32("ibeam", h->ibeam);
  o.i32("nbits", h->nbits);
  o.f64("tstart", h->tstart);
  o.f64("tsamp", h->tsamp);
  o.i32("nifs", h->nifs);
  o.str("HEADER_END");
  return o.ok ? (int64_t)o.n : -1;
}

int b200_sigproc_file_write(const char* path, const b200_sigproc_header* h, const unsigned char* h_bytes, uint64_t nbytes,
                            int append) {
  if (!path || (!append && !h)) return B200_ERR_INVALID;
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) return B200_ERR_INVALID;
  bool ok = true;
  if (!append) {
    unsigned char hdr[1024];
    const int64_t n = b200_sigproc_header_write(h, hdr, sizeof hdr);
    ok = n > 0 && fwrite(hdr, 1, (size_t)n, f) == (size_t)n;
  }
  if (ok && nbytes) ok = fwrite(h_bytes, 1, nbytes, f) == nbytes;
  ok = (fclose(f) == 0) && ok;
  return ok ? B200_OK : B200_ERR_INVALID;
}

}  // extern "C"

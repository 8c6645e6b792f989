// phaseseries.cpp -- the host side of the folded accumulator: attributes, merge rules, unload.
//
// What dsp::PhaseSeries (Signal/Pulsar/PhaseSeries.C), dsp::Observation::combinable
// (Kernel/Classes/Observation.C:139-310) and dsp::Archiver::set (Signal/Pulsar/Archiver.C:773-895) do to a
// folded sub-integration between the fold engine and the file, as plain C functions over a POD
// (include/b200dsp.h b200_phase_series).  No device code: the device arrays are summed / gathered by the caller
// (NCCL, multi.cpp) and the rules here are applied to the attributes and to the host copies.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200dsp.h"

namespace {

std::string tostr(double v) {
  char b[64];
  snprintf(b, sizeof b, "%g", v);
  return b;
}
std::string tostr(int v) { return std::to_string(v); }
std::string tostr(unsigned v) { return std::to_string(v); }

b200_mjd normalise(b200_mjd t) {
  const double whole = std::floor(t.frac);
  t.frac -= whole;
  long long sec = (long long)t.sec + (long long)whole;
  long long day = t.day;
  day += sec / 86400;
  sec %= 86400;
  if (sec < 0) { sec += 86400; day -= 1; }
  t.day = (int)day;
  t.sec = (int)sec;
  return t;
}

bool earlier(const b200_mjd& a, const b200_mjd& b) {
  if (a.day != b.day) return a.day < b.day;
  if (a.sec != b.sec) return a.sec < b.sec;
  return a.frac < b.frac;
}

uint64_t nprofile_floats(const b200_phase_series* ps) {
  return uint64_t(ps->obs.nchan) * ps->obs.npol * ps->nbin * ps->obs.ndim;
}

}  // namespace

extern "C" {

double b200_mjd_diff(const b200_mjd* b, const b200_mjd* a) {
  return double(b->day - a->day) * 86400.0 + double(b->sec - a->sec) + (b->frac - a->frac);
}

b200_mjd b200_mjd_add(const b200_mjd* t, double seconds) {
  b200_mjd r = *t;
  r.frac += seconds;
  return normalise(r);
}

// Observation::combinable (Observation.C:139-310), test by test in the reference's order
int b200_observation_combinable(const b200_observation* a, const b200_observation* o, char* reason, unsigned reason_len) {
  if (!a || !o) return 0;
  bool can = true;
  const double eps = 0.000001;
  std::string why;
  const std::string sep = "\n\t";
  auto differ = [&](const std::string& what, const std::string& x, const std::string& y) {
    why += sep + "different " + what + ":" + x + " != " + y;
    can = false;
  };
  if (strncmp(a->telescope, o->telescope, sizeof a->telescope)) differ("telescopes", a->telescope, o->telescope);
  if (strncmp(a->receiver, o->receiver, sizeof a->receiver)) differ("receivers", a->receiver, o->receiver);
  if (strncmp(a->source, o->source, sizeof a->source)) differ("sources", a->source, o->source);
  if (std::fabs(a->centre_frequency - o->centre_frequency) > eps)
    differ("centre frequencies", tostr(a->centre_frequency), tostr(o->centre_frequency));
  else if (std::fabs(a->bandwidth - o->bandwidth) > eps)
    differ("bandwidths", tostr(a->bandwidth), tostr(o->bandwidth));
  if (a->nchan != o->nchan) differ("nchans", tostr(a->nchan), tostr(o->nchan));
  if (a->npol != o->npol) differ("npols", tostr(a->npol), tostr(o->npol));
  if (a->ndim != o->ndim) differ("ndims", tostr(a->ndim), tostr(o->ndim));
  if (a->nbit != o->nbit) differ("nbits", tostr(a->nbit), tostr(o->nbit));
  if (a->type != o->type) differ("types", tostr(a->type), tostr(o->type));
  if (a->state != o->state) differ("states", tostr(a->state), tostr(o->state));
  if (a->basis != o->basis) differ("bases", tostr(a->basis), tostr(o->basis));
  if (a->rate != o->rate) differ("rates", tostr(a->rate), tostr(o->rate));
  if (std::fabs(a->scale - o->scale) > eps * std::fabs(a->scale)) differ("scales", tostr(a->scale), tostr(o->scale));
  if (a->swap != o->swap) differ("swaps", tostr(a->swap), tostr(o->swap));
  if (a->nsub_swap != o->nsub_swap) differ("nsub_swaps", tostr(a->nsub_swap), tostr(o->nsub_swap));
  if (a->dc_centred != o->dc_centred) differ("dccs", tostr(a->dc_centred), tostr(o->dc_centred));
  if (strncmp(a->mode, o->mode, sizeof a->mode)) {
    // two "2-bit..." modes may differ in their tail (Observation.C:267-279)
    if (!(strncmp(a->mode, o->mode, 5) == 0 && strncmp(a->mode, "2-bit", 5) == 0)) differ("modes", a->mode, o->mode);
  }
  if (strncmp(a->machine, o->machine, sizeof a->machine)) differ("machines", a->machine, o->machine);
  if (strncmp(a->format, o->format, sizeof a->format)) differ("formats", a->format, o->format);
  if (std::fabs(a->dispersion_measure - o->dispersion_measure) > eps)
    differ("dispersion measures", tostr(a->dispersion_measure), tostr(o->dispersion_measure));
  if (std::fabs(a->rotation_measure - o->rotation_measure) > eps)
    differ("rotation measures", tostr(a->rotation_measure), tostr(o->rotation_measure));
  if (reason && reason_len) {
    strncpy(reason, why.c_str(), reason_len - 1);
    reason[reason_len - 1] = 0;
  }
  return can ? 1 : 0;
}

}  // extern "C"

// PhaseSeries::mixable (PhaseSeries.C:336-418) once the bounds of the data to mix in are known
static int mixable_bounds(b200_phase_series* ps, const b200_observation* obs, unsigned nbin, const b200_mjd& obsStart,
                          const b200_mjd& obsEnd) {
  if (ps->integration_length == 0.0) {
    // the integration is currently empty: take over the attributes, keep the record of dropped samples
    const uint64_t backup_ndat_total = ps->ndat_total;
    ps->obs = *obs;
    ps->end_time = obsEnd;
    ps->obs.start_time = obsStart;
    ps->nbin = nbin;
    if (!ps->hits_nchan) ps->hits_nchan = 1;
    if (ps->data) memset(ps->data, 0, nprofile_floats(ps) * sizeof(float));
    if (ps->hits) memset(ps->hits, 0, uint64_t(ps->hits_nchan) * nbin * sizeof(unsigned));
    ps->ndat_total = backup_ndat_total;
    return 1;
  }
  // the PhaseSeries' own start_time was moved by earlier calls: compare everything else
  if (!b200_observation_combinable(&ps->obs, obs, nullptr, 0)) return 0;
  if (ps->nbin != nbin) return 0;
  if (earlier(ps->end_time, obsEnd)) ps->end_time = obsEnd;
  if (earlier(obsStart, ps->obs.start_time)) ps->obs.start_time = obsStart;
  return 1;
}

extern "C" {

int b200_phase_series_mixable(b200_phase_series* ps, const b200_observation* obs, unsigned nbin, int64_t istart,
                              int64_t fold_ndat) {
  if (!ps || !obs || obs->rate <= 0) return 0;
  const b200_mjd obsStart = b200_mjd_add(&obs->start_time, double(istart) / obs->rate);
  b200_mjd obsEnd;
  if (fold_ndat == 0) obsEnd = b200_mjd_add(&obs->start_time, double(obs->ndat) / obs->rate);   // Observation::get_end_time
  else obsEnd = b200_mjd_add(&obsStart, double(fold_ndat) / obs->rate);
  return mixable_bounds(ps, obs, nbin, obsStart, obsEnd);
}

int b200_phase_series_folded(b200_phase_series* ps, uint64_t ndat_folded, uint64_t ndat_fold) {
  if (!ps || ps->obs.rate <= 0) return B200_ERR_INVALID;
  ps->integration_length += double(ndat_folded) / ps->obs.rate;      // Fold.C:789,801
  ps->ndat_total += ndat_fold;                                       // :802
  return B200_OK;
}

// PhaseSeries::combine (PhaseSeries.C:442-480)
int b200_phase_series_combine(b200_phase_series* ps, const b200_phase_series* prof) {
  if (!ps) return B200_ERR_INVALID;
  if (!prof || prof->nbin == 0) return B200_OK;
  if (!ps->integration_length) {
    // "this is empty": *this = *prof (arrays copied into this's own buffers)
    float* data = ps->data;
    unsigned* hits = ps->hits;
    *ps = *prof;
    ps->data = data;
    ps->hits = hits;
    if (data && prof->data) memcpy(data, prof->data, nprofile_floats(prof) * sizeof(float));
    if (hits && prof->hits) memcpy(hits, prof->hits, uint64_t(prof->hits_nchan) * prof->nbin * sizeof(unsigned));
    return B200_OK;
  }
  // mixable(*prof, prof->get_nbin()): istart 0, fold to prof's end (PhaseSeries::get_end_time returns end_time)
  if (!mixable_bounds(ps, &prof->obs, prof->nbin, prof->obs.start_time, prof->end_time)) return B200_ERR_INVALID;
  if (ps->data && prof->data) {
    const uint64_t n = nprofile_floats(ps);
    for (uint64_t i = 0; i < n; i++) ps->data[i] += prof->data[i];          // TimeSeries::operator +=
  }
  if (ps->hits && prof->hits) {
    const unsigned nhits = ps->nbin * ps->hits_nchan;
    for (unsigned i = 0; i < nhits; i++) ps->hits[i] += prof->hits[i];
  }
  ps->integration_length += prof->integration_length;
  ps->ndat_total += prof->ndat_total;
  if (!ps->ndat_expected) ps->ndat_expected = prof->ndat_expected;
  return B200_OK;
}

// Archiver::set (Archiver.C:773-895) for every profile of the sub-integration
int b200_phase_series_normalise(const b200_phase_series* ps, float* out, float* weights, unsigned* corrupted) {
  if (!ps || !ps->data || !ps->hits || !out) return B200_ERR_INVALID;
  const double scale = ps->obs.scale;
  if (scale == 0 || !std::isfinite(scale)) return B200_ERR_INVALID;           // "invalid scale"
  const unsigned nbin = ps->nbin, npol = ps->obs.npol, ndim = ps->obs.ndim, nchan = ps->obs.nchan;
  unsigned bad = 0;
  for (unsigned ichan = 0; ichan < nchan; ichan++)
    for (unsigned ipol = 0; ipol < npol; ipol++)
      for (unsigned idim = 0; idim < ndim; idim++) {
        const unsigned* hits = ps->hits + (ps->hits_nchan > 1 ? uint64_t(ichan) * nbin : 0);
        const float* from = ps->data + (uint64_t(ichan) * npol + ipol) * nbin * ndim + idim;    // OrderFPT
        float* into = out + ((uint64_t(ichan) * npol + ipol) * ndim + idim) * nbin;
        float weight = 1.0f;
        unsigned zeroes = 0, not_finite = 0, hits_sum = 0;
        for (unsigned ibin = 0; ibin < nbin; ibin++) {
          hits_sum += hits[ibin];
          if (hits[ibin] == 0) {
            zeroes++;
            into[ibin] = 0.0;
          } else if (!std::isfinite(*from))
            not_finite++;
          else
            into[ibin] = *from / (scale * double(hits[ibin]));
          from += ndim;
        }
        if (not_finite) {
          for (unsigned ibin = 0; ibin < nbin; ibin++) into[ibin] = 0;
          bad++;
          weight = 0;
        }
        if (zeroes) {
          double sum = 0.0;
          unsigned count = 0;
          for (unsigned ibin = 0; ibin < nbin; ibin++)
            if (hits[ibin] != 0) {
              sum += into[ibin];
              count++;
            }
          if (count == 0) count = 1;
          const double mean = sum / count;
          for (unsigned ibin = 0; ibin < nbin; ibin++)
            if (hits[ibin] == 0) into[ibin] = mean;
        }
        if (ps->hits_nchan > 1) weight = (float)hits_sum / (float)ps->ndat_total;             // zeroed data
        if (weights) weights[(uint64_t(ichan) * npol + ipol) * ndim + idim] = weight;
      }
  if (corrupted) *corrupted = bad;
  return B200_OK;
}

// ---- self-describing dump ------------------------------------------------------------------------
static const unsigned HDR = 4096;

int b200_phase_series_unload(const b200_phase_series* ps, const char* path) {
  if (!ps || !path || !ps->data || !ps->hits) return B200_ERR_INVALID;
  const b200_observation& o = ps->obs;
  const uint64_t nprof = uint64_t(o.nchan) * o.npol * o.ndim;
  std::vector<float> prof(nprof * ps->nbin), w(nprof);
  unsigned bad = 0;
  int rc = b200_phase_series_normalise(ps, prof.data(), w.data(), &bad);
  if (rc != B200_OK) return rc;
  std::string h;
  char line[256];
  auto kv = [&](const char* k, const char* fmt, auto v) {
    char val[160];
    snprintf(val, sizeof val, fmt, v);
    snprintf(line, sizeof line, "%-20s %s\n", k, val);
    h += line;
  };
  kv("HDR_VERSION", "%s", "1.0");
  kv("HDR_SIZE", "%u", HDR);
  kv("FILE_TYPE", "%s", "B200_PHASESERIES");
  kv("TELESCOPE", "%s", o.telescope[0] ? o.telescope : "unknown");
  kv("RECEIVER", "%s", o.receiver[0] ? o.receiver : "unknown");
  kv("SOURCE", "%s", o.source[0] ? o.source : "unknown");
  kv("MODE", "%s", o.mode[0] ? o.mode : "PSR");
  kv("INSTRUMENT", "%s", o.machine[0] ? o.machine : "unknown");
  kv("FORMAT", "%s", o.format[0] ? o.format : "unknown");
  kv("FREQ", "%.17g", o.centre_frequency);
  kv("BW", "%.17g", o.bandwidth);
  kv("RATE", "%.17g", o.rate);
  kv("SCALE", "%.17g", o.scale);
  kv("DM", "%.17g", o.dispersion_measure);
  kv("RM", "%.17g", o.rotation_measure);
  kv("NCHAN", "%u", o.nchan);
  kv("NPOL", "%u", o.npol);
  kv("NDIM", "%u", o.ndim);
  kv("NBIT", "%u", o.nbit);
  kv("STATE", "%d", o.state);
  kv("TYPE", "%d", o.type);
  kv("BASIS", "%d", o.basis);
  kv("SWAP", "%d", o.swap);
  kv("NSUB_SWAP", "%d", o.nsub_swap);
  kv("DC_CENTRED", "%d", o.dc_centred);
  kv("NBIN", "%u", ps->nbin);
  kv("HITS_NCHAN", "%u", ps->hits_nchan);
  kv("MJD_START_DAY", "%d", o.start_time.day);
  kv("MJD_START_SEC", "%d", o.start_time.sec);
  kv("MJD_START_FRAC", "%.17g", o.start_time.frac);
  kv("MJD_END_DAY", "%d", ps->end_time.day);
  kv("MJD_END_SEC", "%d", ps->end_time.sec);
  kv("MJD_END_FRAC", "%.17g", ps->end_time.frac);
  kv("INTEGRATION_LENGTH", "%.17g", ps->integration_length);
  kv("NDAT_TOTAL", "%llu", (unsigned long long)ps->ndat_total);
  kv("NDAT_EXPECTED", "%llu", (unsigned long long)ps->ndat_expected);
  kv("FOLDING_PERIOD", "%.17g", ps->folding_period);
  kv("REFERENCE_PHASE", "%.17g", ps->reference_phase);
  kv("CORRUPTED_PROFILES", "%u", bad);
  kv("ARRAYS", "%s", "profiles:f4[nchan][npol][ndim][nbin],weights:f4[nchan][npol][ndim],hits:u4[hits_nchan][nbin],sums:f4[nchan][npol][nbin][ndim]");
  if (h.size() >= HDR) return B200_ERR_INVALID;
  h.resize(HDR, '\0');
  FILE* f = fopen(path, "wb");
  if (!f) return B200_ERR_INVALID;
  bool ok = fwrite(h.data(), 1, HDR, f) == HDR;
  ok = ok && fwrite(prof.data(), sizeof(float), prof.size(), f) == prof.size();
  ok = ok && fwrite(w.data(), sizeof(float), w.size(), f) == w.size();
  const size_t nh = size_t(ps->hits_nchan) * ps->nbin;
  ok = ok && fwrite(ps->hits, sizeof(unsigned), nh, f) == nh;
  const size_t nd = size_t(nprofile_floats(ps));
  ok = ok && fwrite(ps->data, sizeof(float), nd, f) == nd;
  ok = (fclose(f) == 0) && ok;
  return ok ? B200_OK : B200_ERR_INVALID;
}

int b200_phase_series_load(const char* path, b200_phase_series* ps, float* h_profiles, float* h_weights,
                           unsigned* h_hits, float* h_raw) {
  if (!path || !ps) return B200_ERR_INVALID;
  FILE* f = fopen(path, "rb");
  if (!f) return B200_ERR_INVALID;
  std::vector<char> hdr(HDR + 1, 0);
  if (fread(hdr.data(), 1, HDR, f) != HDR) { fclose(f); return B200_ERR_INVALID; }
  float* data = ps->data;
  unsigned* hits = ps->hits;
  memset(ps, 0, sizeof *ps);
  ps->data = data;
  ps->hits = hits;
  b200_observation& o = ps->obs;
  unsigned long long u;
  char* save = nullptr;
  bool typed = false;
  for (char* ln = strtok_r(hdr.data(), "\n", &save); ln; ln = strtok_r(nullptr, "\n", &save)) {
    char key[64], val[192];
    if (sscanf(ln, "%63s %191s", key, val) != 2) continue;
    const std::string k = key;
    auto str = [&](char* dst, size_t n) { strncpy(dst, val, n - 1); dst[n - 1] = 0; };
    if (k == "FILE_TYPE") typed = !strcmp(val, "B200_PHASESERIES");
    else if (k == "TELESCOPE") str(o.telescope, sizeof o.telescope);
    else if (k == "RECEIVER") str(o.receiver, sizeof o.receiver);
    else if (k == "SOURCE") str(o.source, sizeof o.source);
    else if (k == "MODE") str(o.mode, sizeof o.mode);
    else if (k == "INSTRUMENT") str(o.machine, sizeof o.machine);
    else if (k == "FORMAT") str(o.format, sizeof o.format);
    else if (k == "FREQ") o.centre_frequency = atof(val);
    else if (k == "BW") o.bandwidth = atof(val);
    else if (k == "RATE") o.rate = atof(val);
    else if (k == "SCALE") o.scale = atof(val);
    else if (k == "DM") o.dispersion_measure = atof(val);
    else if (k == "RM") o.rotation_measure = atof(val);
    else if (k == "NCHAN") o.nchan = (unsigned)atoi(val);
    else if (k == "NPOL") o.npol = (unsigned)atoi(val);
    else if (k == "NDIM") o.ndim = (unsigned)atoi(val);
    else if (k == "NBIT") o.nbit = (unsigned)atoi(val);
    else if (k == "STATE") o.state = atoi(val);
    else if (k == "TYPE") o.type = atoi(val);
    else if (k == "BASIS") o.basis = atoi(val);
    else if (k == "SWAP") o.swap = atoi(val);
    else if (k == "NSUB_SWAP") o.nsub_swap = atoi(val);
    else if (k == "DC_CENTRED") o.dc_centred = atoi(val);
    else if (k == "NBIN") ps->nbin = (unsigned)atoi(val);
    else if (k == "HITS_NCHAN") ps->hits_nchan = (unsigned)atoi(val);
    else if (k == "MJD_START_DAY") o.start_time.day = atoi(val);
    else if (k == "MJD_START_SEC") o.start_time.sec = atoi(val);
    else if (k == "MJD_START_FRAC") o.start_time.frac = atof(val);
    else if (k == "MJD_END_DAY") ps->end_time.day = atoi(val);
    else if (k == "MJD_END_SEC") ps->end_time.sec = atoi(val);
    else if (k == "MJD_END_FRAC") ps->end_time.frac = atof(val);
    else if (k == "INTEGRATION_LENGTH") ps->integration_length = atof(val);
    else if (k == "NDAT_TOTAL") { sscanf(val, "%llu", &u); ps->ndat_total = u; }
    else if (k == "NDAT_EXPECTED") { sscanf(val, "%llu", &u); ps->ndat_expected = u; }
    else if (k == "FOLDING_PERIOD") ps->folding_period = atof(val);
    else if (k == "REFERENCE_PHASE") ps->reference_phase = atof(val);
  }
  if (!typed || !ps->nbin || !o.nchan) { fclose(f); return B200_ERR_INVALID; }
  const size_t nprof = size_t(o.nchan) * o.npol * o.ndim;
  const size_t nh = size_t(ps->hits_nchan) * ps->nbin;
  auto rd = [&](void* dst, size_t esz, size_t n) {
    if (dst) return fread(dst, esz, n, f) == n;
    return fseek(f, long(esz * n), SEEK_CUR) == 0;
  };
  bool ok = rd(h_profiles, 4, nprof * ps->nbin) && rd(h_weights, 4, nprof) && rd(h_hits, 4, nh) &&
            rd(h_raw, 4, nprof * ps->nbin);
  fclose(f);
  return ok ? B200_OK : B200_ERR_INVALID;
}

}  // extern "C"

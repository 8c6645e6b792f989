// oracle/orc_fft.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product)
//
// Single-precision, unnormalised FFTs with FFTW sign/layout conventions, standing in for
// PSRCHIVE's FTransform::Plan::{frc1d,fcc1d,bcc1d} -> FFTW3 (un-vendored dependency; call
// sites Signal/General/Filterbank.C:252-261,591-593,642 and Convolution.C:213-218,411-415,446).
//   fcc1d: X[k] = sum_n x[n] exp(-2 pi i k n / N)          (forward,  FFTW_FORWARD)
//   bcc1d: x[n] = sum_k X[k] exp(+2 pi i k n / N)          (backward, FFTW_BACKWARD, no 1/N)
//   frc1d: real input of N points -> N/2+1 complex bins (FFTW r2c "halfcomplex" front half)
// Algorithm: radix-4 Stockham autosort for in-cache sizes, four-step (column tiles of 16)
// above 2^15 points.  Twiddles are computed in double and rounded once to float.
// This file is compiled with FMA contraction allowed (speed of the cpu_baseline);
// the rest of the oracle is compiled with -ffp-contract=off.
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace {

typedef std::complex<float> cf;

struct Plan {
  unsigned n = 0;
  std::vector<cf> tw;          // tw[k] = exp(-2 pi i k / n), k < n (forward sign)
  // four-step
  unsigned n1 = 0, n2 = 0;     // n = n1 * n2 (n1 = column FFT length)
  std::vector<std::complex<double>> tlo, thi;   // W_n^b (b < 2048), W_n^(2048 a)
};

std::mutex plan_mutex;
std::map<unsigned, std::shared_ptr<Plan>> plans;

const unsigned FOURSTEP_MIN = 1u << 16;

std::shared_ptr<Plan> get_plan(unsigned n) {
  std::lock_guard<std::mutex> lock(plan_mutex);
  auto it = plans.find(n);
  if (it != plans.end()) return it->second;
  auto p = std::make_shared<Plan>();
  p->n = n;
  if (n < FOURSTEP_MIN) {
    p->tw.resize(n);
    for (unsigned k = 0; k < n; k++) {
      double a = -2.0 * M_PI * double(k) / double(n);
      p->tw[k] = cf(float(std::cos(a)), float(std::sin(a)));
    }
  } else {
    unsigned lg = 0;
    while ((1u << lg) < n) lg++;
    p->n1 = 1u << (lg / 2);
    p->n2 = n / p->n1;
  }
  // two-level double-precision table: W_n^m = thi[m >> 11] * tlo[m & 2047]
  {
    unsigned nlo = n < 2048 ? n : 2048;
    p->tlo.resize(nlo);
    for (unsigned b = 0; b < nlo; b++) {
      double a = -2.0 * M_PI * double(b) / double(n);
      p->tlo[b] = std::complex<double>(std::cos(a), std::sin(a));
    }
    unsigned nhi = n < 2048 ? 1 : n / 2048;
    p->thi.resize(nhi);
    for (unsigned a_ = 0; a_ < nhi; a_++) {
      double a = -2.0 * M_PI * double(a_) / double(nhi);
      p->thi[a_] = std::complex<double>(std::cos(a), std::sin(a));
    }
  }
  plans[n] = p;
  return p;
}

inline cf mul(cf a, cf b) {
  return cf(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
// multiply by -i (forward) or +i (backward)
template <bool FWD> inline cf rot(cf a) {
  return FWD ? cf(a.imag(), -a.real()) : cf(-a.imag(), a.real());
}
template <bool FWD> inline cf twd(cf w) { return FWD ? w : std::conj(w); }

// Stockham autosort, radix 4 with a trailing radix-2 stage when log2(n) is odd.
// x: input (destroyed), y: scratch; returns pointer to the buffer holding the result.
template <bool FWD>
cf* stockham(unsigned n, cf* x, cf* y, const cf* tw, unsigned tw_n) {
  unsigned Ns = 1;
  cf* in = x;
  cf* out = y;
  const unsigned tstep0 = tw_n / n;   // tw is a table for size tw_n >= n
  while (Ns * 4 <= n) {
    const unsigned q = n / 4;
    const unsigned ts = tstep0 * (n / (Ns * 4));
    if (Ns == 1) {
      for (unsigned j = 0; j < q; j++) {
        cf a = in[j], b = in[j + q], c = in[j + 2 * q], d = in[j + 3 * q];
        cf s0 = a + c, s1 = a - c, s2 = b + d, s3 = rot<FWD>(b - d);
        cf* o = out + 4 * j;
        o[0] = s0 + s2; o[1] = s1 + s3; o[2] = s0 - s2; o[3] = s1 - s3;
      }
    } else {
      for (unsigned jb = 0; jb < q; jb += Ns) {
        cf* o = out + 4 * jb;
        const cf* i0 = in + jb;
        for (unsigned k = 0; k < Ns; k++) {
          cf w1 = twd<FWD>(tw[k * ts]);
          cf w2 = twd<FWD>(tw[2 * k * ts]);
          cf w3 = twd<FWD>(tw[3 * k * ts]);
          cf a = i0[k], b = mul(i0[k + q], w1), c = mul(i0[k + 2 * q], w2), d = mul(i0[k + 3 * q], w3);
          cf s0 = a + c, s1 = a - c, s2 = b + d, s3 = rot<FWD>(b - d);
          o[k] = s0 + s2; o[k + Ns] = s1 + s3; o[k + 2 * Ns] = s0 - s2; o[k + 3 * Ns] = s1 - s3;
        }
      }
    }
    std::swap(in, out);
    Ns *= 4;
  }
  if (Ns < n) {   // one radix-2 stage left
    const unsigned h = n / 2;
    const unsigned ts = tstep0 * (n / (Ns * 2));
    for (unsigned jb = 0; jb < h; jb += Ns) {
      cf* o = out + 2 * jb;
      const cf* i0 = in + jb;
      for (unsigned k = 0; k < Ns; k++) {
        cf w1 = twd<FWD>(tw[k * ts]);
        cf a = i0[k], b = mul(i0[k + h], w1);
        o[k] = a + b; o[k + Ns] = a - b;
      }
    }
    std::swap(in, out);
  }
  return in;
}

template <bool FWD>
void fft_small(unsigned n, cf* out, const cf* in) {
  if (n == 1) { out[0] = in[0]; return; }
  auto p = get_plan(n);
  std::vector<cf> a(in, in + n), b(n);
  cf* r = stockham<FWD>(n, a.data(), b.data(), p->tw.data(), n);
  std::memcpy(out, r, sizeof(cf) * n);
}

// Four-step: n = n1*n2, input index = n2*i1 + i2, output index = k1 + n1*k2.
template <bool FWD>
void fft_large(unsigned n, cf* out, const cf* in) {
  auto p = get_plan(n);
  const unsigned n1 = p->n1, n2 = p->n2;
  auto p1 = get_plan(n1);
  auto p2 = get_plan(n2);
  std::vector<cf> work(n);
  const unsigned W = 16;   // columns per tile
  std::vector<cf> ta(n1 * W), tb(n1 * W);
  // pass A: FFT of length n1 down each column i2; twiddle by W_n^(i2*k1); store work[k1*n2 + i2]
  for (unsigned c0 = 0; c0 < n2; c0 += W) {
    for (unsigned i1 = 0; i1 < n1; i1++)
      for (unsigned c = 0; c < W; c++) ta[c * n1 + i1] = in[uint64_t(i1) * n2 + c0 + c];
    for (unsigned c = 0; c < W; c++) {
      cf* r = stockham<FWD>(n1, ta.data() + c * n1, tb.data() + c * n1, p1->tw.data(), n1);
      const unsigned i2 = c0 + c;
      for (unsigned k1 = 0; k1 < n1; k1++) {
        uint64_t m = (uint64_t(i2) * k1) % n;
        std::complex<double> w = p->thi[m >> 11] * p->tlo[m & 2047];
        if (!FWD) w = std::conj(w);
        std::complex<double> v(r[k1].real(), r[k1].imag());
        v *= w;
        work[uint64_t(k1) * n2 + i2] = cf(float(v.real()), float(v.imag()));
      }
    }
  }
  // pass B: FFT of length n2 along each row k1; out[k1 + n1*k2]
  std::vector<cf> ra(n2), rb(n2);
  for (unsigned k1 = 0; k1 < n1; k1++) {
    std::memcpy(ra.data(), work.data() + uint64_t(k1) * n2, sizeof(cf) * n2);
    cf* r = stockham<FWD>(n2, ra.data(), rb.data(), p2->tw.data(), n2);
    for (unsigned k2 = 0; k2 < n2; k2++) out[k1 + uint64_t(n1) * k2] = r[k2];
  }
}

template <bool FWD>
void fft(unsigned n, cf* out, const cf* in) {
  if (n < FOURSTEP_MIN) fft_small<FWD>(n, out, in);
  else fft_large<FWD>(n, out, in);
}

}  // namespace

extern "C" {

// FTransform::Plan::fcc1d (forward complex-to-complex), Filterbank.C:593, Convolution.C:415
void orc_fft_fcc1d(unsigned n, float* out, const float* in) {
  fft<true>(n, reinterpret_cast<cf*>(out), reinterpret_cast<const cf*>(in));
}

// FTransform::Plan::bcc1d (backward complex-to-complex), Filterbank.C:642, Convolution.C:446
void orc_fft_bcc1d(unsigned n, float* out, const float* in) {
  fft<false>(n, reinterpret_cast<cf*>(out), reinterpret_cast<const cf*>(in));
}

// FTransform::Plan::frc1d (forward real-to-complex, n real points -> n/2+1 complex bins),
// Filterbank.C:591, Convolution.C:412.  Half-length complex FFT of the even/odd packing,
// followed by the standard split (what FFTW's r2c codelets compute for even n).
void orc_fft_frc1d(unsigned n, float* out, const float* in) {
  const unsigned h = n / 2;
  auto pn = get_plan(n);
  std::vector<cf> z(h);
  fft<true>(h, z.data(), reinterpret_cast<const cf*>(in));
  cf* X = reinterpret_cast<cf*>(out);
  for (unsigned k = 0; k <= h; k++) {
    cf zk = z[k % h];
    cf zm = std::conj(z[(h - k) % h]);
    std::complex<double> e(0.5 * (double(zk.real()) + zm.real()), 0.5 * (double(zk.imag()) + zm.imag()));
    std::complex<double> d(0.5 * (double(zk.real()) - zm.real()), 0.5 * (double(zk.imag()) - zm.imag()));
    std::complex<double> o(d.imag(), -d.real());   // -i * d
    std::complex<double> w = (k == h) ? std::complex<double>(-1.0, 0.0)
                                      : pn->thi[k >> 11] * pn->tlo[k & 2047];
    std::complex<double> x = e + w * o;
    X[k] = cf(float(x.real()), float(x.imag()));
  }
}

}  // extern "C"

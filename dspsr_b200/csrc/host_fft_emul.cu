// host_fft_emul.cu -- CPU emulation of the block FFT (fft_core.cuh) used by the CPU-only
// test-suite: runs the very same __host__ __device__ stage code thread-by-thread with an
// emulated shared memory, and compares against a double-precision DFT.
// Build: nvcc -std=c++17 -O2 host_fft_emul.cu -o host_fft_emul   (no GPU needed to run)
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_core.cuh"
#include "fft_c2.cuh"

using namespace b200;

template <int EPT, bool INV, typename Map>
static void run_stage(int R, std::vector<float2>& regs, unsigned T, unsigned Ns, const float2* tw, unsigned NT,
                      std::vector<float2>& smem, std::vector<Map>& maps, bool last, std::vector<float2>& out) {
  for (unsigned j = 0; j < T; j++) {
    float2* v = &regs[size_t(j) * EPT];
    switch (R) {
      case 2: if constexpr (EPT >= 2) stage_compute<EPT, 2, INV>(v, j, T, Ns, tw, NT); break;
      case 4: if constexpr (EPT >= 4) stage_compute<EPT, 4, INV>(v, j, T, Ns, tw, NT); break;
      case 8: if constexpr (EPT >= 8) stage_compute<EPT, 8, INV>(v, j, T, Ns, tw, NT); break;
      case 16: if constexpr (EPT >= 16) stage_compute<EPT, 16, INV>(v, j, T, Ns, tw, NT); break;
      case 32: if constexpr (EPT >= 32) stage_compute<EPT, 32, INV>(v, j, T, Ns, tw, NT); break;
    }
  }
  // "syncthreads", then scatter
  for (unsigned j = 0; j < T; j++) {
    const int NB = EPT / R;
    for (int q = 0; q < NB; q++)
      for (int r = 0; r < R; r++) {
        unsigned b = j + q * T, k = b & (Ns - 1);
        unsigned d = (b - k) * R + k + r * Ns;
        if (last) out[d] = regs[size_t(j) * EPT + q + r * NB];
        else smem[maps[j](d)] = regs[size_t(j) * EPT + q + r * NB];
      }
  }
  if (!last)
    for (unsigned j = 0; j < T; j++)
      for (int e = 0; e < EPT; e++) regs[size_t(j) * EPT + e] = smem[maps[j](j + e * T)];
}

template <int EPT, bool INV>
static double test_one(unsigned N, int mapkind) {
  const unsigned T = N / EPT;
  std::vector<float2> tw(N);
  for (unsigned m = 0; m < N; m++) {
    double a = -2.0 * M_PI * m / N;
    tw[m] = make_float2(float(cos(a)), float(sin(a)));
  }
  std::vector<std::complex<double>> x(N);
  srand(N + EPT);
  for (auto& z : x) z = {rand() / double(RAND_MAX) - 0.5, rand() / double(RAND_MAX) - 0.5};
  std::vector<float2> regs(size_t(T) * EPT);
  for (unsigned j = 0; j < T; j++)
    for (int e = 0; e < EPT; e++) regs[size_t(j) * EPT + e] = make_float2(float(x[j + e * T].real()), float(x[j + e * T].imag()));
  RadixPlan plan = make_radix_plan(N, EPT);
  unsigned sh = 0;
  while ((1u << sh) < (unsigned)plan.radix[0]) sh++;
  std::vector<float2> out(N);
  double err = 0;
  if (mapkind == 0) {
    std::vector<MapRows> maps(T);
    for (auto& m : maps) m = MapRows{16, sh, 5};
    std::vector<float2> smem(N + 64);
    unsigned Ns = 1;
    for (int s = 0; s < plan.nstage; s++) {
      run_stage<EPT, INV>(plan.radix[s], regs, T, Ns, tw.data(), N, smem, maps, s == plan.nstage - 1, out);
      Ns *= plan.radix[s];
    }
  } else {
    // emulate B interleaved transforms, test transform b = 3 of B = 8
    std::vector<MapCols> maps(T);
    for (auto& m : maps) m = MapCols{3, 3, sh};
    std::vector<float2> smem(size_t(N) * 8 + 64);
    unsigned Ns = 1;
    for (int s = 0; s < plan.nstage; s++) {
      run_stage<EPT, INV>(plan.radix[s], regs, T, Ns, tw.data(), N, smem, maps, s == plan.nstage - 1, out);
      Ns *= plan.radix[s];
    }
  }
  // reference DFT (O(N^2) for small N, recursive split for large)
  std::vector<std::complex<double>> X(N);
  if (N <= 2048) {
    for (unsigned k = 0; k < N; k++) {
      std::complex<double> acc = 0;
      for (unsigned n = 0; n < N; n++) {
        double a = (INV ? 2.0 : -2.0) * M_PI * double((uint64_t(k) * n) % N) / N;
        acc += x[n] * std::complex<double>(cos(a), sin(a));
      }
      X[k] = acc;
    }
  } else {
    // iterative radix-2 in double
    std::vector<std::complex<double>> a(x);
    unsigned lg = 0;
    while ((1u << lg) < N) lg++;
    for (unsigned i = 0; i < N; i++) {
      unsigned r = 0;
      for (unsigned bit = 0; bit < lg; bit++) if (i & (1u << bit)) r |= 1u << (lg - 1 - bit);
      if (r > i) std::swap(a[i], a[r]);
    }
    for (unsigned len = 2; len <= N; len <<= 1) {
      double ang = (INV ? 2.0 : -2.0) * M_PI / len;
      for (unsigned i = 0; i < N; i += len)
        for (unsigned k = 0; k < len / 2; k++) {
          std::complex<double> w(cos(ang * k), sin(ang * k));
          auto u = a[i + k], t = a[i + k + len / 2] * w;
          a[i + k] = u + t;
          a[i + k + len / 2] = u - t;
        }
    }
    X = a;
  }
  double rms = 0;
  for (unsigned k = 0; k < N; k++) rms += std::norm(X[k]);
  rms = sqrt(rms / N);
  for (unsigned k = 0; k < N; k++) {
    double e = std::abs(std::complex<double>(out[k].x, out[k].y) - X[k]) / rms;
    if (e > err) err = e;
  }
  return err;
}

template <int EPT>
static int sweep(unsigned nmin, unsigned nmax) {
  int bad = 0;
  for (unsigned N = nmin; N <= nmax; N <<= 1) {
    double e0 = test_one<EPT, false>(N, 0), e1 = test_one<EPT, true>(N, 0), e2 = test_one<EPT, false>(N, 1);
    printf("EPT=%d N=%u fwd %.3e inv %.3e cols %.3e\n", EPT, N, e0, e1, e2);
    if (!(e0 < 2e-6 && e1 < 2e-6 && e2 < 2e-6)) bad++;
  }
  return bad;
}


// ---- compile-time-sized stages (stage_compute_ct / dft_tw) ------------------------------------
template <int EPT, bool INV, unsigned N, unsigned NS>
static void ct_stages(std::vector<float2>& regs, const float2* tw, std::vector<float2>& smem, const MapRows& map,
                      std::vector<float2>& out) {
  constexpr unsigned T = N / EPT;
  constexpr unsigned REM = N / NS;
  constexpr int R = REM >= (unsigned)EPT ? EPT : (int)REM;
  constexpr int NB = EPT / R;
  constexpr bool last = (REM == (unsigned)R);
  for (unsigned j = 0; j < T; j++) stage_compute_ct<EPT, R, INV, N, NS>(&regs[size_t(j) * EPT], j, tw);
  for (unsigned j = 0; j < T; j++)
    for (int q = 0; q < NB; q++)
      for (int r = 0; r < R; r++) {
        unsigned b = j + q * T, k = b & (NS - 1);
        unsigned d = (b - k) * R + k + r * NS;
        if (last) out[d] = regs[size_t(j) * EPT + q + r * NB];
        else smem[map(d)] = regs[size_t(j) * EPT + q + r * NB];
      }
  if constexpr (!last) {
    for (unsigned j = 0; j < T; j++)
      for (int e = 0; e < EPT; e++) regs[size_t(j) * EPT + e] = smem[map(j + e * T)];
    ct_stages<EPT, INV, N, NS * R>(regs, tw, smem, map, out);
  }
}

template <int EPT, bool INV, unsigned N>
static double test_ct() {
  const unsigned T = N / EPT;
  std::vector<float2> tw(stage_table_size(N, EPT));
  {
    unsigned ns = 1, off = 0;
    while (ns < N) {
      int R = stage_radix(N, EPT, ns);
      if (ns > 1) {
        for (int mi = 0; mi < stage_nmult(R); mi++)
          for (unsigned k = 0; k < ns; k++) {
            double a = -2.0 * M_PI * double(stage_mult(R, mi)) * k / (double(ns) * R);
            tw[off + mi * ns + k] = make_float2(float(cos(a)), float(sin(a)));
          }
        off += stage_nmult(R) * ns;
      }
      ns *= R;
    }
  }
  std::vector<std::complex<double>> x(N);
  srand(N * 7 + EPT);
  for (auto& z : x) z = {rand() / double(RAND_MAX) - 0.5, rand() / double(RAND_MAX) - 0.5};
  std::vector<float2> regs(size_t(T) * EPT), smem(N + 64), out(N);
  for (unsigned j = 0; j < T; j++)
    for (int e = 0; e < EPT; e++)
      regs[size_t(j) * EPT + e] = make_float2(float(x[j + e * T].real()), float(x[j + e * T].imag()));
  unsigned sh = 0;
  while ((1u << sh) < (unsigned)(N >= (unsigned)EPT ? EPT : N)) sh++;
  MapRows map{8, sh, 3};
  ct_stages<EPT, INV, N, 1>(regs, tw.data(), smem, map, out);
  // reference: iterative radix-2 in double
  std::vector<std::complex<double>> a(x);
  unsigned lg = 0;
  while ((1u << lg) < N) lg++;
  for (unsigned i = 0; i < N; i++) {
    unsigned r = 0;
    for (unsigned bit = 0; bit < lg; bit++) if (i & (1u << bit)) r |= 1u << (lg - 1 - bit);
    if (r > i) std::swap(a[i], a[r]);
  }
  for (unsigned len = 2; len <= N; len <<= 1) {
    double ang = (INV ? 2.0 : -2.0) * M_PI / len;
    for (unsigned i = 0; i < N; i += len)
      for (unsigned k = 0; k < len / 2; k++) {
        std::complex<double> w(cos(ang * k), sin(ang * k));
        auto u = a[i + k], t = a[i + k + len / 2] * w;
        a[i + k] = u + t;
        a[i + k + len / 2] = u - t;
      }
  }
  double rms = 0, err = 0;
  for (unsigned k = 0; k < N; k++) rms += std::norm(a[k]);
  rms = sqrt(rms / N);
  for (unsigned k = 0; k < N; k++) {
    double e = std::abs(std::complex<double>(out[k].x, out[k].y) - a[k]) / rms;
    if (e > err) err = e;
  }
  return err;
}

#define CT_CASE(EPT, N)                                                     \
  {                                                                         \
    double e0 = test_ct<EPT, false, N>(), e1 = test_ct<EPT, true, N>();     \
    printf("CT EPT=%d N=%d fwd %.3e inv %.3e\n", EPT, N, e0, e1);           \
    if (!(e0 < 2e-6 && e1 < 2e-6)) bad++;                                   \
  }

static int sweep_ct() {
  int bad = 0;
  CT_CASE(16, 16) CT_CASE(16, 256) CT_CASE(16, 1024) CT_CASE(16, 2048) CT_CASE(16, 4096) CT_CASE(16, 8192)
  CT_CASE(32, 32) CT_CASE(32, 64) CT_CASE(32, 512) CT_CASE(32, 1024) CT_CASE(32, 2048) CT_CASE(32, 4096)
  CT_CASE(32, 8192) CT_CASE(32, 16384) CT_CASE(8, 64) CT_CASE(8, 512)
  return bad;
}

// ---- fft_c2.cuh: two sequences per thread, padded affine exchange -----------------------------------
template <unsigned L, bool INV, int S>
static void c2_stages(std::vector<float2>& ra, std::vector<float2>& rb, std::vector<float4>& smem,
                      const float2* tw, std::vector<int>& touched) {
  constexpr unsigned T = L / 16;
  for (unsigned j = 0; j < T; j++) c2::stage_compute<L, S, INV>(&ra[j * 16], &rb[j * 16], j, tw);
  if constexpr (S + 1 < c2::Plan<L>::nstage) {
    std::fill(smem.begin(), smem.end(), make_float4(NAN, NAN, NAN, NAN));
    for (unsigned j = 0; j < T; j++) c2::scatter<L, S>(smem.data(), &ra[j * 16], &rb[j * 16], j);
    for (unsigned j = 0; j < T; j++) c2::gather<L>(smem.data(), &ra[j * 16], &rb[j * 16], j);
    c2_stages<L, INV, S + 1>(ra, rb, smem, tw, touched);
  }
}

template <unsigned L, bool INV>
static int c2_test() {
  constexpr unsigned T = L / 16;
  std::vector<float2> tw(c2::twiddle_count<L>());
  c2::fill_twiddles<L>(tw.data());
  std::vector<std::complex<double>> xa(L), xb(L);
  srand(L * 3 + INV);
  for (unsigned i = 0; i < L; i++) {
    xa[i] = {rand() / double(RAND_MAX) - 0.5, rand() / double(RAND_MAX) - 0.5};
    xb[i] = {rand() / double(RAND_MAX) - 0.5, rand() / double(RAND_MAX) - 0.5};
  }
  std::vector<float2> ra(L), rb(L);
  for (unsigned j = 0; j < T; j++)
    for (int e = 0; e < 16; e++) {
      ra[j * 16 + e] = make_float2(float(xa[j + e * T].real()), float(xa[j + e * T].imag()));
      rb[j * 16 + e] = make_float2(float(xb[j + e * T].real()), float(xb[j + e * T].imag()));
    }
  std::vector<float4> smem(c2::pair_slots<L>());
  std::vector<int> touched;
  c2_stages<L, INV, 0>(ra, rb, smem, tw.data(), touched);
  // reference: double DFT via the O(L^2)-free route: recursive radix-2 in double
  auto dft = [&](std::vector<std::complex<double>> x) {
    const unsigned n = x.size();
    for (unsigned i = 1, jj = 0; i < n; i++) {
      unsigned bit = n >> 1;
      for (; jj & bit; bit >>= 1) jj ^= bit;
      jj ^= bit;
      if (i < jj) std::swap(x[i], x[jj]);
    }
    for (unsigned len = 2; len <= n; len <<= 1) {
      double ang = (INV ? 2.0 : -2.0) * M_PI / len;
      for (unsigned i = 0; i < n; i += len)
        for (unsigned k = 0; k < len / 2; k++) {
          std::complex<double> w(cos(ang * k), sin(ang * k));
          auto u = x[i + k], v = x[i + k + len / 2] * w;
          x[i + k] = u + v;
          x[i + k + len / 2] = u - v;
        }
    }
    return x;
  };
  auto Xa = dft(xa), Xb = dft(xb);
  double err = 0, rms = 0;
  for (unsigned j = 0; j < T; j++)
    for (int e = 0; e < 16; e++) {
      const unsigned k = j + e * T;
      err = std::max(err, std::abs(std::complex<double>(ra[j * 16 + e].x, ra[j * 16 + e].y) - Xa[k]));
      err = std::max(err, std::abs(std::complex<double>(rb[j * 16 + e].x, rb[j * 16 + e].y) - Xb[k]));
      rms += std::norm(Xa[k]);
    }
  rms = sqrt(rms / L);
  const bool ok = err / rms < 2e-6;
  if (!ok) printf("c2 L=%u inv=%d rel err %.3e FAIL\n", L, int(INV), err / rms);
  return ok ? 0 : 1;
}

static int sweep_c2() {
  int bad = 0;
  bad += c2_test<256, false>() + c2_test<256, true>();
  bad += c2_test<512, false>() + c2_test<512, true>();
  bad += c2_test<1024, false>() + c2_test<1024, true>();
  bad += c2_test<2048, false>() + c2_test<2048, true>();
  bad += c2_test<4096, false>() + c2_test<4096, true>();
  bad += c2_test<8192, false>() + c2_test<8192, true>();
  return bad;
}

int main() {
  int bad = 0;
  bad += sweep_ct();
  bad += sweep_c2();
  bad += sweep<2>(2, 2);
  bad += sweep<4>(4, 64);
  bad += sweep<8>(8, 512);
  bad += sweep<16>(16, 16384);
  bad += sweep<32>(32, 16384);
  printf(bad ? "FAIL %d\n" : "OK\n", bad);
  return bad ? 1 : 0;
}

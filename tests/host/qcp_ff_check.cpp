// Host-side check of the float-float quartic builder (qcp_poly_ff) against the fp64 builder
// (qcp_poly) and a long-double evaluation, on integer covariances exactly as the tensor
// kernel's epilogue sees them.  Built and run by tests/test_qcp_host.py (g++, no GPU).
// Output: one line per scenario
//   "name cases max|dq0| max|dy| max|drmsd_A| fallbacks max|drmsd_A|(fp64 coefficients) fallbacks(fp64 coefficients)".
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <cmath>
#include <random>
#include <vector>

#include "../../loos_b200/csrc/qcp.cuh"

using namespace loosb200;

struct Case { double R[9]; double ga, gb; int N; double unit; };

static void rotation(std::mt19937_64& g, double Q[9]) {
  std::normal_distribution<double> n(0, 1);
  double q[4], s = 0;
  for (double& x : q) { x = n(g); s += x * x; }
  s = 1 / std::sqrt(s);
  for (double& x : q) x *= s;
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double M[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                       2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
  memcpy(Q, M, sizeof(M));
}

static Case make_case(std::mt19937_64& g, int N, double extent, double noise, bool unrelated, int shape = 0) {
  std::normal_distribution<double> n(0, 1);
  std::vector<double> A(3 * N), B(3 * N);
  for (double& x : A) x = extent * n(g);
  if (shape == 1) for (int k = 0; k < N; ++k) { A[3 * k + 1] = 0.3 * A[3 * k]; A[3 * k + 2] = -0.7 * A[3 * k]; }  // collinear
  if (shape == 2) for (int k = 0; k < N; ++k) A[3 * k + 2] = 0.5 * A[3 * k] - 0.25 * A[3 * k + 1];                // planar
  double Q[9];
  rotation(g, Q);
  for (int k = 0; k < N; ++k)
    for (int r = 0; r < 3; ++r) {
      double v = 0;
      for (int c = 0; c < 3; ++c) v += Q[3 * r + c] * A[3 * k + c];
      B[3 * k + r] = unrelated ? extent * n(g) : v + noise * n(g);
      if (shape == 3 && r == 0) B[3 * k + r] = -B[3 * k + r];  // mirror image
    }
  // center
  for (auto* X : {&A, &B}) {
    double c[3] = {0, 0, 0};
    for (int k = 0; k < N; ++k) for (int r = 0; r < 3; ++r) c[r] += (*X)[3 * k + r];
    for (int k = 0; k < N; ++k) for (int r = 0; r < 3; ++r) (*X)[3 * k + r] -= c[r] / N;
  }
  double mx = 0;
  for (double x : A) mx = std::fmax(mx, std::fabs(x));
  for (double x : B) mx = std::fmax(mx, std::fabs(x));
  int e = (int)std::ceil(std::log2(mx / 8388607.0));
  if (e < -20) e = -20;
  const double unit = std::ldexp(1.0, e);
  Case cs;
  cs.N = N; cs.unit = unit;
  __int128 R[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long ga = 0, gb = 0;
  for (int k = 0; k < N; ++k) {
    long long a[3], b[3];
    for (int r = 0; r < 3; ++r) { a[r] = std::llround(A[3 * k + r] / unit); b[r] = std::llround(B[3 * k + r] / unit); }
    for (int r = 0; r < 3; ++r) {
      ga += a[r] * a[r]; gb += b[r] * b[r];
      for (int c = 0; c < 3; ++c) R[3 * r + c] += (__int128)a[r] * b[c];
    }
  }
  for (int k = 0; k < 9; ++k) cs.R[k] = (double)R[k];
  cs.ga = (double)ga; cs.gb = (double)gb;
  return cs;
}

// the epilogue's conversion: double -> (hi, lo) floats by mantissa masking, then power-of-two scaling
static void split(double v, float sc, float& h, float& l) {
  uint64_t b;
  memcpy(&b, &v, 8);
  b &= ~((1ull << 29) - 1);
  double hd;
  memcpy(&hd, &b, 8);
  h = (float)hd * sc;
  l = (float)(v - hd) * sc;
}

// Newton in fp32 + one float-float correction, fp64 root finder for the rare cases: the epilogue's solve
static double solve(const QcpPolyF& p, bool* fell_back) {
  float t = p.t0, ls = 0;
  for (int it = 0; it < 12; ++it) {
    ls = qcp_newton_step(p.c0h, p.c1h, p.c2h, p.c3, t);
    if (it >= 1 && std::fabs(ls) <= 4e-7f * std::fabs(t) + 1e-12f) break;
  }
  bool ok;
  const float y = qcp_polish_f32(p, p.xform, t, ls, &ok);
  *fell_back = !ok || p.xform;
  if (*fell_back) return qcp_root_fp64(p, ok ? t : p.t0);
  return y;
}

// coefficients in float-float from split operands (qcp_poly_ff)
static double y_ff(const Case& c, bool* fell_back, double* q0) {
  const double e0 = 0.5 * (c.ga + c.gb);
  int ex;
  std::frexp(e0, &ex);  // e0 = m 2^ex, m in [0.5, 1)
  const float sc = std::ldexp(1.0f, -ex);
  float Rh[9], Rl[9], e0h, e0l;
  for (int k = 0; k < 9; ++k) split(c.R[k], sc, Rh[k], Rl[k]);
  split(e0, sc, e0h, e0l);
  const QcpPolyF p = qcp_poly_ff(Rh, Rl, e0h, e0l);
  *q0 = (double)p.c0h + (double)p.c0l;
  return solve(p, fell_back);
}

// coefficients in fp64 then split (what k_pairs_tensor does inside its fp64 window)
static double y_split(const Case& c, bool* fell_back) {
  const QcpPolyF p = qcp_poly_split(qcp_poly(c.R, 0.5 * (c.ga + c.gb)));
  return solve(p, fell_back);
}

int main() {
  std::mt19937_64 g(20261019);
  struct Sc { const char* name; int N; double extent, noise; bool unrelated; int cases; int shape; };
  const Sc scs[] = {
      {"identical_N1000", 1000, 15.0, 0.0, false, 2000},   {"noise1e-4_N1000", 1000, 15.0, 1e-4, false, 2000},
      {"noise1e-2_N1000", 1000, 15.0, 1e-2, false, 2000},  {"noise1_N1000", 1000, 15.0, 1.0, false, 2000},
      {"noise5_N1000", 1000, 15.0, 5.0, false, 2000},      {"unrelated_N1000", 1000, 15.0, 0.0, true, 2000},
      {"identical_N10000", 10000, 40.0, 0.0, false, 300},  {"noise1e-3_N43000", 43000, 60.0, 1e-3, false, 60},
      {"noise1_N3", 3, 5.0, 1.0, false, 2000},             {"noise0.1_N4", 4, 5.0, 0.1, false, 2000},
      {"noise1_N2", 2, 5.0, 1.0, false, 2000},             {"noise1_N30", 30, 8.0, 1.0, false, 2000},
      {"tiny_extent_N100", 100, 1e-3, 1e-5, false, 2000},  {"huge_extent_N100", 100, 5e3, 1.0, false, 2000},
      {"collinear_N64", 64, 10.0, 0.0, false, 1000, 1},    {"collinear_noise1e-4_N64", 64, 10.0, 1e-4, false, 1000, 1},
      {"collinear_noise1e-2_N64", 64, 10.0, 1e-2, false, 1000, 1}, {"collinear_noise1_N64", 64, 10.0, 1.0, false, 1000, 1},
      {"collinear_noise1e-3_N3", 3, 10.0, 1e-3, false, 2000, 1},   {"planar_N64", 64, 10.0, 0.0, false, 1000, 2},
      {"planar_noise1e-2_N64", 64, 10.0, 1e-2, false, 1000, 2},    {"mirror_N64", 64, 10.0, 0.0, false, 1000, 3},
      {"mirror_noise1_N1000", 1000, 15.0, 1.0, false, 1000, 3},    {"mirror_planar_N64", 64, 10.0, 1e-3, false, 1000, 2},
  };
  for (const Sc& sc : scs) {
    double mq = 0, my = 0, mr = 0, mr2 = 0;
    int fb = 0, fb2 = 0;
    for (int k = 0; k < sc.cases; ++k) {
      const Case c = make_case(g, sc.N, sc.extent, sc.noise, sc.unrelated, sc.shape);
      const double e0 = 0.5 * (c.ga + c.gb);
      if (!(e0 > 0)) continue;
      const QcpResult ref = qcp_lambda(c.R, c.ga, c.gb);
      (void)ref;
      double Rs[9], Ks[10];
      for (int q = 0; q < 9; ++q) Rs[q] = c.R[q] / e0;
      qcp_build_K(Rs, Ks);
      const double y_ref = std::fabs(1.0 - qcp_jacobi_max(Ks, 0));  // cyclic Jacobi: 1e-16 absolute
      const QcpPoly pd = qcp_poly(c.R, e0);
      bool fell;
      double q0;
      const double y = y_ff(c, &fell, &q0);
      fb += fell;
      if (!pd.xform) mq = std::fmax(mq, std::fabs(q0 - pd.c0));
      my = std::fmax(my, std::fabs(y - y_ref));
      const double r_ref = c.unit * std::sqrt(y_ref * 2.0 * e0 / c.N), r = c.unit * std::sqrt(y * 2.0 * e0 / c.N);
      mr = std::fmax(mr, std::fabs(r - r_ref));
      bool fell2;
      const double y2 = y_split(c, &fell2);
      fb2 += fell2;
      mr2 = std::fmax(mr2, std::fabs(c.unit * std::sqrt(y2 * 2.0 * e0 / c.N) - r_ref));
    }
    printf("%s %d %.3e %.3e %.3e %d %.3e %d\n", sc.name, sc.cases, mq, my, mr, fb, mr2, fb2);
  }
  return 0;
}

// Host-side restatement of the per-pair arithmetic of k_pairs_align (loos_b200/csrc/pairs_align.cu)
// on top of the SAME solver header the kernel uses (qcp.cuh compiled with g++): centroids of the
// align selections, R over the align selection -> Z, C over the rmsd selection about the align
// centroids, d^2 = (gx + gy - 2 sum Z_ab C_ba) / Nr.  tests/test_qcp_host.py compares it with the
// oracle's restatement of Packages/Clustering/rmsds-align.py, so that sign / index conventions are
// checked without a GPU.  Inputs are xyz-interleaved float32.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../loos_b200/csrc/qcp.cuh"

using namespace loosb200;

static void centroid(const float* x, int n, double* c) {
  for (int a = 0; a < 3; ++a) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += (double)x[3 * k + a];
    c[a] = s / n;
  }
}

extern "C" double align_pair(const float* alnA, const float* alnB, int Na, const float* rmsA, const float* rmsB, int Nr,
                             int* used_fallback) {
  double ca[3], cb[3];
  centroid(alnA, Na, ca);
  centroid(alnB, Na, cb);
  double R[9] = {0}, C[9] = {0}, ga = 0, gb = 0, gx = 0, gy = 0;
  for (int k = 0; k < Na; ++k) {
    double u[3], v[3];
    for (int a = 0; a < 3; ++a) { u[a] = (double)alnA[3 * k + a] - ca[a]; v[a] = (double)alnB[3 * k + a] - cb[a]; }
    for (int p = 0; p < 3; ++p) {
      ga = fma(u[p], u[p], ga); gb = fma(v[p], v[p], gb);
      for (int q = 0; q < 3; ++q) R[3 * p + q] = fma(u[p], v[q], R[3 * p + q]);
    }
  }
  for (int k = 0; k < Nr; ++k) {
    double u[3], v[3];
    for (int a = 0; a < 3; ++a) { u[a] = (double)rmsA[3 * k + a] - ca[a]; v[a] = (double)rmsB[3 * k + a] - cb[a]; }
    for (int p = 0; p < 3; ++p) {
      gx = fma(u[p], u[p], gx); gy = fma(v[p], v[p], gy);
      for (int q = 0; q < 3; ++q) C[3 * p + q] = fma(u[p], v[q], C[3 * p + q]);
    }
  }
  double Z[9];
  const bool fast = qcp_rotation_fast(R, ga, gb, Z);
  if (!fast) qcp_rotation(R, Z, nullptr);
  if (used_fallback) *used_fallback = fast ? 0 : 1;
  double tr = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) tr = fma(Z[3 * r + c], C[3 * c + r], tr);
  const double ss = gx + gy - 2.0 * tr;
  return sqrt(fmax(ss, 0.0) / (double)Nr);
}

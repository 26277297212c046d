// Per-pair superposition solver shared by every kernel of the RMSD path.
//
// The reference gets the minimal RMSD from the singular values of the 3x3
// covariance R = sum_k u_k v_k^T:  rmsd = sqrt(|Ga + Gb - 2 (s0 + s1 +- s2)| / N),
// the sign being that of det(U) det(V)  (src/alignment.cpp:21-87,115-140).
// s0 + s1 + sign(det R) s2 is exactly the largest eigenvalue of Horn's symmetric
// traceless 4x4 matrix K(R), so we obtain it as the largest root of K's
// characteristic quartic (Theobald's QCP) by Newton iteration from the upper
// bound (Ga+Gb)/2, with a Jacobi eigen-solve of K as the safeguard.  The
// eigenvector (a unit quaternion) gives the rotation Z = V U'^T that
// src/alignment.cpp:183-214 (kabschCentered) returns.
//
// Everything is fp64 and scaled by 1/E0 (E0 = (Ga+Gb)/2) so that the integer
// operands of the tensor-core path (|R| up to 2^57) and Angstrom^2 operands of
// the fp64 path go through the same code.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD static inline
#endif

namespace loosb200 {

struct QcpResult {
  double msd_x2_over_e0;  // (Ga + Gb - 2 lambda) / E0 = 2 (1 - x), clamped at >= 0 by |.|
  double x;               // lambda / E0
  int iters;              // Newton iterations used (<0: Jacobi safeguard was taken)
};

// K(R) for R(a,b) = sum_k u_a v_b (Horn 1987).  Upper triangle, row-major 10 values.
LB_HD void qcp_build_K(const double* R, double* K) {
  const double Sxx = R[0], Sxy = R[1], Sxz = R[2];
  const double Syx = R[3], Syy = R[4], Syz = R[5];
  const double Szx = R[6], Szy = R[7], Szz = R[8];
  K[0] = Sxx + Syy + Szz; K[1] = Syz - Szy; K[2] = Szx - Sxz; K[3] = Sxy - Syx;
  K[4] = Sxx - Syy - Szz; K[5] = Sxy + Syx; K[6] = Szx + Sxz;
  K[7] = -Sxx + Syy - Szz; K[8] = Syz + Szy;
  K[9] = -Sxx - Syy + Szz;
}

// Largest eigenvalue (and optionally eigenvector) of the symmetric 4x4 K by
// cyclic Jacobi.  Safeguard path only; robust for degenerate spectra.
LB_HD double qcp_jacobi_max(const double* Kp, double* qvec) {
  double A[4][4], V[4][4];
  A[0][0] = Kp[0]; A[0][1] = A[1][0] = Kp[1]; A[0][2] = A[2][0] = Kp[2]; A[0][3] = A[3][0] = Kp[3];
  A[1][1] = Kp[4]; A[1][2] = A[2][1] = Kp[5]; A[1][3] = A[3][1] = Kp[6];
  A[2][2] = Kp[7]; A[2][3] = A[3][2] = Kp[8];
  A[3][3] = Kp[9];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < 4; ++i) {
      dg += A[i][i] * A[i][i];
      for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-34 * dg || off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  for (int i = 1; i < 4; ++i)
    if (A[i][i] > A[m][m]) m = i;
  if (qvec)
    for (int k = 0; k < 4; ++k) qvec[k] = V[k][m];
  return A[m][m];
}

// R: row-major 3x3, R[3a+b] = sum_k u_a v_b over CENTRED coordinates.
// ga, gb: sum of squares of each centred frame.  All in the same units.
LB_HD QcpResult qcp_lambda(const double* R, double ga, double gb) {
  QcpResult out;
  const double e0 = 0.5 * (ga + gb);
  if (!(e0 > 0.0)) { out.msd_x2_over_e0 = 0.0; out.x = 1.0; out.iters = 0; return out; }
  const double inv = 1.0 / e0;
  const double Sxx = R[0] * inv, Sxy = R[1] * inv, Sxz = R[2] * inv;
  const double Syx = R[3] * inv, Syy = R[4] * inv, Syz = R[5] * inv;
  const double Szx = R[6] * inv, Szy = R[7] * inv, Szz = R[8] * inv;

  const double Sxx2 = Sxx * Sxx, Syy2 = Syy * Syy, Szz2 = Szz * Szz;
  const double Sxy2 = Sxy * Sxy, Syz2 = Syz * Syz, Sxz2 = Sxz * Sxz;
  const double Syx2 = Syx * Syx, Szy2 = Szy * Szy, Szx2 = Szx * Szx;

  const double SyzSzymSyySzz2 = 2.0 * (Syz * Szy - Syy * Szz);
  const double Sxx2Syy2Szz2Syz2Szy2 = Syy2 + Szz2 - Sxx2 + Syz2 + Szy2;

  const double C2 = -2.0 * (Sxx2 + Syy2 + Szz2 + Sxy2 + Syx2 + Sxz2 + Szx2 + Syz2 + Szy2);
  const double C1 = 8.0 * (Sxx * Syz * Szy + Syy * Szx * Sxz + Szz * Sxy * Syx -
                           Sxx * Syy * Szz - Syz * Szx * Sxy - Szy * Syx * Sxz);

  const double SxzpSzx = Sxz + Szx, SyzpSzy = Syz + Szy, SxypSyx = Sxy + Syx;
  const double SyzmSzy = Syz - Szy, SxzmSzx = Sxz - Szx, SxymSyx = Sxy - Syx;
  const double SxxpSyy = Sxx + Syy, SxxmSyy = Sxx - Syy;
  const double Sxy2Sxz2Syx2Szx2 = Sxy2 + Sxz2 - Syx2 - Szx2;

  const double C0 =
      Sxy2Sxz2Syx2Szx2 * Sxy2Sxz2Syx2Szx2 +
      (Sxx2Syy2Szz2Syz2Szy2 + SyzSzymSyySzz2) * (Sxx2Syy2Szz2Syz2Szy2 - SyzSzymSyySzz2) +
      (-(SxzpSzx) * (SyzmSzy) + (SxymSyx) * (SxxmSyy - Szz)) *
          (-(SxzmSzx) * (SyzpSzy) + (SxymSyx) * (SxxmSyy + Szz)) +
      (-(SxzpSzx) * (SyzpSzy) - (SxypSyx) * (SxxpSyy - Szz)) *
          (-(SxzmSzx) * (SyzmSzy) - (SxypSyx) * (SxxpSyy + Szz)) +
      (+(SxypSyx) * (SyzpSzy) + (SxzpSzx) * (SxxmSyy + Szz)) *
          (-(SxymSyx) * (SyzmSzy) + (SxzpSzx) * (SxxpSyy + Szz)) +
      (+(SxypSyx) * (SyzmSzy) + (SxzmSzx) * (SxxmSyy - Szz)) *
          (-(SxymSyx) * (SyzpSzy) + (SxzmSzx) * (SxxpSyy - Szz));

  // Newton from x0 = 1 >= x_max: monotone from above for a quartic with real roots.
  double x = 1.0;
  int it = 0;
  bool ok = false;
  for (; it < 64; ++it) {
    const double x2 = x * x;
    const double b = (x2 + C2) * x;
    const double a = b + C1;
    const double P = a * x + C0;
    const double dP = 2.0 * x2 * x + b + a;
    if (!(fabs(dP) > 1e-300)) break;
    const double step = P / dP;
    x -= step;
    if (fabs(step) <= 1e-15 * fabs(x)) { ok = true; ++it; break; }
  }
  // x in scaled units must lie in [~0, 1]; anything else (or a stalled Newton on
  // a multiple root) goes to the Jacobi safeguard.
  if (!ok || !(x <= 1.0 + 1e-12) || !(x >= -1e-12)) {
    double Ks[10];
    const double Rs[9] = {Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz};
    qcp_build_K(Rs, Ks);
    x = qcp_jacobi_max(Ks, 0);
    it = -1 - it;
  }
  out.x = x;
  out.iters = it;
  out.msd_x2_over_e0 = fabs(2.0 * (1.0 - x));  // |E0 - 2 ss| / E0  (alignment.cpp:137 abs())
  return out;
}

// rmsd = sqrt(|Ga + Gb - 2 lambda| / n) scaled by `unit` (length of one operand unit).
LB_HD double qcp_rmsd(const double* R, double ga, double gb, double n, double unit) {
  const QcpResult r = qcp_lambda(R, ga, gb);
  const double e0 = 0.5 * (ga + gb);
  return unit * sqrt(r.msd_x2_over_e0 * e0 / n);
}

// Rotation (row-major 3x3, maps u onto v: v ~ Z u) from R, via the eigenvector of
// K for lambda_max.  Always uses the Jacobi solve: it is called once per frame on
// the streaming path, never per pair.
LB_HD void qcp_rotation(const double* R, double* Z, double* lambda_out) {
  double K[10], q[4];
  qcp_build_K(R, K);
  const double lam = qcp_jacobi_max(K, q);
  if (lambda_out) *lambda_out = lam;
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double s = n2 > 0 ? 1.0 / sqrt(n2) : 0.0;
  const double w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
  Z[0] = w * w + x * x - y * y - z * z; Z[1] = 2 * (x * y - w * z);           Z[2] = 2 * (x * z + w * y);
  Z[3] = 2 * (x * y + w * z);           Z[4] = w * w - x * x + y * y - z * z; Z[5] = 2 * (y * z - w * x);
  Z[6] = 2 * (x * z - w * y);           Z[7] = 2 * (y * z + w * x);           Z[8] = w * w - x * x - y * y + z * z;
}

}  // namespace loosb200

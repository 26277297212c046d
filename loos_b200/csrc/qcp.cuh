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
#include <string.h>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define LB_FDIV(a, b) __fdividef((a), (b))
#define LB_FRCP(a) __frcp_rn(a)
#define LB_FMUL(a, b) __fmul_rn((a), (b))  // a rounded product the compiler may not fuse into an FMA
#else
#define LB_FDIV(a, b) ((a) / (b))
#define LB_FRCP(a) (1.0f / (a))
#define LB_FMUL(a, b) ((a) * (b))
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
  double x = fmin(1.0, sqrt(fmax(0.0, -1.5 * C2)) * (1.0 + 1e-12));
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

// Fast path of the same solve for the per-pair epilogue of the tensor kernel.
// With x = 1 - y the quartic becomes Q(y) = P(1 - y) = y^4 + q3 y^3 + q2 y^2 + q1 y + q0
// whose smallest root y* = 1 - lambda_max/E0 = (Ga + Gb - 2 lambda)/(Ga + Gb) is the
// quantity wanted, so the cancellation is confined to q0, q1 (computed in fp64);
// Newton from y = 0 then only needs RELATIVE accuracy in y and runs in fp32, and
// one fp64 correction step (with an fp32 reciprocal) restores full precision.
// Returns y* (>= 0); sets *ok = false when the iteration has not settled
// (multiple root / far-off root), in which case the caller takes qcp_lambda().
// The same fast path split into phases so that a thread can run several solves
// in lock-step (the tensor-core epilogue interleaves four: the Newton recurrences
// are strictly serial, and only independent solves can fill their latency).
struct QcpPoly {
  double c0, c1, c2, c3;  // monic quartic t^4 + c3 t^3 + c2 t^2 + c1 t + c0 in t = x or y = 1 - x
  float t0;               // Newton start (an upper bound of the root in x, 0 in y)
  bool xform;
};

// Quartic coefficients from the invariants of S = R / E0.  The eigenvalues of K are
// {s1+s2+s3, s1-s2-s3, -s1+s2-s3, -s1-s2+s3} for the (signed) singular values s_i of S, hence
//   C2 = -2 F,  C1 = -8 det S,  C0 = F^2 - 4 G   with  F = ||S||_F^2,  G = ||cof S||_F^2
// (sum of the squared 2x2 minors): 45 fp64 operations instead of the ~100 of the
// textbook expansion of det K, and det S reuses the first row of cofactors.
LB_HD QcpPoly qcp_poly(const double* R, double e0) {
  // 1/e0: fp32 seed + one Newton step in fp64 (relative error 2^-46: a uniform scaling
  // of S by (1 + eps) moves y by eps, far below the 1e-13 the coefficients carry)
  double inv = (double)LB_FRCP((float)e0);
  inv = inv * (2.0 - e0 * inv);
  // invariants of the unnormalised R, scaled once at the end
  const double c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
  const double c10 = R[2] * R[7] - R[1] * R[8], c11 = R[0] * R[8] - R[2] * R[6], c12 = R[1] * R[6] - R[0] * R[7];
  const double c20 = R[1] * R[5] - R[2] * R[4], c21 = R[2] * R[3] - R[0] * R[5], c22 = R[0] * R[4] - R[1] * R[3];
  const double FR = (R[0] * R[0] + R[1] * R[1] + R[2] * R[2]) + (R[3] * R[3] + R[4] * R[4] + R[5] * R[5]) +
                    (R[6] * R[6] + R[7] * R[7] + R[8] * R[8]);
  const double GR = (c00 * c00 + c01 * c01 + c02 * c02) + (c10 * c10 + c11 * c11 + c12 * c12) +
                    (c20 * c20 + c21 * c21 + c22 * c22);
  const double DR = R[0] * c00 + R[1] * c01 + R[2] * c02;
  const double inv2 = inv * inv;
  const double F = FR * inv2, G = GR * (inv2 * inv2), D8 = 8.0 * (DR * (inv2 * inv));
  const double C2 = -2.0 * F, C1 = -D8, C0 = F * F - 4.0 * G;
  QcpPoly p;
  const float xb = sqrtf(fmaxf(0.0f, 3.0f * (float)F)) * 1.000001f;  // sqrt(-1.5 C2)
  p.xform = xb < 0.5f;
  // y-form: Q(y) = P(1 - y);  q0 = 1 + C2 + C1 + C0 = (1 - F)^2 - 8 det S - 4 G
  const double omf = 1.0 - F;
  p.c0 = p.xform ? C0 : (omf * omf - D8) - 4.0 * G;
  p.c1 = p.xform ? C1 : (4.0 * F - 4.0) + D8;
  p.c2 = p.xform ? C2 : 6.0 - 2.0 * F;
  p.c3 = p.xform ? 0.0 : -4.0;
  p.t0 = p.xform ? xb : 0.0f;
  return p;
}

// One fp32 Newton step on t; returns the step taken.
LB_HD float qcp_newton_step(float f0, float f1, float f2, float f3, float& t) {
  const float f = fmaf(t, fmaf(t, fmaf(t, t + f3, f2), f1), f0);
  const float df = fmaf(t, fmaf(t, fmaf(t, 4.0f, 3.0f * f3), 2.0f * f2), f1);
  const float st = LB_FDIV(f, df);
  t -= st;
  return st;
}

// One fp64 residual correction of the fp32 root (fp32 slope).  `last_step` is the
// final fp32 Newton step: if the recurrence had settled to ~1e-5 relative, the
// corrected root is good to ~1e-10 relative.  Returns y = 1 - lambda_max / E0 (>= 0).
LB_HD double qcp_polish(const QcpPoly& p, float t, float last_step, bool* ok) {
  const float f1 = (float)p.c1, f2 = (float)p.c2, f3 = (float)p.c3;
  const double td0 = (double)t;
  const double fd = fma(td0, fma(td0, fma(td0, td0 + p.c3, p.c2), p.c1), p.c0);
  const float df = fmaf(t, fmaf(t, fmaf(t, 4.0f, 3.0f * f3), 2.0f * f2), f1);
  const double corr = fd * (double)LB_FRCP(df);
  const double td = td0 - corr;
  const double yn = p.xform ? 1.0 - td : td;
  // slope sign: P' > 0 right of the largest root, Q' < 0 left of the smallest
  *ok = (fabsf(last_step) <= 1e-5f * fabsf(t) + 1e-12f) && (fabs(corr) <= 1e-4 * fabs(td) + 1e-12) &&
        (p.xform ? df > 0.0f : df < 0.0f) && (yn == yn) && (yn <= 1.0 + 1e-9);
  return yn < 0.0 ? -yn : yn;
}

// ---------------------------------------------------------------------------------------
// The same quartic built WITHOUT fp64, in float-float ("two-float") arithmetic: a value is
// an unevaluated sum h + l of two floats (~48 significant bits).  On sm_100 fp64 CUDA-core
// math shares a pipe with the tensor cores and runs at a tenth of its rate while MMAs are in
// flight (tools/ubench_contend.cu); fp32 is unaffected.  With this builder the tensor-core
// epilogue needs fp64 only to drain the exact integer accumulators.
//
// Inputs: R = Rh + Rl (nine entries) and e0 = e0h + e0l, all scaled by one power of two such
// that e0 lies in [0.5, 1) (|R_ij| <= e0 then keeps every intermediate in [-4, 4]).
// Error budget: every product keeps its exact fp32 error term and first-order cross terms,
// every sum is an error-free two-sum with the errors collected in an fp32 bucket, so each of
// F, det S, G carries ~2^-46 and q0 (the coefficient that cancels) ~1e-13 absolute -- the
// accuracy qcp_poly() documents for its fp64 coefficients (tests/test_qcp_host.py).
struct FFAcc { float h, e; };  // running sum: two-sum chain in h, all error terms in e

// s += (ah + al) (bh + bl)   (al bl is below 2^-48 of the product and dropped)
LB_HD void ff_acc_prod(FFAcc& s, float ah, float al, float bh, float bl) {
  const float p = LB_FMUL(ah, bh);
  float pe = fmaf(ah, bh, -p);
  pe = fmaf(ah, bl, pe);
  pe = fmaf(al, bh, pe);
  const float t = s.h + p, bb = t - s.h;
  const float err = (s.h - (t - bb)) + (p - bb);
  s.h = t;
  s.e += err + pe;
}
// s += (ah + al)^2
LB_HD void ff_acc_sq(FFAcc& s, float ah, float al) {
  const float p = LB_FMUL(ah, ah);
  float pe = fmaf(ah, ah, -p);
  pe = fmaf(ah + ah, al, pe);
  const float t = s.h + p, bb = t - s.h;
  const float err = (s.h - (t - bb)) + (p - bb);
  s.h = t;
  s.e += err + pe;
}
// s += (ah + al), error-free in the high word
LB_HD void ff_acc_add(FFAcc& s, float ah, float al) {
  const float t = s.h + ah, bb = t - s.h;
  const float err = (s.h - (t - bb)) + (ah - bb);
  s.h = t;
  s.e += err + al;
}
// normalise h + e into a float-float pair (full two-sum: after cancellation e may exceed h)
LB_HD void ff_norm(const FFAcc& s, float& h, float& l) {
  const float t = s.h + s.e, bb = t - s.h;
  l = (s.h - (t - bb)) + (s.e - bb);
  h = t;
}
// (ah + al)(bh + bl) -> (h, l)
LB_HD void ff_mul(float ah, float al, float bh, float bl, float& h, float& l) {
  const float p = LB_FMUL(ah, bh);
  float pe = fmaf(ah, bh, -p);
  pe = fmaf(ah, bl, pe);
  pe = fmaf(al, bh, pe);
  h = p + pe;
  l = pe - (h - p);
}
// a b - c d -> (h, l), all operands float-float
LB_HD void ff_minor(float ah, float al, float bh, float bl, float ch, float cl, float dh, float dl, float& h, float& l) {
  FFAcc s;
  s.h = LB_FMUL(ah, bh);
  s.e = fmaf(ah, bh, -s.h);
  s.e = fmaf(ah, bl, s.e);
  s.e = fmaf(al, bh, s.e);
  ff_acc_prod(s, -ch, -cl, dh, dl);
  ff_norm(s, h, l);
}

struct QcpPolyF {
  float c0h, c0l, c1h, c1l, c2h, c2l, c3;  // monic quartic in t (y-form, or x-form when xform)
  float t0;                                 // Newton start
  bool xform;
};

LB_HD QcpPolyF qcp_poly_ff(const float* Rh, const float* Rl, float e0h, float e0l) {
  // 1/e0 to second order around the fp32 reciprocal
  const float ih = LB_FRCP(e0h);
  float r = fmaf(-e0h, ih, 1.0f);
  r = fmaf(-e0l, ih, r);
  const float il = ih * fmaf(r, r, r);
  // cofactors of R
  float ch[9], cl[9];
#define LB_MINOR(k, a, b, c, d) ff_minor(Rh[a], Rl[a], Rh[b], Rl[b], Rh[c], Rl[c], Rh[d], Rl[d], ch[k], cl[k])
  LB_MINOR(0, 4, 8, 5, 7); LB_MINOR(1, 5, 6, 3, 8); LB_MINOR(2, 3, 7, 4, 6);
  LB_MINOR(3, 2, 7, 1, 8); LB_MINOR(4, 0, 8, 2, 6); LB_MINOR(5, 1, 6, 0, 7);
  LB_MINOR(6, 1, 5, 2, 4); LB_MINOR(7, 2, 3, 0, 5); LB_MINOR(8, 0, 4, 1, 3);
#undef LB_MINOR
  FFAcc sF = {0.0f, 0.0f}, sG = {0.0f, 0.0f}, sD = {0.0f, 0.0f};
#pragma unroll
  for (int k = 0; k < 9; ++k) { ff_acc_sq(sF, Rh[k], Rl[k]); ff_acc_sq(sG, ch[k], cl[k]); }
#pragma unroll
  for (int k = 0; k < 3; ++k) ff_acc_prod(sD, Rh[k], Rl[k], ch[k], cl[k]);
  float FRh, FRl, GRh, GRl, DRh, DRl;
  ff_norm(sF, FRh, FRl); ff_norm(sG, GRh, GRl); ff_norm(sD, DRh, DRl);
  // scale by powers of 1/e0
  float i2h, i2l, i3h, i3l, i4h, i4l, Fh, Fl, Gh, Gl, Dh, Dl;
  ff_mul(ih, il, ih, il, i2h, i2l);
  ff_mul(i2h, i2l, ih, il, i3h, i3l);
  ff_mul(i2h, i2l, i2h, i2l, i4h, i4l);
  ff_mul(FRh, FRl, i2h, i2l, Fh, Fl);
  ff_mul(GRh, GRl, i4h, i4l, Gh, Gl);
  ff_mul(DRh, DRl, i3h, i3l, Dh, Dl);
  Dh *= 8.0f; Dl *= 8.0f;  // D8 = 8 det S
  QcpPolyF p;
  const float xb = sqrtf(fmaxf(0.0f, 3.0f * Fh)) * 1.000001f;  // sqrt(-1.5 C2) >= lambda_max / e0
  p.xform = xb < 0.5f;
  if (!p.xform) {
    // y-form: q0 = (1 - F)^2 - D8 - 4 G,  q1 = -4 (1 - F) + D8,  q2 = 6 - 2 F,  q3 = -4
    FFAcc o = {1.0f, 0.0f};
    ff_acc_add(o, -Fh, -Fl);
    float oh, ol;
    ff_norm(o, oh, ol);
    FFAcc q = {0.0f, 0.0f};
    ff_acc_sq(q, oh, ol);
    ff_acc_add(q, -Dh, -Dl);
    ff_acc_add(q, -4.0f * Gh, -4.0f * Gl);
    ff_norm(q, p.c0h, p.c0l);
    FFAcc q1 = {-4.0f * oh, -4.0f * ol};
    ff_acc_add(q1, Dh, Dl);
    ff_norm(q1, p.c1h, p.c1l);
    FFAcc q2 = {6.0f, 0.0f};
    ff_acc_add(q2, -2.0f * Fh, -2.0f * Fl);
    ff_norm(q2, p.c2h, p.c2l);
    p.c3 = -4.0f;
    p.t0 = 0.0f;
  } else {
    // x-form: C0 = F^2 - 4 G,  C1 = -D8,  C2 = -2 F,  C3 = 0
    FFAcc q = {0.0f, 0.0f};
    ff_acc_sq(q, Fh, Fl);
    ff_acc_add(q, -4.0f * Gh, -4.0f * Gl);
    ff_norm(q, p.c0h, p.c0l);
    p.c1h = -Dh; p.c1l = -Dl;
    p.c2h = -2.0f * Fh; p.c2l = -2.0f * Fl;
    p.c3 = 0.0f;
    p.t0 = xb;
  }
  return p;
}

// fp64-free version of the residual correction, for code that runs while the tensor
// pipe is busy (on sm_100 fp64 CUDA-core math shares that pipe and crawls under MMA load).
// c0 and c1 -- the two coefficients whose sum cancels at the root -- are carried as
// float-float pairs prepared in fp64 beforehand; the rest of the quartic only needs
// fp32.  Returns the corrected root t' = t - Q(t)/Q'(t) in fp32 (relative accuracy ~1e-7,
// i.e. 5e-8 in the RMSD) and whether the iteration had settled.
// double -> (high float, low float) with the high part TRUNCATED to 24 bits by masking the
// mantissa: the high part is then exactly a float and its double image is known without
// converting it back (conversions run on a 16-lane unit).  |low| < ulp(high).
LB_HD void qcp_split_trunc(double v, float& h, float& l) {
#if defined(__CUDA_ARCH__)
  const double hd = __hiloint2double(__double2hiint(v), __double2loint(v) & (int)0xE0000000u);
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
  b &= ~((1ull << 29) - 1);
  double hd;
  memcpy(&hd, &b, 8);
#endif
  h = (float)hd;
  l = (float)(v - hd);
}

LB_HD QcpPolyF qcp_poly_split(const QcpPoly& p) {
  QcpPolyF f;
  qcp_split_trunc(p.c0, f.c0h, f.c0l);
  qcp_split_trunc(p.c1, f.c1h, f.c1l);
  qcp_split_trunc(p.c2, f.c2h, f.c2l);
  f.c3 = p.xform ? 0.0f : -4.0f;
  f.t0 = p.t0; f.xform = p.xform;
  return f;
}

LB_HD float qcp_polish_f32(const QcpPolyF& c, bool xform, float t, float last_step, bool* ok) {
  // Q(t) = c0 + t c1 + t^2 (c2 + t (c3 + t)); the first two terms in float-float
  const float p = t * c.c1h;
  const float pe = fmaf(t, c.c1h, -p) + t * c.c1l;          // exact product error + low word
  const float s = c.c0h + p;                                 // two-sum
  const float bb = s - c.c0h;
  const float se = (c.c0h - (s - bb)) + (p - bb);
  const float tail = t * t * fmaf(t, c.c3 + t, c.c2h);
  const float f = s + (((se + pe) + c.c0l) + tail);
  const float df = fmaf(t, fmaf(t, fmaf(t, 4.0f, 3.0f * c.c3), 2.0f * c.c2h), c.c1h);
  const float corr = LB_FDIV(f, df);
  const float tn = t - corr;
  const float yn = xform ? 1.0f - tn : tn;
  // |Q'/Q''| below the noise floor of the coefficients: a (numerically) double root, which only
  // qcp_root_fp64() resolves
  const float ddf = fmaf(t, fmaf(t, 12.0f, 6.0f * c.c3), 2.0f * c.c2h);
  *ok = (fabsf(last_step) <= 1e-5f * fabsf(t) + 1e-12f) && (fabsf(corr) <= 1e-4f * fabsf(tn) + 1e-12f) &&
        (xform ? df > 0.0f : df < 0.0f) && (fabsf(df) >= 5e-7f * fabsf(ddf)) &&
        (corr * corr * fabsf(ddf) <= 2e-7f * fabsf(tn) * fabsf(df) + 1e-30f) &&  // what one more step would still move
        (fabsf(tail) <= 3.0f * fabsf(tn * df)) &&  // the fp32 part of Q(t) is accurate enough for this slope
        (yn == yn) && (yn <= 1.0f + 1e-6f);
  return yn < 0.0f ? -yn : yn;
}

// Rare path (dissimilar structures, multiple or far-off root, Newton not settled): the same
// quartic in fp64 from the float-float coefficients.  Returns y = 1 - lambda_max / E0 (>= 0).
// Started outside the outermost root Newton is monotone and its steps -f/f' = 1 / sum_i 1/(r_i - t)
// can only shrink; a step that grows means the outermost roots are a (numerically) double or
// complex pair r +- sqrt(eps) -- collinear atoms, N = 2 -- where Q' has a simple root: finishing on
// Q' recovers the root to O(eps) instead of O(sqrt(eps)).
LB_HD double qcp_root_fp64(const QcpPolyF& p, float start) {
  const double c0 = (double)p.c0h + (double)p.c0l, c1 = (double)p.c1h + (double)p.c1l;
  const double c2 = (double)p.c2h + (double)p.c2l, c3 = (double)p.c3;
  double td = (double)start, prev = 1e300;
  bool multiple = false;
  for (int it = 0; it < 200; ++it) {
    const double fd = fma(td, fma(td, fma(td, td + c3, c2), c1), c0);
    const double dd = fma(td, fma(td, fma(td, 4.0, 3.0 * c3), 2.0 * c2), c1);
    if (dd == 0.0) { multiple = true; break; }
    const double stp = fd / dd;
    // rounding noise of the Horner evaluation, as a step
    const double at = fabs(td);
    const double noise = 4e-16 * (fabs(c0) + at * (fabs(c1) + at * (fabs(c2) + at * (fabs(c3) + at)))) / fabs(dd);
    if (fabs(stp) <= 4.0 * noise) break;  // at the floor
    if (fabs(stp) > prev) { multiple = true; break; }
    prev = fabs(stp);
    td -= stp;
    if (fabs(stp) <= 1e-15 * fabs(td) + 1e-300) break;
  }
  if (!multiple) {
    // a converged root closer than the coefficients' noise floor (sqrt(5e-14)) to a critical
    // point cannot be told from a double root: treat it as one
    const double d1 = fma(td, fma(td, fma(td, 4.0, 3.0 * c3), 2.0 * c2), c1);
    const double d2 = fma(td, fma(td, 12.0, 6.0 * c3), 2.0 * c2);
    multiple = fabs(d1) < 5e-7 * fabs(d2);
  }
  if (multiple) {
    for (int it = 0; it < 60; ++it) {
      const double d1 = fma(td, fma(td, fma(td, 4.0, 3.0 * c3), 2.0 * c2), c1);
      const double d2 = fma(td, fma(td, 12.0, 6.0 * c3), 2.0 * c2);
      if (d2 == 0.0) break;
      const double stp = d1 / d2;
      td -= stp;
      if (fabs(stp) <= 1e-15 * fabs(td) + 1e-300) break;
    }
  }
  const double yd = p.xform ? 1.0 - td : td;
  return yd < 0.0 ? -yd : yd;
}

#ifndef QCP_FAST_ITERS
#define QCP_FAST_ITERS 10
#endif
// Scalar composition of the phases above (used by host-side unit tests).
LB_HD double qcp_y_fast(const double* R, double e0, bool* ok) {
  const QcpPoly p = qcp_poly(R, e0);
  const float f0 = (float)p.c0, f1 = (float)p.c1, f2 = (float)p.c2, f3 = (float)p.c3;
  float t = p.t0, st = 0.0f;
  for (int it = 0; it < QCP_FAST_ITERS; ++it) st = qcp_newton_step(f0, f1, f2, f3, t);
  return qcp_polish(p, t, st, ok);
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
// Unit quaternion (w, x, y, z) -> rotation matrix, row-major.
LB_HD void qcp_quat_to_rot(const double* q, double* Z) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  Z[0] = w * w + x * x - y * y - z * z; Z[1] = 2 * (x * y - w * z);           Z[2] = 2 * (x * z + w * y);
  Z[3] = 2 * (x * y + w * z);           Z[4] = w * w - x * x + y * y - z * z; Z[5] = 2 * (y * z - w * x);
  Z[6] = 2 * (x * z - w * y);           Z[7] = 2 * (y * z + w * x);           Z[8] = w * w - x * x - y * y + z * z;
}

LB_HD void qcp_rotation(const double* R, double* Z, double* lambda_out, double* quat_out = 0) {
  double K[10], q[4];
  qcp_build_K(R, K);
  const double lam = qcp_jacobi_max(K, q);
  if (lambda_out) *lambda_out = lam;
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double s = n2 > 0 ? 1.0 / sqrt(n2) : 0.0;
  const double u[4] = {q[0] * s, q[1] * s, q[2] * s, q[3] * s};
  if (quat_out) { quat_out[0] = u[0]; quat_out[1] = u[1]; quat_out[2] = u[2]; quat_out[3] = u[3]; }
  if (Z) qcp_quat_to_rot(u, Z);
}

// Rotation for the streaming path without the Jacobi sweeps: lambda_max from the
// quartic (fp64 coefficients, fp32 Newton, fp64 correction), then the eigenvector as the
// best-conditioned column of adj(K - lambda I) = prod(lambda - lambda_i) q q^T.
// Returns false (caller falls back to qcp_rotation) when the top eigenvalue is not
// well separated -- collinear / planar / N < 3 inputs -- or the solve did not settle.
// *resid (optional) receives ga + gb - 2 lambda_max = 2 e0 y, free of cancellation.
// quat_out (optional) receives the unit quaternion; Z may then be null.
LB_HD bool qcp_rotation_fast(const double* R, double ga, double gb, double* Z, double* resid = 0, double* quat_out = 0) {
  const double e0 = 0.5 * (ga + gb);
  if (!(e0 > 0.0)) return false;
  const QcpPoly p = qcp_poly(R, e0);
  const float f0 = (float)p.c0, f1 = (float)p.c1, f2 = (float)p.c2, f3 = (float)p.c3;
  float t = p.t0, st = 0.0f;
  for (int it = 0; it < 12; ++it) st = qcp_newton_step(f0, f1, f2, f3, t);
  bool ok;
  double y = qcp_polish(p, t, st, &ok);
  if (!ok) return false;
  // a second fp64 Newton step on the same quartic: the eigenvector needs lambda to ~1e-15
  {
    double td = p.xform ? 1.0 - y : y;
    const double fd = fma(td, fma(td, fma(td, td + p.c3, p.c2), p.c1), p.c0);
    const double dd = fma(td, fma(td, fma(td, 4.0, 3.0 * p.c3), 2.0 * p.c2), p.c1);
    if (dd != 0.0) td -= fd / dd;
    y = p.xform ? 1.0 - td : td;
  }
  if (resid) *resid = 2.0 * e0 * y;
  const double lam = (1.0 - y);  // in units of e0
  const double inv = 1.0 / e0;
  double S[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) S[k] = R[k] * inv;
  double Kd[10];
  qcp_build_K(S, Kd);
  // A = K - lam I (symmetric), row-major upper triangle -> full
  const double a00 = Kd[0] - lam, a01 = Kd[1], a02 = Kd[2], a03 = Kd[3];
  const double a11 = Kd[4] - lam, a12 = Kd[5], a13 = Kd[6];
  const double a22 = Kd[7] - lam, a23 = Kd[8];
  const double a33 = Kd[9] - lam;
  // 2x2 minors of rows (2,3) and of rows (0,1)
  const double m23_01 = a02 * a13 - a03 * a12, m23_02 = a02 * a23 - a03 * a22, m23_03 = a02 * a33 - a03 * a23;
  const double m23_12 = a12 * a23 - a13 * a22, m23_13 = a12 * a33 - a13 * a23, m23_23 = a22 * a33 - a23 * a23;
  const double m01_01 = a00 * a11 - a01 * a01, m01_02 = a00 * a12 - a01 * a02, m01_03 = a00 * a13 - a01 * a03;
  const double m01_12 = a01 * a12 - a11 * a02, m01_13 = a01 * a13 - a11 * a03, m01_23 = a02 * a13 - a12 * a03;
  // adjugate (symmetric) by Laplace expansion over row pairs; here m23_xy is the minor of
  // rows 2,3 and columns x,y, m01_xy that of rows 0,1
  const double d00 = a11 * m23_23 - a12 * m23_13 + a13 * m23_12;
  const double d01 = -(a01 * m23_23 - a12 * m23_03 + a13 * m23_02);
  const double d02 = a01 * m23_13 - a11 * m23_03 + a13 * m23_01;
  const double d03 = -(a01 * m23_12 - a11 * m23_02 + a12 * m23_01);
  const double d11 = a00 * m23_23 - a02 * m23_03 + a03 * m23_02;
  const double d12 = -(a00 * m23_13 - a01 * m23_03 + a03 * m23_01);
  const double d13 = a00 * m23_12 - a01 * m23_02 + a02 * m23_01;
  const double d22 = a33 * m01_01 - a13 * m01_03 + a03 * m01_13;
  const double d23 = -(a23 * m01_01 - a13 * m01_02 + a03 * m01_12);
  const double d33 = a22 * m01_01 - a12 * m01_02 + a02 * m01_12;
  (void)m01_23;
  // best-conditioned column: the largest diagonal entry ~ q_k^2
  double q0 = d00, q1 = d01, q2 = d02, q3 = d03, dm = fabs(d00);
  if (fabs(d11) > dm) { dm = fabs(d11); q0 = d01; q1 = d11; q2 = d12; q3 = d13; }
  if (fabs(d22) > dm) { dm = fabs(d22); q0 = d02; q1 = d12; q2 = d22; q3 = d23; }
  if (fabs(d33) > dm) { dm = fabs(d33); q0 = d03; q1 = d13; q2 = d23; q3 = d33; }
  // |adj| = |P'(lambda)| = prod of gaps: tiny => (near-)multiple eigenvalue
  if (!(dm > 1e-7)) return false;
  const double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
  const double sN = 1.0 / sqrt(n2);
  const double u[4] = {q0 * sN, q1 * sN, q2 * sN, q3 * sN};
  if (quat_out) { quat_out[0] = u[0]; quat_out[1] = u[1]; quat_out[2] = u[2]; quat_out[3] = u[3]; }
  if (Z) qcp_quat_to_rot(u, Z);
  return true;
}

}  // namespace loosb200

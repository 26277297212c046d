// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called
// from the product path (loos_b200/, include/); only tests/, smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load the library built
// from this file.
//
// This translation unit wraps the REFERENCE'S OWN arithmetic, compiled from the
// sources where they lie under /root/reference (never copied into the repo):
//   * src/alignment.cpp:12-281  (namespace alignment: kabschCore, centerAtOrigin,
//     centeredRMSD, alignedRMSD, kabschCentered, kabsch, applyTransform,
//     averageCoords, rmsd) -- extracted by oracle/Makefile into the git-ignored
//     oracle/_ref/alignment_core.inc and #included below;
//   * src/Coord.hpp, src/Matrix44.hpp, src/XForm.hpp, src/XForm.cpp -- compiled
//     directly via -I/root/reference/src (behind the shim loos_defs.hpp).
// BLAS/LAPACK (dgemm_, dgesvj_) come from OpenBLAS 0.3.15 found in the image
// (the reference pins no version: CMakeLists.txt:44-46).
//
// Everything below the #include is OUR restatement of the tool-level drivers
// that surround that arithmetic (thread pool, loops), with file:line citations.

#include <loos_defs.hpp>          // oracle/shim
#include <boost/tuple/tuple.hpp>  // oracle/shim
#include <Coord.hpp>              // /root/reference/src
#include <Matrix44.hpp>           // /root/reference/src
#include <XForm.hpp>              // /root/reference/src

#include <atomic>
#include <cstring>
#include <thread>

namespace loos {
namespace alignment {
typedef std::vector<double> vecDouble;
typedef std::vector<vecDouble> vecMatrix;
typedef boost::tuple<vecDouble, vecDouble, vecDouble> SVDTupleVec;  // alignment.hpp:42-44
}  // namespace alignment
}  // namespace loos

#include "_ref/alignment_core.inc"  // verbatim reference lines, built by Makefile
// src/alignment.cpp:288-318: the reference's own iterativeAlignment over a vecMatrix (frames as plain vectors).
// rmsd2ref calls the trajectory-based overload (:355-404, restated below as ref_rmsd2ref_iterative because it needs
// AtomicGroup / Trajectory); this overload is the same algorithm on the same arithmetic and is the cross-check for
// that restatement (tests/test_oracle.py).
namespace loos {
#include "_ref/alignment_iterative.inc"
}

using loos::GCoord;
using loos::GMatrix;
using loos::XForm;
using loos::alignment::vecDouble;

namespace {

// Tools/rmsds.cpp:219-250 (Master::workAvailable): rows handed out one at a time
// from a shared counter.  The mutex is replaced by an atomic fetch-add; the
// hand-out order and granularity are the same.
struct RowMaster {
  std::atomic<unsigned> top{0};
  unsigned maxrow;
  explicit RowMaster(unsigned n) : maxrow(n) {}
  bool workAvailable(unsigned* ip) {
    unsigned i = top.fetch_add(1);
    if (i >= maxrow) return false;
    *ip = i;
    return true;
  }
};

std::vector<vecDouble> to_rows(const double* T, int F, int n3) {
  std::vector<vecDouble> rows(F);
  for (int i = 0; i < F; ++i) rows[i].assign(T + (size_t)i * n3, T + (size_t)(i + 1) * n3);
  return rows;
}

unsigned pick_threads(int nthreads) {
  // Tools/rmsds.cpp:509: 0 => hardware_concurrency()
  unsigned n = nthreads > 0 ? (unsigned)nthreads : std::thread::hardware_concurrency();
  return n ? n : 1;
}

}  // namespace

extern "C" {

int ref_available(void) { return 1; }
void ref_set_blas_threads(int n) { openblas_set_num_threads(n); }

// src/alignment.cpp:90-109
void ref_center_at_origin(double* v, int n3, double* c3) {
  vecDouble x(v, v + n3);
  GCoord c = loos::alignment::centerAtOrigin(x);
  std::memcpy(v, x.data(), sizeof(double) * n3);
  if (c3) { c3[0] = c.x(); c3[1] = c.y(); c3[2] = c.z(); }
}

// src/alignment.cpp:115-140
double ref_centered_rmsd(const double* u, const double* v, int n3) {
  vecDouble U(u, u + n3), V(v, v + n3);
  return loos::alignment::centeredRMSD(U, V);
}

// src/alignment.cpp:146-176
double ref_aligned_rmsd(const double* u, const double* v, int n3) {
  vecDouble U(u, u + n3), V(v, v + n3);
  return loos::alignment::alignedRMSD(U, V);
}

// src/alignment.cpp:21-87: S (signed singular values) for a centred pair
void ref_kabsch_core_S(const double* u, const double* v, int n3, double* S3) {
  vecDouble U(u, u + n3), V(v, v + n3);
  loos::alignment::SVDTupleVec svd = loos::alignment::kabschCore(U, V);
  vecDouble S(boost::get<1>(svd));
  S3[0] = S[0]; S3[1] = S[1]; S3[2] = S[2];
}

// src/alignment.cpp:217-233 == AtomicGroup::superposition (src/AG_linalg.cpp:188-197)
// M16: row-major 4x4 (Matrix44.hpp:67-78), maps u onto v.
void ref_kabsch(const double* u, const double* v, int n3, double* M16) {
  vecDouble U(u, u + n3), V(v, v + n3);
  GMatrix M = loos::alignment::kabsch(U, V);
  for (int i = 0; i < 16; ++i) M16[i] = M[i];
}

// Tools/rmsds.cpp:460-463 (centerTrajectory) + :345-373 (SingleWorker) + :387-419
// (Threader).  T: F rows of 3N doubles, xyz interleaved (src/ensembles.cpp:275-279).
// M: float, column-major F x F (MatrixOrder.hpp:101-103), zero-initialised here
// as Matrix's constructor does (MatrixStorage.hpp:126-130).
void ref_rmsds_self(const double* T, int F, int n3, float* M, int nthreads) {
  std::vector<vecDouble> rows = to_rows(T, F, n3);
  for (int i = 0; i < F; ++i) loos::alignment::centerAtOrigin(rows[i]);
  std::memset(M, 0, sizeof(float) * (size_t)F * F);
  RowMaster master(F);
  auto worker = [&]() {
    unsigned i;
    while (master.workAvailable(&i))
      for (unsigned j = 0; j < i; ++j) {  // rmsds.cpp:359-365
        double d = loos::alignment::centeredRMSD(rows[i], rows[j]);
        M[(size_t)i * F + j] = M[(size_t)j * F + i] = (float)d;
      }
  };
  unsigned np = pick_threads(nthreads);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < np; ++t) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

// Tools/rmsds.cpp:302-338 (DualWorker) + :537-563.  M: F1 x F2 column-major,
// M(i,j) at j*F1 + i.
void ref_rmsds_cross(const double* T1, int F1, const double* T2, int F2, int n3, float* M,
                     int nthreads) {
  std::vector<vecDouble> A = to_rows(T1, F1, n3), B = to_rows(T2, F2, n3);
  for (auto& r : A) loos::alignment::centerAtOrigin(r);
  for (auto& r : B) loos::alignment::centerAtOrigin(r);
  std::memset(M, 0, sizeof(float) * (size_t)F1 * F2);
  RowMaster master(F1);
  auto worker = [&]() {
    unsigned i;
    while (master.workAvailable(&i))
      for (int j = 0; j < F2; ++j)  // rmsds.cpp:316-322
        M[(size_t)j * F1 + i] = (float)loos::alignment::centeredRMSD(A[i], B[j]);
  };
  unsigned np = pick_threads(nthreads);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < np; ++t) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

// Tools/rmsd2ref.cpp:205-216 (target given, --align non-empty) + :238-245.
// fa: F x 3Na align-selection coords per frame; fr: F x 3Nr rmsd-selection
// coords per frame; ta/tr: the target's selections.  xf16 (optional): F x 16.
void ref_rmsd2ref_target(const double* fa, const double* fr, const double* ta, const double* tr,
                         int F, int na3, int nr3, double* out, double* xf16) {
  vecDouble TA(ta, ta + na3);
  for (int f = 0; f < F; ++f) {
    vecDouble U(fa + (size_t)f * na3, fa + (size_t)(f + 1) * na3);
    GMatrix W = loos::alignment::kabsch(U, TA);  // AG_linalg.cpp:188-197
    if (xf16) for (int k = 0; k < 16; ++k) xf16[(size_t)f * 16 + k] = W[k];
    // AG_numerical.cpp:363-370 (applyTransform) then :301-318 (rmsd)
    double d = 0.0;
    int n = nr3 / 3;
    for (int i = 0; i < n; ++i) {
      const double* p = fr + (size_t)f * nr3 + 3 * i;
      GCoord x(p[0], p[1], p[2]);
      x = W * x;
      GCoord y(tr[3 * i], tr[3 * i + 1], tr[3 * i + 2]);
      d += y.distance2(x);
    }
    out[f] = sqrt(d / n);
  }
}

// src/alignment.cpp:355-404 (iterativeAlignment over a trajectory) followed by
// src/ensembles.cpp:91-121 (averageStructure with transforms) and
// Tools/rmsd2ref.cpp:238-245: rmsd2ref's default (no --target) mode.
// Returns the iteration count; *final_rms gets the last avg-vs-target rms.
int ref_rmsd2ref_iterative(const double* fa, const double* fr, int F, int na3, int nr3,
                           double tol, int maxiter, double* out, double* xf16,
                           double* avg_out, double* final_rms) {
  int na = na3 / 3, nr = nr3 / 3;
  auto frame = [&](int f) {
    std::vector<GCoord> g(na);
    for (int i = 0; i < na; ++i)
      g[i] = GCoord(fa[(size_t)f * na3 + 3 * i], fa[(size_t)f * na3 + 3 * i + 1],
                    fa[(size_t)f * na3 + 3 * i + 2]);
    return g;
  };
  auto as_vec = [&](const std::vector<GCoord>& g) {
    vecDouble v(3 * g.size());
    for (size_t i = 0; i < g.size(); ++i) { v[3*i] = g[i].x(); v[3*i+1] = g[i].y(); v[3*i+2] = g[i].z(); }
    return v;
  };
  // target = first frame, centred via AtomicGroup::centerAtOrigin
  // (AG_numerical.cpp: centroid = sum/size, then translate(-c)).
  std::vector<GCoord> target = frame(0);
  {
    GCoord c(0, 0, 0);
    for (auto& a : target) c += a;
    c /= target.size();
    for (auto& a : target) a -= c;
  }
  std::vector<XForm> xforms(F);
  std::vector<GCoord> avg(na);
  int iter = 0;
  double rms;
  do {
    for (auto& a : avg) a = GCoord(0, 0, 0);
    vecDouble tv = as_vec(target);
    for (int f = 0; f < F; ++f) {
      std::vector<GCoord> fr_ = frame(f);
      GMatrix M = loos::alignment::kabsch(as_vec(fr_), tv);  // alignOnto: AG_linalg.cpp:200-208
      for (auto& a : fr_) a = M * a;
      xforms[f].load(M);
      for (int j = 0; j < na; ++j) avg[j] += fr_[j];
    }
    for (auto& a : avg) a /= F;
    double d = 0.0;  // avg.rmsd(target): AG_numerical.cpp:301-318
    for (int j = 0; j < na; ++j) d += avg[j].distance2(target[j]);
    rms = sqrt(d / na);
    target = avg;
    ++iter;
  } while (rms > tol && iter <= maxiter);
  if (final_rms) *final_rms = rms;

  // averageStructure(subset, transforms, traj, indices): ensembles.cpp:91-121
  std::vector<GCoord> tavg(nr, GCoord(0, 0, 0));
  for (int f = 0; f < F; ++f) {
    GMatrix W = xforms[f].current();
    for (int i = 0; i < nr; ++i) {
      const double* p = fr + (size_t)f * nr3 + 3 * i;
      tavg[i] += W * GCoord(p[0], p[1], p[2]);
    }
  }
  for (auto& a : tavg) a /= F;
  for (int f = 0; f < F; ++f) {
    GMatrix W = xforms[f].current();
    if (xf16) for (int k = 0; k < 16; ++k) xf16[(size_t)f * 16 + k] = W[k];
    double d = 0.0;
    for (int i = 0; i < nr; ++i) {
      const double* p = fr + (size_t)f * nr3 + 3 * i;
      GCoord x = W * GCoord(p[0], p[1], p[2]);
      d += tavg[i].distance2(x);
    }
    out[f] = sqrt(d / nr);
  }
  if (avg_out)
    for (int i = 0; i < nr; ++i) { avg_out[3*i] = tavg[i].x(); avg_out[3*i+1] = tavg[i].y(); avg_out[3*i+2] = tavg[i].z(); }
  return iter;
}


// The reference's own iterativeAlignment(vecMatrix&, threshold, maxiter) (src/alignment.cpp:288-318): frames F x n3
// doubles (transformed in place there: a copy here), xf16 = F row-major 4x4 accumulated transforms.  Returns iterations.
int ref_iterative_vecmatrix(const double* frames, int F, int n3, double tol, int maxiter, double* xf16, double* final_rms) {
  loos::alignment::vecMatrix ens = to_rows(frames, F, n3);
  boost::tuple<std::vector<XForm>, loos::greal, int> res = loos::iterativeAlignment(ens, tol, maxiter);
  const std::vector<XForm>& xf = boost::get<0>(res);
  for (int f = 0; f < F; ++f) {
    GMatrix W = xf[f].current();
    for (int k = 0; k < 16; ++k) xf16[(size_t)f * 16 + k] = W[k];
  }
  if (final_rms) *final_rms = boost::get<1>(res);
  return boost::get<2>(res);
}

}  // extern "C"

/* loosb200.h -- C ABI of the B200-native optimal-superposition RMSD path.
 *
 * This is the drop-in boundary for the hot path behind LOOS's `rmsds`,
 * `multi-rmsds` and `rmsd2ref` tools.  The reference has no plugin/FFI layer;
 * the seam is the L3 function boundary those tools call:
 *
 *   readCoords + centerTrajectory      src/ensembles.cpp:253-296, Tools/rmsds.cpp:460-463
 *   alignment::centeredRMSD  (F^2/2 x) src/alignment.cpp:115-140, Tools/rmsds.cpp:359-365,316-322
 *   AtomicGroup::superposition / rmsd  src/AG_linalg.cpp:188-197, src/AG_numerical.cpp:301-318
 *   iterativeAlignment/averageStructure src/alignment.cpp:355-404, src/ensembles.cpp:91-121
 *
 * A per-pair call cannot feed a GPU, so the entry points are batch-level: stage a
 * set of frames once, then ask for the whole matrix / the whole per-frame series.
 * Plain pointers and sizes only; no C++ or torch types; nothing throws across the
 * boundary.  Every function returns LOOSB200_OK (0) or a negative error code
 * (loosb200_strerror() gives the text; the reference's LOOSError /
 * NumericalError cases map onto these codes).
 *
 * Ownership: all host buffers are caller-owned and may be freed as soon as the
 * call returns.  Handles own their device memory until loosb200_free().
 * Threading: calls on one handle are not re-entrant; different handles are
 * independent.  CUDA streams are internal.  There is no CPU fallback: without a
 * usable sm_100 device every compute entry point fails with LOOSB200_ERR_NO_DEVICE.
 */
#ifndef LOOSB200_H
#define LOOSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define LOOSB200_OK                0
#define LOOSB200_ERR_NO_DEVICE    -1  /* no CUDA device / not sm_100 */
#define LOOSB200_ERR_CUDA         -2  /* a CUDA runtime/driver call failed (see loosb200_last_error) */
#define LOOSB200_ERR_INVALID      -3  /* bad argument (null pointer, zero size, bad flag) */
#define LOOSB200_ERR_SIZE_MISMATCH -4 /* atom counts differ: LOOSError in AG_linalg.cpp:189-192, AG_numerical.cpp:303-304 */
#define LOOSB200_ERR_RANGE        -5  /* coordinates exceed the fixed-point range / index out of bounds */
#define LOOSB200_ERR_NOMEM        -6  /* host or device allocation failed */
#define LOOSB200_ERR_NUMERICAL    -7  /* solver failure: NumericalError in alignment.cpp:66-67 */
#define LOOSB200_ERR_UNSUPPORTED  -8  /* valid request this build cannot serve */

/* ---- coordinate layouts accepted by loosb200_stage_frames ---- */
/* F rows of natoms*3 floats, xyz interleaved per atom: the order readCoords()
 * produces (src/ensembles.cpp:275-279). */
#define LOOSB200_LAYOUT_XYZ    0
/* per frame X[natoms] Y[natoms] Z[natoms]: the order of the DCD coordinate
 * records (src/dcd.cpp:362-368), so decoded frames need no transposition. */
#define LOOSB200_LAYOUT_PLANES 1

/* ---- which all-to-all kernel to use ---- */
#define LOOSB200_ENGINE_AUTO   0  /* tensor-core path when the shape allows (selections of up to 100,000 atoms,
                                     finite coordinates), else fp64 */
#define LOOSB200_ENGINE_TENSOR 1  /* tcgen05 int8 sliced (exact int32) contraction + fused QCP */
#define LOOSB200_ENGINE_FP64   2  /* CUDA-core fp64 contraction + the same fused QCP */

typedef struct loosb200_frames loosb200_frames; /* opaque: one staged frame set on one GPU */

/* Stats over the result, as showStatsHalf / showStatsWhole print them
 * (Tools/rmsds.cpp:425-456): max and sum over the strict lower triangle (self) or
 * the whole matrix (cross), accumulated in double from the float32 entries.
 * `count` is 64-bit (the reference's `uint total` overflows for F > 65536). */
typedef struct loosb200_stats {
  double max;
  double sum;
  uint64_t count;
} loosb200_stats;

/* ---- library / device ---- */
const char* loosb200_version(void);
const char* loosb200_strerror(int code);
/* Text of the last CUDA error seen by the calling thread ("" if none). */
const char* loosb200_last_error(void);
/* Number of usable (compute capability 10.x) devices; 0 if none. Never fails. */
int loosb200_device_count(void);

/* ---- frame staging  (readCoords + centerAtOrigin, device side) ----
 * Copies F frames to `device` (pinned double-buffered async H2D), gathers the
 * `nsel` selected atoms (sel[i] = Atom::index() into the frame, model order --
 * Trajectory::updateGroupCoords, src/dcd.cpp:413-425; sel == NULL means all
 * `natoms`), and keeps them as fp32 SoA planes with the per-frame fp64 centroid
 * (alignment.cpp:90-109) and centred sum of squares.  The fixed-point operand
 * planes of the tensor-core path are derived lazily on first use.
 * host_xyz: F * natoms * 3 floats in `layout`. */
int loosb200_stage_frames(loosb200_frames** out, const float* host_xyz, size_t F, size_t natoms,
                          const uint32_t* sel, size_t nsel, int layout, int device);
/* Incremental form, for overlapping trajectory decode with the upload: announce F
 * frames, append them in any number of pieces (each call copies its frames to a
 * pinned bounce buffer and returns; H2D and the gather kernel run behind the
 * caller's next decode), then finish.  The handle may only be used for compute
 * after loosb200_stage_end() returned LOOSB200_OK. */
int loosb200_stage_begin(loosb200_frames** out, size_t F, size_t natoms, const uint32_t* sel,
                         size_t nsel, int layout, int device);
int loosb200_stage_append(loosb200_frames* h, const float* host_xyz, size_t nframes);
int loosb200_stage_end(loosb200_frames* h);
/* Same as loosb200_stage_frames, but `dev_xyz` already lives in the HBM of `device` (used to time the
 * device-resident path; also lets a decoder running on the GPU hand frames over). */
int loosb200_stage_frames_device(loosb200_frames** out, const float* dev_xyz, size_t F,
                                 size_t natoms, const uint32_t* sel, size_t nsel, int layout,
                                 int device);
void loosb200_free(loosb200_frames* h);
/* Freed handles and finished calls leave their device buffers (the 40 GB bounce matrix of the
 * host-output path included) in a size-keyed cache so that the next call of the same shape skips
 * cudaMalloc/cudaFree (hundreds of ms at 40 GB).  This returns the cache to the driver: call it before
 * another allocator of the same process needs the memory.  LOOSB200_CACHE_GB caps the cache
 * (default 96, 0 disables it). */
void loosb200_trim(void);
int loosb200_frames_info(const loosb200_frames* h, size_t* F, size_t* N, int* device);
/* Centroids as centerAtOrigin returns them: F x 3 doubles (host). */
int loosb200_get_centroids(const loosb200_frames* h, double* out_F3);
/* log2 of the fixed-point unit in Angstrom chosen for the tensor-core operands
 * (e.g. -16 => 2^-16 A); forces the lazy quantisation. */
int loosb200_get_unit_log2(loosb200_frames* h, int* e);

/* ---- all-to-all RMSD  (Tools/rmsds.cpp:515-535, Tools/multi-rmsds.cpp:395-412) ----
 * out: float32, COLUMN-major with leading dimension ld >= F
 * (Math::Matrix<float,ColMajor>, src/MatrixOrder.hpp:101-103): out[j*ld + i] = M(i,j).
 * M(i,j) == M(j,i), diagonal exactly 0.0f, as SingleWorker::calc leaves it.
 * `out` may be NULL (the `--noout 1` mode): only `stats` is produced and the
 * matrix never leaves the SMs.  `stats` may be NULL. */
int loosb200_rmsds_self(loosb200_frames* h, float* out, size_t ld, loosb200_stats* stats,
                        int engine);
/* Same with `out_dev` in the HBM of the handle's device. */
int loosb200_rmsds_self_device(loosb200_frames* h, float* out_dev, size_t ld,
                               loosb200_stats* stats, int engine);

/* Two-trajectory mode (DualWorker, Tools/rmsds.cpp:302-338,537-563):
 * out[j*ld + i] = rmsd(a_i, b_j), i < F_a, j < F_b, ld >= F_a.  Both handles must
 * be on the same device and hold the same number of atoms. */
int loosb200_rmsds_cross(loosb200_frames* a, loosb200_frames* b, float* out, size_t ld,
                         loosb200_stats* stats, int engine);
int loosb200_rmsds_cross_device(loosb200_frames* a, loosb200_frames* b, float* out_dev, size_t ld,
                                loosb200_stats* stats, int engine);

/* ---- result as text, formatted on the GPU  (src/MatrixImpl.hpp:210-222, Tools/rmsds.cpp:566-567) ----
 * `cout << setprecision(p) << M`: "# m n (0)\n", then one line per row, every value as "%.{p}g"
 * followed by a space.  The float32 matrix stays in HBM; rows are formatted there (exact integer
 * arithmetic, the same bytes as printf) and only the text crosses PCIe, in row order, through
 * `sink(user, data, nbytes)` (return non-zero to abort).  At 100 000 frames that is 45 GB of text,
 * which costs a minute of host time even on all cores.  precision 0..9.  If the F x F float
 * matrix does not fit in device memory, bands of rows are computed as rectangles and streamed one
 * after the other (same bytes, same stats). */
typedef int (*loosb200_sink)(void* user, const char* data, size_t nbytes);
int loosb200_rmsds_self_text(loosb200_frames* h, unsigned precision, loosb200_sink sink, void* user,
                             loosb200_stats* stats, int engine);
int loosb200_rmsds_cross_text(loosb200_frames* a, loosb200_frames* b, unsigned precision,
                              loosb200_sink sink, void* user, loosb200_stats* stats, int engine);

/* ---- multi-GPU work splitting ----
 * The upper-triangular tile space is cut statically into `nparts` interleaved
 * sets of 40-frame tile rows; part `part` computes its rows against all columns
 * j <= i.  No inter-GPU exchange is needed: every part holds a replica of the
 * staged frames.  Results come back as a packed list of the part's rows
 * (loosb200_part_rows() frames, ascending): out[r*ld + j] = M(row_r, j) for
 * j < row_r, ld >= F  (i.e. the part's rows of the lower triangle, row-major; the
 * mirror image is the caller's to fill -- see loos_b200/multigpu.py).
 * stats covers exactly the pairs of this part; sums over parts give the total. */
int loosb200_part_rows(size_t F, int part, int nparts, uint32_t* rows_or_null, size_t* nrows);
int loosb200_rmsds_self_part(loosb200_frames* h, int part, int nparts, float* out, size_t ld,
                             loosb200_stats* stats, int engine);
int loosb200_rmsds_self_part_device(loosb200_frames* h, int part, int nparts, float* out_dev,
                                    size_t ld, loosb200_stats* stats, int engine);

/* ---- several GPUs of one box, ONE process, ONE result matrix ----
 * The reference fills one RealMatrix from `np` worker threads inside one process
 * (Master / Threader, Tools/rmsds.cpp:219-285,387-419,515-535; Tools/multi-rmsds.cpp:395-417).
 * Here the workers are GPUs: stage the frames once, replicate the staged set to the other
 * devices (device-to-device over NVLink, no second pass over PCIe), and hand the replicas to the
 * *_multi calls.  The lower triangle is cut statically into `n` contiguous equal-work ranges of
 * 40-frame row tiles; nothing is exchanged while computing.  replicas[0] is "rank 0".
 *   _multi          out = HOST matrix as in loosb200_rmsds_self: every device copies its own row
 *                   strips and their mirror images straight into it over its own PCIe link, band by
 *                   band behind its kernel (pinned memory from loosb200_host_alloc() recommended:
 *                   it is pinned for all devices).
 *   _multi_device   out_dev = F x F matrix in the HBM of replicas[0]'s device: gather of the row
 *                   blocks to rank 0 over NVLink.  Default: the other devices' kernels store their
 *                   tiles there directly (peer stores, the transfer rides under the contraction);
 *                   LOOSB200_GATHER=copy: local strips pushed by the copy engines band by band.
 *   _multi_text     the text of loosb200_rmsds_self_text: row bands dealt round-robin to the
 *                   devices, each formats its bands and ships the text over its own link; the
 *                   sink receives them in row order.
 * Results are bit-identical to the one-device calls.  out == NULL: stats only. */
int loosb200_replicate(const loosb200_frames* src, int device, loosb200_frames** out);
/* loosb200_stage_frames + loosb200_replicate in one, with the upload spread over the devices: device k
 * stages the k-th slice of the frames over its own PCIe link, then every device collects the other
 * slices device-to-device.  out[0..n-1] receive the replicas (out[k] on devices[k]). */
int loosb200_stage_frames_multi(loosb200_frames** out, const int* devices, int n, const float* host_xyz,
                                size_t F, size_t natoms, const uint32_t* sel, size_t nsel, int layout);
int loosb200_rmsds_self_multi(loosb200_frames* const* replicas, int n, float* out, size_t ld,
                              loosb200_stats* stats, int engine);
int loosb200_rmsds_self_multi_device(loosb200_frames* const* replicas, int n, float* out_dev, size_t ld,
                                     loosb200_stats* stats, int engine);
/* a[d] and b[d] on the same device; rows of `a` are split over the devices */
int loosb200_rmsds_cross_multi(loosb200_frames* const* a, loosb200_frames* const* b, int n, float* out,
                               size_t ld, loosb200_stats* stats, int engine);
int loosb200_rmsds_self_multi_text(loosb200_frames* const* replicas, int n, unsigned precision,
                                   loosb200_sink sink, void* user, loosb200_stats* stats, int engine);
int loosb200_rmsds_cross_multi_text(loosb200_frames* const* a, loosb200_frames* const* b, int n,
                                    unsigned precision, loosb200_sink sink, void* user,
                                    loosb200_stats* stats, int engine);
/* Page-locked host memory, pinned for every device of the process (result matrices of the host-output
 * calls: pageable memory works but the copies then run at a fraction of the PCIe rate). */
void* loosb200_host_alloc(size_t bytes);
void loosb200_host_free(void* p);

/* ---- single-reference streaming RMSD  (Tools/rmsd2ref.cpp:205-216,238-245) ----
 * For every frame f of `align`: W_f = superposition of the frame's align
 * selection onto ref_align (AtomicGroup::superposition == alignment::kabsch,
 * src/alignment.cpp:217-233), then rmsd_f = sqrt(sum |W_f x - y|^2 / n) over the
 * rmsd selection (AtomicGroup::applyTransform + ::rmsd).
 *   align      staged align-selection coordinates (N_a atoms)
 *   rmsd_set   staged rmsd-selection coordinates (N_r atoms) of the SAME frames,
 *              or NULL to use `align` for both
 *   ref_align  3*N_a doubles, xyz interleaved (the target's --talign selection), or
 *              NULL: no alignment, W_f = identity (rmsd2ref --align '')
 *   ref_rmsd   3*N_r doubles (the target's --trmsd selection); NULL => ref_align
 *   out_rmsd   F doubles
 *   out_xforms F x 16 doubles, row-major 4x4 GMatrix (src/Matrix44.hpp:67-78,
 *              maps frame -> reference), or NULL */
int loosb200_rmsd2ref(loosb200_frames* align, loosb200_frames* rmsd_set, const double* ref_align,
                      const double* ref_rmsd, double* out_rmsd, double* out_xforms);

/* rmsd2ref's default mode (no --target): iterativeAlignment over the staged
 * align selection (src/alignment.cpp:355-404: target_0 = centred first frame;
 * align all, average, until rms(avg,target) <= tol or iter > maxiter), then
 * averageStructure of the rmsd selection under the final transforms
 * (src/ensembles.cpp:91-121) and the per-frame RMSD to that average.
 *   out_avg   3*N_r doubles (the average structure) or NULL
 *   iters, final_rms: as the reference's result tuple; may be NULL */
int loosb200_rmsd2ref_iterative(loosb200_frames* align, loosb200_frames* rmsd_set, double tol,
                                int maxiter, double* out_rmsd, double* out_xforms,
                                double* out_avg, int* iters, double* final_rms);

/* ---- all-to-all RMSD with different alignment and RMSD selections
 *      (Packages/Clustering/rmsds-align.py:236-256) ----
 * out(i,j) = rmsd of rmsd2's frame j against rmsd1's frame i moved by the superposition of
 * align1's frame i onto align2's frame j (AtomicGroup::alignOnto / applyTransform / rmsd:
 * src/AG_linalg.cpp:188-197, src/AG_numerical.cpp:363-370,301-318).  Not symmetric in general.
 *   align1, rmsd1  align- and rmsd-selection coordinates of the SAME F1 frames (trajectory 1);
 *                  rmsd1 == NULL: the align selection serves as both
 *   align2, rmsd2  the same for trajectory 2 (F2 frames); pass the handles of trajectory 1
 *                  again for the one-trajectory case
 *   out            host, F1 x F2 doubles column-major: out[j*ld + i], ld >= F1
 * Selections of the two trajectories must agree in size (LOOSB200_ERR_SIZE_MISMATCH). */
int loosb200_rmsds_align(loosb200_frames* align1, loosb200_frames* rmsd1, loosb200_frames* align2,
                         loosb200_frames* rmsd2, double* out, size_t ld);
/* The same with a choice of engine.  LOOSB200_ENGINE_FP64 (what loosb200_rmsds_align uses): both 3x3
 * contractions of a pair in double on the CUDA cores, results within 1e-8 A of the reference's.
 * LOOSB200_ENGINE_TENSOR: both contractions on the int8 tensor path (exact integer sums of coordinates
 * on a 24-bit fixed-point grid; the rmsd selection is put on its grid about the ALIGN centroid), the
 * rotation from the quartic's root as in the fp64 engine; results within ~1e-6 A, an order of magnitude
 * faster; selections of up to 100,000 atoms.  LOOSB200_ENGINE_AUTO: tensor when it applies. */
int loosb200_rmsds_align_engine(loosb200_frames* align1, loosb200_frames* rmsd1, loosb200_frames* align2,
                                loosb200_frames* rmsd2, double* out, size_t ld, int engine);

/* One pass of iterativeAlignment's loop body over THIS handle's frames (src/alignment.cpp:379-390:
 * alignOnto(target) for every frame, transformed coordinates summed), for callers that shard a
 * trajectory over several GPUs / processes: sum the per-shard results (the one collective of the
 * whole RMSD path), divide by the total frame count, compare with the target and
 * iterate -- loos_b200/multigpu.py::rmsd2ref_iterative_sharded is that loop.
 *   target     3*N_align doubles (need not be centred: kabsch centres it, src/alignment.cpp:217-233)
 *   accum_set  frames whose transformed coordinates are summed (NULL: the align set itself);
 *              with the rmsd selection this is averageStructure's numerator (src/ensembles.cpp:108-118)
 *   sum_out    3*N_accum doubles, xyz interleaved: sum over the handle's frames of W_f x_f
 *   out_xforms F x 16 doubles or NULL */
int loosb200_align_accumulate(loosb200_frames* align, loosb200_frames* accum_set, const double* target,
                              double* sum_out, double* out_xforms);

/* How the one-process multi-GPU calls cut the 40-frame row tiles [0, T), T = ceil(F / 40), into `nparts`
 * contiguous ranges: equal work (row tile t of the triangle costs t + 1 tiles, of a rectangle a constant),
 * or, with `weights` (one per part, NULL = equal), work proportional to them -- the host-output calls
 * weight a device by 1 / max(compute, bytes / its PCIe rate) per pair.  cuts_out[nparts + 1]:
 * part p owns row tiles [cuts[p], cuts[p + 1]).  Host logic only (no device needed). */
int loosb200_share_cuts(size_t F, int nparts, int self, const double* weights, uint32_t* cuts_out);

/* ---- timing hooks (bench.py) ----
 * Device time in milliseconds of the kernels of the most recent compute call on
 * this handle, measured with CUDA events on the stream they were launched on:
 * [0] contraction+solver kernel(s), [1] everything else in the call (staging
 * of operands, reductions, copies).  Returns the number of kernel launches the
 * call made (>= 1), or a negative error code. */
int loosb200_last_timing(const loosb200_frames* h, double* ms_main, double* ms_other);
/* Sustained dense int8 tensor-pipe rate of `device` (the roofline denominator of the all-to-all
 * kernel): every SM pair issues back-to-back tcgen05.mma.cta_group::2.kind::i8,
 * M = 256, N = 256, K = 32, operands resident in shared memory, for about `seconds` (>= 2: under
 * the power cap).  tops_out: tera-operations per second (2 per multiply-add); ms_out: duration of
 * the measured launch, may be NULL. */
int loosb200_measure_int8_peak(int device, double seconds, double* tops_out, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* LOOSB200_H */

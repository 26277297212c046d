// Shared declarations of the loosb200 CUDA implementation (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in nsys / ncu timelines, free otherwise

#include "../../include/loosb200.h"

namespace loosb200 {

// NVTX range for the duration of a scope: the stages of a call (stage / quantise / contract / copy)
// are named in profiler timelines.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Frames are grouped in blocks of FRAMES_PER_BLOCK; one block of frames occupies
// ROWS_PER_BLOCK rows of the tensor-core operand planes (row = 3*j + axis for
// frame j of the block, rows 30 and 31 are zero) so that the three axes of one
// frame always live in the same warp's 32 TMEM lanes.  A tile is 4 blocks =
// 40 frames = 128 operand rows.
constexpr int FRAMES_PER_BLOCK = 10;
constexpr int ROWS_PER_BLOCK = 32;
constexpr int BLOCKS_PER_TILE = 4;
constexpr int TILE_FRAMES = FRAMES_PER_BLOCK * BLOCKS_PER_TILE;  // 40
constexpr int TILE_ROWS = ROWS_PER_BLOCK * BLOCKS_PER_TILE;      // 128
// Tensor path: an accumulator tile is 40 "a" frames (128 TMEM lanes) x 32 "b"
// frames (96 TMEM columns per weight class, 5 classes = 480 of 512 columns).  The "b" operand
// is a second copy of the digit planes with DENSE rows (row = 3*frame + axis, no padding): TMEM
// columns have no warp structure to respect, so all 96 columns of a class carry data.
constexpr int COL_TILE_FRAMES = 32;
constexpr int COL_TILE_ROWS = 3 * COL_TILE_FRAMES;  // 96
constexpr int K_CHUNK = 128;   // atoms (bytes of int8) per pipeline stage = one 128-byte swizzle row
constexpr int NUM_SLICES = 3;  // signed 8-bit digits of the 24-bit fixed-point coordinate
// int32 accumulators hold up to 3 slice products of magnitude <= 2^14 per atom: 43,000 atoms per accumulation
// segment.  Larger selections are contracted in several segments whose class sums the epilogue adds up in
// 64-bit integers (pairs_tensor.cu, SEG); the exact sum of squares of a frame (int64, k_quantise) bounds them
// at 100,000 atoms: 8355711^2 x 100,000 = 6.98e18 < 2^63.
constexpr size_t TENSOR_SEG_ATOMS = 43000;
constexpr size_t TENSOR_MAX_ATOMS = 100000;
// largest magnitude three balanced base-256 digits in [-128, 127] can hold on the positive side
constexpr long long Q_MAX = 127LL * 65793LL;  // 8355711

void set_last_error(const std::string& s);

// Size-keyed cache of device buffers (per device).  A staged frame set of C4 size owns
// ~2.2 GB and the host-output path a 40 GB bounce matrix; cudaMalloc/cudaFree of such
// blocks costs tens to hundreds of ms and synchronises the device, so freed blocks are
// kept for the next call of the same shape (loosb200_trim() returns them to the driver).
cudaError_t dev_alloc(void** p, size_t bytes);
void dev_release(void* p);
cudaError_t host_alloc(void** p, size_t bytes);  // pinned host blocks, same cache
void host_release(void* p);
template <class T> inline cudaError_t dev_alloc_t(T** p, size_t bytes) { return dev_alloc(reinterpret_cast<void**>(p), bytes); }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
// text.cu: "# m n (0)" + the rows of a device matrix (column-major) as "%.{p}g " text through `sink`
int stream_matrix_text(const float* d_M, size_t m, size_t n, size_t ld, bool symmetric, unsigned precision,
                       loosb200_sink sink, void* user, cudaStream_t st);
// the pieces, for results that are produced band by band: header line; m rows of n values with row j
// contiguous at rows_src + j * pitch; exact zeros on the diagonal of rows [R0, R0 + nrows) of a band
int matrix_text_header(size_t m, size_t n, loosb200_sink sink, void* user);
int stream_rows_text(const float* rows_src, size_t pitch, size_t m, size_t n, unsigned precision,
                     loosb200_sink sink, void* user, cudaStream_t st);
void zero_band_diagonal(float* band, size_t pitch, size_t R0, uint32_t nrows, cudaStream_t st);

#define LB_CUDA(call)                                                       \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) return ::loosb200::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

struct Timing {
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, main0, main1, end
  int launches = 0;
  bool valid = false;
  // per-band events recorded right around the contraction kernel (ms_main = their sum)
  std::vector<cudaEvent_t> kev;
  int kev_used = 0;
};

}  // namespace loosb200

namespace loosb200 { struct Stager; }

// The opaque handle of include/loosb200.h.
struct loosb200_frames {
  loosb200::Stager* stager = nullptr;  // non-null between stage_begin and stage_end
  int device = 0;
  size_t F = 0, N = 0;
  size_t Npad = 0;  // N rounded up to 4 floats (16 B) -- row pitch of `raw`
  // fp32 SoA planes of the selected atoms, uncentred: raw[(f*3 + axis)*Npad + i]
  float* raw = nullptr;
  double* centroid = nullptr;  // [F][3]
  double* gd = nullptr;        // [F] sum |x - c|^2 in A^2
  float* maxabs = nullptr;     // [F] max |x - c| over atoms and axes

  // tensor-core operands (lazy): q8[(slice*rows_total + row)*Kpad + atom]
  int8_t* q8 = nullptr;
  int8_t* q8b = nullptr;    // the same digits as the "b" operand: q8b[(slice*rows_total_b + 3*frame + axis)*Kpad + atom]
  double* gq = nullptr;     // [F] sum q^2 of the quantised frame (exact integer, rounded once to double)
  int* unit_log2_dev = nullptr;
  int unit_log2 = 0;
  bool quantised = false;
  size_t rows_total = 0, rows_total_b = 0, Kpad = 0;
  CUtensorMap tmapA;  // 3-D map over q8 (atom, row, slice), box 128 x 128 x 3
  CUtensorMap tmapB;  // 3-D map over q8b, box 128 x 96 x 3
  CUtensorMap tmapB96, tmapB48;  // q8b, boxes 128 x 96 x 1 and 128 x 48 x 1: the halves of a "b" stage a CTA of a pair loads
  bool tmap_valid = false;

  cudaStream_t stream = nullptr;
  loosb200::Timing timing;
};

namespace loosb200 {

// What one all-to-all launch needs to know about its share of the tile space.
struct PairJob {
  const loosb200_frames* a;
  const loosb200_frames* b;   // == a for self mode
  bool self;                  // lower triangle only (j < i), else full rectangle
  bool symmetric;             // self mode: also store M(j,i); rows unpacked
  float* out;                 // device (or peer-mapped), may be null (stats only)
  size_t ld;
  // Strip delivery (self + symmetric, contiguous rows [row_off, B) of the matrix): `out` is the compact
  // row strip S, column-major with ld = B - row_off: M(i,j) at out[j*ld + (i - row_off)], and the mirror
  // entry M(j,i) goes into S as well when j >= row_off (the band's diagonal block), else into the
  // column strip out_t[(i - row_off)*ld_t + j].  row_off = 0 / out_t = null is the plain F x F matrix.
  size_t row_off = 0;
  float* out_t = nullptr;
  size_t ld_t = 0;
  // row blocks (of TILE_FRAMES frames of `a`) this launch covers, ascending
  const uint32_t* rows_host;       // [n_row_tiles] same list on the host
  const uint32_t* row_tiles_dev;   // [n_row_tiles]
  const uint32_t* tile_prefix_dev; // [n_row_tiles+1] prefix sum of column tiles per row tile
  uint32_t n_row_tiles;
  uint64_t n_tiles;
  bool packed_rows;           // out row index = position in the launch's row list
  bool canon = false;         // rectangle over ONE frame set (a and b hold the same frames): the entry (i, j > i) is
                              // evaluated exactly as its mirror image (j, i) would be, so that a matrix assembled
                              // from such rectangles is bit-for-bit the symmetric one of the triangle walk
  // rmsds-align on the tensor engine: two launches over the same (packed, rectangular) tile space --
  // Rotation over the align selections fills `quat` ([column j][local row][4] doubles, ldq local rows),
  // Trace over the rmsd selections (quantised about the ALIGN centroids) turns it into the result band
  // out_d[j * ldq + local row] (pairs_tensor.cu, KIND_ROT / KIND_TRACE).  `out` and stats are unused.
  enum Kind { Rmsd = 0, Rotation = 1, Trace = 2 };
  Kind kind = Rmsd;
  double* quat = nullptr;
  double* out_d = nullptr;
  size_t ldq = 0;
  double* stats_dev;          // [2]: sum, max-as-bits ; may be null
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // optional: recorded right around the contraction kernel
  std::vector<void*>* scratch = nullptr;  // device scratch (dev_alloc) the engine leaves for the caller to
                                          // dev_release once the stream has been synchronised
};

int ensure_quantised(loosb200_frames* h, int force_unit_log2 /* INT_MIN: auto */);
int alloc_frames(loosb200_frames** out, size_t F, size_t N, int device);  // stage.cu: empty frame set
// capi.cu: the driver behind every all-to-all entry point (see there), and the work split of multi.cu
enum class OutKind { None, Copy, Device };
int run_pairs(loosb200_frames* a, loosb200_frames* b, bool self, bool symmetric, bool packed,
              const std::vector<uint32_t>& rows, float* out, OutKind okind, size_t ld,
              loosb200_stats* stats, int engine, bool canon = false);
std::vector<uint32_t> contiguous_cuts(size_t F, int nparts, bool self, const double* weights = nullptr);
std::vector<uint32_t> row_tile_range(uint32_t t0, uint32_t t1);
int launch_pairs_fp64(const PairJob& job, cudaStream_t st);
int launch_pairs_tensor(const PairJob& job, cudaStream_t st, int* launches);
bool tensor_path_available();

}  // namespace loosb200

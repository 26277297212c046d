// Matrix text on the GPU.
//
// The reference prints the result with `cout << setprecision(p) << M` (src/MatrixImpl.hpp:210-222,
// Tools/rmsds.cpp:566-567): "# m n (0)\n", then one line per row, every value as printf's "%.{p}g"
// followed by a space.  At BASELINE's C4 size that is 10^10 values / 45 GB of text: formatting it
// on the host costs a minute even with the exact integer formatter on all cores
// (loos_b200/frontend/frontend.cpp), against 0.43 s for computing the matrix.  Here the float32
// matrix never leaves HBM: rows are formatted by CTAs (same exact integer arithmetic as the host
// formatter: a float times a power of ten, rounded half-even), packed through shared memory into
// one contiguous byte stream per batch of rows, and only the text crosses PCIe.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace loosb200 {
namespace {

constexpr int TX_THREADS = 256;
constexpr int TX_VPT = 4;                          // values per thread and chunk
constexpr int TX_CHUNK = TX_THREADS * TX_VPT;      // 1024 values
constexpr int TX_MAXLEN = 16;                      // "-d.dddddddde+XX " : sign, 9 digits, point, exponent, space

__constant__ double c_pow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
__constant__ unsigned long long c_pow10u[20] = {
    1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
    10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull, 1000000000000000ull,
    10000000000000000ull, 100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull};

struct Txt {  // up to 16 characters packed little-endian into two registers
  unsigned long long lo, hi;
  int len;
  __device__ __forceinline__ void put(char c) {
    if (len < 8) lo |= (unsigned long long)(unsigned char)c << (8 * len);
    else hi |= (unsigned long long)(unsigned char)c << (8 * (len - 8));
    ++len;
  }
  __device__ __forceinline__ char at(int k) const {
    return (char)((k < 8 ? lo >> (8 * k) : hi >> (8 * (k - 8))) & 0xff);
  }
};

// "%.{p}g " of v.  Exact for 0 and for 1e-5 <= |v| < 1e9 with 1 <= p <= 9; anything else sets
// *slow (the row is then formatted on the host with snprintf).
__device__ __forceinline__ Txt format_g(float v, unsigned p, bool* slow) {
  Txt t;
  t.lo = t.hi = 0ull; t.len = 0;
  const uint32_t bits = __float_as_uint(v), mag = bits & 0x7fffffffu;
  if (mag == 0) {
    if (bits >> 31) t.put('-');
    t.put('0'); t.put(' ');
    return t;
  }
  const double av = fabs((double)v);
  if (!(av >= 1e-5) || !(av < 1e9)) { *slow = true; t.put('?'); t.put(' '); return t; }
  int X;  // decimal exponent of the leading digit (exact comparisons, see the host formatter)
  if (av >= 1.0) {
    X = 0;
    while (X < 8 && av >= c_pow10[X + 1]) ++X;
  } else {
    X = -1;
    while (X > -5 && av * c_pow10[-X] < 1.0) --X;
  }
  const int e_raw = (int)(mag >> 23);
  unsigned long long m = mag & 0x7fffffu;
  int e;
  if (e_raw == 0) e = -149; else { m |= 0x800000u; e = e_raw - 150; }
  const int k = (int)p - 1 - X;  // -8 .. 13
  if (k > 12) { *slow = true; t.put('?'); t.put(' '); return t; }  // p = 9 below 1e-4: needs more than 64 bits
  unsigned long long n;
  if (k >= 0 && k <= 11 && e < 0 && e > -64) {
    const unsigned long long num = m * c_pow10u[k];
    const int sh = -e;
    const unsigned long long q = num >> sh, r = num & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    n = q + ((r > half || (r == half && (q & 1))) ? 1 : 0);
  } else {
    // rational m 2^e 10^k with both parts in 64 bits for every value of the fast range:
    // k >= 0 with e >= 0 only for 2^24 <= av < 10^p, where k <= 1 and e <= 6;  k in 12..13 only for
    // av < 1e-3 at p = 9, handled by dividing the power of ten into both sides
    unsigned long long num = m, den = 1;
    int kk = k;
    if (kk > 11) {  // k = 12: 10^12 = 2 * 5 * 10^11 -- the factor 2 goes into the binary exponent
      num *= 5ull;
      e += 1;
      kk = 11;
    }
    if (kk >= 0) num *= c_pow10u[kk]; else den *= c_pow10u[-kk];
    if (e >= 0) num <<= e; else den <<= -e;
    const unsigned long long q = num / den, r = num - q * den;
    n = q + ((2 * r > den || (2 * r == den && (q & 1))) ? 1 : 0);
  }
  if (n >= c_pow10u[p]) { n /= 10; ++X; }
  char dig[10];
#pragma unroll
  for (int i = 8; i >= 0; --i)
    if (i < (int)p) { dig[i] = (char)('0' + (int)(n % 10)); n /= 10; }
  int nd = (int)p;
  while (nd > 1 && dig[nd - 1] == '0') --nd;
  if (bits >> 31) t.put('-');
  if (X < -4 || X >= (int)p) {
    t.put(dig[0]);
    if (nd > 1) { t.put('.'); for (int i = 1; i < nd; ++i) t.put(dig[i]); }
    t.put('e');
    int ax = X;
    if (ax < 0) { t.put('-'); ax = -ax; } else t.put('+');
    t.put((char)('0' + ax / 10)); t.put((char)('0' + ax % 10));
  } else if (X >= 0) {
    for (int i = 0; i <= X; ++i) t.put(i < nd ? dig[i] : '0');
    if (nd > X + 1) { t.put('.'); for (int i = X + 1; i < nd; ++i) t.put(dig[i]); }
  } else {
    t.put('0'); t.put('.');
    for (int i = 0; i < -X - 1; ++i) t.put('0');
    for (int i = 0; i < nd; ++i) t.put(dig[i]);
  }
  t.put(' ');
  return t;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < TX_THREADS / 32 ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < TX_THREADS / 32) warp_sums[lane] = w;  // inclusive
  }
  __syncthreads();
  const uint32_t before = warp ? warp_sums[warp - 1] : 0u;
  *total = warp_sums[TX_THREADS / 32 - 1];
  __syncthreads();
  return before + x - v;
}

// rows [row0, row0 + nrows) of an m x n matrix whose row j is CONTIGUOUS at M + j * pitch (the
// symmetric self matrix read by columns, or a transposed copy).  WRITE = false: row_len[r] = bytes of
// row r including '\n'.  WRITE = true: the text of row r goes to text + row_off[r].
template <bool WRITE>
__global__ void __launch_bounds__(TX_THREADS)
k_text_rows(const float* __restrict__ M, size_t pitch, size_t n, size_t row0, uint32_t nrows, unsigned p,
            unsigned long long* __restrict__ row_len, const unsigned long long* __restrict__ row_off,
            char* __restrict__ text, int* __restrict__ slow_rows) {
  __shared__ uint32_t warp_sums[TX_THREADS / 32];
  __shared__ __align__(16) char s_txt[TX_CHUNK * TX_MAXLEN + 32];
  for (uint32_t r = blockIdx.x; r < nrows; r += gridDim.x) {
    const float* row = M + (row0 + r) * pitch;
    unsigned long long written = 0;  // bytes of this row so far (same in every thread)
    bool slow = false;
    for (size_t c0 = 0; c0 < n; c0 += TX_CHUNK) {
      Txt t[TX_VPT];
      uint32_t mine = 0;
#pragma unroll
      for (int q = 0; q < TX_VPT; ++q) {
        const size_t i = c0 + (size_t)threadIdx.x * TX_VPT + q;
        if (i < n) { t[q] = format_g(__ldg(row + i), p, &slow); mine += (uint32_t)t[q].len; }
        else { t[q].lo = t[q].hi = 0ull; t[q].len = 0; }
      }
      uint32_t total;
      const uint32_t off = block_exclusive_scan(mine, &total, warp_sums);
      if (WRITE) {
        char* dst = text + row_off[r] + written;
        const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);  // keep shared and global offsets congruent mod 16
        uint32_t o = mis + off;
#pragma unroll
        for (int q = 0; q < TX_VPT; ++q)
          for (int k = 0; k < t[q].len; ++k) s_txt[o++] = t[q].at(k);
        __syncthreads();
        const uint32_t end = mis + total;
        // head bytes up to the first 16-byte boundary, whole uint4 in the middle, tail bytes
        const uint32_t a0 = mis ? 16u : 0u, a1 = end & ~15u;
        if (a1 > a0) {
          for (uint32_t b = mis + threadIdx.x; b < (a0 < end ? a0 : end); b += TX_THREADS) dst[b - mis] = s_txt[b];
          for (uint32_t b = a0 / 16 + threadIdx.x; b < a1 / 16; b += TX_THREADS)
            *reinterpret_cast<uint4*>(dst - mis + 16 * (size_t)b) = *reinterpret_cast<const uint4*>(s_txt + 16 * (size_t)b);
          for (uint32_t b = a1 + threadIdx.x; b < end; b += TX_THREADS) dst[b - mis] = s_txt[b];
        } else {
          for (uint32_t b = mis + threadIdx.x; b < end; b += TX_THREADS) dst[b - mis] = s_txt[b];
        }
        __syncthreads();
      }
      written += total;
    }
    if (WRITE) {
      if (threadIdx.x == 0) text[row_off[r] + written] = '\n';
    } else if (threadIdx.x == 0) {
      row_len[r] = written + 1;
    }
    if (__syncthreads_or(slow ? 1 : 0) && threadIdx.x == 0 && slow_rows) slow_rows[r] = 1;
  }
}

// out[j * m_in + i] = in[i * ld + j]  (an n_in-column, column-major matrix -> its rows contiguous)
__global__ void k_transpose(const float* __restrict__ in, size_t ld, size_t rows, size_t cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const size_t bx = (size_t)blockIdx.x * 32, by = (size_t)blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const size_t i = bx + threadIdx.x, j = by + k;  // element (row i, col j) at in[j * ld + i]
    tile[k][threadIdx.x] = (i < rows && j < cols) ? in[j * ld + i] : 0.0f;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const size_t i = bx + k, j = by + threadIdx.x;
    if (i < rows && j < cols) out[i * cols + j] = tile[threadIdx.x][k];
  }
}

// rows [R0, R0 + nrows) of a self matrix computed as a rectangle: M(i, i) = 0 exactly
// (SingleWorker::calc never writes the diagonal, Tools/rmsds.cpp:359-365)
__global__ void k_zero_diag(float* band, size_t pitch, size_t R0, uint32_t nrows) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nrows) band[(size_t)r * pitch + R0 + r] = 0.0f;
}

}  // namespace

void zero_band_diagonal(float* band, size_t pitch, size_t R0, uint32_t nrows, cudaStream_t st) {
  k_zero_diag<<<(nrows + 255) / 256, 256, 0, st>>>(band, pitch, R0, nrows);
}

int matrix_text_header(size_t m, size_t n, loosb200_sink sink, void* user) {
  char head[64];
  const int hl = snprintf(head, sizeof(head), "# %zu %zu (0)\n", m, n);
  return sink(user, head, (size_t)hl) ? LOOSB200_ERR_INVALID : LOOSB200_OK;
}

// Streams "# m n (0)\n" and the m rows of the device matrix (column-major, M(i,j) at d_M[j*ld + i])
// to `sink`.  `symmetric`: M(i,j) == M(j,i), so row j is read from column j.
int stream_matrix_text(const float* d_M, size_t m, size_t n, size_t ld, bool symmetric, unsigned precision,
                       loosb200_sink sink, void* user, cudaStream_t st) {
  int rc = matrix_text_header(m, n, sink, user);
  if (rc) return rc;
  if (symmetric) return stream_rows_text(d_M, ld, m, n, precision, sink, user, st);
  float* d_T = nullptr;
  if (dev_alloc_t(&d_T, sizeof(float) * m * n) != cudaSuccess) { cudaGetLastError(); return LOOSB200_ERR_NOMEM; }
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
  k_transpose<<<grid, block, 0, st>>>(d_M, ld, m, n, d_T);
  rc = stream_rows_text(d_T, n, m, n, precision, sink, user, st);
  dev_release(d_T);
  return rc;
}

// m rows of n values, row j contiguous at rows_src + j * pitch, as text lines through `sink`.
int stream_rows_text(const float* rows_src, size_t pitch, size_t m, size_t n, unsigned precision,
                     loosb200_sink sink, void* user, cudaStream_t st) {
  if (precision == 0) precision = 1;  // "%.0g" formats like "%.1g"
  if (precision > 9) return LOOSB200_ERR_UNSUPPORTED;
  // Batches of rows (<= 256 MB of text), two buffers deep: while batch b is being formatted, the text
  // of batch b-1 is on its way to the host on a second stream and the sink consumes batch b-2.
  const size_t worst_row = n * (size_t)(precision + 7) + 1;
  const size_t batch = std::max<size_t>(1, std::min<size_t>(m, ((size_t)256 << 20) / worst_row));
  const size_t cap = batch * worst_row + 64;
  char* d_text[2] = {nullptr, nullptr};
  char* h_text[2] = {nullptr, nullptr};
  unsigned long long *d_len = nullptr, *d_off = nullptr;
  int* d_slow = nullptr;
  cudaStream_t cst = nullptr;
  cudaEvent_t formatted[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  bool ok = dev_alloc_t(&d_len, sizeof(unsigned long long) * batch) == cudaSuccess &&
            dev_alloc_t(&d_off, sizeof(unsigned long long) * batch) == cudaSuccess &&
            dev_alloc_t(&d_slow, sizeof(int) * batch) == cudaSuccess &&
            cudaStreamCreateWithFlags(&cst, cudaStreamNonBlocking) == cudaSuccess;
  for (int k = 0; k < 2 && ok; ++k)
    ok = dev_alloc_t(&d_text[k], cap) == cudaSuccess && host_alloc(reinterpret_cast<void**>(&h_text[k]), cap) == cudaSuccess &&
         cudaEventCreateWithFlags(&formatted[k], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming) == cudaSuccess;
  auto cleanup = [&]() {
    for (int k = 0; k < 2; ++k) {
      dev_release(d_text[k]); host_release(h_text[k]);
      if (formatted[k]) cudaEventDestroy(formatted[k]);
      if (copied[k]) cudaEventDestroy(copied[k]);
    }
    dev_release(d_len); dev_release(d_off); dev_release(d_slow);
    if (cst) cudaStreamDestroy(cst);
  };
  if (!ok) { cudaGetLastError(); cleanup(); return LOOSB200_ERR_NOMEM; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  struct Batch { size_t r0; uint32_t nr; unsigned long long total; std::vector<unsigned long long> len, off; std::vector<int> slow; };
  Batch bt[2];
  std::vector<float> rowbuf;
  int rc = LOOSB200_OK;

  auto issue = [&](size_t r0, int k) -> int {  // format rows [r0, r0 + nr) into d_text[k], start its copy to h_text[k]
    Batch& B = bt[k];
    B.r0 = r0; B.nr = (uint32_t)std::min(batch, m - r0);
    B.len.resize(B.nr); B.off.resize(B.nr); B.slow.resize(B.nr);
    const unsigned grid = (unsigned)std::min<size_t>(B.nr, (size_t)sms * 8);
    cudaMemsetAsync(d_slow, 0, sizeof(int) * B.nr, st);
    k_text_rows<false><<<grid, TX_THREADS, 0, st>>>(rows_src, pitch, n, r0, B.nr, precision, d_len, nullptr, nullptr, d_slow);
    cudaMemcpyAsync(B.len.data(), d_len, sizeof(unsigned long long) * B.nr, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(B.slow.data(), d_slow, sizeof(int) * B.nr, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "matrix text (lengths)", __FILE__, __LINE__);
    B.total = 0;
    for (uint32_t r = 0; r < B.nr; ++r) { B.off[r] = B.total; B.total += B.len[r]; }
    cudaMemcpyAsync(d_off, B.off.data(), sizeof(unsigned long long) * B.nr, cudaMemcpyHostToDevice, st);
    k_text_rows<true><<<grid, TX_THREADS, 0, st>>>(rows_src, pitch, n, r0, B.nr, precision, nullptr, d_off, d_text[k], nullptr);
    cudaEventRecord(formatted[k], st);
    cudaStreamWaitEvent(cst, formatted[k], 0);
    cudaMemcpyAsync(h_text[k], d_text[k], B.total, cudaMemcpyDeviceToHost, cst);
    cudaEventRecord(copied[k], cst);
    e = cudaGetLastError();
    return e == cudaSuccess ? LOOSB200_OK : cuda_fail(e, "matrix text", __FILE__, __LINE__);
  };
  auto retire = [&](int k) -> int {  // hand the text of the batch in buffer k to the sink
    Batch& B = bt[k];
    const cudaError_t e = cudaEventSynchronize(copied[k]);
    if (e != cudaSuccess) return cuda_fail(e, "matrix text (copy)", __FILE__, __LINE__);
    bool any_slow = false;
    for (uint32_t r = 0; r < B.nr; ++r) any_slow = any_slow || B.slow[r];
    if (!any_slow) return sink(user, h_text[k], (size_t)B.total) ? LOOSB200_ERR_INVALID : LOOSB200_OK;
    // rows holding values outside the exact device range (|v| < 1e-5, >= 1e9, inf, nan): libc formats them
    for (uint32_t r = 0; r < B.nr; ++r) {
      if (!B.slow[r]) { if (sink(user, h_text[k] + B.off[r], (size_t)B.len[r])) return LOOSB200_ERR_INVALID; continue; }
      rowbuf.resize(n);
      if (cudaMemcpy(rowbuf.data(), rows_src + (B.r0 + r) * pitch, sizeof(float) * n, cudaMemcpyDeviceToHost) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "matrix text (row)", __FILE__, __LINE__);
      std::string line;
      char buf[48];
      for (size_t i = 0; i < n; ++i) { const int l = snprintf(buf, sizeof(buf), "%.*g ", (int)precision, (double)rowbuf[i]); line.append(buf, (size_t)l); }
      line += '\n';
      if (sink(user, line.data(), line.size())) return LOOSB200_ERR_INVALID;
    }
    return LOOSB200_OK;
  };

  size_t nb = 0;
  for (size_t r0 = 0; r0 < m && rc == LOOSB200_OK; r0 += batch, ++nb) {
    rc = issue(r0, (int)(nb & 1));
    if (rc == LOOSB200_OK && nb >= 1) rc = retire((int)((nb - 1) & 1));
  }
  if (rc == LOOSB200_OK && nb >= 1) rc = retire((int)((nb - 1) & 1));
  cudaStreamSynchronize(cst);
  cudaStreamSynchronize(st);
  cleanup();
  return rc;
}

}  // namespace loosb200

// All-to-all RMSD, CUDA-core fp64 variant.
//
// Same tile space, same epilogue (fused QCP solve, float32 store, symmetric fill,
// stats) as the tensor-core kernel in pairs_tensor.cu, but the 3x3 covariance of
// every frame pair is accumulated with DFMA from the fp32 planes centred in fp64
// -- i.e. the arithmetic of dgemm_('N','T',3,3,n,...) at src/alignment.cpp:34
// done for a 40 x 40 tile of pairs at once.  It serves selections larger than
// the int32 accumulation bound of the tensor path, and is the on-device
// cross-check for it.  Roofline: fp64 pipe (18 N flop per pair).
#include "common.cuh"
#include "qcp.cuh"

namespace loosb200 {

namespace {

constexpr int TF = TILE_FRAMES;  // 40
constexpr int KC = 16;           // atoms per smem chunk
constexpr int TFP = 42;          // row pitch of the smem tiles in doubles: 16-byte aligned pairs, and the k-fastest fill
                                 // hits 8 distinct bank pairs instead of 2 (pitch 40: 16-way conflicts on every store)
constexpr int NT = 416;          // 13 warps; threads 0..399 own a 2x2 block of pairs

__device__ __forceinline__ uint32_t find_row_tile(const uint32_t* __restrict__ prefix, uint32_t n,
                                                  uint64_t tile) {
  uint32_t lo = 0, hi = n;  // prefix[lo] <= tile < prefix[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if ((uint64_t)prefix[mid] <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(NT)
k_pairs_fp64(const float* __restrict__ rawA, const double* __restrict__ cenA,
             const double* __restrict__ gdA, uint32_t FA, size_t NpadA,
             const float* __restrict__ rawB, const double* __restrict__ cenB,
             const double* __restrict__ gdB, uint32_t FB, size_t NpadB, uint32_t N,
             const uint32_t* __restrict__ row_tiles, const uint32_t* __restrict__ prefix,
             uint32_t n_row_tiles, int self, int symmetric, int packed, int canon, float* __restrict__ out,
             size_t ld, float* __restrict__ out_t, size_t ld_t, uint32_t row_off,
             double* __restrict__ stats) {
  __shared__ __align__(16) double As[3][KC][TFP];
  __shared__ __align__(16) double Bs[3][KC][TFP];
  __shared__ double cA[TF][3], cB[TF][3];
  __shared__ double red_sum[13], red_max[13];

  const uint64_t tile = blockIdx.x;
  const uint32_t rt = find_row_tile(prefix, n_row_tiles, tile);
  const uint32_t I = row_tiles[rt];
  const uint32_t J = (uint32_t)(tile - prefix[rt]);
  const int tid = threadIdx.x;
  const bool worker = tid < TF * TF / 4;
  const int ty = tid / 20, tx = tid % 20;

  for (int t = tid; t < TF * 3; t += NT) {
    const uint32_t fa = I * TF + t / 3, fb = J * TF + t / 3;
    cA[t / 3][t % 3] = fa < FA ? cenA[(size_t)fa * 3 + t % 3] : 0.0;
    cB[t / 3][t % 3] = fb < FB ? cenB[(size_t)fb * 3 + t % 3] : 0.0;
  }
  __syncthreads();

  double acc[2][2][9];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[a][b][k] = 0.0;

  // the chunk after the one being multiplied is already on its way from global memory (registers):
  // with one CTA per SM nothing else would hide that latency.  A chunk is 120 rows (frame, axis) of
  // 16 atoms per operand = 2 x 480 float4; item = tid + r * NT names (operand, row, quarter).
  constexpr int ITEMS = 2 * TF * 3 * (KC / 4);       // 960
  constexpr int PF = (ITEMS + NT - 1) / NT;          // 3 float4 per thread and chunk
  float4 pv[PF];
  auto prefetch = [&](uint32_t k0) {
#pragma unroll
    for (int r = 0; r < PF; ++r) {
      const int item = tid + r * NT;
      const int op = item / (ITEMS / 2), within = item % (ITEMS / 2), row = within >> 2, qd = within & 3;
      const uint32_t fr = (op ? J : I) * TF + row / 3, at = k0 + 4 * qd;
      // Npad is a multiple of 4, so a quarter that starts below N lies inside the row
      const bool in = item < ITEMS && at < N && fr < (op ? FB : FA);
      const float* src = (op ? rawB : rawA) + ((size_t)fr * 3 + row % 3) * (op ? NpadB : NpadA) + at;
      pv[r] = in ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  prefetch(0);
  for (uint32_t k0 = 0; k0 < N; k0 += KC) {
#pragma unroll
    for (int r = 0; r < PF; ++r) {
      const int item = tid + r * NT;
      if (item < ITEMS) {
        const int op = item / (ITEMS / 2), within = item % (ITEMS / 2), row = within >> 2, qd = within & 3;
        const int f = row / 3, ax = row % 3;
        const bool fr_ok = (op ? J : I) * TF + f < (op ? FB : FA);
        const double c = op ? cB[f][ax] : cA[f][ax];
        double (*dst)[TFP] = op ? Bs[ax] : As[ax];
        const float v[4] = {pv[r].x, pv[r].y, pv[r].z, pv[r].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)  // padded atoms and frames beyond the sets contribute exact zeros
          dst[4 * qd + e][f] = fr_ok && k0 + 4 * qd + e < N ? (double)v[e] - c : 0.0;
      }
    }
    if (k0 + KC < N) prefetch(k0 + KC);
    __syncthreads();
    if (worker) {
#pragma unroll 4
      for (int k = 0; k < KC; ++k) {
        double a0[3], a1[3], b0[3], b1[3];
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {  // frames 2 ty, 2 ty + 1 and 2 tx, 2 tx + 1: one 16-byte load each
          const double2 av = *reinterpret_cast<const double2*>(&As[ax][k][2 * ty]);
          const double2 bv = *reinterpret_cast<const double2*>(&Bs[ax][k][2 * tx]);
          a0[ax] = av.x; a1[ax] = av.y;
          b0[ax] = bv.x; b1[ax] = bv.y;
        }
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            acc[0][0][3 * p + q] = fma(a0[p], b0[q], acc[0][0][3 * p + q]);
            acc[0][1][3 * p + q] = fma(a0[p], b1[q], acc[0][1][3 * p + q]);
            acc[1][0][3 * p + q] = fma(a1[p], b0[q], acc[1][0][3 * p + q]);
            acc[1][1][3 * p + q] = fma(a1[p], b1[q], acc[1][1][3 * p + q]);
          }
      }
    }
    __syncthreads();
  }

  double ssum = 0.0, smax = 0.0;
  if (worker) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const uint32_t la = 2 * ty + a, lb = 2 * tx + b;
        const uint32_t i = I * TF + la, j = J * TF + lb;
        if (i >= FA || j >= FB) continue;
        if (self && j >= i) {
          // the diagonal is never computed and stays 0 (Tools/rmsds.cpp:359-365)
          if (j == i && out) {
            if (packed) out[((size_t)rt * TF + la) * ld + j] = 0.f;
            else out[(size_t)j * ld + (i - row_off)] = 0.f;
          }
          continue;
        }
        double d;
        if (canon && j == i) {
          d = 0.0;  // the diagonal of a rectangle over one set: exact zero, as in the triangle walk
        } else if (canon && j > i) {
          // rectangle over one frame set: the entry above the diagonal is evaluated as its mirror image
          // (transposed covariance, swapped G), so that both are the same float (see PairJob::canon)
          const double* R = acc[a][b];
          const double Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
          d = qcp_rmsd(Rt, gdB[j], gdA[i], (double)N, 1.0);
        } else {
          d = qcp_rmsd(acc[a][b], gdA[i], gdB[j], (double)N, 1.0);
        }
        const float v = (float)d;  // Tools/rmsds.cpp:363: double -> float on store
        if (!canon || j < i) {  // canon: only the triangle counts (several devices' row ranges add up)
          ssum += (double)v;
          smax = fmax(smax, (double)v);
        }
        if (out) {
          if (packed) {
            out[((size_t)rt * TF + la) * ld + j] = v;
          } else {
            out[(size_t)j * ld + (i - row_off)] = v;
            if (symmetric) {  // strip delivery: see PairJob
              if (j >= row_off) out[(size_t)i * ld + (j - row_off)] = v;
              else out_t[(size_t)(i - row_off) * ld_t + j] = v;
            }
          }
        }
      }
  }
  if (stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
    if ((tid & 31) == 0) { red_sum[tid >> 5] = ssum; red_max[tid >> 5] = smax; }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0, m = 0.0;
      for (int w = 0; w < NT / 32; ++w) { s += red_sum[w]; m = fmax(m, red_max[w]); }
      atomicAdd(&stats[0], s);
      // non-negative doubles order like their bit patterns
      atomicMax(reinterpret_cast<unsigned long long*>(&stats[1]), (unsigned long long)__double_as_longlong(m));
    }
  }
}

}  // namespace

int launch_pairs_fp64(const PairJob& job, cudaStream_t st) {
  if (job.n_tiles == 0) return LOOSB200_OK;
  if (job.n_tiles > 0x7fffffffull) return LOOSB200_ERR_UNSUPPORTED;
  const loosb200_frames* a = job.a;
  const loosb200_frames* b = job.b;
  if (job.ev_k0) cudaEventRecord(job.ev_k0, st);
  k_pairs_fp64<<<(unsigned)job.n_tiles, NT, 0, st>>>(
      a->raw, a->centroid, a->gd, (uint32_t)a->F, a->Npad, b->raw, b->centroid, b->gd,
      (uint32_t)b->F, b->Npad, (uint32_t)a->N, job.row_tiles_dev, job.tile_prefix_dev,
      job.n_row_tiles, job.self ? 1 : 0, job.symmetric ? 1 : 0, job.packed_rows ? 1 : 0,
      (job.canon && !job.self) ? 1 : 0, job.out,
      job.ld, job.out_t, job.ld_t, (uint32_t)job.row_off, job.stats_dev);
  if (job.ev_k1) cudaEventRecord(job.ev_k1, st);
  LB_CUDA(cudaGetLastError());
  return LOOSB200_OK;
}

}  // namespace loosb200

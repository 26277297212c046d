// All-to-all RMSD on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// The F x F frame-pair covariance (nine N-long inner products per pair, the
// dgemm_ of src/alignment.cpp:34 for every pair at once) is one dense
// contraction  C = A A^T  with A = (3F) x N.  fp32 accumulation cannot hold the
// cancellation in Ga + Gb - 2 lambda, so the operands are
// 24-bit fixed point split into three signed 8-bit digits and multiplied with
// tcgen05.mma kind::i8: int32 accumulation is EXACT.  The nine digit products
// fall into five weight classes s = p + q (weight 2^(8s)); products of one class
// share an accumulator, so a tile needs 5 accumulators.
//
// Tile = 128 TMEM lanes x 96 columns per class:
//   lane   = 32*blk + 3*jf + axis   -> "a" frame 10*blk + jf of the row tile (40 frames)
//   column = 96*s + 3*jb + axis'    -> "b" frame jb of the column tile (32 frames, dense)
// 5 x 96 = 480 of the 512 TMEM columns.  The three axes of one "a" frame sit in
// three adjacent lanes of one warp, so the epilogue assembles each 3x3 with a few
// warp shuffles and solves one pair per lane (QCP, qcp.cuh) straight out of TMEM:
// covariance tiles never touch HBM.
//
// Warp roles (512 threads, 1 CTA per SM, persistent over a tile list):
//   warp 0   TMA producer: per 128-atom K chunk one 3-D box for A (128 B x 128 rows x
//            3 digits = 48 KB) and one for B (128 B x 96 x 3 = 36 KB), SWIZZLE_128B
//   warp 1   TMEM allocator + MMA issuer: per 32-atom K step six tcgen05.mma
//            (three with N = 192 spanning two adjacent classes, three with N = 96)
//            (both control warps run their loops with all lanes, so that addresses, coordinates and
//            descriptors are warp-uniform and sit in uniform registers; one elected lane issues)
//   warps 2-3 idle (they complete the control warp group, see setmaxnreg below)
//   warps 4-15 epilogue, one (lane quarter, 11/11/10-frame column group = 33/33/30 columns) each:
//            tcgen05.ld -> class sums in 64-bit integers -> row exchange by shuffles -> quartic
//            coefficients (fp64) -> TMEM handed back to the MMA warp -> fp32 Newton + float-float
//            correction -> float32 tile in smem -> coalesced (and mirrored) stores.
//            The fp64 part is serial with the next tile's MMAs on purpose: fp64 CUDA-core
//            math and the tensor pipe contend on sm_100 (see the epilogue).
// Default launch shape: CTA pairs (CL = 2 below), two row tiles against one column tile.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "qcp.cuh"

namespace loosb200 {

namespace {

// Two launch shapes (template parameter CL):
//   CL = 1  one CTA per tile, tcgen05 cta_group::1: stage = A 48 KB + B 36 KB, 2 stages
//   CL = 2  a CTA PAIR per two row tiles x one column tile, tcgen05 cta_group::2 (UMMA M = 256):
//           each CTA loads its own "a" tile and HALF of the shared "b" operand (digit 1 or 2 in
//           full and half the rows of digit 0); the tensor cores of both SMs read both halves.
//           Stage = A 48 KB + B 18 KB, 3 stages, and 21 % fewer operand bytes leave L2 per pair.
constexpr int MAX_STAGES = 3;
constexpr int A_STAGE_BYTES = NUM_SLICES * TILE_ROWS * K_CHUNK;      // 49152
constexpr int A_SLICE_BYTES = TILE_ROWS * K_CHUNK;                   // 16384
constexpr int B_SLICE_BYTES = COL_TILE_ROWS * K_CHUNK;               // 12288
template <int CL> struct Shape;
template <> struct Shape<1> {
  static constexpr int STAGES = 2;
  static constexpr int B_STAGE_BYTES = NUM_SLICES * B_SLICE_BYTES;   // 36864
};
template <> struct Shape<2> {
  static constexpr int STAGES = 3;
  static constexpr int B_STAGE_BYTES = B_SLICE_BYTES + B_SLICE_BYTES / 2;  // 18432: one digit + half of digit 0
};
template <int CL> __host__ __device__ constexpr int stage_bytes() { return A_STAGE_BYTES + Shape<CL>::B_STAGE_BYTES; }
// Solve rounds inside the drain, i.e. before TMEM is handed back (fp64 runs at full rate there, but the MMAs
// wait); the others run next to the next tile's MMAs at a fraction of that rate.  C4, same box, with the
// warp-uniform MMA issue: 1 round 405 ms, 2 rounds 368 ms, 3 rounds 372 ms, all four (RELEASE_AT=1) 376 ms.
#ifndef LB_INDRAIN
#define LB_INDRAIN 2
#endif
constexpr int NUM_CLASSES = 5;
constexpr int GROUP_FRAMES = 11;  // "b" frames per epilogue column group (33 columns); the third group has 10
constexpr int CLASS_COLS = COL_TILE_ROWS;  // 96
constexpr int TMEM_COLS = 512;
// Warp group 0 = the two control warps (+ two idle ones), warp groups 1-3 = the twelve epilogue warps.
// A sub-partition holds 512 registers per lane and always hosts four of these warps, so an even split
// would leave every warp 128 -- the epilogue then spills.  setmaxnreg moves the control group's
// share over: 24 + 3 x 160 = 504 registers per lane and sub-partition.
constexpr int NTHREADS = 512;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NWARPS = NTHREADS / 32;
constexpr int REGS_CONTROL = 32, REGS_EPILOGUE = 160;
constexpr int NUM_EPI_WARPS = 12;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;  // 384
constexpr int TPAD = TILE_FRAMES + 1;  // out tile row pitch (floats)
constexpr int NUM_SCHED = 4;           // depth of the tile-slot ring
constexpr uint32_t SLOT_END = 0xffffffffu;

struct __align__(8) SmemCtl {
  unsigned long long full[MAX_STAGES];
  unsigned long long empty[MAX_STAGES];
  unsigned long long tmem_full;
  unsigned long long tmem_empty;
  unsigned long long sched_full[NUM_SCHED];   // ring of tile slots handed out by the cluster leader
  unsigned long long sched_empty[NUM_SCHED];  // (only the leader CTA's copy is used)
  uint32_t sched_slot[NUM_SCHED];
  uint32_t tmem_base;
  uint32_t pad;
  float T[2][COL_TILE_FRAMES][TPAD];  // double-buffered out tile, T[.][b][a]
  double red_sum[NUM_EPI_WARPS], red_max[NUM_EPI_WARPS];
};

template <int CL> __host__ __device__ constexpr int smem_bytes() { return Shape<CL>::STAGES * stage_bytes<CL>() + (int)sizeof(SmemCtl) + 1024; }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU box.  A waiting warp backs
// off with nanosleep so that its polling does not eat the issue slots of the epilogue
// warps sharing its scheduler (the two control warps executed 39 % of all
// instructions of the kernel before this was added).
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if (spin > (1u << 24)) __trap();  // seconds
  }
}
// Same, acquiring at cluster scope: the data guarded by the barrier was written by another CTA.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if (spin > (1u << 24)) __trap();
  }
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_remote_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// CTA-pair forms (cute SM100_TMA_2SM_LOAD_3D, umma_arrive_multicast_2x1SM): both CTAs issue their
// loads against the LEADER's barrier (its shared::cluster address, from mapa); the leader's
// commit arrives on the same barrier in both CTAs.
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, unsigned long long* bar,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(map_to_cta(bar, 0)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int CL>
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CL == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// wait for the outstanding TMEM loads; the registers are listed so that no use of them is
// scheduled above the wait
__device__ __forceinline__ void tc_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

__device__ __forceinline__ void tc_ld2(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld18(uint32_t (&r)[18]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17])
               :: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor,
// cutlass mma_sm100_desc.hpp:98-121): start>>4 | LBO=1<<16 | SBO=(8 rows*128 B)>>4<<32
// | version 1<<46 | layout SWIZZLE_128B(2)<<61.  A K step of 32 int8 inside the 128-byte
// swizzle row is addressed by advancing the start address by 32 bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = signed
// int8, both K-major, M = m (128, or 256 across a CTA pair), N = n.
__host__ __device__ constexpr uint32_t make_idesc(int n, int m) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

#define LB_TICK(slot)                                               \
  do {                                                              \
    if (a.prof) {                                                   \
      const long long now__ = clock64();                           \
      pacc[slot] += now__ - tprev;                                  \
      tprev = now__;                                                \
    }                                                               \
  } while (0)

struct TensorArgs {
  const uint4* sblocks;        // super-blocks: {first row index, #rows, first column tile, #columns}
  const uint32_t* sb_prefix;   // [n_sb + 1] prefix sum of #rows * #columns (tile slots)
  uint32_t n_slots;            // total slots; slots above the diagonal are skipped by the leader
  uint32_t* counter;           // next slot to hand out (zeroed per launch)
  const uint32_t* row_tiles;   // row tile index -> I
  const double* gA; const double* gB;  // exact sum of squares of every quantised frame, as double
  uint32_t FA, FB, N, n_kchunks;
  uint32_t seg_chunks;         // K chunks per accumulation segment (SEG kernels; n_kchunks otherwise)
  int self, symmetric, packed;
  int canon;                   // rectangle over ONE frame set: entries above the diagonal are evaluated as their mirror image
  double unit;                 // Angstrom per fixed-point unit
  float* out; size_t ld;
  float* out_t; size_t ld_t; uint32_t row_off;  // strip delivery, see PairJob
  double* stats;
  // rmsds-align (KIND_ROT / KIND_TRACE, see the kernel): unit quaternions of the align-selection pass,
  // [column j][local row][4], and the double result band, column-major; both with `ldq` local rows
  double* quat; double* out_d; size_t ldq;
  int w256;        // 256: the weight of one digit, as a run-time value (see the epilogue's combine)
  int release_at;  // when the epilogue hands TMEM back (see kernel)
  long long* prof;  // optional [grid][NWARPS][8] cycle counters (LOOSB200_PROF=1)
};


// Tile slots are numbered super-block by super-block (row-major inside).  A slot is one tile
// (CL = 1) or, for a CTA pair (CL = 2), two consecutive row tiles of the job against one column
// tile: the pair shares the "b" operand, CTA r takes row r of the two.  Slots are handed out
// DYNAMICALLY: the producer thread of the leader CTA takes the next slot from a global counter
// and publishes it through a small shared-memory ring (one copy per CTA, written remotely for the
// partner); every other role of every CTA consumes the ring in order, so the roles stay in step
// without further talk, and all slots in flight on the GPU are consecutive -- they share one
// super-block's few MB of operands in L2.  (With a static slot = cluster + k * #clusters walk the
// CTAs drift apart over the ~28000 tiles each computes, and 55 % of the operand reads missed
// L2: profiles/r1_summary.md.)  A CTA of a pair whose own tile does not exist (above the
// diagonal, or an odd row count) while its partner's does still loads its half of "b" and some
// valid "a" rows and takes part in every barrier; its results are masked.
template <int CL>
struct TileWalk {
  uint32_t sb, crank, ridx, rphase;
  __device__ __forceinline__ TileWalk(uint32_t rank) : sb(0), crank(rank), ridx(0), rphase(0) {}
  __device__ __forceinline__ bool above_diagonal(const TensorArgs& a, uint32_t I, uint32_t J) const {
    return a.self && (uint64_t)J * COL_TILE_FRAMES >= (uint64_t)I * TILE_FRAMES + TILE_FRAMES;
  }
  // slot -> tile of this CTA; false when no CTA of the cluster has a tile there.
  // Slots arrive in increasing order, so the super-block search only moves forward.
  __device__ __forceinline__ bool decode(const TensorArgs& a, uint32_t slot, uint32_t& rt, uint32_t& I, uint32_t& J,
                                         bool& masked) {
    while (slot >= a.sb_prefix[sb + 1]) ++sb;
    const uint4 b = a.sblocks[sb];  // {first row index, #rows, first column tile, #columns}
    const uint32_t local = slot - a.sb_prefix[sb];
    const uint32_t lr = local / b.w, lc = local - lr * b.w;
    J = b.z + lc;
    masked = false;
    if (CL == 1) {
      rt = b.x + lr;
      I = a.row_tiles[rt];
      return !above_diagonal(a, I, J);
    }
    // rows of the job are ascending, so the second row of a pair reaches further right
    const uint32_t r0 = b.x + 2 * lr, r1 = r0 + 1 < b.x + b.y ? r0 + 1 : r0;
    const uint32_t I1 = a.row_tiles[r1];
    if (above_diagonal(a, I1, J)) return false;
    rt = crank == 0 ? r0 : r1;
    I = a.row_tiles[rt];
    masked = (crank == 1 && r1 == r0) || above_diagonal(a, I, J);
    return true;
  }
  // Leader producer thread: hand `slot` (or SLOT_END) to every consumer of the cluster.
  // Called by ALL lanes of the producer warp (the walk's state stays warp-uniform); lane 0 publishes.
  __device__ __forceinline__ void publish(SmemCtl* ctl, uint32_t slot) {
    mbar_wait_cluster<32>(&ctl->sched_empty[ridx], rphase ^ 1);
    if ((threadIdx.x & 31) == 0) {
      ctl->sched_slot[ridx] = slot;
      mbar_arrive(&ctl->sched_full[ridx]);
#pragma unroll
      for (uint32_t r = 1; r < (uint32_t)CL; ++r) {
        st_remote_u32(map_to_cta(&ctl->sched_slot[ridx], r), slot);
        mbar_arrive_remote(map_to_cta(&ctl->sched_full[ridx], r));
      }
    }
    __syncwarp();
    if (++ridx == NUM_SCHED) { ridx = 0; rphase ^= 1; }
  }
  // Consumers: next tile, false when the cluster is done.  WARP: called by all lanes of a warp
  // (one release per warp), otherwise by a single thread.
  template <bool WARP>
  __device__ __forceinline__ bool next(const TensorArgs& a, SmemCtl* ctl, uint32_t& rt, uint32_t& I, uint32_t& J,
                                       bool& masked) {
    mbar_wait_cluster<32>(&ctl->sched_full[ridx], rphase);
    const uint32_t slot = *(volatile uint32_t*)&ctl->sched_slot[ridx];
    if (WARP) __syncwarp();
    if (!WARP || (threadIdx.x & 31) == 0) {
      if (CL == 1 || crank == 0) mbar_arrive(&ctl->sched_empty[ridx]);
      else mbar_arrive_remote(map_to_cta(&ctl->sched_empty[ridx], 0));
    }
    if (++ridx == NUM_SCHED) { ridx = 0; rphase ^= 1; }
    if (slot == SLOT_END) return false;
    decode(a, slot, rt, I, J, masked);  // published slots always decode
    return true;
  }
};

// An epilogue warp hands the accumulators back: to the MMA thread of its own CTA, or of the leader.
template <int CL>
__device__ __forceinline__ void release_tmem(SmemCtl* ctl, uint32_t crank) {
  if (CL == 1 || crank == 0) mbar_arrive(&ctl->tmem_empty);
  else mbar_arrive_remote(map_to_cta(&ctl->tmem_empty, 0));
}

// What the epilogue makes of a pair's 3 x 3 sums (one instantiation each: a branch inside the coefficient
// rounds would keep the compiler from interleaving them):
//   KIND_RMSD   the minimal RMSD (float32 tile, symmetric fill, stats)
//   KIND_CANON  the same with mirror-canonical evaluation, see PairJob::canon
//   KIND_ROT    rmsds-align, pass 1 over the ALIGN selections: the optimal rotation as a unit quaternion
//               (eigenvector of Horn's K for the quartic's root, qcp_rotation_fast; Jacobi for degenerate
//               selections), written to a.quat
//   KIND_TRACE  rmsds-align, pass 2 over the RMSD selections quantised about the align centroids:
//               d^2 = (Gx + Gy - 2 sum_ab Z_ab C_ba) / N with Z from a.quat, double result in a.out_d
// Both passes walk the same tile geometry, so a pair meets the same lane -- and with it the same cyclic
// relabelling of the rows of R and C, which the trace is invariant under when both carry it.
constexpr int KIND_RMSD = 0, KIND_CANON = 1, KIND_ROT = 2, KIND_TRACE = 3;
// SEG: selections above 43,000 atoms.  An int32 accumulator holds three digit products of up to 2^14 per atom,
// i.e. 43,000 atoms; beyond that the K loop is cut into segments of a.seg_chunks chunks: at the end of a
// segment the epilogue drains the five classes into its 64-bit sums and hands TMEM back, the MMA thread
// starts the next segment with overwriting MMAs, and only the last drain is followed by the solve.  Its own
// instantiation: the common case keeps its instruction schedule.
template <int CL, int KIND, bool SEG = false>
__global__ void __launch_bounds__(NTHREADS, 1)
k_pairs_tensor(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapB96, const __grid_constant__ CUtensorMap mapB48,
               const TensorArgs a) {
  constexpr int NUM_STAGES = Shape<CL>::STAGES;
  constexpr int STAGE_BYTES = stage_bytes<CL>();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  SmemCtl* ctl = (SmemCtl*)(smem + NUM_STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;

  if (threadIdx.x == 0) {
    // pair: the leader's full[s] collects the bytes of both CTAs, its commits arrive on empty[s]
    // and tmem_full of both, and its tmem_empty collects the epilogue warps of both
    for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->tmem_full, 1);
    mbar_init(&ctl->tmem_empty, CL * NUM_EPI_WARPS);
    // ring consumers: MMA thread + epilogue warps of every CTA, producer threads of the non-leaders
    for (int s = 0; s < NUM_SCHED; ++s) {
      mbar_init(&ctl->sched_full[s], 1);
      mbar_init(&ctl->sched_empty[s], CL * (1 + NUM_EPI_WARPS) + (CL - 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CL == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "n"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {  // the same warp of both CTAs, same destination address
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "n"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // barriers of both CTAs exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp < FIRST_EPI_WARP) {
  // the register budgets hold for the code DOMINATED by the instruction: the two groups only meet again at the end
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
  if (warp == 0) {
    // ================= TMA producer =================
    // The whole warp walks the tiles (uniform control flow and coordinates, as in the MMA warp below); lane 0
    // talks to the global tile counter, one elected lane issues the TMA loads.
    {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        if (CL == 1) {
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        } else {
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB96) : "memory");
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB48) : "memory");
        }
      }
      uint32_t stage = 0, phase = 0;
      TileWalk<CL> walk(crank);
      uint32_t rt, I, J;
      bool masked;
      const bool leader = (CL == 1 || crank == 0);
      // the slot after the current one, fetched early
      auto take_slot = [&]() {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(a.counter, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
      };
      uint32_t ahead = leader ? take_slot() : 0u;
      while (true) {
        if (leader) {
          const uint32_t slot = ahead;
          if (slot >= a.n_slots) { walk.publish(ctl, SLOT_END); break; }
          ahead = take_slot();
          if (!walk.decode(a, slot, rt, I, J, masked)) continue;  // above the diagonal
          walk.publish(ctl, slot);
        } else if (!walk.template next<true>(a, ctl, rt, I, J, masked)) {
          break;
        }
        const int rowA = (int)I * TILE_ROWS;
        const int rowB = (int)J * COL_TILE_ROWS;
        for (uint32_t kc = 0; kc < a.n_kchunks; ++kc) {
          mbar_wait<32>(&ctl->empty[stage], phase ^ 1);
          unsigned char* sA = smem + stage * STAGE_BYTES;
          const int k0 = (int)(kc * K_CHUNK);
          if (elect_one()) {
            if (CL == 1) {
              mbar_expect_tx(&ctl->full[stage], STAGE_BYTES);
              tma_load_3d(sA, &mapA, &ctl->full[stage], k0, rowA, 0);
              tma_load_3d(sA + A_STAGE_BYTES, &mapB, &ctl->full[stage], k0, rowB, 0);
            } else {
              if (crank == 0) mbar_expect_tx(&ctl->full[stage], 2 * STAGE_BYTES);
              tma_load_3d_pair(sA, &mapA, &ctl->full[stage], k0, rowA, 0);
              // this CTA's half of "b": digit 1 + crank in full (the N = 192 operand is [digit 1 | digit 2],
              // one digit per CTA) and rows [48 crank, +48) of digit 0 (the N = 96 operand)
              tma_load_3d_pair(sA + A_STAGE_BYTES, &mapB96, &ctl->full[stage], k0, rowB, 1 + (int)crank);
              tma_load_3d_pair(sA + A_STAGE_BYTES + B_SLICE_BYTES, &mapB48, &ctl->full[stage], k0,
                               rowB + (int)crank * (COL_TILE_ROWS / 2), 0);
            }
          }
          __syncwarp();
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The WHOLE warp walks the tiles, waits on the barriers and forms the descriptors, so that all of
    // that is warp-uniform and lives in uniform registers; one elected lane issues.  (With the loop
    // inside `if (lane == 0)` every tcgen05.mma cost an ELECT and five R2UR.BROADCAST -- operands of
    // UTCIMMA are uniform registers -- about 16 instructions or ~83 cycles per MMA on this one thread,
    // more than the 72 cycles the average MMA of this kernel executes in: the issue thread, not the
    // tensor pipe, the shared-memory feed or the TMA ring, was what bound the MMA phase.)
    {
      constexpr uint32_t ID192 = make_idesc(2 * CLASS_COLS, 128 * CL), ID96 = make_idesc(CLASS_COLS, 128 * CL);
      uint32_t stage = 0, phase = 0, tphase = 0;
      long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long tprev = a.prof ? clock64() : 0;
      TileWalk<CL> walk(crank);
      uint32_t rt, I, J;
      bool masked;
      while (walk.template next<true>(a, ctl, rt, I, J, masked)) {
        if (CL > 1 && crank != 0) continue;  // only the leader of a pair issues; the other keeps the ring moving
        // the epilogue (of both CTAs) has drained the previous tile
        if (!SEG) {
          if (CL == 1) mbar_wait<32>(&ctl->tmem_empty, tphase ^ 1); else mbar_wait_cluster<32>(&ctl->tmem_empty, tphase ^ 1);
          tc_fence_after();
          LB_TICK(0);  // waiting for TMEM
        }
        uint32_t seg_left = 0;  // SEG: chunks left in the current accumulation segment
        for (uint32_t kc = 0; kc < a.n_kchunks; ++kc) {
          bool seg_first = kc == 0;
          if (SEG) {
            seg_first = seg_left == 0;
            if (seg_first) {  // the epilogue has drained the previous segment (or tile)
              mbar_wait<32>(&ctl->tmem_empty, tphase ^ 1);
              tc_fence_after();
              LB_TICK(0);
              seg_left = a.seg_chunks;
            }
            --seg_left;
          }
          mbar_wait<0>(&ctl->full[stage], phase);
          tc_fence_after();
          LB_TICK(kc == 0 ? 3 : (kc == 1 ? 4 : 1));  // waiting for operands (first, second, later chunks of the tile)
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + A_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < K_CHUNK / 32; ++ks) {
            const uint32_t first = (seg_first && ks == 0) ? 0u : 1u;
            const uint64_t A0 = make_desc(sA + 0 * A_SLICE_BYTES + ks * 32);
            const uint64_t A1 = make_desc(sA + 1 * A_SLICE_BYTES + ks * 32);
            const uint64_t A2 = make_desc(sA + 2 * A_SLICE_BYTES + ks * 32);
            // single CTA: digits 0 | 1 | 2 of "b" back to back, the N = 192 operand spans digits 1, 2.
            // pair: [this CTA's digit (96 rows) | its 48 rows of digit 0]; the hardware reads the
            // other half of N at the same offsets of the partner's shared memory.
            const uint64_t B0 = make_desc(sB + (CL == 1 ? 0 : B_SLICE_BYTES) + ks * 32);
            const uint64_t B1 = make_desc(sB + (CL == 1 ? B_SLICE_BYTES : 0) + ks * 32);
            if (elect_one()) {
              // class s lives at columns [96 s, 96 s + 96); first touch of a class overwrites
              tc_mma_i8<CL>(tmem_base + 3 * CLASS_COLS, A2, B1, ID192, first);  // a2.b1 -> s3, a2.b2 -> s4
              tc_mma_i8<CL>(tmem_base + 1 * CLASS_COLS, A0, B1, ID192, first);  // a0.b1 -> s1, a0.b2 -> s2
              tc_mma_i8<CL>(tmem_base + 0 * CLASS_COLS, A0, B0, ID96, first);   // a0.b0 -> s0
              tc_mma_i8<CL>(tmem_base + 2 * CLASS_COLS, A1, B1, ID192, 1u);     // a1.b1 -> s2, a1.b2 -> s3
              tc_mma_i8<CL>(tmem_base + 2 * CLASS_COLS, A2, B0, ID96, 1u);      // a2.b0 -> s2
              tc_mma_i8<CL>(tmem_base + 1 * CLASS_COLS, A1, B0, ID96, 1u);      // a1.b0 -> s1
            }
          }
          const bool seg_last = kc + 1 == a.n_kchunks || (SEG && seg_left == 0);
          if (elect_one()) {
            // smem slot reusable once these MMAs have read it (in both CTAs of a pair)
            if (CL == 1) tc_commit(&ctl->empty[stage]); else tc_commit_pair(&ctl->empty[stage]);
            if (seg_last) { if (CL == 1) tc_commit(&ctl->tmem_full); else tc_commit_pair(&ctl->tmem_full); }
          }
          __syncwarp();
          if (SEG && seg_last) tphase ^= 1;
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
          LB_TICK(2);  // issuing
        }
        if (!SEG) tphase ^= 1;
      }
      if (a.prof && lane == 0)
        for (int k = 0; k < 8; ++k) a.prof[((size_t)blockIdx.x * NWARPS + warp) * 8 + k] = pacc[k];
    }
  }
  } else {
    // ================= epilogue (warps 4..15) =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPILOGUE));
    const int q = warp & 3;                 // TMEM lane quarter == frame block of the row tile
    const int g = (warp - FIRST_EPI_WARP) >> 2;          // column group of the column tile: frames [11 g, 11 g + 11) (10 in the last)
    const int nfr = g < 2 ? GROUP_FRAMES : COL_TILE_FRAMES - 2 * GROUP_FRAMES;
    const int et = (warp - FIRST_EPI_WARP) * 32 + lane;  // 0..383
    const int jf = lane / 3, al = lane - 3 * jf;
    const bool active = lane < 30;
    const int a_local = q * FRAMES_PER_BLOCK + jf;
    const int tri = lane - al;              // first lane of this frame's lane triple
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 3 * GROUP_FRAMES);
    uint32_t tphase = 0, buf = 0;
    float sum_h = 0.0f, sum_l = 0.0f, tmax = 0.0f;
    const double inv_n = 1.0 / (double)a.N;
    const int release_at = a.release_at;  // hand TMEM back: 0 right after the drain, 1 after the fp64 coefficient phase
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = a.prof ? clock64() : 0;

    TileWalk<CL> walk(crank);
    uint32_t rt, I, J;
    bool masked;
    for (; walk.template next<true>(a, ctl, rt, I, J, masked); buf ^= 1) {
      const uint32_t i = I * TILE_FRAMES + a_local;
      const bool i_ok = active && i < a.FA && !masked;  // a shadow tile contributes nothing
      const double ga = i_ok ? __ldg(a.gA + i) : 0.0;
      // The eleven (ten) "b" frames of this warp's column group are solved in four rounds: the lanes of
      // a triple (al = 0, 1, 2) own frames base + al for base = 0, 3, 6, 9 (3, 3, 3, 2 frames; the last
      // group has no frame 10).  Rounds 0 and 1 only need the group's columns 0..17, so they run while
      // the loads of columns 18..32 are in flight: the latency-bound coefficient chains and the
      // issue-bound class sums fill each other's gaps.
      constexpr int RBASE[4] = {0, 3, 6, 9}, RCNT[4] = {3, 3, 3, 2};
      double gbq[4];  // G of the four "b" frames this lane solves against (prefetched)
#pragma unroll
      for (int round = 0; round < 4; ++round) {
        const uint32_t j = J * COL_TILE_FRAMES + g * GROUP_FRAMES + RBASE[round] + (RBASE[round] + al < nfr ? al : 0);
        gbq[round] = j < a.FB ? __ldg(a.gB + j) : 0.0;
      }
      // v[3 jb + axis'] = sum_s 2^(8s) D_s as a 64-bit integer (|v| < 2^58): one IMAD.WIDE per class and
      // element.  (The first version summed in fp64 through the exponent-bias trick: four instructions per
      // class and element -- constant high word, xor, two DADDs -- a third of all the epilogue issued.)
      long long v[3 * GROUP_FRAMES];
#pragma unroll
      for (int c = 0; c < 3 * GROUP_FRAMES; ++c) v[c] = 0;
      const int w8[4] = {a.w256 >> 8, a.w256, a.w256 << 8, a.w256 << 16};  // 1, 2^8, 2^16, 2^24
      QcpPolyF polf[4];
      float e0s[4];
      int state[4];  // 0: no entry, 1: zero, 2: solve
      bool in_stats[4] = {true, true, true, true};

      auto combine = [&](const uint32_t (&r)[18], int s, int c0, int count) {
#pragma unroll
        for (int c = 0; c < 18; ++c)
          if (c < count) {
            // (the weights come from a kernel argument: with an immediate power of two ptxas splits the
            // multiply-add into shifts, high products and carry chains)
            if (s < 4) asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(v[c0 + c]) : "r"((int)r[c]), "r"(w8[s]));
            else asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tadd.s32 hi, hi, %1;\n\tmov.b64 %0, {lo, hi};\n\t}"
                     : "+l"(v[c0 + c]) : "r"((int)r[c]));  // + D_4 2^32: the high word alone
          }
      };
      // 3-way selection by a lane-dependent index on the 32-bit halves (two SEL each; left to the compiler
      // a double-valued ?: chain became divergent branches)
      auto pick = [](bool k0, bool k1, long long x0, long long x1, long long x2) {
        const int lo = k0 ? (int)x0 : (k1 ? (int)x1 : (int)x2);
        const int hi = k0 ? (int)(x0 >> 32) : (k1 ? (int)(x1 >> 32) : (int)(x2 >> 32));
        return ((long long)hi << 32) | (unsigned int)lo;
      };
      // exchange rows inside each lane triple and build the quartic of this lane's pair (fp64)
      auto do_round = [&](const int round) {
        const int cb = 3 * RBASE[round];  // first column of the round's frames
        long long own[3], s1[3], s2[3];
        double R[9];
        const int d1 = (al + 2) % 3, d2 = (al + 1) % 3;  // destination axis served at shift 1, 2
#pragma unroll
        for (int be = 0; be < 3; ++be) {
          const long long x0 = v[cb + be], x1 = v[cb + 3 + be];
          const long long x2 = RCNT[round] == 3 ? v[cb + 6 + be] : x0;  // two-frame rounds: lane al = 2 idles
          own[be] = pick(al == 0, al == 1, x0, x1, x2);
          s1[be] = pick(d1 == 0, d1 == 1, x0, x1, x2);
          s2[be] = pick(d2 == 0, d2 == 1, x0, x1, x2);
        }
        const int src1 = tri + (al + 1) % 3, src2 = tri + (al + 2) % 3;
#pragma unroll
        for (int be = 0; be < 3; ++be) {
          // rows of R end up as (al, al+1, al+2) mod 3: a cyclic relabelling of the axes of
          // frame a is a proper rotation and leaves lambda_max unchanged
          R[be] = (double)own[be];  // one rounding at 2^-53, only for |v| >= 2^53
          R[3 + be] = (double)__shfl_sync(0xffffffffu, s1[be], src1);
          R[6 + be] = (double)__shfl_sync(0xffffffffu, s2[be], src2);
        }
        const int jb = RBASE[round] + al;
        const uint32_t j = J * COL_TILE_FRAMES + g * GROUP_FRAMES + jb;
        const bool owner = active && al < RCNT[round] && jb < nfr;
        if (KIND == KIND_CANON) {
          // A rectangle over one frame set (the matrix text computed in bands of rows, capi.cu): the
          // reference's matrix is symmetric by construction (Tools/rmsds.cpp:359-365), so the entry
          // above the diagonal must be the very float its mirror image below it is.  The quartic's
          // coefficients are not bit-invariant under R -> R^T or under the cyclic relabelling of the
          // rows, so (i, j > i) is evaluated on exactly the matrix the lane that owns (j, i) in the
          // triangle walk would hold: R of the swapped pair, rows relabelled by THAT lane's offset
          // (the position of frame i inside its column group, modulo 3).
          const uint32_t ib = i % COL_TILE_FRAMES;
          const uint32_t gm = ib / GROUP_FRAMES < 2 ? ib / GROUP_FRAMES : 2;
          const int alm = (int)((ib - gm * GROUP_FRAMES) % 3);
          auto sel3 = [](int k, double x0, double x1, double x2) { return k == 0 ? x0 : (k == 1 ? x1 : x2); };
          double T[9], Mr[9];
#pragma unroll
          for (int x = 0; x < 3; ++x)  // undo this lane's relabelling: T = the covariance of (i, j) in x, y, z order
#pragma unroll
            for (int y = 0; y < 3; ++y) T[3 * x + y] = sel3((x - al + 3) % 3, R[y], R[3 + y], R[6 + y]);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Mr[3 * r + c] = sel3((alm + r) % 3, T[3 * c + 0], T[3 * c + 1], T[3 * c + 2]);
          const bool flip = j > i;
#pragma unroll
          for (int k = 0; k < 9; ++k) R[k] = flip ? Mr[k] : R[k];
        }
        // 0: no entry (idle lane, out of range, or j > i of the self matrix), 1: exact zero (the
        // diagonal, Tools/rmsds.cpp:359-365, or every atom at the centroid in both frames), 2: solve.
        // Idle lanes run the same arithmetic on whatever their registers hold; their results are
        // never stored and they are excluded from the convergence vote below.
        const bool in_range = owner && i_ok && j < a.FB;
        const bool upper = a.self && j >= i;
        if (KIND == KIND_ROT || KIND == KIND_TRACE) {
          if (in_range) {
            const size_t cell = (size_t)j * a.ldq + (size_t)rt * TILE_FRAMES + (size_t)a_local;
            if (KIND == KIND_ROT) {
              double qv[4];
              if (!qcp_rotation_fast(R, ga, gbq[round], nullptr, nullptr, qv)) qcp_rotation(R, nullptr, nullptr, qv);
              double2* dst = reinterpret_cast<double2*>(a.quat + 4 * cell);
              dst[0] = make_double2(qv[0], qv[1]);
              dst[1] = make_double2(qv[2], qv[3]);
            } else {
              const double2* src = reinterpret_cast<const double2*>(a.quat + 4 * cell);
              const double2 q01 = src[0], q23 = src[1];
              const double qv[4] = {q01.x, q01.y, q23.x, q23.y};
              double z[9];
              qcp_quat_to_rot(qv, z);
              double tr = 0.0;  // sum_k (Z x'_k) . y'_k = sum_ab Z_ab C_ba
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) tr = fma(z[3 * r + c], R[3 * c + r], tr);
              const double ss = (ga + gbq[round]) - 2.0 * tr;
              a.out_d[cell] = a.unit * sqrt(fmax(ss, 0.0) * inv_n);
            }
          }
          return;
        }
        double e0 = 0.5 * (ga + gbq[round]);
        const bool pos = e0 > 0.0;
        e0 = pos ? e0 : 1.0;
        state[round] = !in_range ? 0 : (upper ? (j == i ? 1 : 0) : (pos ? 2 : 1));
        if (KIND == KIND_CANON) {
          // the rectangle covers the diagonal and both triangles of ONE set: exact zeros on the diagonal
          // (Tools/rmsds.cpp:359-365 never computes it), and only j < i counts in the statistics, so that
          // the row ranges of several devices add up to showStatsHalf's sum over the triangle
          if (in_range && j == i) state[round] = 1;
          in_stats[round] = j < i;
        }
        e0s[round] = (float)(2.0 * e0 * inv_n);  // msd = y * this
        polf[round] = qcp_poly_split(qcp_poly(R, e0));
      };

      // ---- drain, software-pipelined: columns 0..17 of the five classes first (x16 + x2), then 18..32 ----
      // (The x16 load of the second half reads one column past the group; for the last group that is
      // the next class's first columns -- loaded, never used.)
      // On sm_100 CUDA-core fp64 shares the tensor pipe with tcgen05.mma and gets ~10 % of its
      // throughput while MMAs run (tools/ubench_contend.cu), so the accumulators are handed back
      // only AFTER the fp64-heavy part; the fp32 Newton recurrences, the final correction and the
      // stores then overlap the next tile's MMAs.
      uint32_t rb[2][18];
      auto load_a = [&](uint32_t col, uint32_t (&r)[18]) {
        tc_ld16(col, reinterpret_cast<uint32_t(&)[16]>(r));
        tc_ld2(col + 16, r[16], r[17]);
      };
      if (SEG) {
        // all accumulation segments but the last: drain into the 64-bit sums, hand TMEM back, nothing else
        const uint32_t nseg = (a.n_kchunks + a.seg_chunks - 1) / a.seg_chunks;
        for (uint32_t sg = 0; sg + 1 < nseg; ++sg) {
          mbar_wait<32>(&ctl->tmem_full, tphase);
          tphase ^= 1;
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int s5 = 0; s5 < NUM_CLASSES; ++s5) {
              uint32_t r[18];
              r[16] = 0u; r[17] = 0u;
              if (half == 0) load_a(tlane + (uint32_t)(s5 * CLASS_COLS), r);
              else tc_ld16(tlane + (uint32_t)(s5 * CLASS_COLS + 18), reinterpret_cast<uint32_t(&)[16]>(r));
              tc_wait_ld18(r);
              combine(r, s5, half ? 18 : 0, half ? 15 : 18);
            }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_tmem<CL>(ctl, crank);
        }
      }
      mbar_wait<32>(&ctl->tmem_full, tphase);
      tphase ^= 1;
      tc_fence_after();
      LB_TICK(0);  // waiting for the accumulators
      load_a(tlane, rb[0]);
#pragma unroll
      for (int s5 = 0; s5 < NUM_CLASSES; ++s5) {
        tc_wait_ld18(rb[s5 & 1]);
        if (s5 + 1 < NUM_CLASSES) load_a(tlane + (uint32_t)((s5 + 1) * CLASS_COLS), rb[(s5 + 1) & 1]);
        else tc_ld16(tlane + 18u, reinterpret_cast<uint32_t(&)[16]>(rb[(s5 + 1) & 1]));
        combine(rb[s5 & 1], s5, 0, 18);
      }
      // columns 18..32: class 0 is in flight in rb[1]
      do_round(0);
#pragma unroll
      for (int s5 = 0; s5 < NUM_CLASSES; ++s5) {
        tc_wait_ld18(rb[(s5 + 1) & 1]);
        if (s5 + 1 < NUM_CLASSES) tc_ld16(tlane + (uint32_t)((s5 + 1) * CLASS_COLS + 18), reinterpret_cast<uint32_t(&)[16]>(rb[s5 & 1]));
        combine(rb[(s5 + 1) & 1], s5, 18, 15);
#if LB_INDRAIN >= 2
        if (s5 == 1) do_round(1);
#endif
      }
#if LB_INDRAIN >= 3
      do_round(2);
#endif
      tc_fence_before();  // this warp's TMEM reads are complete
      if (release_at == 0) { __syncwarp(); if (lane == 0) release_tmem<CL>(ctl, crank); }
      LB_TICK(1);  // drain (+ the rounds inside it)
#if LB_INDRAIN < 2
      do_round(1);
#endif
#if LB_INDRAIN < 3
      do_round(2);
#endif
      do_round(3);
      if (KIND == KIND_RMSD || KIND == KIND_CANON) {
      // pin every fp64-derived value on this side of the hand-over: without this the compiler
      // is free to sink the fp64 -> float conversions below it, into the contended region
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4)
        asm volatile("" : "+f"(polf[r4].c0h), "+f"(polf[r4].c0l), "+f"(polf[r4].c1h), "+f"(polf[r4].c1l),
                          "+f"(polf[r4].c2h), "+f"(polf[r4].c2l), "+f"(e0s[r4]), "+f"(polf[r4].t0));
      }
      if (release_at == 1) { __syncwarp(); if (lane == 0) release_tmem<CL>(ctl, crank); }
      if (KIND == KIND_RMSD || KIND == KIND_CANON) {
      LB_TICK(2);  // exchange + quartic coefficients (fp64)

      // ---- four fp32 Newton recurrences in lock-step, warp-uniform exit ----
      float tt[4], ls[4];
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) { tt[r4] = polf[r4].t0; ls[r4] = 0.0f; }
      // A step is applied only while it still matters (|step| > 4e-7 |t|): a settled solve stands still
      // while the warp iterates on for slower ones, so the result of a pair does not depend on which
      // other pairs share its warp (the same pair is met again in other launch shapes, bands and as
      // its mirror image).  The step not taken is second order after the correction below.
      for (int it = 0; it < 12; ++it) {
        bool conv = true;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          float t_new = tt[r4];
          ls[r4] = qcp_newton_step(polf[r4].c0h, polf[r4].c1h, polf[r4].c2h, polf[r4].c3, t_new);
          const bool c = fabsf(ls[r4]) <= 4e-7f * fabsf(tt[r4]) + 1e-12f;
          if (!c) tt[r4] = t_new;
          conv = conv && (c || state[r4] != 2);
        }
        if (__all_sync(0xffffffffu, conv)) break;
      }
      LB_TICK(3);  // Newton (fp32)

      // ---- one float-float residual correction per pair, float32 result ----
#pragma unroll
      for (int round = 0; round < 4; ++round) {
        const int jb = RBASE[round] + al;
        const int b_local = g * GROUP_FRAMES + jb;
        const bool owner = active && al < RCNT[round] && jb < nfr;
        float val = state[round] == 1 ? 0.0f : -1.0f;  // -1: "no entry"
        if (state[round] == 2) {
          bool ok;
          // no fp64 here in the common case: this code overlaps the next tile's MMAs
          float y = qcp_polish_f32(polf[round], polf[round].xform, tt[round], ls[round], &ok);
          if (!ok || polf[round].xform || !(fabsf(polf[round].c0h) < 1e30f)) {
            // rare (dissimilar or mirror-image structures, multiple / far-off root, non-finite
            // input): the same quartic in fp64, see qcp_root_fp64
            y = (float)qcp_root_fp64(polf[round], ok ? tt[round] : polf[round].t0);
          }
          val = (float)a.unit * sqrtf(y * e0s[round]);
          // running sum of the float32 results as a float-float pair (error-free two-sum): the
          // reference accumulates in double (Tools/rmsds.cpp:425-439); no fp64 in this loop
          if (KIND != KIND_CANON || in_stats[round]) {
            const float s2 = sum_h + val, bb2 = s2 - sum_h;
            sum_l += (sum_h - (s2 - bb2)) + (val - bb2);
            sum_h = s2;
            tmax = fmaxf(tmax, val);
          }
        }
        if (owner) ctl->T[buf][b_local][a_local] = val;
      }
      LB_TICK(4);  // polish
      asm volatile("bar.sync 1, 384;" ::: "memory");  // tile T[buf] complete
      LB_TICK(5);  // barrier
      if (a.out) {
        const size_t i0 = (size_t)I * TILE_FRAMES, j0 = (size_t)J * COL_TILE_FRAMES;
        if (a.packed) {
          // out[(rt*40 + a) * ld + j]: rows of the lower triangle, row-major
          for (int idx = et; idx < TILE_FRAMES * COL_TILE_FRAMES; idx += NUM_EPI_THREADS) {
            const int al_ = idx / COL_TILE_FRAMES, bl = idx - al_ * COL_TILE_FRAMES;
            const float val = ctl->T[buf][bl][al_];
            if (val >= 0.0f) a.out[((size_t)rt * TILE_FRAMES + al_) * a.ld + j0 + bl] = val;
          }
        } else {
          for (int idx = et; idx < TILE_FRAMES * COL_TILE_FRAMES; idx += NUM_EPI_THREADS) {
            const int bl = idx / TILE_FRAMES, al_ = idx - bl * TILE_FRAMES;
            const float val = ctl->T[buf][bl][al_];
            if (val >= 0.0f) a.out[(j0 + bl) * a.ld + (i0 - a.row_off) + al_] = val;  // M(i,j), column-major
          }
          if (a.symmetric) {
            for (int idx = et; idx < TILE_FRAMES * COL_TILE_FRAMES; idx += NUM_EPI_THREADS) {
              const int al_ = idx / COL_TILE_FRAMES, bl = idx - al_ * COL_TILE_FRAMES;
              const float val = ctl->T[buf][bl][al_];
              if (val >= 0.0f) {  // M(j,i): inside the band's own rows it lives in the row strip too
                const size_t j = j0 + bl;
                float* dst = j >= a.row_off ? a.out + (i0 + al_) * a.ld + (j - a.row_off)
                                            : a.out_t + (i0 - a.row_off + al_) * a.ld_t + j;
                *dst = val;
              }
            }
          }
        }
      }
      }
      LB_TICK(6);  // stores
      // no second barrier: the next tile writes T[buf ^ 1]; T[buf] is rewritten two tiles
      // later, after every thread has passed the next tile's barrier
    }
    if (a.prof && lane == 0)
      for (int k = 0; k < 8; ++k) a.prof[((size_t)blockIdx.x * NWARPS + warp) * 8 + k] = pacc[k];
    if (a.stats) {
      double ssum = (double)sum_h + (double)sum_l, smax = (double)tmax;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, o));
      }
      if (lane == 0) { ctl->red_sum[warp - FIRST_EPI_WARP] = ssum; ctl->red_max[warp - FIRST_EPI_WARP] = smax; }
      asm volatile("bar.sync 1, 384;" ::: "memory");
      if (et == 0) {
        double sm = 0.0, m = 0.0;
        for (int w = 0; w < NUM_EPI_WARPS; ++w) { sm += ctl->red_sum[w]; m = fmax(m, ctl->red_max[w]); }
        atomicAdd(&a.stats[0], sm);
        atomicMax(reinterpret_cast<unsigned long long*>(&a.stats[1]), (unsigned long long)__double_as_longlong(m));
      }
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // no CTA leaves while its partner may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (CL == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace

bool tensor_path_available() { return true; }

// Tile order: super-blocks of SB_R row tiles x SB_C column tiles so that the ~150
// tiles in flight share a few MB of operands in L2 (operand sets are larger than L2).
int launch_pairs_tensor(const PairJob& job, cudaStream_t st, int* launches) {
  // A super-block is SB_R row tiles x SB_C column tiles.  The "a" operands of its row group
  // (SB_R x 393 KB at N = 1000) stay L2-resident while the column super-blocks stream past, so every
  // row group reads all "b" operands left of the diagonal once from HBM: DRAM reads ~ F^2 N 3 / (2 SB_R 40).
  // 12 rows cost 105 GB per C4 launch, 48 rows a quarter of that (LOOSB200_SB_R / LOOSB200_SB_C override).
  uint32_t SB_R = 48, SB_C = 16;
  if (const char* e = getenv("LOOSB200_SB_R")) SB_R = (uint32_t)std::max(1, atoi(e));
  if (const char* e = getenv("LOOSB200_SB_C")) SB_C = (uint32_t)std::max(1, atoi(e));
  const char* cl_env = getenv("LOOSB200_CLUSTER");
  // 2 (default): CTA pairs (tcgen05 cta_group::2) sharing the "b" operand -- a quarter less shared-memory
  // traffic per MAC, which is what bounds the MMA phase of independent CTAs (572 against 432 cycles per
  // K step), and 21 % fewer operand bytes out of L2.  An fp64 instruction on either SM stalls the pair's
  // MMAs, so the accumulators are held through the whole fp64 part (release_at = 1) and the leader waits
  // for the slower of two epilogues before every tile.  1: independent CTAs.  Same box, after the MMA issue
  // became warp-uniform: C4 376 -> 362 ms, 20,000 x 3,000 38.6 -> 36.0 ms, 20,000 x 5,000 62 -> 56.8 ms
  // (profiles/r2_phase_counters.txt; before that change the issue thread hid the difference).
  // Mirror-canonical rectangles, rmsds-align passes and segmented selections run as independent CTAs.
  const uint32_t CLs = (job.kind != PairJob::Rmsd || job.canon || job.a->N > TENSOR_SEG_ATOMS) ? 1u
                       : cl_env ? (atoi(cl_env) == 1 ? 1u : 2u) : 2u;
  const loosb200_frames* A = job.a;
  const loosb200_frames* B = job.b;
  if (!A->quantised || !B->quantised || A->unit_log2 != B->unit_log2 || A->Kpad != B->Kpad)
    return LOOSB200_ERR_INVALID;
  const uint32_t FB = (uint32_t)B->F;
  const uint32_t ncol = (FB + COL_TILE_FRAMES - 1) / COL_TILE_FRAMES;
  std::vector<uint4> sbl;
  std::vector<uint32_t> prefix(1, 0u);
  uint64_t slots = 0;
  for (uint32_t r0 = 0; r0 < job.n_row_tiles; r0 += SB_R) {
    const uint32_t r1 = std::min(job.n_row_tiles, r0 + SB_R);
    // columns needed by the last (largest) row tile of the group
    uint32_t cmax = ncol;
    if (job.self) {
      const uint64_t last_i = (uint64_t)job.rows_host[r1 - 1] * TILE_FRAMES + TILE_FRAMES - 1;
      cmax = std::min<uint32_t>(ncol, (uint32_t)(last_i / COL_TILE_FRAMES) + 1);
    }
    for (uint32_t c0 = 0; c0 < cmax; c0 += SB_C) {
      const uint32_t nc = std::min(cmax, c0 + SB_C) - c0;
      sbl.push_back(make_uint4(r0, r1 - r0, c0, nc));
      slots += (uint64_t)((r1 - r0 + CLs - 1) / CLs) * nc;  // pair: two rows per slot
      prefix.push_back((uint32_t)slots);
    }
  }
  if (slots == 0) return LOOSB200_OK;
  if (slots > 0xfffffff0ull) return LOOSB200_ERR_UNSUPPORTED;

  uint4* d_sbl = nullptr;
  uint32_t* d_prefix = nullptr;
  // scratch from the buffer cache; the caller releases it after synchronising the stream
  if (!job.scratch) return LOOSB200_ERR_INVALID;
  LB_CUDA(dev_alloc_t(&d_sbl, sizeof(uint4) * sbl.size()));
  job.scratch->push_back(d_sbl);
  LB_CUDA(dev_alloc_t(&d_prefix, sizeof(uint32_t) * prefix.size()));
  job.scratch->push_back(d_prefix);
  uint32_t* d_counter = nullptr;
  LB_CUDA(dev_alloc_t(&d_counter, sizeof(uint32_t)));
  job.scratch->push_back(d_counter);
  LB_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), st));
  LB_CUDA(cudaMemcpyAsync(d_sbl, sbl.data(), sizeof(uint4) * sbl.size(), cudaMemcpyHostToDevice, st));
  LB_CUDA(cudaMemcpyAsync(d_prefix, prefix.data(), sizeof(uint32_t) * prefix.size(), cudaMemcpyHostToDevice, st));
  // pageable sources: the copies have been staged by the time the calls return

  TensorArgs ta;
  ta.sblocks = d_sbl; ta.sb_prefix = d_prefix; ta.n_slots = (uint32_t)slots; ta.counter = d_counter;
  ta.row_tiles = job.row_tiles_dev;
  ta.gA = A->gq; ta.gB = B->gq;
  ta.FA = (uint32_t)A->F; ta.FB = FB; ta.N = (uint32_t)A->N;
  ta.n_kchunks = (uint32_t)(A->Kpad / K_CHUNK);
  const bool seg = A->N > TENSOR_SEG_ATOMS;
  ta.seg_chunks = seg ? (uint32_t)(TENSOR_SEG_ATOMS / K_CHUNK) : ta.n_kchunks;  // 335 chunks = 42,880 atoms
  ta.self = job.self; ta.symmetric = job.symmetric; ta.packed = job.packed_rows;
  ta.canon = (job.canon && !job.self) ? 1 : 0;
  ta.unit = exp2((double)A->unit_log2);
  ta.out = job.out; ta.ld = job.ld; ta.stats = job.stats_dev;
  ta.out_t = job.out_t; ta.ld_t = job.ld_t; ta.row_off = (uint32_t)job.row_off;
  ta.prof = nullptr;
  ta.w256 = 256;
  ta.quat = job.quat; ta.out_d = job.out_d; ta.ldq = job.ldq;
  if (job.kind != PairJob::Rmsd && (!job.quat || job.self || !job.packed_rows || (job.kind == PairJob::Trace && !job.out_d)))
    return LOOSB200_ERR_INVALID;
  // The accumulators are handed back right after the drain (0): rounds 2 and 3 of the coefficient
  // phase then run next to the MMAs (fp64 crawls there, but the tensor pipe is not left idle).
  // LOOSB200_RELEASE_AT=1 holds them until the whole fp64 phase is done: 3 % slower for independent
  // CTAs, but what the CTA-pair shape needs (fp64 on either SM stalls the pair's MMAs: 462 ms against
  // 411 ms at C4).
  ta.release_at = getenv("LOOSB200_RELEASE_AT") ? (atoi(getenv("LOOSB200_RELEASE_AT")) ? 1 : 0) : (CLs == 2 ? 1 : 0);
  const bool want_prof = getenv("LOOSB200_PROF") != nullptr;

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)sms / CLs, slots) * CLs;
  if (want_prof) {
    LB_CUDA(dev_alloc_t(&ta.prof, sizeof(long long) * grid * NWARPS * 8));
    LB_CUDA(cudaMemsetAsync(ta.prof, 0, sizeof(long long) * grid * NWARPS * 8, st));
  }
  if (job.ev_k0) cudaEventRecord(job.ev_k0, st);
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TensorArgs);
  Kern kern;
  if (seg) kern = job.kind == PairJob::Rotation ? k_pairs_tensor<1, KIND_ROT, true>
                  : job.kind == PairJob::Trace  ? k_pairs_tensor<1, KIND_TRACE, true>
                  : ta.canon                    ? k_pairs_tensor<1, KIND_CANON, true>
                                                : k_pairs_tensor<1, KIND_RMSD, true>;
  else if (job.kind == PairJob::Rotation) kern = k_pairs_tensor<1, KIND_ROT>;
  else if (job.kind == PairJob::Trace) kern = k_pairs_tensor<1, KIND_TRACE>;
  else if (CLs == 2) kern = k_pairs_tensor<2, KIND_RMSD>;
  else if (ta.canon) kern = k_pairs_tensor<1, KIND_CANON>;
  else kern = k_pairs_tensor<1, KIND_RMSD>;
  const int smem = CLs == 2 ? smem_bytes<2>() : smem_bytes<1>();
  LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = CLs == 2 ? 1 : 0;
  LB_CUDA(cudaLaunchKernelEx(&cfg, kern, A->tmapA, B->tmapB, B->tmapB96, B->tmapB48, ta));
  if (job.ev_k1) cudaEventRecord(job.ev_k1, st);
  LB_CUDA(cudaGetLastError());
  if (want_prof) {
    std::vector<long long> hp((size_t)grid * NWARPS * 8);
    LB_CUDA(cudaMemcpyAsync(hp.data(), ta.prof, sizeof(long long) * hp.size(), cudaMemcpyDeviceToHost, st));
    LB_CUDA(cudaStreamSynchronize(st));
    dev_release(ta.prof);
    const char* names_e[7] = {"wait_acc", "drain", "exch+poly", "newton", "polish", "barrier", "stores"};
    const char* names_m[5] = {"wait_tmem", "wait_operands(chunks 2+)", "issue", "wait_operands(chunk 0)", "wait_operands(chunk 1)"};
    const double ntl = (double)job.n_tiles * TILE_FRAMES / COL_TILE_FRAMES / grid;  // approximate
    fprintf(stderr, "[loosb200 prof] tiles/CTA %.1f  cycles per tile (mean over CTAs):\n  MMA warp:", ntl);
    for (int k = 0; k < 5; ++k) { double s = 0; for (unsigned c = 0; c < grid; ++c) s += hp[((size_t)c * NWARPS + 1) * 8 + k]; fprintf(stderr, " %s=%.0f", names_m[k], s / grid / ntl); }
    fprintf(stderr, "\n  epilogue warps:");
    for (int k = 0; k < 7; ++k) { double s = 0; for (unsigned c = 0; c < grid; ++c) for (int w = FIRST_EPI_WARP; w < NWARPS; ++w) s += hp[((size_t)c * NWARPS + w) * 8 + k]; fprintf(stderr, " %s=%.0f", names_e[k], s / grid / 12 / ntl); }
    fprintf(stderr, "\n");
  }
  if (launches) *launches = 1;
  return LOOSB200_OK;
}

}  // namespace loosb200

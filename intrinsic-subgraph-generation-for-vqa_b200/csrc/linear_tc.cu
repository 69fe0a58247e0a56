// linear_tc.cu — kernel (d), modes 1 and 2: the dense projections on the 5th-gen tensor cores.
//
//   mode 1  3xTF32 split (fp32-grade):  D = A_hi*B_hi  (+)  [A_hi*B_lo + A_lo*B_hi]
//   mode 2  single-pass TF32 (~2^-11 per product; reported separately, not the parity configuration)
//
// Numerics of mode 1.  kind::tf32 ignores the low 13 mantissa bits of its fp32 operands, so the raw
// tile already is `hi`; `lo = x - hi` is exact.  The tensor core TRUNCATES (rounds toward zero) when it
// adds an MMA result into the TMEM accumulator — measured on B200: -0.5 ulp per accumulation step,
// error growing linearly with the reduction length (scripts/diag_tc_rounding.py).  Two measures keep the
// result at fp32-FFMA accuracy: (i) the two correction products go to a SEPARATE accumulator, so only
// one MMA per k-step touches the main chain; (ii) the main accumulator is drained every DRAIN_KB
// k-blocks (K = 128): the epilogue warps add the partial into fp32 registers with round-to-nearest and
// the next chunk restarts from zero in the other TMEM stage.  Chains are 16 steps long instead of K/8.
//
// One persistent, warp-specialised kernel (one CTA per SM, 512 threads) covers the three products of a
// Linear layer (reference call sites: models/mgat_v2_conv.py:177,181,259, models/mgat.py:156,
// models/masking.py:137,152 and their autograd backward):
//   fwd    y[m,n]  = sum_k x[m,k]  W[n,k]    A K-major,  B K-major
//   dgrad  gx[m,k] = sum_n gy[m,n] W[n,k]    A K-major,  B MN-major (W read in place, no transpose)
//   wgrad  gW[n,k] = sum_m gy[m,n] x[m,k]    A MN-major, B MN-major, deterministic split over m
//
//   warp 0       TMA producer: cp.async.bulk.tensor.2d of raw fp32 operand tiles (BK = 16 floats per
//                k-block: 64-byte swizzle for K-major, 128B/32B-atom swizzle for MN-major operands).
//                CTAs run as CLUSTERS OF TWO on adjacent M tiles of the same N block: each CTA fetches
//                its own A tile and HALF of the shared B tile, multicast into both CTAs' shared memory
//                (the mainloop is bound by L2->SM operand traffic: 16 KB per k-block per CTA alone,
//                12 KB as a pair); smem stages are released by both CTAs' MMA commits (multicast arrive)
//   warps 4-7    splitters (mode 1): thread m owns row m of the A tile — it gathers the row's 16 k-values
//                from the swizzled smem tile and writes A_hi (raw) and A_lo = x - (x & ~0x1fff) into a
//                4-slot A ring IN TENSOR MEMORY (tcgen05.st), so the MMAs take A from TMEM and only B
//                from shared memory (with SS operands all three products re-read A and B from smem and
//                the mainloop is shared-memory-bandwidth bound: 96 KB per k-block; TS form: 64 KB);
//                B_lo is written to a twin smem buffer at the same swizzled offsets; fence, arrive
//   warp 1       one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN<=128, K=8) into
//                TMEM; tcgen05.commit releases smem stages and publishes accumulator chunks
//   warps 8-15   epilogue: tcgen05.ld 32x32b of the chunk partials into register accumulators (two
//                column halves x four lane quarters), at tile end + correction accumulator, then a padded
//                smem transpose and coalesced float4 rows with the fused bias / exact GELU /
//                pre-activation side output / GELU-derivative / accumulate
//   warp 2       TMEM allocation, 512 columns.  mode 1: main[2 chunk stages] 0..255, correction 256..383,
//                A ring (4 slots x {hi 16, lo 16} columns) 384..511; mode 2: main[2] only, SS operands
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "linear_tc.h"
#include "tc_ptx.cuh"

namespace {

using namespace isg;
using namespace isg_tc;

constexpr int BM = 128;      // UMMA M (cta_group::1)
// Default ("wide") build: 32-wide k-blocks (128-byte K-major rows) with FOUR A slots in tensor memory.  To make room
// the two correction products accumulate into the main accumulator (no separate correction accumulator: 48 instead
// of 16 truncating steps per drained chunk -> 1.4-2.0e-6 instead of 6.5e-7 against fp64, the FFMA kernel's range)
// and the epilogue staging tile is halved so that four 48 KiB stages fit.  Halves the synchronisation rounds per
// reduction element: fwd 0.188 -> 0.157, dgrad 0.204 -> 0.155, wgrad 0.215 -> 0.162 ms on [39809,300]x[300,1200].
// -DISG_TC_NARROW builds the previous kernel (16-wide k-blocks, separate correction accumulator, 6.5e-7).
#if !defined(ISG_TC_NARROW) && !defined(ISG_TC_BK)
#define ISG_TC_WIDE 1
#endif
#ifdef ISG_TC_WIDE
#define ISG_TC_BK 32
#define ISG_TC_MERGE 1
#else
#define ISG_TC_MERGE 0
#endif
#ifndef ISG_TC_BK
#define ISG_TC_BK 16
#endif
constexpr bool MERGE = ISG_TC_MERGE != 0;
// Diagnostics only (WRONG RESULTS): -DISG_TC_DIAG=<bits> removes parts of the splitter's per-k-block work to see what the
// mainloop waits for.  1: no B_lo split (and no proxy fence), 2: no A_lo tensor-memory stores, 4: no A_hi stores either.
#ifndef ISG_TC_DIAG
#define ISG_TC_DIAG 0
#endif
constexpr int BK = ISG_TC_BK;  // fp32 elements per k-block: 16 (64-byte K-major rows, SWIZZLE_64B) or 32 (128-byte, SWIZZLE_128B)
static_assert(BK == 16 || BK == 32, "BK must be 16 or 32");
constexpr int UMMA_K = 8;    // kind::tf32
constexpr int MAX_BN = 128;
#ifndef ISG_TC_DRAIN_K
#define ISG_TC_DRAIN_K 128
#endif
// k-blocks per accumulator chunk (K = 128): 16 truncating steps per chain with a separate correction accumulator,
// 48 with the merged one.  Measured with the merged accumulator: K = 64 -> 1.1e-6 but fwd 0.157 -> 0.184 ms,
// K = 32 -> 8e-7 and 0.207 ms; K = 256 -> 2.4-3.1e-6 and no faster (0.160 ms): the chunk stays at 128.
constexpr int DRAIN_KB = ISG_TC_DRAIN_K / BK;
constexpr int NTHREADS = 512;
constexpr int A_TILE_BYTES = BM * BK * 4;          // 8 KiB
constexpr int EPI_WARPS = 8;
constexpr int EPI_ROW_BYTES = 144;                 // 32 floats + 16 B pad (conflict-free float4 transpose)
constexpr int EPI_ROWS = MERGE ? 16 : 32;           // rows staged per pass by one epilogue warp
constexpr int EPI_BYTES = EPI_WARPS * EPI_ROWS * EPI_ROW_BYTES;
constexpr int BAR_BYTES = 512;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int MAX_STAGES = 8;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_MAIN = 0, TM_CORR = 256;     // column bases; main stage s adds 128*s
constexpr uint32_t TM_A = MERGE ? 256 : 384;       // mode 1: A ring, slot s at TM_A + 2*BK*s (hi BK cols, lo BK cols)
constexpr int TS_STAGES = (512 - (int)TM_A) / (2 * BK);  // A-ring slots == smem stages in mode 1

enum Epi { EPI_FWD = 0, EPI_DGRAD = 1, EPI_PLAIN = 2 };

struct TcArgs {
  float* C;
  int64_t ldc;
  int64_t rows;  // output rows
  int cols;      // output cols
  int64_t R;     // reduction length
  int BN;
  int stages;
  int stage_bytes;
  int m_tiles, n_tiles, splits;
  int64_t r_chunk;  // reduction elements per split (multiple of BK)
  int64_t c_split_stride;
  const float* bias;
  float* Z;
  int64_t ldz;
  const float* Zprev;
  int act;
  int accumulate;
};

// PTX wrappers: tc_ptx.cuh
// shared-memory matrix descriptors (sm_100 format: version 1)
//   K-major : SWIZZLE_64B (4): rows of 64 B (BK = 16 fp32 along K), 8-row groups 512 B apart (SBO).
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (1) — 32-byte swizzle chunks, the
//             pattern TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  A tile is a row of
//             32-wide MN chunks, each [BK k-rows][128 B]: chunk stride = LBO = BK*128 B, groups of
//             4 k-rows 512 B apart (SBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? (uint64_t)((BK * 128) >> 4) : 1ull;
  // K-major: 8-row groups of BK*4-byte rows; swizzle mode = row length (64B: code 4, 128B: code 2)
  const uint64_t sbo = mn_major ? (512 >> 4) : ((8 * BK * 4) >> 4);
  const uint64_t layout = mn_major ? 1ull : (BK == 16 ? 4ull : 2ull);
  return (uint64_t)((saddr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ float4 hi_part(float4 v) {
  return make_float4(__uint_as_float(__float_as_uint(v.x) & 0xffffe000u),
                     __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                     __uint_as_float(__float_as_uint(v.z) & 0xffffe000u),
                     __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
}
__device__ __forceinline__ float4 lo_part(float4 v) {
  const float4 h = hi_part(v);
  return make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// ------------------------------------------------------------------------------ the kernel
// SPLIT (mode 1 vs mode 2), PAIR (cluster of two sharing the B tile by multicast) and FULLBN (BN == 128) are
// compile-time: the per-k-block role loops are latency chains in which every runtime conditional costs ~1 %
// (measured by bisecting an instrumented build), so the hot instantiation carries none.
// BLO: the lo plane of B comes pre-split from global memory (a second tensor map): the splitter warps then only
// move A into tensor memory — no shared-memory stores and no generic->async proxy fence on the per-k-block chain.
// P2 (mode 1, BN == 128; selected by launch()): the CTAs run as cta_group::2 MMA PAIRS on adjacent M tiles of one N block.
// The leader issues M = 256 instructions; each CTA stages its own A tile (and moves it to ITS tensor memory) and only
// HALF of the B tile (64 of the 128 rows / columns), which the pair's MMA reads from both shared memories.  Per CTA
// and 32-wide k-block the shared-memory traffic drops from 128 KB (TMA 32 + splitter 48 + MMA operand reads 48) to
// 80 KB (24 + 32 + 24) and the L2 -> SM operand traffic from 32 to 24 KB: 5-9 % faster on the E-sized products.
// DEEP (with P2): the shared-memory ring is decoupled from the four-slot A ring in tensor memory.  A pair's stage is
// only 32 KiB (A 16 + B_hi half 8 + B_lo half 8), so SIX stages fit where four 48 KiB ones did: the TMA producer runs
// up to six k-blocks ahead of the MMAs (its loop: commit -> free stage -> TMA latency -> full), the splitter up to four
// (its loop: commit -> free TMEM slot -> convert -> MMA).  Two commits per k-block instead of one.
template <bool A_MN, bool B_MN, int EPI, bool SPLIT, bool PAIR, bool FULLBN, bool BLO, bool P2 = false, bool DEEP = false>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_blo, const TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(!P2 || (SPLIT && MERGE && FULLBN && !PAIR && !BLO), "P2: mode 1, merged accumulator, BN = 128");
  static_assert(!DEEP || P2, "DEEP needs the half-sized B tiles of a pair");
  constexpr int TSLOTS = TS_STAGES;  // A-ring slots in tensor memory
  constexpr uint32_t CL = (PAIR || P2) ? 2u : 1u;  // plain launch, or a cluster pair sharing the B tile
  const uint32_t cta_rank = cluster_ctarank();

  const int S = DEEP ? 6 : (SPLIT ? TS_STAGES : g.stages);  // mode 1: smem ring == TMEM A ring (DEEP: six smem stages)
  const int BN = FULLBN ? MAX_BN : g.BN;
  const uint32_t b_tile_bytes = (uint32_t)BN * (BK * 4);
  // stage layout: mode 1 [A raw 8K][B BN*64][B_lo BN*64] (A goes on to TMEM); mode 2 [A 8K][B BN*64]
  const uint32_t off_b_hi = A_TILE_BYTES;
  const uint32_t off_b_lo = DEEP ? off_b_hi + b_tile_bytes / 2 : off_b_hi + b_tile_bytes;
  // stage size: a compile-time constant in the hot instantiation (A 8 KiB + B 8 KiB + B_lo 8 KiB)
  const uint32_t stage_bytes = DEEP ? (uint32_t)(A_TILE_BYTES + MAX_BN * BK * 4)
                             : (SPLIT && FULLBN) ? (uint32_t)(A_TILE_BYTES + 2 * MAX_BN * BK * 4) : (uint32_t)g.stage_bytes;
  const uint32_t epi_base = smem_base + (uint32_t)S * stage_bytes;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto mfull_bar = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + s); };
  auto mempty_bar = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + 2 + s); };
  auto cfull_bar = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + 4 + s); };
  auto cempty_bar = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + 6 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * MAX_STAGES + 8);
  auto tfree_bar = [&](int s) { return bar_base + 8u * (3 * MAX_STAGES + 10 + s); };  // DEEP: TMEM A slot s is free
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), P2 ? 8 : 128);  // P2: one arrival per splitter warp of BOTH CTAs, on the leader's barrier
      mbar_init(empty_bar(s), P2 ? 1 : CL);  // released by the MMA commits of every CTA of the cluster (P2: the leader's)
    }
    if (DEEP)
      for (int s = 0; s < TSLOTS; ++s) mbar_init(tfree_bar(s), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(mfull_bar(s), 1);
      mbar_init(mempty_bar(s), P2 ? 2 * EPI_WARPS : EPI_WARPS);
      mbar_init(cfull_bar(s), 1);
      mbar_init(cempty_bar(s), EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (P2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // peers' barriers must be initialised before any multicast / remote arrive
  else __syncthreads();
  tc_fence_after();
  // everything above is on-chip set-up and overlaps the previous kernel's tail; the operands are read from here on
  pdl_enter();
  const uint32_t tmem_base = *tmem_slot_gen;

  // Work items are (split, M-tile group of CL adjacent tiles, N block); CTA `cta_rank` of the cluster takes
  // M tile  group*CL + cta_rank  (possibly past the end: zero operands, nothing stored).
  const int m_groups = (g.m_tiles + (int)CL - 1) / (int)CL;
  const int tiles_mn = m_groups * g.n_tiles;
  const int total_tiles = tiles_mn * g.splits;
  const int work0 = (int)(blockIdx.x / CL), work_stride = (int)(gridDim.x / CL);
  auto tile_kb = [&](int t) {
    const int split = t / tiles_mn;
    const int64_t r_beg = (int64_t)split * g.r_chunk;
    const int64_t r_end = min(g.R, r_beg + g.r_chunk);
    return (int)((r_end - r_beg + BK - 1) / BK);
  };

  if (warp == 0) {
   if (elect_one()) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = work0; t < total_tiles; t += work_stride) {
      const int split = t / tiles_mn, rem = t - split * tiles_mn;
      const int m_grp = rem / g.n_tiles, n_blk = rem - m_grp * g.n_tiles;
      const int m_blk = m_grp * (int)CL + (int)cta_rank;
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int64_t r_beg = (int64_t)split * g.r_chunk;
      const int num_kb = tile_kb(t);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
        const uint32_t fb = full_bar(stage);
        mbar_arrive_expect_tx(fb, (uint32_t)A_TILE_BYTES + (P2 ? b_tile_bytes / 2 : (BLO ? 2u : 1u) * b_tile_bytes));
        const int r0 = (int)(r_beg + (int64_t)kb * BK);
        if (!A_MN) {
          tma_load_2d(sa, &map_a, fb, r0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), &map_a, fb, m0 + 32 * c, r0);
        }
        if (P2) {
          // this CTA's half of the B tile, at the START of the B region (the pair's MMA reads N/2 from each CTA)
          // (a ragged last N tile is multiplied at its own width, see n_eff in the MMA issuer: the halves split there)
          const int half = min(MAX_BN, ((g.cols - n0) + 15) & ~15) / 2;
          if (!B_MN) {
            tma_load_2d(sa + off_b_hi, &map_b, fb, r0, n0 + (int)cta_rank * half);
          } else {
#pragma unroll
            for (int c = 0; c < MAX_BN / 64; ++c)
              tma_load_2d(sa + off_b_hi + c * (BK * 128), &map_b, fb, n0 + (int)cta_rank * half + 32 * c, r0);
          }
        } else if (CL == 1) {
          if (!B_MN) {
            tma_load_2d(sa + off_b_hi, &map_b, fb, r0, n0);
            if (BLO) tma_load_2d(sa + off_b_lo, &map_blo, fb, r0, n0);
          } else {
            for (int c = 0; c < BN / 32; ++c)
              tma_load_2d(sa + off_b_hi + c * (BK * 128), &map_b, fb, n0 + 32 * c, r0);
            if (BLO)
              for (int c = 0; c < BN / 32; ++c)
                tma_load_2d(sa + off_b_lo + c * (BK * 128), &map_blo, fb, n0 + 32 * c, r0);
          }
        } else {
          // this CTA fetches half of the B tile and multicasts it into both CTAs of the pair
          if (!B_MN) {
            const int half_rows = BN / 2;  // the tensor map's box is BN/2 rows in cluster mode
            tma_load_2d_mc(sa + off_b_hi + cta_rank * (uint32_t)half_rows * (BK * 4), &map_b, fb, r0,
                           n0 + (int)cta_rank * half_rows, (uint16_t)0x3);
          } else {
            const int nch = BN / 32, h = nch / 2;
            for (int c = (int)cta_rank * h; c < ((int)cta_rank + 1) * h; ++c)
              tma_load_2d_mc(sa + off_b_hi + c * (BK * 128), &map_b, fb, n0 + 32 * c, r0, (uint16_t)0x3);
          }
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
   }
  } else if (warp == 1) {
   if ((!P2 || cta_rank == 0) && elect_one()) {
    // ===================================================================== MMA issuer (one elected lane; P2: leader CTA)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                           ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)((P2 ? 2 * BM : BM) >> 4) << 24);
    // TS form: the splitters have already laid A out row-per-lane / k-per-column, i.e. K-major
    const uint32_t idesc_ts = idesc & ~(1u << 15);
    const uint32_t a_kstep = A_MN ? (1024u >> 4) : ((UMMA_K * 4u) >> 4);  // descriptor units of 16 B
    const uint32_t b_kstep = B_MN ? (1024u >> 4) : ((UMMA_K * 4u) >> 4);
    const uint64_t a_hi0 = make_desc(smem_base, A_MN), b_hi0 = make_desc(smem_base + off_b_hi, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int ts = 0;  // TMEM A slot (== stage unless DEEP)
    uint32_t gchunk = 0;
    int it = 0;
    for (int t = work0; t < total_tiles; t += work_stride, ++it) {
      const int num_kb = tile_kb(t);
      // k-steps of the LAST k-block that hold reduction elements (TMA zero-fills the rest: K = 300 leaves 12 of 32,
      // i.e. two of four k-steps; products with an all-zero operand add +0 and are not issued)
      const int last_ksteps = [&] {
        const int split = t / tiles_mn;
        const int64_t r_beg = (int64_t)split * g.r_chunk;
        const int64_t r_len = min(g.R, r_beg + g.r_chunk) - r_beg;
        const int rem = (int)(r_len - (int64_t)(num_kb - 1) * BK);
        return (rem + UMMA_K - 1) / UMMA_K;
      }();
      const int tp = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      // A ragged last N tile (1200 = 9 x 128 + 48; 300 = 2 x 128 + 44) is multiplied at ITS width, rounded up to 16,
      // instead of the full 128: the tensor pipe's time per instruction is proportional to N (the shared-memory tile
      // and the accumulator keep their size; the epilogue ignores the columns past the edge anyway).
      const int n0_t = ((t % tiles_mn) % g.n_tiles) * BN;
      const int n_eff = (SPLIT && FULLBN) ? min(BN, ((g.cols - n0_t) + 15) & ~15) : BN;
      const uint32_t idesc_t = (idesc_ts & ~(0x3fu << 17)) | ((uint32_t)(n_eff >> 3) << 17);
      const uint32_t d_corr = tmem_base + TM_CORR;  // single stage: drained once per tile by the epilogue
      if (SPLIT && !MERGE) {
        mbar_wait(cempty_bar(0), (uint32_t)(it & 1) ^ 1u);
        tc_fence_after();
      }
      for (int kb0 = 0; kb0 < num_kb; kb0 += DRAIN_KB, ++gchunk) {
        const int ms = gchunk & 1;
        const uint32_t mphase = (gchunk >> 1) & 1u;
        if (P2) mbar_wait_cluster(mempty_bar(ms), mphase ^ 1u);
        else mbar_wait(mempty_bar(ms), mphase ^ 1u);
        tc_fence_after();
        const uint32_t d_main = tmem_base + TM_MAIN + 128u * ms;
        const int kb1 = min(num_kb, kb0 + DRAIN_KB);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (P2) mbar_wait_cluster(conv_bar(stage), phase);
          else mbar_wait(SPLIT ? conv_bar(stage) : full_bar(stage), phase);
          tc_fence_after();
          // descriptors of stage s = those of stage 0 + s * stage_bytes/16 (the 14-bit address field cannot
          // carry: all operand addresses are below 227 KiB)
          const uint32_t sdelta = (uint32_t)stage * (stage_bytes >> 4);
          const uint64_t b_hi = b_hi0 + sdelta, b_lo = b_hi + ((off_b_lo - off_b_hi) >> 4);
          const int ksteps = kb == num_kb - 1 ? last_ksteps : BK / UMMA_K;
          if (SPLIT) {
            const uint32_t a_t = tmem_base + TM_A + (uint32_t)(2 * BK) * (uint32_t)(DEEP ? ts : stage);  // hi at +0, lo at +BK
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k >= ksteps) break;
              const uint64_t bk = (uint64_t)(b_kstep * k);
              if (P2) {
                umma_tf32_ts_2cta(d_main, a_t + 8u * k, b_hi + bk, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
                umma_tf32_ts_2cta(d_main, a_t + 8u * k, b_lo + bk, idesc_t, 1u);
                umma_tf32_ts_2cta(d_main, a_t + (uint32_t)BK + 8u * k, b_hi + bk, idesc_t, 1u);
                continue;
              }
              umma_tf32_ts(d_main, a_t + 8u * k, b_hi + bk, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(MERGE ? d_main : d_corr, a_t + 8u * k, b_lo + bk, idesc_t,
                           (MERGE || kb > 0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(MERGE ? d_main : d_corr, a_t + (uint32_t)BK + 8u * k, b_hi + bk, idesc_t, 1u);
            }
          } else {
            const uint64_t a_hi = a_hi0 + sdelta;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k >= ksteps) break;
              const uint64_t ak = (uint64_t)(a_kstep * k), bk = (uint64_t)(b_kstep * k);
              umma_tf32(d_main, a_hi + ak, b_hi + bk, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          // frees the smem stage (in every CTA of the cluster: the peer's multicast writes land here too)
          if (P2) umma_commit_2cta_mc(empty_bar(stage), (uint16_t)0x3);
          else if (CL > 1) umma_commit_mc(empty_bar(stage), (uint16_t)0x3);
          else umma_commit(empty_bar(stage));
          if (DEEP) {
            umma_commit_2cta_mc(tfree_bar(ts), (uint16_t)0x3);
            if (++ts == TSLOTS) ts = 0;
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (P2) umma_commit_2cta_mc(mfull_bar(ms), (uint16_t)0x3);  // both CTAs' epilogues drain their 128 rows
        else umma_commit(mfull_bar(ms));  // chunk partial complete -> epilogue drains it
      }
      if (SPLIT && !MERGE) umma_commit(cfull_bar(0));
    }
   }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================================== lo-plane splitters (mode 1)
    // (A software-pipelined variant that prefetched the next k-block's operands while the TMEM stores drained
    // measured 5-9 % SLOWER; the straightforward per-k-block sequence below is the faster one.)
    if (SPLIT) {
      const int tid = threadIdx.x - 128;
      const int nb = (int)(b_tile_bytes / 16);  // <= 32*BK float4
      int stage = 0;
      uint32_t phase = 0;
      int ts = 0;
      uint32_t tphase = 0;
      for (int t = work0; t < total_tiles; t += work_stride) {
        const int num_kb = tile_kb(t);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (DEEP) {  // the MMAs that read TMEM slot ts four k-blocks ago have completed
            mbar_wait(tfree_bar(ts), tphase ^ 1u);
            tc_fence_after();
          }
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          // ---- A: row `tid` of the tile -> BK k-values -> TMEM (hi = raw, lo = exact remainder), 16 at a time
          const uint32_t a_t = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + TM_A +
                               (uint32_t)(2 * BK) * (uint32_t)(DEEP ? ts : stage);
#pragma unroll
          for (int hh = 0; hh < BK / 16; ++hh) {
            float av[16];
            if (!A_MN) {
              // K-major tile: rows of BK*4 bytes; 16-byte chunk c of row r sits at c ^ (r & 7) (SWIZZLE_128B, BK = 32)
              // or c ^ ((r >> 1) & 3) (SWIZZLE_64B, BK = 16)
              const uint32_t rowa = sa + (uint32_t)tid * (uint32_t)(BK * 4);
              const uint32_t x = BK == 16 ? (((uint32_t)tid >> 1) & 3u) : ((uint32_t)tid & 7u);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 v = lds_f4(rowa + ((((uint32_t)(4 * hh + c)) ^ x) << 4));
                av[4 * c] = v.x; av[4 * c + 1] = v.y; av[4 * c + 2] = v.z; av[4 * c + 3] = v.w;
              }
            } else {
              // MN-major tile: chunk (tid / 32) of [BK k-rows][128 B], 32-byte swizzle atoms:
              // element (ml, k) sits at k*128 + (((ml >> 3) ^ (k & 3)) << 5) + ((ml & 7) << 2)
              const uint32_t ml = (uint32_t)tid & 31u;
              const uint32_t ca = sa + ((uint32_t)tid >> 5) * (BK * 128u) + ((ml & 7u) << 2) + (uint32_t)hh * (16u * 128u);
#pragma unroll
              for (int k = 0; k < 16; ++k) av[k] = lds_f1(ca + (uint32_t)k * 128u + (((ml >> 3) ^ ((uint32_t)k & 3u)) << 5));
            }
            if (!(ISG_TC_DIAG & 4)) tmem_st16(a_t + 16u * hh, av);
            float al[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) al[k] = lo1(av[k]);
            if (!(ISG_TC_DIAG & 2)) tmem_st16(a_t + (uint32_t)BK + 16u * hh, al);
          }
          if (!BLO && !(ISG_TC_DIAG & 1)) {
#pragma unroll
            for (int u = 0; u < (P2 ? BK / 8 : BK / 4); ++u) {  // b_tile_bytes / 16 / 128 float4 per thread at BN = 128
              const int i = tid + 128 * u;                      // (P2: this CTA's half of the tile)
              if (FULLBN || i < nb) sts_f4(sa + off_b_lo + (uint32_t)i * 16u, lo_part(lds_f4(sa + off_b_hi + (uint32_t)i * 16u)));
            }
          }
          tmem_st_wait();
          tc_fence_before();
          if (!BLO && !(ISG_TC_DIAG & 1)) fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
          if (P2) {
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(conv_bar(stage), 0);  // the leader's barrier: 4 + 4 warp arrivals
          } else {
            mbar_arrive(conv_bar(stage));
          }
          if (DEEP && ++ts == TSLOTS) { ts = 0; tphase ^= 1u; }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================================================================== epilogue (8 warps)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 8) >> 2;  // column half: [64*half, 64*half + 64)
    const uint32_t stg = epi_base + (uint32_t)(warp - 8) * EPI_ROWS * EPI_ROW_BYTES;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t gchunk = 0;
    int it = 0;
    for (int t = work0; t < total_tiles; t += work_stride, ++it) {
      const int split = t / tiles_mn, rem = t - split * tiles_mn;
      const int m_grp = rem / g.n_tiles, n_blk = rem - m_grp * g.n_tiles;
      const int m_blk = m_grp * (int)CL + (int)cta_rank;
      const int64_t m0 = (int64_t)m_blk * BM + q * 32;
      const int n0 = n_blk * BN;
      const int n_lim = min(g.cols, n0 + BN);
      const int num_kb = tile_kb(t);
      const int tp = it & 1;
      const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
      float* Cb = g.C + (int64_t)split * g.c_split_stride;
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      for (int kb0 = 0; kb0 < num_kb; kb0 += DRAIN_KB, ++gchunk) {
        const int ms = gchunk & 1;
        const uint32_t mphase = (gchunk >> 1) & 1u;
        mbar_wait(mfull_bar(ms), mphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_sel + TM_MAIN + 128u * ms + 64u * half;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // one 32-column load at a time keeps the epilogue under 128 registers
          uint32_t r[32];
          tmem_ld32_nowait(taddr + 32u * hh, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[32 * hh + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // this warp is done with the chunk's TMEM stage (P2: tell the leader, whose MMAs write both)
          if (P2) mbar_arrive_remote(mempty_bar(ms), 0);
          else mbar_arrive(mempty_bar(ms));
        }
      }
      if (SPLIT && !MERGE) {
        mbar_wait(cfull_bar(0), (uint32_t)(it & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_sel + TM_CORR + 64u * half;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld32_nowait(taddr + 32u * hh, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[32 * hh + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(cempty_bar(0));
      }
      // ---- store: 2 x (32 rows x 32 columns) through the padded staging tile
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = 64 * half + 32 * cc;
        if (n0 + c0 >= n_lim) continue;  // warp-uniform
        const int col = n0 + c0 + (lane & 7) * 4;
#pragma unroll
        for (int rh = 0; rh < 32 / EPI_ROWS; ++rh) {  // EPI_ROWS rows of the warp's 32 per pass
        __syncwarp();
        if (EPI_ROWS == 32 || (lane >> 4) == rh) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_f4(stg + (lane & (EPI_ROWS - 1)) * EPI_ROW_BYTES + j * 16,
                   make_float4(acc[32 * cc + 4 * j], acc[32 * cc + 4 * j + 1], acc[32 * cc + 4 * j + 2],
                               acc[32 * cc + 4 * j + 3]));
        }
        __syncwarp();
        // all shared-memory reads of the pass first: with one register quad per row the read of row i + 1 would wait
        // for the store of row i
        float4 vs[EPI_ROWS / 4];
#pragma unroll
        for (int i = 0; i < EPI_ROWS / 4; ++i)
          vs[i] = lds_f4(stg + (i * 4 + (lane >> 3)) * EPI_ROW_BYTES + (lane & 7) * 16);
#pragma unroll
        for (int i = 0; i < EPI_ROWS / 4; ++i) {
          const int rl = i * 4 + (lane >> 3);
          const int64_t row = m0 + rh * EPI_ROWS + rl;
          float4 v = vs[i];
          if (row < g.rows && col < n_lim) {
            if (EPI == EPI_FWD) {
              if (g.bias) v = f4_add(v, Vec4<float>::ld(g.bias + col));
              if (g.Z) Vec4<float>::st(g.Z + row * g.ldz + col, v);
              if (g.act == ISG_ACT_GELU) v = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
            } else if (EPI == EPI_DGRAD) {
              if (g.Zprev) {
                const float4 z = Vec4<float>::ld(g.Zprev + row * g.ldz + col);
                v = make_float4(v.x * gelu_grad_f(z.x), v.y * gelu_grad_f(z.y), v.z * gelu_grad_f(z.z),
                                v.w * gelu_grad_f(z.w));
              }
              if (g.accumulate) v = f4_add(v, *reinterpret_cast<const float4*>(Cb + row * g.ldc + col));
            }
            Vec4<float>::st(Cb + row * g.ldc + col, v);
          }
        }
        }
      }
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it / arrive on it
  else __syncthreads();
  if (warp == 2) {
    if (P2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;  // resolved once; a function pointer, not device state
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2-D fp32 tensor map: dim0 (contiguous) x dim1 with row pitch `ld` elements; out-of-range elements
// read as zero (this is what pads K = 300 to the 16-wide k-blocks).
int make_map(CUtensorMap* m, const float* base, int64_t dim0, int64_t dim1, int64_t ld, int box0, int box1,
             bool mn_major) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return ISG_EUNSUPPORTED;
  // cuTensorMapEncodeTiled is a DRIVER call and needs a current context.  A thread that has only ever been handed
  // cached allocations (an autograd worker whose first CUDA action is this GEMM) has none bound yet and the encode
  // fails with CUDA_ERROR_INVALID_CONTEXT; one runtime call per thread binds the primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    if (cudaFree(nullptr) != cudaSuccess) return ISG_EINVAL;
    ctx_bound = true;
  }
  if (((uintptr_t)base & 15) || (ld % 4)) return ISG_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : (BK == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("ISG_TC_VERBOSE"))
    fprintf(stderr, "[isg] cuTensorMapEncodeTiled failed: CUresult %d base %p dims %llu x %llu pitch %llu B box %u x %u mn %d\n",
            (int)r, (const void*)base, (unsigned long long)dims[0], (unsigned long long)dims[1],
            (unsigned long long)strides[0], box[0], box[1], (int)mn_major);
  return r == CUDA_SUCCESS ? ISG_OK : ISG_EINVAL;
}

int pick_bn(int cols, bool mn_major) {
  const int q = mn_major ? 32 : 16;
  const int parts = (cols + MAX_BN - 1) / MAX_BN;
  int bn = (cols + parts - 1) / parts;
  bn = ((bn + q - 1) / q) * q;
  return bn > MAX_BN ? MAX_BN : bn;
}

template <bool A_MN, bool B_MN, int EPI>
int launch(const isg::TcGemm& p, cudaStream_t stream) {
  if (p.split3 != 0 && p.split3 != 1) return ISG_EUNSUPPORTED;
  TcArgs g{};
  g.C = p.C; g.ldc = p.ldc; g.rows = p.rows; g.cols = p.cols; g.R = p.R;
  g.BN = pick_bn(p.cols, B_MN);
  g.stage_bytes = A_TILE_BYTES + (p.split3 ? 2 : 1) * g.BN * BK * 4;
  g.stages = (SMEM_LIMIT - 1024 - EPI_BYTES - BAR_BYTES) / g.stage_bytes;
  if (g.stages > MAX_STAGES) g.stages = MAX_STAGES;
  if (p.split3) {  // one TMEM A slot per smem stage
    if (g.stages < TS_STAGES) return ISG_EUNSUPPORTED;
    g.stages = TS_STAGES;
  }
  if (g.stages < 2) return ISG_EUNSUPPORTED;
  g.m_tiles = ceil_div(p.rows, BM);
  g.n_tiles = ceil_div(p.cols, g.BN);
  g.splits = p.splits;
  g.r_chunk = p.r_chunk;
  g.c_split_stride = p.c_split_stride;
  g.bias = p.bias; g.Z = p.Z; g.ldz = p.ldz; g.Zprev = p.Zprev; g.act = p.act; g.accumulate = p.accumulate;
  if (p.rows >= (1ll << 31) || p.R >= (1ll << 31)) return ISG_EUNSUPPORTED;  // TMA coordinates are int32

  // cluster pairs whenever there are at least two M tiles to pair and the B tile splits evenly
  // Measured (B200, c3 step): the single-pass mode is bound by L2->SM operand traffic and gains 1.33x from the
  // pair (0.160 -> 0.120 ms on [39809,300]x[300,1200]); the 3xTF32 mode is bound by its splitter / MMA chain
  // and runs 1.7 % SLOWER as pairs (lock-step coupling), so it always launches single CTAs.
  static const bool env_no_cluster = getenv("ISG_TC_NO_CLUSTER") != nullptr;
  const bool want_pair = p.split3 == 0 && !p.no_cluster && !env_no_cluster;
  const bool pair = want_pair && g.m_tiles >= 2 && (B_MN ? ((g.BN / 32) % 2 == 0) : true);
  // mode 1 at BN == 128 (default build): cta_group::2 MMA pairs — see the kernel's header comment.  Measured on B200
  // (r2, scripts/gemm_probe.py, profiles/r2_gemm_probe.txt), bit-identical results: [39809,300]x[300,1200] fwd 145 ->
  // 133 us, dgrad 154 -> 142, wgrad 161 -> 153; node-sized fwd / dgrad 2-5 % faster, node-sized wgrad 20 % SLOWER (the
  // reduction split leaves the 74 pairs one short wave), so wgrad pairs only from 16 K reduction rows.  ISG_TC_PAIR2=0
  // keeps single CTAs everywhere.  (The first version waited / arrived on the cross-CTA barriers with
  // .acquire.cluster / .release.cluster and was 1.5-1.7x SLOWER than single CTAs: a cluster-scope acquire on every
  // per-k-block wait of the MMA thread.  The default-scope forms — what CUTLASS's cluster barriers use — are enough:
  // the tensor-memory and async-proxy hand-offs are ordered by the tcgen05 / proxy fences next to them.)
  static const bool env_no_pair2 = getenv("ISG_TC_PAIR2") != nullptr && atoi(getenv("ISG_TC_PAIR2")) == 0;
  const bool pair2 = MERGE && p.split3 == 1 && g.BN == MAX_BN && p.B_lo == nullptr && g.m_tiles >= 2 && !env_no_pair2 &&
                     !p.no_cluster && (!A_MN || p.R >= 16384);
  const int CLh = (pair || pair2) ? 2 : 1;

  CUtensorMap ma, mb;
  int rc;
  if (!A_MN) rc = make_map(&ma, p.A, p.R, p.rows, p.lda, BK, BM, false);
  else rc = make_map(&ma, p.A, p.rows, p.R, p.lda, 32, BK, true);
  if (rc != ISG_OK) return rc;
  if (!B_MN) rc = make_map(&mb, p.B, p.R, p.cols, p.ldb, BK, g.BN / CLh, false);
  else rc = make_map(&mb, p.B, p.cols, p.R, p.ldb, 32, BK, true);
  if (rc != ISG_OK) return rc;
  const bool blo = p.split3 && p.B_lo != nullptr;
  CUtensorMap mblo = mb;
  if (blo) {
    if (!B_MN) rc = make_map(&mblo, p.B_lo, p.R, p.cols, p.ldb, BK, g.BN, false);
    else rc = make_map(&mblo, p.B_lo, p.cols, p.R, p.ldb, 32, BK, true);
    if (rc != ISG_OK) return rc;
  }

  const int smem = 1024 + g.stages * g.stage_bytes + EPI_BYTES + BAR_BYTES;
  // instantiations: mode 1 single CTA (BN == 128 specialised, generic BN), mode 2 single CTA, mode 2 pair
  const bool fullbn = g.BN == MAX_BN;
  // (+ the pre-split-B variants of mode 1; wgrad's B is an activation and is always split in the kernel)
  constexpr bool CAN_BLO = !A_MN;
  constexpr bool CAN_P2 = MERGE;
  static const bool env_no_deep = getenv("ISG_TC_P2_DEEP") != nullptr && atoi(getenv("ISG_TC_P2_DEEP")) == 0;
  auto kern = (pair2 && !env_no_deep) ? tc_gemm_kernel<A_MN, B_MN, EPI, true, false, true, false, CAN_P2, CAN_P2>
            : pair2   ? tc_gemm_kernel<A_MN, B_MN, EPI, true, false, true, false, CAN_P2>
            : p.split3 ? (blo && CAN_BLO ? (fullbn ? tc_gemm_kernel<A_MN, B_MN, EPI, true, false, true, CAN_BLO>
                                                   : tc_gemm_kernel<A_MN, B_MN, EPI, true, false, false, CAN_BLO>)
                                         : (fullbn ? tc_gemm_kernel<A_MN, B_MN, EPI, true, false, true, false>
                                                   : tc_gemm_kernel<A_MN, B_MN, EPI, true, false, false, false>))
                       : (pair ? tc_gemm_kernel<A_MN, B_MN, EPI, false, true, false, false>
                               : tc_gemm_kernel<A_MN, B_MN, EPI, false, false, false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  const int m_groups = (g.m_tiles + CLh - 1) / CLh;
  const int total = m_groups * g.n_tiles * g.splits;
  int grid = total * CLh < ISG_NUM_SMS ? total * CLh : ISG_NUM_SMS;
  grid -= grid % CLh;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CLh;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  pdl_attribute(&attr[1]);
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (CLh > 1 && getenv("ISG_TC_VERBOSE")) {
    int nclusters = -1;
    cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    fprintf(stderr, "[isg] tc_gemm cluster launch: grid %d, cluster %d, smem %d, max active clusters %d (pair2 %d)\n", grid,
            CLh, smem, nclusters, (int)pair2);
  }
  e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mblo, g);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

}  // namespace

namespace isg {

int tc_gemm(const TcGemm& p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!p.a_mn && !p.b_mn && p.epi == 0) return launch<false, false, EPI_FWD>(p, stream);
  if (!p.a_mn && p.b_mn && p.epi == 0) return launch<false, true, EPI_FWD>(p, stream);  // fwd on a transposed weight
  if (!p.a_mn && p.b_mn && p.epi == 1) return launch<false, true, EPI_DGRAD>(p, stream);
  if (p.a_mn && p.b_mn && p.epi == 2) return launch<true, true, EPI_PLAIN>(p, stream);
  return ISG_EUNSUPPORTED;
}

int tc_wgrad_splits(int64_t M, int Nout, int K, int64_t* r_chunk) {
  // The reduction over rows is split for parallelism only (accuracy is handled by the in-kernel chunk drain);
  // partials are summed in a fixed order.  With t output tiles and s splits the persistent grid runs
  // ceil(t*s / SMs) waves of M/s rows each (+ a per-tile prologue/epilogue worth ~256 rows): pick the s that
  // minimises that, every split at least 32 k-blocks long.  (One wave of floor(SMs / t) splits — the first
  // heuristic — left 19-23 % of the SMs idle on the [E,1200]x[E,300] and [N,2400]x[N,300] gradients.)
  const int64_t tiles = (int64_t)ceil_div(Nout, BM) * ceil_div(K, pick_bn(K, true));
  int64_t max_s = (M + 32 * BK - 1) / (32 * BK);
  if (max_s > 64) max_s = 64;
  if (max_s < 1) max_s = 1;
  int64_t best_s = 1;
  double best_cost = 0.0;
  for (int64_t s = 1; s <= max_s; ++s) {
    const int64_t waves = (tiles * s + ISG_NUM_SMS - 1) / ISG_NUM_SMS;
    const double cost = (double)waves * ((double)M / (double)s + 256.0);
    if (s == 1 || cost < best_cost * 0.995) {  // prefer fewer splits on near-ties (fewer partials to reduce)
      best_cost = cost;
      best_s = s;
    }
  }
  int64_t chunk = (M + best_s - 1) / best_s;
  chunk = ((chunk + BK - 1) / BK) * BK;
  if (chunk < BK) chunk = BK;
  int64_t s = (M + chunk - 1) / chunk;  // every split non-empty
  if (s < 1) s = 1;
  *r_chunk = chunk;
  return (int)s;
}

}  // namespace isg

// linear_bf16.cu — kernel (d) for the bf16 configuration (BASELINE.json config 3, "fp32 vs bf16"): the dense
// projections with bf16 operands on the 5th-gen tensor cores (tcgen05.mma kind::f16, fp32 accumulation in tensor
// memory), ONE product instead of the three of the fp32-grade 3xTF32 split.  Reference call sites: lin_l / lin_r /
// lin_edge (models/mgat_v2_conv.py:177,181,259), x_proj (models/mgat.py:156) and their autograd backward; the
// reference itself never enters autocast (training/train_epoch.py:7 imports it only), so this is the configuration
// torch.autocast(bfloat16) WOULD give those Linear layers: bf16 inputs and weights, fp32 accumulate, bf16 outputs.
//
//   fwd    y[m,n]  = act(sum_k x[m,k] W[n,k] + b[n])     A = x   K-major,  B = W    K-major ([Nout, K] bf16 copy)
//   dgrad  gx[m,k] = sum_n gy[m,n] Wt[k,n] (* gelu'(z))  A = gy  K-major,  B = W^T  K-major ([K, Nout] bf16 copy)
//   wgrad  gW[n,k] = sum_m gy[m,n] x[m,k]                A = gy  MN-major, B = x    MN-major, deterministic split
//
// One persistent warp-specialised kernel, one CTA per SM, 384 threads:
//   warp 0      TMA producer (cp.async.bulk.tensor.2d, SWIZZLE_128B; 64-element = 128-byte rows per k-block;
//               out-of-range elements read as zero, which pads K = 300 and ragged tile edges)
//   warp 1      one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN <= 256, K = 16), A and B
//               from shared memory; tcgen05.commit releases the smem stage / publishes the accumulator
//   warp 2      TMEM allocation: 512 columns = two accumulator stages of 256 (tile i+1's mainloop overlaps tile i's
//               epilogue)
//   warps 4-11  epilogue: tcgen05.ld 32x32b.x32 (thread = row, 32 consecutive columns), fused bias / exact GELU /
//               pre-activation side output / GELU-derivative / accumulate, 16-byte stores straight from registers
//               (bf16 or fp32 output; every thread writes whole 32-byte sectors of its own row)
// These products are HBM-bound, not tensor-bound: [39809,300]x[300,1200] moves 24 + 96 MB for 28.7 GFLOP.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace isg;
using namespace isg_tc;

#ifndef ISG_BF16_DIAG
#define ISG_BF16_DIAG 0  // scripts/gemm_diag.sh: 1 = epilogue without its global stores, 2 = without the staging round trip too
#endif
constexpr int BM = 128;
constexpr int BK = 64;       // bf16 elements per k-block: 128-byte rows, SWIZZLE_128B
constexpr int UMMA_K = 16;   // kind::f16
constexpr int MAX_BN = 256;
constexpr int NTHREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB (K-major: 128 rows x 128 B; MN-major: 2 chunks of 64 x 64)
constexpr int CHUNK_BYTES = 64 * BK * 2;   // one MN-major chunk: 64 k-rows x 128 B
constexpr int BAR_BYTES = 256;
constexpr int EPI_ROW_BYTES = 144;                     // 128 B of output per row + 16 B pad (conflict-free)
constexpr int EPI_STG_BYTES = 32 * EPI_ROW_BYTES;      // per epilogue warp
constexpr int EPI_BYTES = EPI_WARPS * EPI_STG_BYTES;   // 36 KiB
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int MAX_STAGES = 6;
constexpr uint32_t TMEM_COLS = 512;

enum Epi { EPI_FWD = 0, EPI_DGRAD = 1, EPI_PLAIN = 2 };

struct BfArgs {
  void* C;
  int64_t ldc;
  int64_t rows;
  int cols;
  int64_t R;
  int BN;
  int stages;
  int stage_bytes;
  int res_kb;  // > 0 (K-major, no split): the CTA keeps its column block's B tile (res_kb k-blocks) in shared memory
  int m_tiles, n_tiles, splits;
  int64_t r_chunk;
  int64_t c_split_stride;
  const float* bias;
  void* Z;
  int64_t ldz;
  const void* Zprev;
  int act;
  int accumulate;
};

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// sm_100 shared-memory matrix descriptor, SWIZZLE_128B (layout code 2), 16-bit operands:
//   K-major : rows of 128 B (64 elements along K), 8-row groups 1024 B apart (SBO); LBO unused
//   MN-major: a row of 64-wide MN chunks, each [64 k-rows][128 B]; 8-k-row groups 1024 B apart (SBO), next MN
//             chunk CHUNK_BYTES further (LBO)  — ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units
__device__ __forceinline__ uint64_t make_desc16(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? (uint64_t)(CHUNK_BYTES >> 4) : 1ull;
  const uint64_t sbo = 1024 >> 4;
  return (uint64_t)((saddr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}

template <bool MN, int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(NTHREADS, 1)
bf16_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const BfArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = g.stages;
  const int BN = g.BN;
  const uint32_t stage_bytes = (uint32_t)g.stage_bytes;
  const uint32_t b_tile_bytes = stage_bytes - A_TILE_BYTES;
  // B-resident form (short reductions: K <= 5 k-blocks): [B tile, res_kb k-blocks][A-only ring of S stages]; the CTA
  // works on ONE column block and walks the row blocks, so the weight tile is read from L2 once per CTA instead of
  // once per output tile (the [39809,304]x[304,1200] product pulled 350 MB through L2 for 24 + 96 MB of HBM traffic)
  const int res_kb = MN ? 0 : g.res_kb;
  const uint32_t ring_base = smem_base + (uint32_t)res_kb * b_tile_bytes;
  const uint32_t ring_stride = res_kb ? (uint32_t)A_TILE_BYTES : stage_bytes;
  const uint32_t epi_base = ring_base + (uint32_t)S * ring_stride;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto mfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto mempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
  const uint32_t bres_bar = bar_base + 8u * (2 * MAX_STAGES + 5);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(mfull_bar(s), 1);
      mbar_init(mempty_bar(s), EPI_WARPS);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();  // on-chip set-up above overlaps the previous kernel's tail; operands are read from here on
  const uint32_t tmem_base = *tmem_slot_gen;

  const int tiles_mn = g.m_tiles * g.n_tiles;
  const int total_tiles = tiles_mn * g.splits;
  // tile walk: round-robin over all tiles, or (B-resident) the row blocks of this CTA's column block
  int t_first = blockIdx.x, t_step = gridDim.x;
  if (res_kb) {
    const int n_blk = (int)blockIdx.x % g.n_tiles, j = (int)blockIdx.x / g.n_tiles;
    const int group = ((int)gridDim.x - n_blk + g.n_tiles - 1) / g.n_tiles;  // CTAs that share this column block
    t_first = j * g.n_tiles + n_blk;
    t_step = group * g.n_tiles;
  }
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
      const int b_chunks = (BN + 63) / 64;
      if (res_kb && t_first < total_tiles) {
        const int n0 = (t_first % g.n_tiles) * BN;
        mbar_arrive_expect_tx(bres_bar, (uint32_t)res_kb * b_tile_bytes);
        for (int kb = 0; kb < res_kb; ++kb) tma_load_2d(smem_base + (uint32_t)kb * b_tile_bytes, &map_b, bres_bar, kb * BK, n0);
      }
      for (int t = t_first; t < total_tiles; t += t_step) {
        const int split = t / tiles_mn, rem = t - split * tiles_mn;
        const int m_blk = rem / g.n_tiles, n_blk = rem - m_blk * g.n_tiles;
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        const int64_t r_beg = (int64_t)split * g.r_chunk;
        const int num_kb = tile_kb(t);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = ring_base + (uint32_t)stage * ring_stride;
          const uint32_t fb = full_bar(stage);
          mbar_arrive_expect_tx(fb, (uint32_t)A_TILE_BYTES + (res_kb ? 0u : b_tile_bytes));
          const int r0 = (int)(r_beg + (int64_t)kb * BK);
          if (!MN) {
            tma_load_2d(sa, &map_a, fb, r0, m0);
            if (!res_kb) tma_load_2d(sa + A_TILE_BYTES, &map_b, fb, r0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * CHUNK_BYTES, &map_a, fb, m0 + 64 * c, r0);
            for (int c = 0; c < b_chunks; ++c)
              tma_load_2d(sa + A_TILE_BYTES + c * CHUNK_BYTES, &map_b, fb, n0 + 64 * c, r0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===================================================================== MMA issuer
      // instruction descriptor: D = f32 (1 << 4), A = B = bf16 (1 << 7, 1 << 10), majors, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((MN ? 1u : 0u) << 15) | ((MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t kstep = MN ? ((UMMA_K * 128u) >> 4) : ((UMMA_K * 2u) >> 4);  // 16-byte units per K = 16
      const uint64_t a0 = make_desc16(ring_base, MN), b0 = make_desc16(res_kb ? smem_base : smem_base + A_TILE_BYTES, MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (res_kb && t_first < total_tiles) mbar_wait(bres_bar, 0u);
      for (int t = t_first; t < total_tiles; t += t_step, ++it) {
        const int num_kb = tile_kb(t);
        const int ms = it & 1;
        const uint32_t mphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(mempty_bar(ms), mphase ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + 256u * ms;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sdelta = (uint32_t)stage * (ring_stride >> 4);
          const uint64_t a = a0 + sdelta, b = b0 + (res_kb ? (uint32_t)kb * (b_tile_bytes >> 4) : sdelta);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(d, a + (uint64_t)(kstep * k), b + (uint64_t)(kstep * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(mfull_bar(ms));
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (8 warps)
    // Thread = accumulator row.  A pass covers 128 bytes of output per row (64 bf16 or 32 fp32 columns): the values
    // get their epilogue math in registers, go through a padded per-warp staging tile and leave as whole 128-byte row
    // segments (8 lanes x 16 B per row, 4 rows per store instruction).  (The first version stored 16-byte pieces of
    // 32 different rows per instruction: 1.7 TB/s on the [39809,300]x[300,1200] product, a third of what HBM takes.)
    constexpr int PASS_COLS = OUT_BF16 ? 64 : 32;
    constexpr int PIECE = OUT_BF16 ? 8 : 4;  // columns per 16-byte piece
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // passes are dealt alternately to the two halves
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t stg = epi_base + (uint32_t)(warp - 4) * EPI_STG_BYTES;
    const int n_pass = (BN + PASS_COLS - 1) / PASS_COLS;
    int it = 0;
    for (int t = t_first; t < total_tiles; t += t_step, ++it) {
      const int split = t / tiles_mn, rem = t - split * tiles_mn;
      const int m_blk = rem / g.n_tiles, n_blk = rem - m_blk * g.n_tiles;
      const int64_t row0 = (int64_t)m_blk * BM + q * 32;
      const int64_t row = row0 + lane;
      const int n0 = n_blk * BN;
      // bf16 outputs are written in whole 8-column groups: a 300-wide output fills its pitch-304 row, pads = 0
      const int c_lim = g.cols;
      const int n_lim = min(OUT_BF16 ? ((g.cols + 7) & ~7) : g.cols, n0 + BN);
      const int ms = it & 1;
      const uint32_t mphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(mfull_bar(ms), mphase);
      tc_fence_after();
      // stage `v` (this thread's row, PASS_COLS columns from col0) and write it to `base` (row pitch ld) coalesced
      auto store_pass = [&](void* base, int64_t ld, int64_t extra, const float* v, int col0, bool accumulate) {
        if (ISG_BF16_DIAG & 2) return;
        __syncwarp();
        if (OUT_BF16) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_u4(stg + lane * EPI_ROW_BYTES + j * 16,
                   make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                              pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7])));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_u4(stg + lane * EPI_ROW_BYTES + j * 16,
                   make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                              __float_as_uint(v[4 * j + 3])));
        }
        __syncwarp();
        const int piece = lane & 7;
        const int col = col0 + piece * PIECE;
        // all eight shared-memory reads first, then the stores: with one register quad re-used per row the LDS of row
        // i + 1 waited for the STG of row i (ncu: those two instructions held 53 % of the epilogue warps' samples)
        uint4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = lds_u4(stg + (4 * i + (lane >> 3)) * EPI_ROW_BYTES + piece * 16);
        if (ISG_BF16_DIAG & 1) return;  // diagnostic build: no global stores (wrong results)
        if (col < n_lim) {
          const int64_t r_first = row0 + (lane >> 3);
          if (OUT_BF16) {
            __nv_bfloat16* dst = (__nv_bfloat16*)base + r_first * ld + col;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (r_first + 4 * i < g.rows) *reinterpret_cast<uint4*>(dst + (int64_t)(4 * i) * ld) = o[i];
          } else {
            float* dst = (float*)base + extra + r_first * ld + col;
            if (accumulate) {
              float4 c[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                c[i] = (r_first + 4 * i < g.rows) ? Vec4<float>::ld(dst + (int64_t)(4 * i) * ld) : f4_zero();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                o[i] = make_uint4(__float_as_uint(__uint_as_float(o[i].x) + c[i].x), __float_as_uint(__uint_as_float(o[i].y) + c[i].y),
                                  __float_as_uint(__uint_as_float(o[i].z) + c[i].z), __float_as_uint(__uint_as_float(o[i].w) + c[i].w));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (r_first + 4 * i < g.rows) *reinterpret_cast<uint4*>(dst + (int64_t)(4 * i) * ld) = o[i];
          }
        }
      };
      for (int c = half; c < n_pass; c += 2) {
        const int col0 = n0 + PASS_COLS * c;
        if (col0 >= n_lim) break;  // warp-uniform
        float v[PASS_COLS];
        {
          uint32_t r[32], r2[32];
          tmem_ld32_nowait(tmem_base + lane_sel + 256u * ms + (uint32_t)(PASS_COLS * c), r);
          if (OUT_BF16) tmem_ld32_nowait(tmem_base + lane_sel + 256u * ms + (uint32_t)(PASS_COLS * c) + 32u, r2);
          tmem_ld_wait();  // both loads in flight
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (OUT_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[(OUT_BF16 ? 32 : 0) + j] = __uint_as_float(r2[j]);
          }
        }
        if (EPI == EPI_FWD) {
          if (g.bias) {
#pragma unroll
            for (int j = 0; j < PASS_COLS; j += 4) {
              if (col0 + j < c_lim) {  // cols % 4 == 0
                const float4 b = Vec4<float>::ld(g.bias + col0 + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
          if (g.Z) store_pass(g.Z, g.ldz, 0, v, col0, false);
          if (g.act == ISG_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < PASS_COLS; ++j)  // (bf16: the backward differentiates at the STORED pre-activation)
              v[j] = gelu_f((OUT_BF16 && g.Z) ? __bfloat162float(__float2bfloat16_rn(v[j])) : v[j]);
          }
        } else if (EPI == EPI_DGRAD) {
          if (g.Zprev && row < g.rows) {
            if (OUT_BF16) {
              const __nv_bfloat16* zr = (const __nv_bfloat16*)g.Zprev + row * g.ldz + col0;
#pragma unroll
              for (int j = 0; j < PASS_COLS; j += 8) {
                if (col0 + j < n_lim) {
                  const uint4 z = *reinterpret_cast<const uint4*>(zr + j);
                  const float2 z0 = unpack_bf16(z.x), z1 = unpack_bf16(z.y), z2 = unpack_bf16(z.z), z3 = unpack_bf16(z.w);
                  v[j] *= gelu_grad_f(z0.x); v[j + 1] *= gelu_grad_f(z0.y); v[j + 2] *= gelu_grad_f(z1.x);
                  v[j + 3] *= gelu_grad_f(z1.y); v[j + 4] *= gelu_grad_f(z2.x); v[j + 5] *= gelu_grad_f(z2.y);
                  v[j + 6] *= gelu_grad_f(z3.x); v[j + 7] *= gelu_grad_f(z3.y);
                }
              }
            } else {
              const float* zr = (const float*)g.Zprev + row * g.ldz + col0;
#pragma unroll
              for (int j = 0; j < PASS_COLS; j += 4) {
                if (col0 + j < n_lim) {
                  const float4 z = Vec4<float>::ld(zr + j);
                  v[j] *= gelu_grad_f(z.x); v[j + 1] *= gelu_grad_f(z.y); v[j + 2] *= gelu_grad_f(z.z);
                  v[j + 3] *= gelu_grad_f(z.w);
                }
              }
            }
          }
        }
        if (OUT_BF16 && col0 + PASS_COLS > c_lim) {  // pad columns: exact zeros whatever z_prev holds there
#pragma unroll
          for (int j = 0; j < PASS_COLS; ++j)
            if (col0 + j >= c_lim) v[j] = 0.f;
        }
        store_pass(g.C, g.ldc, (int64_t)split * g.c_split_stride, v, col0, EPI == EPI_DGRAD && g.accumulate);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(mempty_bar(ms));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) x dim1, row pitch `ld` ELEMENTS (must be a multiple of 8 = 16 bytes)
int make_map16(CUtensorMap* m, const void* base, int64_t dim0, int64_t dim1, int64_t ld, int box0, int box1) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return ISG_EUNSUPPORTED;
  static thread_local bool ctx_bound = false;  // see linear_tc.cu: the driver call needs a bound context
  if (!ctx_bound) {
    if (cudaFree(nullptr) != cudaSuccess) return ISG_EINVAL;
    ctx_bound = true;
  }
  if (((uintptr_t)base & 15) || (ld % 8)) return ISG_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("ISG_TC_VERBOSE"))
    fprintf(stderr, "[isg] bf16 cuTensorMapEncodeTiled failed: CUresult %d base %p dims %llu x %llu pitch %llu B box %u x %u\n",
            (int)r, base, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides[0],
            box[0], box[1]);
  return r == CUDA_SUCCESS ? ISG_OK : ISG_EINVAL;
}

constexpr int RES_MAX_KB = 5;
// ISG_BF16_BRES=0: never keep the B tile resident; n >= 2: resident form with at least n A stages (default 4).
// Measured on [39809,304]x[304,1200] (ncu): 43.9 -> 40.0 us, L2 sectors 13.9 M -> 10.2 M; the product stays bound by its
// epilogue warps (84 % busy) and the shared-memory pipe, so the gain is small.
inline int bres_mode() {
  static const int v = getenv("ISG_BF16_BRES") ? atoi(getenv("ISG_BF16_BRES")) : 1;
  return v;
}

int pick_bn16(int cols) {
  const int parts = (cols + MAX_BN - 1) / MAX_BN;
  int bn = (cols + parts - 1) / parts;
  bn = ((bn + 15) / 16) * 16;
  return bn > MAX_BN ? MAX_BN : bn;
}

struct BfGemm {
  const void *A, *B;
  int64_t lda, ldb;
  void* C;
  int64_t ldc, rows;
  int cols;
  int64_t R;
  int splits;
  int64_t r_chunk, c_split_stride;
  const float* bias;
  void* Z;
  int64_t ldz;
  const void* Zprev;
  int act, accumulate;
};

template <bool MN, int EPI, bool OUT_BF16>
int launch16(const BfGemm& p, cudaStream_t stream) {
  BfArgs g{};
  g.C = p.C; g.ldc = p.ldc; g.rows = p.rows; g.cols = p.cols; g.R = p.R;
  g.BN = pick_bn16(p.cols);
  int b_bytes = MN ? ((g.BN + 63) / 64) * CHUNK_BYTES : g.BN * BK * 2;
  g.stage_bytes = A_TILE_BYTES + b_bytes;
  g.stages = (SMEM_LIMIT - 1024 - BAR_BYTES - EPI_BYTES) / g.stage_bytes;
  if (g.stages > MAX_STAGES) g.stages = MAX_STAGES;
  if (g.stages < 2) return ISG_EUNSUPPORTED;
  g.m_tiles = ceil_div(p.rows, BM);
  g.n_tiles = ceil_div(p.cols, g.BN);
  g.res_kb = 0;
  if (!MN && p.splits == 1 && bres_mode() != 0) {
    // B-resident form: the reduction fits in RES_MAX_KB k-blocks, at least `min_stages` A stages remain beside the
    // resident tile, and every CTA gets at least two row blocks of its column block (otherwise nothing is re-used)
    const int kbs = (int)ceil_div(p.R, (int64_t)BK);
    const int min_stages = bres_mode() > 1 ? bres_mode() : 4;
    if (kbs <= RES_MAX_KB) {
      for (int parts = (p.cols + MAX_BN - 1) / MAX_BN; parts <= 16; ++parts) {
        int bn = (p.cols + parts - 1) / parts;
        bn = ((bn + 15) / 16) * 16;
        const int bb = bn * BK * 2;
        int st = (SMEM_LIMIT - 1024 - BAR_BYTES - EPI_BYTES - kbs * bb) / A_TILE_BYTES;
        if (st < min_stages) continue;
        if (st > MAX_STAGES) st = MAX_STAGES;
        const int nt = ceil_div(p.cols, bn);
        if (nt <= ISG_NUM_SMS && (int64_t)g.m_tiles * nt >= 2ll * ISG_NUM_SMS) {
          g.BN = bn;
          b_bytes = bb;
          g.stage_bytes = A_TILE_BYTES + bb;
          g.stages = st;
          g.n_tiles = nt;
          g.res_kb = kbs;
        }
        break;
      }
    }
  }
  g.splits = p.splits; g.r_chunk = p.r_chunk; g.c_split_stride = p.c_split_stride;
  g.bias = p.bias; g.Z = p.Z; g.ldz = p.ldz; g.Zprev = p.Zprev; g.act = p.act; g.accumulate = p.accumulate;
  if (p.rows >= (1ll << 31) || p.R >= (1ll << 31)) return ISG_EUNSUPPORTED;
  CUtensorMap ma, mb;
  int rc;
  if (!MN) {
    if ((rc = make_map16(&ma, p.A, p.R, p.rows, p.lda, BK, BM)) != ISG_OK) return rc;
    if ((rc = make_map16(&mb, p.B, p.R, p.cols, p.ldb, BK, g.BN)) != ISG_OK) return rc;
  } else {
    if ((rc = make_map16(&ma, p.A, p.rows, p.R, p.lda, 64, BK)) != ISG_OK) return rc;
    if ((rc = make_map16(&mb, p.B, p.cols, p.R, p.ldb, 64, BK)) != ISG_OK) return rc;
  }
  const int smem = 1024 + (g.res_kb ? g.res_kb * b_bytes + g.stages * A_TILE_BYTES : g.stages * g.stage_bytes) + EPI_BYTES +
                   BAR_BYTES;
  auto kern = bf16_gemm_kernel<MN, EPI, OUT_BF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  const int total = g.m_tiles * g.n_tiles * g.splits;
  const int grid = total < ISG_NUM_SMS ? total : ISG_NUM_SMS;
  e = launch_pdl(kern, dim3((unsigned)grid), dim3(NTHREADS), (size_t)smem, stream, ma, mb, g);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

// reduction split of the bf16 wgrad (same cost model as tc_wgrad_splits in linear_tc.cu)
int wgrad_splits16(int64_t M, int Nout, int K, int64_t* r_chunk) {
  const int64_t tiles = (int64_t)ceil_div(Nout, BM) * ceil_div(K, pick_bn16(K));
  int64_t max_s = (M + 16 * BK - 1) / (16 * BK);
  if (max_s > 64) max_s = 64;
  if (max_s < 1) max_s = 1;
  int64_t best_s = 1;
  double best_cost = 0.0;
  for (int64_t s = 1; s <= max_s; ++s) {
    const int64_t waves = (tiles * s + ISG_NUM_SMS - 1) / ISG_NUM_SMS;
    const double cost = (double)waves * ((double)M / (double)s + 512.0);
    if (s == 1 || cost < best_cost * 0.995) {
      best_cost = cost;
      best_s = s;
    }
  }
  int64_t chunk = (M + best_s - 1) / best_s;
  chunk = ((chunk + BK - 1) / BK) * BK;
  if (chunk < BK) chunk = BK;
  int64_t s = (M + chunk - 1) / chunk;
  if (s < 1) s = 1;
  *r_chunk = chunk;
  return (int)s;
}

__global__ void split_reduce16_kernel(const float* __restrict__ part, int splits, int64_t stride, int64_t n,
                                      float* __restrict__ out) {
  pdl_enter();
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 s = f4_zero();
  for (int p = 0; p < splits; ++p) s = f4_add(s, Vec4<float>::ld(part + (int64_t)p * stride + i));
  Vec4<float>::st(out + i, s);
}

// ---- fp32 -> bf16 copies -------------------------------------------------------------------------------------
// rows x cols fp32 (pitch ld_in) -> bf16 (pitch ld_out >= cols, the pad columns are written as zero)
__global__ void to_bf16_kernel(const float* __restrict__ in, int64_t ld_in, int64_t rows, int cols,
                               __nv_bfloat16* __restrict__ out, int64_t ld_out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // one 8-column group per thread
  const int g8 = (int)(ld_out >> 3);
  const int64_t r = i / g8;
  const int c = (int)(i - r * g8) * 8;
  if (r >= rows) return;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? in[r * ld_in + c + j] : 0.f;
  *reinterpret_cast<uint4*>(out + r * ld_out + c) =
      make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// Weights of one step in one launch: job j converts W_j [n, k] fp32 into W_j bf16 [n, ldw] and W_j^T bf16 [k, ldt]
// (pad columns zero).  32x32 tiles through shared memory.
constexpr int WB_MAX_JOBS = 24;
struct WeightJob {
  const float* w;
  __nv_bfloat16 *wb, *wt;
  int n, k, ldw, ldt, tiles_k, tile0;
};
struct WeightBatch {
  WeightJob job[WB_MAX_JOBS];
  int n;
};
__global__ void __launch_bounds__(256) weights_to_bf16_kernel(const __grid_constant__ WeightBatch batch) {
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < batch.n && (int)blockIdx.x >= batch.job[j + 1].tile0) ++j;
  const WeightJob& J = batch.job[j];
  const int local = blockIdx.x - J.tile0;
  const int n0 = (local / J.tiles_k) * 32, k0 = (local % J.tiles_k) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    const float v = (n < J.n && k < J.k) ? J.w[(int64_t)n * J.k + k] : 0.f;
    tile[r][tx] = v;
    if (n < J.n && k < J.ldw) J.wb[(int64_t)n * J.ldw + k] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    if (k < J.k && n < J.ldt) J.wt[(int64_t)k * J.ldt + n] = __float2bfloat16_rn(tile[tx][r]);
  }
}

}  // namespace

// ------------------------------------------------------------------------------ C ABI
extern "C" int isg_to_bf16(const float* in, int64_t ld_in, int64_t rows, int cols, void* out, int64_t ld_out,
                           void* stream_) {
  if (rows < 0 || cols <= 0 || ld_out < cols || ld_out % 8) return ISG_EINVAL;
  if (rows == 0) return ISG_OK;
  if (!in || !out) return ISG_EINVAL;
  if ((uintptr_t)out & 15) return ISG_EUNSUPPORTED;
  const int64_t n = rows * (ld_out / 8);
  to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(in, ld_in, rows, cols,
                                                                                (__nv_bfloat16*)out, ld_out);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_weights_to_bf16(int n, const float* const* w, const int* rows, const int* cols, void* const* w_bf16,
                                   const int* ld_w, void* const* w_t_bf16, const int* ld_t, void* stream_) {
  if (n < 0 || n > WB_MAX_JOBS) return ISG_EINVAL;
  if (n == 0) return ISG_OK;
  if (!w || !rows || !cols || !w_bf16 || !ld_w || !w_t_bf16 || !ld_t) return ISG_EINVAL;
  WeightBatch batch;
  batch.n = n;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    if (!w[i] || !w_bf16[i] || !w_t_bf16[i] || rows[i] <= 0 || cols[i] <= 0) return ISG_EINVAL;
    if (ld_w[i] < cols[i] || ld_t[i] < rows[i] || ld_w[i] % 8 || ld_t[i] % 8) return ISG_EINVAL;
    WeightJob& J = batch.job[i];
    J.w = w[i];
    J.wb = (__nv_bfloat16*)w_bf16[i];
    J.wt = (__nv_bfloat16*)w_t_bf16[i];
    J.n = rows[i];
    J.k = cols[i];
    J.ldw = ld_w[i];
    J.ldt = ld_t[i];
    J.tiles_k = ceil_div(ld_w[i], 32);          // covers the pad columns of W
    const int tiles_n = ceil_div(ld_t[i], 32);  // covers the pad columns of W^T
    J.tile0 = tiles;
    tiles += J.tiles_k * tiles_n;
  }
  weights_to_bf16_kernel<<<tiles, 256, 0, (cudaStream_t)stream_>>>(batch);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_linear_bf16_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* y,
                                   int64_t ldy, void* z_pre, int64_t ldz, int64_t M, int Nout, int K, int act,
                                   int out_dtype, void* stream_) {
  if (M < 0 || Nout <= 0 || K <= 0) return ISG_EINVAL;
  if (M == 0) return ISG_OK;
  if (!x || !w || !y) return ISG_EINVAL;
  const bool ob = out_dtype == ISG_BF16;
  if (!ob && out_dtype != ISG_F32) return ISG_EUNSUPPORTED;
  const int q = ob ? 8 : 4;
  const int need_ld = ob ? ((Nout + 7) & ~7) : Nout;  // bf16 rows are written in whole 8-column groups
  if (Nout % 4 || ldy % q || ldy < need_ld || (z_pre && (ldz % q || ldz < need_ld)) || ((uintptr_t)y & 15) ||
      ((uintptr_t)z_pre & 15) || (bias && ((uintptr_t)bias & 15)))
    return ISG_EUNSUPPORTED;
  BfGemm t{};
  t.A = x; t.lda = ldx; t.B = w; t.ldb = ldw; t.C = y; t.ldc = ldy; t.rows = M; t.cols = Nout; t.R = K;
  t.splits = 1; t.r_chunk = ((int64_t)K + BK - 1) / BK * BK; t.bias = bias; t.Z = z_pre; t.ldz = ldz; t.act = act;
  return ob ? launch16<false, EPI_FWD, true>(t, (cudaStream_t)stream_)
            : launch16<false, EPI_FWD, false>(t, (cudaStream_t)stream_);
}

extern "C" int isg_linear_bf16_dgrad(const void* g_y, int64_t ldg, const void* w_t, int64_t ldwt, const void* z_prev,
                                     int64_t ldz, void* g_x, int64_t ldgx, int accumulate, int64_t M, int Nout, int K,
                                     int out_dtype, void* stream_) {
  if (M < 0 || Nout <= 0 || K <= 0) return ISG_EINVAL;
  if (M == 0) return ISG_OK;
  if (!g_y || !w_t || !g_x) return ISG_EINVAL;
  const bool ob = out_dtype == ISG_BF16;
  if (!ob && out_dtype != ISG_F32) return ISG_EUNSUPPORTED;
  if (ob && accumulate) return ISG_EUNSUPPORTED;
  const int q = ob ? 8 : 4;
  const int need_ld = ob ? ((K + 7) & ~7) : K;
  if (K % 4 || ldgx % q || ldgx < need_ld || (z_prev && (ldz % q || ldz < need_ld)) || ((uintptr_t)g_x & 15) ||
      ((uintptr_t)z_prev & 15))
    return ISG_EUNSUPPORTED;
  BfGemm t{};
  t.A = g_y; t.lda = ldg; t.B = w_t; t.ldb = ldwt; t.C = g_x; t.ldc = ldgx; t.rows = M; t.cols = K; t.R = Nout;
  t.splits = 1; t.r_chunk = ((int64_t)Nout + BK - 1) / BK * BK; t.Zprev = z_prev; t.ldz = ldz; t.accumulate = accumulate;
  return ob ? launch16<false, EPI_DGRAD, true>(t, (cudaStream_t)stream_)
            : launch16<false, EPI_DGRAD, false>(t, (cudaStream_t)stream_);
}

extern "C" size_t isg_linear_bf16_wgrad_workspace_bytes(int64_t M, int Nout, int K) {
  if (M <= 0 || Nout <= 0 || K <= 0) return 0;
  int64_t chunk = 0;
  const int s = wgrad_splits16(M > 0 ? M : 1, Nout, K, &chunk);
  return s > 1 ? (size_t)s * (size_t)Nout * (size_t)K * sizeof(float) : 0;
}

extern "C" int isg_linear_bf16_wgrad(const void* g_y, int64_t ldg, const void* x, int64_t ldx, float* g_w, int64_t M,
                                     int Nout, int K, void* workspace, size_t ws_bytes, void* stream_) {
  if (M < 0 || Nout <= 0 || K <= 0 || !g_w) return ISG_EINVAL;
  if (K % 4 || ((uintptr_t)g_w & 15)) return ISG_EUNSUPPORTED;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M == 0) {
    cudaError_t e = cudaMemsetAsync(g_w, 0, (size_t)Nout * K * sizeof(float), stream);
    return e == cudaSuccess ? ISG_OK : (int)e;
  }
  if (!g_y || !x) return ISG_EINVAL;
  const size_t need = isg_linear_bf16_wgrad_workspace_bytes(M, Nout, K);
  if (need > 0 && (ws_bytes < need || !workspace)) return ISG_EWORKSPACE;
  int64_t chunk = 0;
  const int ts = wgrad_splits16(M, Nout, K, &chunk);
  BfGemm t{};
  t.A = g_y; t.lda = ldg; t.B = x; t.ldb = ldx; t.rows = Nout; t.cols = K; t.R = M; t.splits = ts; t.r_chunk = chunk;
  if (ts > 1) { t.C = workspace; t.ldc = K; t.c_split_stride = (int64_t)Nout * K; }
  else { t.C = g_w; t.ldc = K; t.c_split_stride = 0; }
  const int rc = launch16<true, EPI_PLAIN, false>(t, stream);
  if (rc != ISG_OK) return rc;
  if (ts > 1) {
    const int64_t n = (int64_t)Nout * K;
    cudaError_t le = isg::launch_pdl(split_reduce16_kernel, dim3((unsigned)isg::ceil_div(n / 4, 256)), dim3(256), 0, stream,
                                     (const float*)workspace, ts, n, n, g_w);
    if (le != cudaSuccess) return (int)le;
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

// simple.cu — kernel (c), SIMPLE sampler: exact k-subset marginals + Gumbel top-k sample with the
// straight-through estimator, forward and backward.
//
// Reference: EdgeSIMPLEBatched.forward (sampling/methods/simple_scheme.py:44-162, policy 'edge_candid',
// ensemble 1, no logits activation) -> Layer.log_pr / Layer.sample (sampling/methods/simple.py:113-252)
// over the "exactly k of n" SDD built by create_simple_constraint.py:34-73.
//
// The SDD is a balanced binary tree: node (level L, position i, count j) decomposes into the pairs
// (left count jj, right count j - jj).  simple.py evaluates it level by level in log space
// (levelwiseSL bottom-up, levelwiseMars top-down) and pads every element list to `max_elements` and
// every parent list to `max_parents` with a dummy of log-value -1000.  Those pads are NOT neutral:
// to_dense_batch pads and dropped-out logits are exactly 0, whose negative-literal weight is
// log(1 - e^0) = -inf, and then the -2000 / -1000 pad terms decide the result.  The pad counts are
// therefore part of the function and are reproduced (plan built on the host by simple_plan()).
//
// One warp per graph; the whole circuit (2n(k+1) values, 2n(k+1)^2 conditionals) lives in shared
// memory.  The reference differentiates through both passes with autograd; the backward kernel
// recomputes the forward and runs the two reverse sweeps (gather form, no atomics).
// Latency-bound: ~8 B per node of HBM traffic.
#include "common.cuh"

namespace {

using namespace isg;

constexpr int SP_MAX_LEVELS = 12;
constexpr float SP_DUMMY = -1000.0f;       // simple.py:222
constexpr float SP_PAD_SCORE = -1.0e10f;   // simple_scheme.py:16,97-106

struct SimplePlan {
  int p, k, n, me, mp, kp1;
  int cap[SP_MAX_LEVELS];         // max count a node of level L can hold
  unsigned reach[SP_MAX_LEVELS];  // bitmask of counts reachable from the root at level L
  int off[SP_MAX_LEVELS];         // float offset of level L in the value arrays (unit: kp1 floats per node)
  int total_nodes;                // sum over levels of n >> L
};

__host__ int simple_plan(int n, int k, SimplePlan* pl) {
  int p = 0;
  while ((1 << p) < n) ++p;
  if ((1 << p) != n || p < 1 || p >= SP_MAX_LEVELS || k < 1 || k > 7 || k > n) return ISG_EUNSUPPORTED;
  pl->p = p; pl->k = k; pl->n = n; pl->kp1 = k + 1;
  pl->cap[0] = 1;
  for (int L = 1; L <= p; ++L) pl->cap[L] = k < (1 << L) ? k : (1 << L);
  for (int L = 0; L <= p; ++L) pl->reach[L] = 0;
  pl->reach[p] = 1u << k;
  for (int L = p; L >= 1; --L)
    for (int j = 0; j <= pl->cap[L]; ++j) {
      if (!(pl->reach[L] >> j & 1)) continue;
      for (int jj = 0; jj <= j; ++jj)
        if (jj <= pl->cap[L - 1] && j - jj <= pl->cap[L - 1]) pl->reach[L - 1] |= (1u << jj) | (1u << (j - jj));
    }
  int me = 0, mp = 0;
  for (int L = 1; L <= p; ++L)
    for (int j = 0; j <= pl->cap[L]; ++j) {
      if (!(pl->reach[L] >> j & 1)) continue;
      int cnt = 0;
      for (int jj = 0; jj <= j; ++jj) cnt += (jj <= pl->cap[L - 1] && j - jj <= pl->cap[L - 1]);
      me = cnt > me ? cnt : me;
    }
  for (int L = 0; L < p; ++L)
    for (int c = 0; c <= pl->cap[L]; ++c) {
      if (!(pl->reach[L] >> c & 1)) continue;
      int cnt = 0;
      for (int j = 0; j <= pl->cap[L + 1]; ++j)
        cnt += ((pl->reach[L + 1] >> j & 1) && j - c >= 0 && j - c <= pl->cap[L]);
      mp = cnt > mp ? cnt : mp;
    }
  pl->me = me; pl->mp = mp;
  int off = 0;
  for (int L = 0; L <= p; ++L) { pl->off[L] = off; off += n >> L; }
  pl->total_nodes = off;
  return ISG_OK;
}

__device__ __forceinline__ float log1mexp_f(float x) {  // simple.py:44-56: log(1 - exp(-|x|))
  x = -fabsf(x);
  return x > -0.6931471805599453094f ? logf(-expm1f(x)) : log1pf(-expf(x));
}

// torch.logsumexp over `cnt` values plus `pads` copies of `padval` (NaN-propagating max, +-inf max -> 0)
__device__ __forceinline__ float lse_pad(const float* v, int cnt, int pads, float padval) {
  float m = -INFINITY;
  for (int e = 0; e < cnt; ++e) m = (v[e] > m || v[e] != v[e]) ? v[e] : m;
  if (pads > 0 && padval > m) m = padval;
  const float mm = isinf(m) ? 0.f : m;
  float s = 0.f;
  for (int e = 0; e < cnt; ++e) s += expf(v[e] - mm);
  for (int e = 0; e < pads; ++e) s += expf(padval - mm);
  return logf(s) + mm;
}

struct Circuit {
  float* D;  // [total_nodes][kp1]            log-values (bottom-up); reused for gM in the backward
  float* M;  // [total_nodes][kp1]            log-marginals (top-down); reused for gD in the backward
  float* C;  // [total_nodes][kp1][kp1]       log-conditionals of the elements
};

// first valid left count of node (L, j): elements are jj in [lo, hi]
__device__ __forceinline__ void elem_range(const SimplePlan& pl, int L, int j, int& lo, int& hi) {
  const int cap = pl.cap[L - 1];
  lo = j - cap > 0 ? j - cap : 0;
  hi = j < cap ? j : cap;
}

// forward: fills D, C, M.  theta_dense value for slot i is produced by `leaf(i)`.
template <typename Leaf>
__device__ void circuit_forward(const SimplePlan& pl, const Circuit& c, int lane, Leaf leaf) {
  const int kp1 = pl.kp1, n = pl.n;
  for (int i = lane; i < n; i += 32) {
    const float th = leaf(i);
    c.D[(size_t)i * kp1 + 0] = log1mexp_f(th);
    c.D[(size_t)i * kp1 + 1] = th;
  }
  __syncwarp();
  for (int L = 1; L <= pl.p; ++L) {
    const int width = n >> L, nj = pl.cap[L] + 1;
    const float* prev = c.D + (size_t)pl.off[L - 1] * kp1;
    float* cur = c.D + (size_t)pl.off[L] * kp1;
    float* curC = c.C + (size_t)pl.off[L] * kp1 * kp1;
    for (int w = lane; w < width * nj; w += 32) {
      const int i = w / nj, j = w - i * nj;
      if (!(pl.reach[L] >> j & 1)) continue;
      int lo, hi;
      elem_range(pl, L, j, lo, hi);
      float t[8];
      const int cnt = hi - lo + 1;
      for (int e = 0; e < cnt; ++e)
        t[e] = prev[(size_t)(2 * i) * kp1 + lo + e] + prev[(size_t)(2 * i + 1) * kp1 + (j - lo - e)];
      const float val = lse_pad(t, cnt, pl.me - cnt, 2.f * SP_DUMMY);
      cur[(size_t)i * kp1 + j] = val;
      for (int e = 0; e < cnt; ++e) curC[((size_t)i * kp1 + j) * kp1 + e] = t[e] - val;
    }
    __syncwarp();
  }
  if (lane == 0) {
    const float root = c.D[(size_t)pl.off[pl.p] * kp1 + pl.k];
    c.M[(size_t)pl.off[pl.p] * kp1 + pl.k] = root - root;  // simple.py:229 (NaN when the root is +-inf)
  }
  __syncwarp();
  for (int L = pl.p - 1; L >= 0; --L) {
    const int width = n >> L, nc = pl.cap[L] + 1;
    float* cur = c.M + (size_t)pl.off[L] * kp1;
    const float* parM = c.M + (size_t)pl.off[L + 1] * kp1;
    const float* parC = c.C + (size_t)pl.off[L + 1] * kp1 * kp1;
    for (int w = lane; w < width * nc; w += 32) {
      const int i = w / nc, cc = w - i * nc;
      if (!(pl.reach[L] >> cc & 1)) continue;
      const int ip = i >> 1, side = i & 1;
      float t[8];
      int cnt = 0;
      for (int j = cc; j <= pl.cap[L + 1]; ++j) {
        if (!(pl.reach[L + 1] >> j & 1) || j - cc > pl.cap[L]) continue;
        int lo, hi;
        elem_range(pl, L + 1, j, lo, hi);
        const int e = (side == 0 ? cc : j - cc) - lo;
        t[cnt++] = parC[((size_t)ip * kp1 + j) * kp1 + e] + parM[(size_t)ip * kp1 + j];
      }
      cur[(size_t)i * kp1 + cc] = lse_pad(t, cnt, pl.mp - cnt, SP_DUMMY);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ size_t circuit_floats(const SimplePlan& pl) {
  return (size_t)pl.total_nodes * pl.kp1 * (2 + pl.kp1);
}

__global__ void simple_fwd_kernel(const float* __restrict__ theta, const float* __restrict__ gumbel,
                                  const int* __restrict__ gptr, int64_t B, int nmax, SimplePlan pl,
                                  int warps_per_cta, size_t floats_per_warp, float* __restrict__ mask,
                                  float* __restrict__ marg_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * warps_per_cta + warp;
  if (b >= B) return;
  float* base = smem + (size_t)warp * floats_per_warp;
  Circuit c;
  c.D = base;
  c.M = c.D + (size_t)pl.total_nodes * pl.kp1;
  c.C = c.M + (size_t)pl.total_nodes * pl.kp1;
  float* key = c.C + (size_t)pl.total_nodes * pl.kp1 * pl.kp1;  // [n] perturbed scores for the sample
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  auto leaf = [&](int i) -> float { return i < nb ? theta[n0 + i] : (i < nmax ? 0.f : SP_PAD_SCORE); };
  circuit_forward(pl, c, lane, leaf);
  // sample: one-hot of topk(theta_padded + Gumbel(0,1), k)  (simple.py:91-110, 246-252; no grad)
  for (int i = lane; i < pl.n; i += 32) key[i] = __fadd_rn(leaf(i), gumbel[b * pl.n + i]);
  __syncwarp();
  unsigned hot_lo = 0;  // lanes keep hot flags for their own slots: slot i -> lane i%32, bit i/32 ... up to 32*32 slots
  for (int r = 0; r < pl.k; ++r) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < pl.n; i += 32) {
      const float v = key[i];
      if (v > best || bi == 0x7fffffff) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(ISG_FULL_MASK, best, o);
      const int oi = __shfl_xor_sync(ISG_FULL_MASK, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((bi & 31) == lane) {
      hot_lo |= 1u << (bi >> 5);
      key[bi] = -INFINITY;
    }
    __syncwarp();
  }
  for (int i = lane; i < nmax; i += 32) {
    const float m = expf(c.M[(size_t)i * pl.kp1 + 1]);
    const float hot = (hot_lo >> (i >> 5) & 1u) ? 1.f : 0.f;
    if (i < nb) mask[n0 + i] = __fadd_rn(__fsub_rn(hot, m), m);  // (samples - marginals) + marginals
    if (marg_out) marg_out[b * nmax + i] = m;
  }
}

__global__ void simple_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dmarg,
                                  const float* __restrict__ theta, const int* __restrict__ gptr, int64_t B,
                                  int nmax, SimplePlan pl, int warps_per_cta, size_t floats_per_warp,
                                  float* __restrict__ g_theta) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * warps_per_cta + warp;
  if (b >= B) return;
  const int kp1 = pl.kp1, n = pl.n;
  float* base = smem + (size_t)warp * floats_per_warp;
  Circuit c;
  c.D = base;
  c.M = c.D + (size_t)pl.total_nodes * kp1;
  c.C = c.M + (size_t)pl.total_nodes * kp1;
  float* G = c.C + (size_t)pl.total_nodes * kp1 * kp1;  // [total_nodes][kp1][kp1]: gC, then gt
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  auto leaf = [&](int i) -> float { return i < nb ? theta[n0 + i] : (i < nmax ? 0.f : SP_PAD_SCORE); };
  circuit_forward(pl, c, lane, leaf);
  float* gM = c.D;  // D is dead after the forward (only C and M are needed below)
  // ---- adjoint of exp and of the top-down pass (children before parents)
  for (int i = lane; i < n; i += 32) {
    float g = 0.f;
    if (i < nb) g = dy[n0 + i];
    if (dmarg && i < nmax) g += dmarg[b * nmax + i];
    gM[(size_t)i * kp1 + 0] = 0.f;
    gM[(size_t)i * kp1 + 1] = g * expf(c.M[(size_t)i * kp1 + 1]);
  }
  __syncwarp();
  for (int L = 1; L <= pl.p; ++L) {
    const int width = n >> L, nj = pl.cap[L] + 1;
    const float* chM = c.M + (size_t)pl.off[L - 1] * kp1;
    const float* chG = gM + (size_t)pl.off[L - 1] * kp1;
    const float* curM = c.M + (size_t)pl.off[L] * kp1;
    const float* curC = c.C + (size_t)pl.off[L] * kp1 * kp1;
    float* curG = gM + (size_t)pl.off[L] * kp1;
    float* curGC = G + (size_t)pl.off[L] * kp1 * kp1;
    for (int w = lane; w < width * nj; w += 32) {
      const int i = w / nj, j = w - i * nj;
      if (!(pl.reach[L] >> j & 1)) continue;
      int lo, hi;
      elem_range(pl, L, j, lo, hi);
      float acc = 0.f;
      for (int e = 0; e <= hi - lo; ++e) {
        const int jl = lo + e, jr = j - jl;
        const float a = curC[((size_t)i * kp1 + j) * kp1 + e] + curM[(size_t)i * kp1 + j];
        const float wl = expf(a - chM[(size_t)(2 * i) * kp1 + jl]);
        const float wr = expf(a - chM[(size_t)(2 * i + 1) * kp1 + jr]);
        const float gc = chG[(size_t)(2 * i) * kp1 + jl] * wl + chG[(size_t)(2 * i + 1) * kp1 + jr] * wr;
        curGC[((size_t)i * kp1 + j) * kp1 + e] = gc;
        acc += gc;
      }
      curG[(size_t)i * kp1 + j] = acc;  // adjoint of M[L][i][j] (the root's is unused: M_root = D - D)
    }
    __syncwarp();
  }
  // ---- adjoint of the bottom-up pass (parents before children); gD overwrites M
  float* gD = c.M;
  if (lane == 0) gD[(size_t)pl.off[pl.p] * kp1 + pl.k] = 0.f;
  __syncwarp();
  for (int L = pl.p; L >= 1; --L) {
    const int width = n >> L, nj = pl.cap[L] + 1;
    const float* curC = c.C + (size_t)pl.off[L] * kp1 * kp1;
    float* curGC = G + (size_t)pl.off[L] * kp1 * kp1;
    const float* curGD = gD + (size_t)pl.off[L] * kp1;
    // gt_e = gc_e + (gD - sum gc) * exp(c_e)
    for (int w = lane; w < width * nj; w += 32) {
      const int i = w / nj, j = w - i * nj;
      if (!(pl.reach[L] >> j & 1)) continue;
      int lo, hi;
      elem_range(pl, L, j, lo, hi);
      float s = 0.f;
      for (int e = 0; e <= hi - lo; ++e) s += curGC[((size_t)i * kp1 + j) * kp1 + e];
      const float r = curGD[(size_t)i * kp1 + j] - s;
      for (int e = 0; e <= hi - lo; ++e) {
        const size_t idx = ((size_t)i * kp1 + j) * kp1 + e;
        curGC[idx] = curGC[idx] + r * expf(curC[idx]);
      }
    }
    __syncwarp();
    // children gather their gD from the parents' gt
    const int cw = n >> (L - 1), nc = pl.cap[L - 1] + 1;
    float* chGD = gD + (size_t)pl.off[L - 1] * kp1;
    for (int w = lane; w < cw * nc; w += 32) {
      const int i = w / nc, cc = w - i * nc;
      if (!(pl.reach[L - 1] >> cc & 1)) continue;
      const int ip = i >> 1, side = i & 1;
      float acc = 0.f;
      for (int j = cc; j <= pl.cap[L]; ++j) {
        if (!(pl.reach[L] >> j & 1) || j - cc > pl.cap[L - 1]) continue;
        int lo, hi;
        elem_range(pl, L, j, lo, hi);
        const int e = (side == 0 ? cc : j - cc) - lo;
        acc += curGC[((size_t)ip * kp1 + j) * kp1 + e];
      }
      chGD[(size_t)i * kp1 + cc] = acc;
    }
    __syncwarp();
  }
  // theta enters through the positive literals only (the negative weight is detached, simple.py:215-217)
  for (int i = lane; i < nb; i += 32) g_theta[n0 + i] = gD[(size_t)i * kp1 + 1];
}

int pad_pow2(int nmax) {
  if (nmax > (1 << 30)) return 0;  // no power of two fits an int (and the doubling below would never terminate)
  int n = 1;
  while (n < nmax) n <<= 1;
  return n;
}

}  // namespace

extern "C" int isg_simple_npad(int nmax) { return nmax >= 1 ? pad_pow2(nmax) : 0; }

extern "C" int isg_simple_marginals_fwd(const float* theta, const float* gumbel, const int32_t* graph_ptr,
                                        int64_t B, int nmax, int k, float* mask, float* marginals,
                                        void* stream_) {
  if (B < 0 || nmax < 0 || k < 1) return ISG_EINVAL;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!theta || !gumbel || !graph_ptr || !mask) return ISG_EINVAL;
  SimplePlan pl;
  const int lk = k < nmax ? k : nmax;
  int rc = simple_plan(pad_pow2(nmax), lk, &pl);
  if (rc != ISG_OK) return rc;
  const size_t fpw = (size_t)pl.total_nodes * pl.kp1 * (2 + pl.kp1) + pl.n;
  const size_t limit = 200 * 1024;
  if (fpw * 4 > limit || pl.n > 1024) return ISG_EUNSUPPORTED;
  int wpc = (int)(limit / (fpw * 4));
  if (wpc > 4) wpc = 4;
  const size_t smem = fpw * 4 * wpc;
  cudaError_t e = cudaFuncSetAttribute(simple_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  simple_fwd_kernel<<<isg::ceil_div(B, wpc), wpc * 32, smem, (cudaStream_t)stream_>>>(
      theta, gumbel, graph_ptr, B, nmax, pl, wpc, fpw, mask, marginals);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_simple_marginals_bwd(const float* dy, const float* d_marginals, const float* theta,
                                        const int32_t* graph_ptr, int64_t B, int nmax, int k, float* g_theta,
                                        void* stream_) {
  if (B < 0 || nmax < 0 || k < 1) return ISG_EINVAL;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!dy || !theta || !graph_ptr || !g_theta) return ISG_EINVAL;
  SimplePlan pl;
  const int lk = k < nmax ? k : nmax;
  int rc = simple_plan(pad_pow2(nmax), lk, &pl);
  if (rc != ISG_OK) return rc;
  const size_t fpw = (size_t)pl.total_nodes * pl.kp1 * (2 + 2 * pl.kp1);
  const size_t limit = 200 * 1024;
  if (fpw * 4 > limit || pl.n > 1024) return ISG_EUNSUPPORTED;
  int wpc = (int)(limit / (fpw * 4));
  if (wpc > 4) wpc = 4;
  const size_t smem = fpw * 4 * wpc;
  cudaError_t e = cudaFuncSetAttribute(simple_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  simple_bwd_kernel<<<isg::ceil_div(B, wpc), wpc * 32, smem, (cudaStream_t)stream_>>>(
      dy, d_marginals, theta, graph_ptr, B, nmax, pl, wpc, fpw, g_theta);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

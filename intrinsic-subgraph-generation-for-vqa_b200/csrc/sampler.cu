// sampler.cu — kernel (c): fused noise-perturbation + per-graph top-k MAP + mask, with the
// perturbation-based gradients of IMLE / AIMLE and the relaxed Gumbel top-k.
//
// One warp per graph; the graph's Nmax dense slots (reference: torch_geometric.utils.to_dense_batch,
// models/masking.py:162 — pads are 0.0 and COMPETE in top-k) live in shared memory.  The k-th
// largest score is found by k rounds of warp arg-max (k is 2..5 in the reference scripts), which
// reproduces `torch.topk(...).values[:, -1]` exactly (multiset semantics), and the mask is
// `score >= thresh` (sampling/methods/deterministic_scheme.py:36-43).  All score arithmetic uses
// explicit round-to-nearest intrinsics in the reference's operation order (no FMA contraction)
// so masks are bit-exact given the same noise tensor.
//
// These kernels are latency-bound at every BASELINE size (≈20 B per node).
#include "common.cuh"

namespace {

using namespace isg;

constexpr int SAMP_WARPS = 4;

// k-th largest of work[0..n) (destroyed), 1 <= k <= n.  All lanes return the same value.
__device__ float kth_largest_warp(float* work, int n, int k, int lane) {
  float thr = 0.f;
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      const float v = work[i];
      if (bi == 0x7fffffff || v > best) {
        best = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(ISG_FULL_MASK, best, o);
      const int oi = __shfl_xor_sync(ISG_FULL_MASK, bi, o);
      const bool take = (oi != 0x7fffffff) && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi));
      if (take) {
        best = ob;
        bi = oi;
      }
    }
    thr = best;
    if ((bi & 31) == lane && bi < n) work[bi] = -INFINITY;
    __syncwarp();
  }
  return thr;
}

// z = MAP(theta_dense + noise*tau)
__global__ void __launch_bounds__(SAMP_WARPS * 32)
topk_mask_fwd_kernel(const float* __restrict__ theta, const float* __restrict__ noise,
                     const int* __restrict__ gptr, int64_t B, int nmax, int k, float tau,
                     float* __restrict__ mask, float* __restrict__ zd) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (b >= B) return;
  float* orig = smem + (size_t)warp * 2 * nmax;
  float* work = orig + nmax;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  for (int i = lane; i < nmax; i += 32) {
    const float th = (i < nb) ? theta[n0 + i] : 0.f;
    const float nz = noise ? noise[b * nmax + i] : 0.f;
    const float sc = __fadd_rn(th, __fmul_rn(nz, tau));
    orig[i] = sc;
    work[i] = sc;
  }
  __syncwarp();
  const bool all = k >= nmax;
  const float thr = all ? 0.f : kth_largest_warp(work, nmax, k, lane);
  for (int i = lane; i < nmax; i += 32) {
    const float z = (all || orig[i] >= thr) ? 1.f : 0.f;
    zd[b * nmax + i] = z;
    if (i < nb) mask[n0 + i] = z;
  }
}

// IMLE: z' = MAP(alpha*theta - beta*dy + noise*tau); g = z - z'
__global__ void __launch_bounds__(SAMP_WARPS * 32)
imle_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ theta,
                const float* __restrict__ noise, const float* __restrict__ zd,
                const int* __restrict__ gptr, int64_t B, int nmax, int k, float alpha, float beta,
                float tau, float* __restrict__ g_theta) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (b >= B) return;
  float* orig = smem + (size_t)warp * 2 * nmax;
  float* work = orig + nmax;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  for (int i = lane; i < nmax; i += 32) {
    const float th = (i < nb) ? theta[n0 + i] : 0.f;
    const float d = (i < nb) ? dy[n0 + i] : 0.f;
    const float nz = noise ? noise[b * nmax + i] : 0.f;
    const float tgt = __fsub_rn(__fmul_rn(alpha, th), __fmul_rn(beta, d));
    const float sc = __fadd_rn(tgt, __fmul_rn(nz, tau));
    orig[i] = sc;
    work[i] = sc;
  }
  __syncwarp();
  const bool all = k >= nmax;
  const float thr = all ? 0.f : kth_largest_warp(work, nmax, k, lane);
  for (int i = lane; i < nb; i += 32) {
    const float z2 = (all || orig[i] >= thr) ? 1.f : 0.f;
    g_theta[n0 + i] = zd[b * nmax + i] - z2;
  }
}

// ---- AIMLE ------------------------------------------------------------------------------
// workspace floats: [0] sum theta^2, [1] sum dy^2 ; ints: [2] nnz
__global__ void aimle_norms_kernel(const float* __restrict__ theta, const float* __restrict__ dy, int64_t N,
                                   float* __restrict__ ws) {
  __shared__ float red[32];
  float st = 0.f, sd = 0.f;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float t = theta[i], d = dy[i];
    st = fmaf(t, t, st);
    sd = fmaf(d, d, sd);
  }
  st = block_sum(st, red);
  sd = block_sum(sd, red);
  if (threadIdx.x == 0) {
    ws[0] = st;
    ws[1] = sd;
    reinterpret_cast<int*>(ws)[2] = 0;
  }
}

__device__ __forceinline__ float aimle_pm(const float* ws, const double* state, int adaptive) {
  if (!adaptive) return (float)state[0];  // fixed TargetDistribution: beta
  const float norm_dy = sqrtf(ws[1]);
  if (!(norm_dy > 0.f)) return 0.f;  // target_aimle.py:113-115
  return __fmul_rn((float)state[0], __fdiv_rn(sqrtf(ws[0]), norm_dy));
}

__global__ void __launch_bounds__(SAMP_WARPS * 32)
aimle_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ theta,
                 const float* __restrict__ noise, const int* __restrict__ gptr, int64_t B, int nmax,
                 int k, float tau, int adaptive, const double* __restrict__ state, float* __restrict__ ws,
                 float* __restrict__ g_theta) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (b >= B) return;
  float* sr = smem + (size_t)warp * 4 * nmax;  // scores R, work R, scores L, work L
  float* wr = sr + nmax;
  float* sl = wr + nmax;
  float* wl = sl + nmax;
  const float pm = aimle_pm(ws, state, adaptive);
  const float alpha = (float)state[3];
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  for (int i = lane; i < nmax; i += 32) {
    const float th = (i < nb) ? theta[n0 + i] : 0.f;
    const float d = (i < nb) ? dy[n0 + i] : 0.f;
    const float nz = noise ? noise[b * nmax + i] : 0.f;
    const float eps = __fmul_rn(nz, tau);
    const float at = __fmul_rn(alpha, th);
    const float r = __fadd_rn(__fsub_rn(at, __fmul_rn(pm, d)), eps);   // theta'_R = a*theta - pm*dy
    const float l = __fadd_rn(__fsub_rn(at, __fmul_rn(pm, -d)), eps);  // theta'_L = a*theta - pm*(-dy)
    sr[i] = r;
    wr[i] = r;
    sl[i] = l;
    wl[i] = l;
  }
  __syncwarp();
  const bool all = k >= nmax;
  const float thr_r = all ? 0.f : kth_largest_warp(wr, nmax, k, lane);
  const float thr_l = all ? 0.f : kth_largest_warp(wl, nmax, k, lane);
  const float div = (adaptive && pm > 0.f) ? pm : 1.f;
  int nnz = 0;
  for (int i = lane; i < nmax; i += 32) {
    const float zr = (all || sr[i] >= thr_r) ? 1.f : 0.f;
    const float zl = (all || sl[i] >= thr_l) ? 1.f : 0.f;
    const float g = (zl - zr) / 2.0f;
    nnz += (g != 0.f);
    if (i < nb) g_theta[n0 + i] = __fdiv_rn(g, div);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nnz += __shfl_xor_sync(ISG_FULL_MASK, nnz, o);
  if (lane == 0 && nnz) atomicAdd(reinterpret_cast<int*>(ws) + 2, nnz);
}

// AdaptiveTargetDistribution.process (target_aimle.py:130-162) on the device: no host sync.
__global__ void aimle_state_update_kernel(double* __restrict__ state, const float* __restrict__ ws, int64_t nb) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nnz = reinterpret_cast<const int*>(ws)[2];
  const float ratio = __fdiv_rn((float)nnz, (float)nb);
  const float decay = (float)state[5];
  const float one_minus = (float)(1.0 - state[5]);
  const float gn = __fadd_rn(__fmul_rn(decay, (float)state[1]), __fmul_rn(one_minus, ratio));
  state[1] = (double)gn;
  double upd = ((gn < (float)state[6]) ? 1.0 : -1.0) * state[4];
  upd = state[7] * state[2] + upd;
  const double nb_ = state[0] + upd;
  state[0] = nb_ > 0.0 ? nb_ : 0.0;
  state[2] = upd;
}

// ---- Gumbel relaxed top-k -----------------------------------------------------------------
__device__ __forceinline__ float warp_sum_strided(const float* a, int n, int lane) {
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += a[i];
  return warp_sum(s);
}

__global__ void __launch_bounds__(SAMP_WARPS * 32)
gumbel_topk_fwd_kernel(const float* __restrict__ theta, const float* __restrict__ gum,
                       const int* __restrict__ gptr, int64_t B, int nmax, int k, float tau,
                       float* __restrict__ mask, float* __restrict__ saved) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (b >= B) return;
  float* s = smem + (size_t)warp * 4 * nmax;
  float* khot = s + nmax;
  float* onehot = khot + nmax;
  float* work = onehot + nmax;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  const int lk = min(k, nmax);
  const float tiny = 1.17549435e-38f;  // np.finfo(np.float32).tiny (gumbel_scheme.py:9)
  for (int i = lane; i < nmax; i += 32) {
    const float th = (i < nb) ? theta[n0 + i] : 0.f;
    s[i] = __fadd_rn(th, gum[b * nmax + i]);
    khot[i] = 0.f;
    onehot[i] = 0.f;
  }
  __syncwarp();
  for (int r = 0; r < lk; ++r) {
    float mx = -INFINITY;
    for (int i = lane; i < nmax; i += 32) {
      const float km = fmaxf(__fsub_rn(1.0f, onehot[i]), tiny);
      const float v = __fadd_rn(s[i], logf(km));
      s[i] = v;
      const float x = __fdiv_rn(v, tau);
      work[i] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < nmax; i += 32) {
      const float e = expf(work[i] - mx);
      work[i] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int i = lane; i < nmax; i += 32) {
      const float o = __fdiv_rn(work[i], sum);
      onehot[i] = o;
      khot[i] = __fadd_rn(khot[i], o);
      saved[((int64_t)b * k + r) * nmax + i] = o;
    }
    __syncwarp();
  }
  for (int r = lk; r < k; ++r)
    for (int i = lane; i < nmax; i += 32) saved[((int64_t)b * k + r) * nmax + i] = 0.f;
  // hard top-k of khot (straight-through): res = (hard - khot) + khot  (gumbel_scheme.py:84-88)
  for (int i = lane; i < nmax; i += 32) work[i] = khot[i];
  __syncwarp();
  const float thr = kth_largest_warp(work, nmax, lk, lane);
  // ties at the threshold: keep the lowest indices so exactly lk entries are hot
  int above = 0;
  for (int i = lane; i < nmax; i += 32) above += (khot[i] > thr);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(ISG_FULL_MASK, above, o);
  int quota = lk - above;  // how many == thr entries to take, in index order
  for (int base = 0; base < nmax; base += 32) {
    const int i = base + lane;
    const bool eq = (i < nmax) && (khot[i] == thr);
    const unsigned bal = __ballot_sync(ISG_FULL_MASK, eq);
    const int rank = __popc(bal & ((1u << lane) - 1u));
    if (i < nmax) {
      const float kh = khot[i];
      const float hard = (kh > thr || (eq && rank < quota)) ? 1.f : 0.f;
      const float res = __fadd_rn(__fsub_rn(hard, kh), kh);
      if (i < nb) mask[n0 + i] = res;
    }
    quota -= __popc(bal);
    if (quota < 0) quota = 0;
  }
}

__global__ void __launch_bounds__(SAMP_WARPS * 32)
gumbel_topk_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ saved,
                       const int* __restrict__ gptr, int64_t B, int nmax, int k, float tau,
                       float* __restrict__ g_theta) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (b >= B) return;
  float* A = smem + (size_t)warp * 2 * nmax;  // adjoint of the scores carried backwards
  float* go = A + nmax;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  const int lk = min(k, nmax);
  const float tiny = 1.17549435e-38f;
  for (int i = lane; i < nmax; i += 32) A[i] = 0.f;
  __syncwarp();
  for (int r = lk - 1; r >= 0; --r) {
    const float* o = saved + ((int64_t)b * k + r) * nmax;
    float dotp = 0.f;
    for (int i = lane; i < nmax; i += 32) {
      const float oi = o[i];
      float g = (i < nb) ? dy[n0 + i] : 0.f;
      if (r < lk - 1) {  // through log(max(1 - o_r, tiny)) of the next round
        const float om = 1.0f - oi;
        if (om > tiny) g -= A[i] / om;
      }
      go[i] = g;
      dotp += oi * g;
    }
    dotp = warp_sum(dotp);
    for (int i = lane; i < nmax; i += 32) A[i] += o[i] * (go[i] - dotp) / tau;
    __syncwarp();
  }
  for (int i = lane; i < nb; i += 32) g_theta[n0 + i] = A[i];
}

// ---- fused forward (IMLE / AIMLE): gate logit -> dropout -> + tau*noise -> per-graph top-k -> node mask ->
// edge mask, ONE launch (SURVEY.md §8d "fused variant"; reference: models/masking.py:151-176 followed by
// mgat_v2_conv.py:166-171 / sampling/node_edge_masks.py:5-12).  One CTA per graph: the logits of the graph's
// nodes are warp dot products (one warp per node row) <xn[n], q[batch[batch[n]]]> in exactly the order of gate_theta_fwd_kernel
// (segment.cu), so theta — and hence the mask — is bit-identical to the unfused kernels; the graph's mask stays
// in shared memory for the edge-mask pass over the graph's in-edges (needs every edge inside one graph).
constexpr int FUSED_WARPS = 8;  // one CTA per graph: the gate dot products (one warp per node row) dominate
__global__ void __launch_bounds__(FUSED_WARPS * 32)
sampler_fused_fwd_kernel(const float* __restrict__ xn, const float* __restrict__ q, const float* __restrict__ keep,
                         const float* __restrict__ noise, const int* __restrict__ batch, const int* __restrict__ gptr,
                         const int* __restrict__ dst_ptr, const int* __restrict__ dst_nbr,
                         const int* __restrict__ dst_eid, int64_t B, int D, int dbl, int nmax, int k, float tau,
                         float* __restrict__ theta, float* __restrict__ mask, float* __restrict__ zd,
                         float* __restrict__ emask) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  float* orig = smem;
  float* work = orig + nmax;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  const float rs = sqrtf((float)D);
  for (int i = warp; i < nb; i += FUSED_WARPS) {  // same per-row arithmetic order as gate_theta_fwd_kernel
    const int n = n0 + i;
    const int qi = dbl ? batch[batch[n]] : batch[n];  // quirk Q1: double gather (see gate_theta_fwd_kernel)
    const float* xr = xn + (int64_t)n * D;
    const float* qr = q + (int64_t)qi * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(xr[c], qr[c], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float g = gelu_f(s / rs);
      const float th = keep ? __fmul_rn(g, keep[n]) : g;
      theta[n] = th;
      orig[i] = th;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nmax; i += FUSED_WARPS * 32) {
    const float th = (i < nb) ? orig[i] : 0.f;
    const float nz = noise ? noise[b * nmax + i] : 0.f;
    const float sc = __fadd_rn(th, __fmul_rn(nz, tau));
    orig[i] = sc;
    work[i] = sc;
  }
  __syncthreads();
  const bool all = k >= nmax;
  if (warp == 0 && !all) {
    const float t = kth_largest_warp(work, nmax, k, lane);
    __syncwarp();
    if (lane == 0) work[0] = t;  // work[] is dead after the selection: slot 0 carries the threshold
  }
  __syncthreads();
  const float thr = all ? 0.f : work[0];
  __syncthreads();
  for (int i = threadIdx.x; i < nmax; i += FUSED_WARPS * 32) {
    const float z = (all || orig[i] >= thr) ? 1.f : 0.f;
    zd[b * nmax + i] = z;
    work[i] = z;
    if (i < nb) mask[n0 + i] = z;
  }
  __syncthreads();
  if (emask != nullptr && nb > 0) {  // edge_mask[e] = mask[src] * mask[dst] over the in-edges of the graph's nodes
    const int p0 = dst_ptr[n0], p1 = dst_ptr[n0 + nb];
    int node = 0;  // threads walk the edge range; the owning node is found by advancing over dst_ptr
    for (int p = p0 + threadIdx.x; p < p1; p += FUSED_WARPS * 32) {
      while (dst_ptr[n0 + node + 1] <= p) ++node;
      const int src = dst_nbr[p] - n0;
      emask[dst_eid[p]] = __fmul_rn(work[src], work[node]);
    }
  }
}

inline int check_samp(const int32_t* gptr, int64_t B, int nmax, int k) {
  if (B < 0 || nmax < 0 || k < 0) return ISG_EINVAL;
  if (B > 0 && !gptr) return ISG_EINVAL;
  return ISG_OK;
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 200 * 1024) return ISG_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return ISG_OK;
}

}  // namespace

extern "C" int isg_topk_mask_fwd(const float* theta, const float* noise, const int32_t* gptr, int64_t B,
                                 int nmax, int k, float tau, float* mask, float* z_dense, void* stream_) {
  int rc = check_samp(gptr, B, nmax, k);
  if (rc) return rc;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!theta || !mask || !z_dense || k < 1) return ISG_EINVAL;
  const size_t smem = (size_t)SAMP_WARPS * 2 * nmax * sizeof(float);
  if ((rc = set_smem(topk_mask_fwd_kernel, smem))) return rc;
  topk_mask_fwd_kernel<<<isg::ceil_div(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream_>>>(
      theta, noise, gptr, B, nmax, k, tau, mask, z_dense);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_sampler_fused_fwd(const float* xn, const float* q, const float* keep, const float* noise,
                                     const int32_t* batch32, const int32_t* gptr, const int32_t* dst_ptr,
                                     const int32_t* dst_nbr, const int32_t* dst_eid, int64_t B, int D,
                                     int double_gather, int nmax, int k, float tau, float* theta, float* mask,
                                     float* z_dense, float* edge_mask, void* stream_) {
  int rc = check_samp(gptr, B, nmax, k);
  if (rc) return rc;
  if (D <= 0) return ISG_EINVAL;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!xn || !q || !batch32 || !theta || !mask || !z_dense || k < 1) return ISG_EINVAL;
  if (edge_mask && (!dst_ptr || !dst_nbr || !dst_eid)) return ISG_EINVAL;
  const size_t smem = (size_t)2 * nmax * sizeof(float);
  if ((rc = set_smem(sampler_fused_fwd_kernel, smem))) return rc;
  sampler_fused_fwd_kernel<<<(unsigned)B, FUSED_WARPS * 32, smem, (cudaStream_t)stream_>>>(
      xn, q, keep, noise, batch32, gptr, dst_ptr, dst_nbr, dst_eid, B, D, double_gather, nmax, k, tau, theta, mask,
      z_dense, edge_mask);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_imle_bwd(const float* dy, const float* theta, const float* noise, const float* z_dense,
                            const int32_t* gptr, int64_t B, int nmax, int k, float alpha, float beta,
                            float tau_target, float* g_theta, void* stream_) {
  int rc = check_samp(gptr, B, nmax, k);
  if (rc) return rc;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!dy || !theta || !z_dense || !g_theta || k < 1) return ISG_EINVAL;
  const size_t smem = (size_t)SAMP_WARPS * 2 * nmax * sizeof(float);
  if ((rc = set_smem(imle_bwd_kernel, smem))) return rc;
  imle_bwd_kernel<<<isg::ceil_div(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream_>>>(
      dy, theta, noise, z_dense, gptr, B, nmax, k, alpha, beta, tau_target, g_theta);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" size_t isg_aimle_workspace_bytes(void) { return 256; }

extern "C" int isg_aimle_bwd(const float* dy, const float* theta, const float* noise, const int32_t* gptr,
                             int64_t N, int64_t B, int nmax, int k, float tau_target, int adaptive,
                             double* state, float* g_theta, void* workspace, size_t ws_bytes, void* stream_) {
  int rc = check_samp(gptr, B, nmax, k);
  if (rc) return rc;
  if (!state) return ISG_EINVAL;
  if (ws_bytes < isg_aimle_workspace_bytes() || !workspace) return ISG_EWORKSPACE;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!dy || !theta || !g_theta || k < 1 || N < 0) return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  float* ws = (float*)workspace;
  aimle_norms_kernel<<<1, 1024, 0, stream>>>(theta, dy, N, ws);
  ISG_CHECK_LAUNCH();
  const size_t smem = (size_t)SAMP_WARPS * 4 * nmax * sizeof(float);
  if ((rc = set_smem(aimle_bwd_kernel, smem))) return rc;
  aimle_bwd_kernel<<<isg::ceil_div(B, SAMP_WARPS), SAMP_WARPS * 32, smem, stream>>>(
      dy, theta, noise, gptr, B, nmax, k, tau_target, adaptive, state, ws, g_theta);
  ISG_CHECK_LAUNCH();
  if (adaptive) {
    aimle_state_update_kernel<<<1, 32, 0, stream>>>(state, ws, B);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

extern "C" int isg_gumbel_topk_fwd(const float* theta, const float* gumbel, const int32_t* gptr, int64_t B,
                                   int nmax, int k, float tau, float* mask, float* saved, void* stream_) {
  int rc = check_samp(gptr, B, nmax, k);
  if (rc) return rc;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!theta || !gumbel || !mask || !saved || k < 1) return ISG_EINVAL;
  const size_t smem = (size_t)SAMP_WARPS * 4 * nmax * sizeof(float);
  if ((rc = set_smem(gumbel_topk_fwd_kernel, smem))) return rc;
  gumbel_topk_fwd_kernel<<<isg::ceil_div(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream_>>>(
      theta, gumbel, gptr, B, nmax, k, tau, mask, saved);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_gumbel_topk_bwd(const float* dy, const float* saved, const int32_t* gptr, int64_t B,
                                   int nmax, int k, float tau, float* g_theta, void* stream_) {
  int rc = check_samp(gptr, B, nmax, k);
  if (rc) return rc;
  if (B == 0 || nmax == 0) return ISG_OK;
  if (!dy || !saved || !g_theta || k < 1) return ISG_EINVAL;
  const size_t smem = (size_t)SAMP_WARPS * 2 * nmax * sizeof(float);
  if ((rc = set_smem(gumbel_topk_bwd_kernel, smem))) return rc;
  gumbel_topk_bwd_kernel<<<isg::ceil_div(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream_>>>(
      dy, saved, gptr, B, nmax, k, tau, g_theta);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

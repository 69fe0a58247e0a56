// optim.cu — SURVEY.md §8 row f4: the optimizer tail of the training step as three launches with no host
// synchronisation.  Reference (training/train_epoch.py:111-118):
//     gradscaler.scale(loss).backward(); gradscaler.unscale_(optimizer)
//     torch.nn.utils.clip_grad_norm_(model_params, max_norm=2.0)
//     gradscaler.step(optimizer)            # torch.optim.Adam(lr), skipped when a gradient is inf/nan (host sync)
//     gradscaler.update()
// i.e. per step: a multi-tensor unscale + inf check, a multi-tensor L2 norm, a multi-tensor scale, one `.item()`
// sync, and the Adam update.  Here: (1) per-tensor-chunk partial sums of (g * inv_scale)^2 with a non-finite flag,
// (2) a fixed-order final reduction -> total norm, clip coefficient, found_inf — all device scalars, (3) the Adam
// update with the combined factor inv_scale * clip_coef, skipped on the device when found_inf is set (the step
// counter then does not advance either, like GradScaler.step).  Deterministic.
//
// Tensors are passed BY VALUE in the kernel parameter block (<= ISG_OPT_MAX_TENSORS per launch), so no metadata
// upload is needed even though the gradients' addresses change every step.
#include "common.cuh"

namespace {

using namespace isg;

constexpr int OPT_THREADS = 256;
constexpr int OPT_ILP = 4;                              // float4 per thread per iteration
constexpr int OPT_CHUNK = OPT_THREADS * OPT_ILP * 4;    // elements per block: 4096

struct TensorList {
  float* p[ISG_OPT_MAX_TENSORS];
  const float* g[ISG_OPT_MAX_TENSORS];
  float* m[ISG_OPT_MAX_TENSORS];
  float* v[ISG_OPT_MAX_TENSORS];
  int64_t n[ISG_OPT_MAX_TENSORS];
  int first_block[ISG_OPT_MAX_TENSORS + 1];  // prefix sum of ceil(n / OPT_CHUNK)
  int count;
};

__device__ __forceinline__ int find_tensor(const TensorList& tl, int block) {
  int lo = 0, hi = tl.count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tl.first_block[mid] <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// partial[block] = sum over the block's chunk of (g * inv_scale)^2 (double); flag |= any non-finite g
__global__ void __launch_bounds__(OPT_THREADS)
grad_sq_partial_kernel(TensorList tl, const float* __restrict__ inv_scale_p, double* __restrict__ partial,
                       int partial_off, int* __restrict__ nonfinite) {
  __shared__ double red[OPT_THREADS / 32];
  const int t = find_tensor(tl, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - tl.first_block[t]) * OPT_CHUNK;
  const int64_t n = tl.n[t];
  const float* g = tl.g[t];
  const float inv = inv_scale_p ? *inv_scale_p : 1.0f;
  float acc = 0.f;
  bool bad = false;
  for (int it = 0; it < OPT_ILP * 4; ++it) {
    const int64_t i = base + (int64_t)it * OPT_THREADS + threadIdx.x;
    if (i < n) {
      const float x = g[i] * inv;
      bad |= !isfinite(x);
      acc = fmaf(x, x, acc);
    }
  }
  double s = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(ISG_FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  const unsigned anybad = __syncthreads_or(bad ? 1 : 0);
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) tsum += red[w];
    partial[partial_off + blockIdx.x] = tsum;
    if (anybad) atomicOr(nonfinite, 1);
  }
}

// state (float[4]): [0] total grad norm (unscaled), [1] clip coefficient, [2] found_inf (0/1), [3] steps taken
__global__ void clip_finalize_kernel(const double* __restrict__ partial, int nparts, const int* __restrict__ nonfinite,
                                     float max_norm, float* __restrict__ state, float* __restrict__ found_inf_out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];  // fixed assignment -> deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(ISG_FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float norm = (float)sqrt(tot);
    const bool bad = (*nonfinite != 0) || !isfinite(norm);
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float coef = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
    state[0] = norm;
    state[1] = coef;
    state[2] = bad ? 1.f : 0.f;
    if (!bad) state[3] += 1.f;
    if (found_inf_out) *found_inf_out = bad ? 1.f : 0.f;
  }
}

// torch.optim.Adam (no weight decay, no amsgrad): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(OPT_THREADS)
adam_update_kernel(TensorList tl, const float* __restrict__ inv_scale_p, const float* __restrict__ state,
                   const float* __restrict__ lr_p, float lr, float beta1, float beta2, float eps) {
  if (state[2] != 0.f) return;  // found_inf: skip the step (GradScaler.step semantics)
  const int t = find_tensor(tl, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - tl.first_block[t]) * OPT_CHUNK;
  const int64_t n = tl.n[t];
  const float gscale = (inv_scale_p ? *inv_scale_p : 1.0f) * state[1];
  const float step = state[3];  // already counts this step
  const float bc1 = 1.0f - powf(beta1, step), bc2 = 1.0f - powf(beta2, step);
  const float lr_eff = lr_p ? *lr_p : lr;
  const float step_size = lr_eff / bc1, bc2_sqrt = sqrtf(bc2);
  float *p = tl.p[t], *m = tl.m[t], *v = tl.v[t];
  const float* g = tl.g[t];
  for (int it = 0; it < OPT_ILP * 4; ++it) {
    const int64_t i = base + (int64_t)it * OPT_THREADS + threadIdx.x;
    if (i < n) {
      const float gi = g[i] * gscale;
      const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);            // lerp, as torch's _single_tensor_adam
      const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);    // mul_(beta2).addcmul_(g, g, 1 - beta2)
      m[i] = mi;
      v[i] = vi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      p[i] = p[i] - step_size * (mi / denom);
    }
  }
}

int fill(TensorList& tl, void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
         const int64_t* numel, int count) {
  if (count < 1 || count > ISG_OPT_MAX_TENSORS) return ISG_EINVAL;
  int blocks = 0;
  for (int i = 0; i < count; ++i) {
    if (numel[i] < 0 || !grads[i]) return ISG_EINVAL;
    tl.p[i] = params ? (float*)params[i] : nullptr;
    tl.g[i] = (const float*)grads[i];
    tl.m[i] = exp_avg ? (float*)exp_avg[i] : nullptr;
    tl.v[i] = exp_avg_sq ? (float*)exp_avg_sq[i] : nullptr;
    tl.n[i] = numel[i];
    tl.first_block[i] = blocks;
    blocks += (int)((numel[i] + OPT_CHUNK - 1) / OPT_CHUNK);
  }
  tl.first_block[count] = blocks;
  tl.count = count;
  return ISG_OK;
}

}  // namespace

extern "C" int isg_opt_max_tensors(void) { return ISG_OPT_MAX_TENSORS; }

extern "C" int64_t isg_opt_blocks(const int64_t* numel, int count) {
  int64_t b = 0;
  if (!numel) return 0;
  for (int i = 0; i < count; ++i) b += numel[i] > 0 ? (numel[i] + OPT_CHUNK - 1) / OPT_CHUNK : 0;
  return b;
}

// One group (<= ISG_OPT_MAX_TENSORS tensors) of the squared-norm pass; partial sums go to partial[partial_off ...).
extern "C" int isg_grad_sq_partials(const void* const* grads, const int64_t* numel, int count, const float* inv_scale,
                                    double* partial, int partial_off, int32_t* nonfinite, void* stream_) {
  if (!grads || !numel || !partial || !nonfinite) return ISG_EINVAL;
  TensorList tl;
  int rc = fill(tl, nullptr, grads, nullptr, nullptr, numel, count);
  if (rc) return rc;
  const int blocks = tl.first_block[count];
  if (blocks == 0) return ISG_OK;
  grad_sq_partial_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream_>>>(tl, inv_scale, partial, partial_off, nonfinite);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_clip_finalize(const double* partial, int nparts, const int32_t* nonfinite, float max_norm,
                                 float* state, float* found_inf_out, void* stream_) {
  if (!partial || !nonfinite || !state || nparts < 0) return ISG_EINVAL;
  clip_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(partial, nparts, nonfinite, max_norm, state, found_inf_out);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_adam_update(void* const* params, const void* const* grads, void* const* exp_avg,
                               void* const* exp_avg_sq, const int64_t* numel, int count, const float* inv_scale,
                               const float* state, const float* lr_dev, float lr, float beta1, float beta2, float eps,
                               void* stream_) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !state) return ISG_EINVAL;
  TensorList tl;
  int rc = fill(tl, params, grads, exp_avg, exp_avg_sq, numel, count);
  if (rc) return rc;
  for (int i = 0; i < count; ++i)
    if (!tl.p[i] || !tl.m[i] || !tl.v[i]) return ISG_EINVAL;
  const int blocks = tl.first_block[count];
  if (blocks == 0) return ISG_OK;
  adam_update_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream_>>>(tl, inv_scale, state, lr_dev, lr, beta1, beta2, eps);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

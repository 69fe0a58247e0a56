// sgenc.cu — SURVEY.md §8 row f2: the gather / segment kernels of the scene-graph encoding layer that feeds MGAT
// (reference: torch_geometric.nn.MetaLayer(EdgeModel, NodeModel), models/scene_graph_encoder.py:107-146, and the
// GraphNorm that SceneGraphEncoder.forward evaluates in float64 ON THE CPU, :99-102).
//
// The reference concatenates [x[src], x[dst], e] into an [E, 900] tensor and multiplies it by a [300, 900] weight.
// Here the weight is split by column block: the node blocks are applied ONCE PER NODE (N rows instead of E rows:
// ~8x fewer FLOPs at GQA's 8 edges per node) by the tcgen05 projections, and these kernels add the gathered node
// terms to the per-edge term, apply GELU, and do the scatter_mean / its transpose as deterministic CSR segment
// sums (no atomics).  All HBM-bound streaming over [E, D] fp32 rows with 16-byte accesses.
#include "common.cuh"

namespace {

using namespace isg;

// z[e] = a[src[e]] + b[dst[e]] + q[e]  (a, b optional);  y = act(z).  One thread per float4.
__global__ void gather_add_act_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                      const float* __restrict__ q, const int64_t* __restrict__ ei, int64_t E, int D4,
                                      int act, float* __restrict__ z, float* __restrict__ y) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= E * D4) return;
  const int64_t e = idx / D4;
  const int c = (int)(idx - e * D4);
  float4 v = Vec4<float>::ld_stream(q + idx * 4);
  if (a) v = f4_add(v, Vec4<float>::ld(a + (ei[e] * D4 + c) * 4));
  if (b) v = f4_add(v, Vec4<float>::ld(b + (ei[E + e] * D4 + c) * 4));
  if (z) Vec4<float>::st_stream(z + idx * 4, v);
  if (act == ISG_ACT_GELU) v = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
  Vec4<float>::st_stream(y + idx * 4, v);
}

// out[n] = scale_n * sum_{p in [ptr[n], ptr[n+1])} in[eid[p]],  scale_n = mean ? 1/max(deg,1) : 1.
// One warp per (node, 128-column group); fixed CSR order -> deterministic.
__global__ void segment_sum_kernel(const float* __restrict__ in, const int* __restrict__ ptr,
                                   const int* __restrict__ eid, int64_t N, int D4, int mean, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int groups = (D4 + 31) / 32;
  if (w >= N * groups) return;
  const int64_t n = w / groups;
  const int c = (int)(w - n * groups) * 32 + lane;
  const int beg = ptr[n], end = ptr[n + 1];
  float4 acc = f4_zero();
  if (c < D4) {
    int p = beg;
    for (; p + 3 < end; p += 4) {  // four independent row loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = Vec4<float>::ld_stream(in + ((int64_t)eid[p + u] * D4 + c) * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = f4_add(acc, v[u]);
    }
    for (; p < end; ++p) acc = f4_add(acc, Vec4<float>::ld_stream(in + ((int64_t)eid[p] * D4 + c) * 4));
    if (mean) acc = f4_scale(acc, 1.0f / (float)max(end - beg, 1));
    Vec4<float>::st(out + (n * D4 + c) * 4, acc);
  }
}

// out[e] = scale * in[idx[e]],  scale = ptr ? 1/max(deg(idx[e]),1) : 1   (backward of the segment mean)
__global__ void gather_rows_kernel(const float* __restrict__ in, const int64_t* __restrict__ idx,
                                   const int* __restrict__ ptr, int64_t E, int D4, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E * D4) return;
  const int64_t e = i / D4;
  const int c = (int)(i - e * D4);
  const int64_t n = idx[e];
  float4 v = Vec4<float>::ld(in + (n * D4 + c) * 4);
  if (ptr) v = f4_scale(v, 1.0f / (float)max(ptr[n + 1] - ptr[n], 1));
  Vec4<float>::st_stream(out + i * 4, v);
}

// ---- GraphNorm with float64 arithmetic, one CTA per graph, thread per channel (coalesced across channels).
// Reference: x.type(DoubleTensor) -> torch_geometric GraphNorm(eps 1e-5) -> back to float
// (models/scene_graph_encoder.py:99-102): the statistics and the normalisation are evaluated in double, the
// result rounded to float once.  No PCIe round trip: everything stays on the device.
__global__ void graphnorm64_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                       const float* __restrict__ b, const float* __restrict__ ms,
                                       const int* __restrict__ gptr, int D, double eps, float* __restrict__ y,
                                       double* __restrict__ mean_o, double* __restrict__ rstd_o) {
  const int g = blockIdx.x;
  const int n0 = gptr[g], n1 = gptr[g + 1];
  const double inv = 1.0 / (double)max(n1 - n0, 1);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    double s = 0.0;
    for (int n = n0; n < n1; ++n) s += (double)x[(int64_t)n * D + c];
    const double mean = s * inv, msc = (double)ms[c];
    double v = 0.0;
    for (int n = n0; n < n1; ++n) {
      const double o = (double)x[(int64_t)n * D + c] - mean * msc;
      v += o * o;
    }
    const double rstd = 1.0 / sqrt(v * inv + eps);
    const double wc = (double)w[c], bc = (double)b[c];
    for (int n = n0; n < n1; ++n) {
      const double o = (double)x[(int64_t)n * D + c] - mean * msc;
      y[(int64_t)n * D + c] = (float)(wc * o * rstd + bc);
    }
    mean_o[(int64_t)g * D + c] = mean;
    rstd_o[(int64_t)g * D + c] = rstd;
  }
}

// backward (double accumulation).  With o = x - mean*ms, r = rstd, xh = o*r, y = w*xh + b, n = nodes in graph:
//   g_xh = g_y*w;  g_o = r*(g_xh - xh*mean_n(g_xh*xh));  g_x = g_o - ms*mean_n(g_o)
//   g_w += sum g_y*xh;  g_b += sum g_y;  g_ms += -mean * sum g_o       (per graph partials, summed by isg_colsum)
__global__ void graphnorm64_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                       const float* __restrict__ w, const float* __restrict__ ms,
                                       const double* __restrict__ mean_i, const double* __restrict__ rstd_i,
                                       const int* __restrict__ gptr, int D, float* __restrict__ gx,
                                       float* __restrict__ gw_part, float* __restrict__ gb_part,
                                       float* __restrict__ gms_part) {
  const int g = blockIdx.x;
  const int n0 = gptr[g], n1 = gptr[g + 1];
  const double inv = 1.0 / (double)max(n1 - n0, 1);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const double mean = mean_i[(int64_t)g * D + c], r = rstd_i[(int64_t)g * D + c];
    const double wc = (double)w[c], msc = (double)ms[c];
    double s_gy = 0.0, s_gyxh = 0.0;
    for (int n = n0; n < n1; ++n) {
      const double xh = ((double)x[(int64_t)n * D + c] - mean * msc) * r;
      const double g_ = (double)gy[(int64_t)n * D + c];
      s_gy += g_;
      s_gyxh += g_ * xh;
    }
    const double m_gxhxh = wc * s_gyxh * inv;  // mean_n(g_xh * xh)
    // sum_n g_o = r * (w*s_gy - m_gxhxh * sum_n xh);  sum_n xh = r * (sum x - n*mean*ms) = r * n * mean * (1 - ms)
    double s_go = 0.0;
    for (int n = n0; n < n1; ++n) {
      const double xh = ((double)x[(int64_t)n * D + c] - mean * msc) * r;
      s_go += r * ((double)gy[(int64_t)n * D + c] * wc - xh * m_gxhxh);
    }
    const double m_go = s_go * inv;
    for (int n = n0; n < n1; ++n) {
      const double xh = ((double)x[(int64_t)n * D + c] - mean * msc) * r;
      const double go = r * ((double)gy[(int64_t)n * D + c] * wc - xh * m_gxhxh);
      gx[(int64_t)n * D + c] = (float)(go - msc * m_go);
    }
    gw_part[(int64_t)g * D + c] = (float)s_gyxh;
    gb_part[(int64_t)g * D + c] = (float)s_gy;
    gms_part[(int64_t)g * D + c] = (float)(-mean * s_go);
  }
}

}  // namespace

extern "C" int isg_gather_add_act_fwd(const float* a, const float* b, const float* q, const int64_t* edge_index,
                                      int64_t E, int D, int act, float* z_pre, float* y, void* stream_) {
  if (E < 0 || D <= 0) return ISG_EINVAL;
  if (D % 4 != 0) return ISG_EUNSUPPORTED;
  if (E == 0) return ISG_OK;
  if (!q || !y || ((a || b) && !edge_index)) return ISG_EINVAL;
  const int64_t total = E * (D / 4);
  gather_add_act_kernel<<<isg::ceil_div(total, 256), 256, 0, (cudaStream_t)stream_>>>(a, b, q, edge_index, E, D / 4,
                                                                                      act, z_pre, y);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_segment_sum(const float* in, const int32_t* ptr, const int32_t* eid, int64_t N, int D, int mean,
                               float* out, void* stream_) {
  if (N < 0 || D <= 0) return ISG_EINVAL;
  if (D % 4 != 0) return ISG_EUNSUPPORTED;
  if (N == 0) return ISG_OK;
  if (!ptr || !out) return ISG_EINVAL;
  const int D4 = D / 4, groups = (D4 + 31) / 32;
  const int64_t warps = N * groups;
  segment_sum_kernel<<<isg::ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream_>>>(in, ptr, eid, N, D4, mean, out);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_gather_rows(const float* in, const int64_t* idx, const int32_t* ptr, int64_t E, int D, float* out,
                               void* stream_) {
  if (E < 0 || D <= 0) return ISG_EINVAL;
  if (D % 4 != 0) return ISG_EUNSUPPORTED;
  if (E == 0) return ISG_OK;
  if (!in || !idx || !out) return ISG_EINVAL;
  const int64_t total = E * (D / 4);
  gather_rows_kernel<<<isg::ceil_div(total, 256), 256, 0, (cudaStream_t)stream_>>>(in, idx, ptr, E, D / 4, out);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_graphnorm64_fwd(const float* x, const float* weight, const float* bias, const float* mean_scale,
                                   const int32_t* graph_ptr, int64_t B, int D, double eps, float* y, double* mean,
                                   double* rstd, void* stream_) {
  if (B < 0 || D <= 0) return ISG_EINVAL;
  if (B == 0) return ISG_OK;
  if (!x || !weight || !bias || !mean_scale || !graph_ptr || !y || !mean || !rstd) return ISG_EINVAL;
  graphnorm64_fwd_kernel<<<(unsigned)B, 320, 0, (cudaStream_t)stream_>>>(x, weight, bias, mean_scale, graph_ptr, D, eps,
                                                                         y, mean, rstd);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_graphnorm64_bwd(const float* g_y, const float* x, const float* weight, const float* mean_scale,
                                   const double* mean, const double* rstd, const int32_t* graph_ptr, int64_t B, int D,
                                   float* g_x, float* gw_part, float* gb_part, float* gms_part, void* stream_) {
  if (B < 0 || D <= 0) return ISG_EINVAL;
  if (B == 0) return ISG_OK;
  if (!g_y || !x || !weight || !mean_scale || !mean || !rstd || !graph_ptr || !g_x || !gw_part || !gb_part ||
      !gms_part)
    return ISG_EINVAL;
  graphnorm64_bwd_kernel<<<(unsigned)B, 320, 0, (cudaStream_t)stream_>>>(g_y, x, weight, mean_scale, mean, rstd,
                                                                         graph_ptr, D, g_x, gw_part, gb_part, gms_part);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

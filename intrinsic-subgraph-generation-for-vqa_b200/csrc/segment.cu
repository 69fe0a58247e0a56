// segment.cu — node-side fused segment ops of the hot path (all fp32, HBM/latency-bound):
//   * instruction gating           y = gelu(x * ins[batch])            (mgat_v2_conv.py:156-157)
//   * gate logits theta            gelu(<xn, q[batch[batch]]>/sqrt(D)) (masking.py:151-155)
//   * scatter-SDPA + GraphNorm + residual, one CTA per graph           (mgat.py:168-172)
//   * gelu backward, deterministic column sums
// Every per-graph reduction is done by one CTA over the graph's contiguous node range
// (batch is sorted), so there are no floating-point atomics and results are run-to-run stable.
#include "common.cuh"

namespace {

using namespace isg;

// ---------------------------------------------------------------- instruction gating
__global__ void instr_gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ins,
                                      const int* __restrict__ batch, int64_t N, int D4,
                                      float* __restrict__ y) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // over N * D4 float4s
  if (idx >= N * D4) return;
  const int64_t n = idx / D4;
  const int c = (int)(idx - n * D4);
  const float4 xv = Vec4<float>::ld(x + idx * 4);
  const float4 iv = Vec4<float>::ld(ins + ((int64_t)batch[n] * D4 + c) * 4);
  float4 o;
  o.x = gelu_f(xv.x * iv.x);
  o.y = gelu_f(xv.y * iv.y);
  o.z = gelu_f(xv.z * iv.z);
  o.w = gelu_f(xv.w * iv.w);
  Vec4<float>::st(y + idx * 4, o);
}

// one CTA per graph; thread c loops over the graph's nodes (coalesced across c) — any D
__global__ void instr_gate_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                      const float* __restrict__ ins, const int* __restrict__ gptr, int D,
                                      const float* __restrict__ gres, int acc_ins, float* __restrict__ gx,
                                      float* __restrict__ gins) {
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float iv = ins[(int64_t)b * D + c];
    float acc = 0.f;
    for (int n = n0; n < n1; ++n) {
      const int64_t o = (int64_t)n * D + c;
      const float xv = x[o];
      const float gp = gy[o] * gelu_grad_f(xv * iv);
      gx[o] = gres ? __fadd_rn(__fmul_rn(gp, iv), gres[o]) : gp * iv;  // + the residual branch's gradient
      acc = fmaf(gp, xv, acc);
    }
    const int64_t oi = (int64_t)b * D + c;
    gins[oi] = acc_ins ? gins[oi] + acc : acc;
  }
}

// D % 4 == 0: float4 columns x IG_LANES row lanes (the serial walk above is a chain of ~20-40 dependent DRAM
// round trips per thread: 32 us at the c3 size for 24 MB); per-lane partial sums of g_ins are folded in a fixed
// order, so the result is deterministic.
constexpr int IG_COLS = 80, IG_LANES = 8;  // 640 threads (4 lanes: 15.2 us per launch at c3, 8 lanes: 11.1)
__global__ void __launch_bounds__(IG_COLS * IG_LANES)
instr_gate_bwd_v4_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ ins,
                         const int* __restrict__ gptr, int D, const float* __restrict__ gres, int acc_ins,
                         float* __restrict__ gx, float* __restrict__ gins) {
  pdl_enter();
  __shared__ float4 red[IG_LANES][IG_COLS];
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1];
  const int cx = threadIdx.x % IG_COLS, ly = threadIdx.x / IG_COLS;
  const int d4 = D >> 2;
  for (int c0 = 0; c0 < d4; c0 += IG_COLS) {
    const int c4 = c0 + cx;
    float4 acc = f4_zero();
    if (c4 < d4) {
      const float4 iv = Vec4<float>::ld(ins + (int64_t)b * D + 4 * c4);
      for (int n = n0 + ly; n < n1; n += IG_LANES) {
        const int64_t o = (int64_t)n * D + 4 * c4;
        const float4 xv = Vec4<float>::ld(x + o), g = Vec4<float>::ld(gy + o);
        float4 gp, r;
        gp.x = g.x * gelu_grad_f(xv.x * iv.x);
        gp.y = g.y * gelu_grad_f(xv.y * iv.y);
        gp.z = g.z * gelu_grad_f(xv.z * iv.z);
        gp.w = g.w * gelu_grad_f(xv.w * iv.w);
        if (gres) {
          const float4 gr = Vec4<float>::ld(gres + o);
          r.x = __fadd_rn(__fmul_rn(gp.x, iv.x), gr.x);
          r.y = __fadd_rn(__fmul_rn(gp.y, iv.y), gr.y);
          r.z = __fadd_rn(__fmul_rn(gp.z, iv.z), gr.z);
          r.w = __fadd_rn(__fmul_rn(gp.w, iv.w), gr.w);
        } else {
          r.x = gp.x * iv.x;
          r.y = gp.y * iv.y;
          r.z = gp.z * iv.z;
          r.w = gp.w * iv.w;
        }
        Vec4<float>::st(gx + o, r);
        acc.x = fmaf(gp.x, xv.x, acc.x);
        acc.y = fmaf(gp.y, xv.y, acc.y);
        acc.z = fmaf(gp.z, xv.z, acc.z);
        acc.w = fmaf(gp.w, xv.w, acc.w);
      }
    }
    red[ly][cx] = acc;
    __syncthreads();
    if (ly == 0 && c4 < d4) {
      float4 t = red[0][cx];
#pragma unroll
      for (int y = 1; y < IG_LANES; ++y) t = f4_add(t, red[y][cx]);
      float* go = gins + (int64_t)b * D + 4 * c4;
      if (acc_ins) t = f4_add(Vec4<float>::ld(go), t);
      Vec4<float>::st(go, t);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- concat_instr variant of the gating
// y[n] = [x[n], ins[batch[n]]]  (mgat_v2_conv.py:153-154, `--concat_instr 1`; the conv then has in_channels = 2C)
__global__ void concat_instr_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ins,
                                        const int* __restrict__ batch, int64_t N, int D, float* __restrict__ y) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N * 2 * D) return;
  const int64_t n = i / (2 * D);
  const int c = (int)(i - n * 2 * D);
  y[i] = c < D ? x[n * D + c] : ins[(int64_t)batch[n] * D + (c - D)];
}
// gx[n] = gy[n, :D] (+ residual gradient);  gins[b] (+)= sum over the graph's nodes of gy[n, D:]  (fixed order)
__global__ void concat_instr_bwd_kernel(const float* __restrict__ gy, const int* __restrict__ gptr, int D,
                                        const float* __restrict__ gres, int acc_ins, float* __restrict__ gx,
                                        float* __restrict__ gins) {
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int n = n0; n < n1; ++n) {
      const float* row = gy + (int64_t)n * 2 * D;
      const int64_t o = (int64_t)n * D + c;
      gx[o] = gres ? __fadd_rn(row[c], gres[o]) : row[c];
      acc += row[D + c];
    }
    const int64_t oi = (int64_t)b * D + c;
    gins[oi] = acc_ins ? gins[oi] + acc : acc;
  }
}

// ---------------------------------------------------------------- gate logits theta
// warp per node
__global__ void gate_theta_fwd_kernel(const float* __restrict__ xn, const float* __restrict__ q,
                                      const int* __restrict__ batch, int64_t N, int D, int dbl,
                                      const float* __restrict__ keep, float* __restrict__ theta) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const int qi = dbl ? batch[batch[n]] : batch[n];  // quirk Q1: double gather
  const float* xr = xn + n * D;
  const float* qr = q + (int64_t)qi * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s = fmaf(xr[c], qr[c], s);
  s = warp_sum(s);
  if (lane == 0) {
    const float th = gelu_f(s / sqrtf((float)D));
    theta[n] = keep ? __fmul_rn(th, keep[n]) : th;  // dropout keep-mask (0 or 1/(1-p)), masking.py:159
  }
}

// g_xn: warp per node (recomputes the pre-activation)
__global__ void gate_theta_bwd_xn_kernel(const float* __restrict__ gth, const float* __restrict__ xn,
                                         const float* __restrict__ q, const int* __restrict__ batch,
                                         int64_t N, int D, int dbl, const float* __restrict__ keep,
                                         float* __restrict__ gxn, float* __restrict__ gpre) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const int qi = dbl ? batch[batch[n]] : batch[n];
  const float* xr = xn + n * D;
  const float* qr = q + (int64_t)qi * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s = fmaf(xr[c], qr[c], s);
  s = warp_sum(s);
  const float rs = 1.0f / sqrtf((float)D);
  const float gth_n = keep ? __fmul_rn(gth[n], keep[n]) : gth[n];
  const float gp = gth_n * gelu_grad_f(s * rs) * rs;  // d loss / d <xn,q>
  for (int c = lane; c < D; c += 32) gxn[n * D + c] = gp * qr[c];
  if (lane == 0) gpre[n] = gp;
}

// g_q[r] = sum over nodes n with batch[batch[n]] == r of gpre[n]*xn[n].  Nodes of graph g all use
// row batch[g]; the graphs g with batch[g] == r are g in [gptr[r], gptr[r+1]) ∩ [0,B) — a
// contiguous range of graphs, hence a contiguous range of nodes.  One CTA per row r.
constexpr int GQ_COLS = 80, GQ_LANES = 12;  // 960 threads: 80 float4 column groups (D <= 320) x 12 row lanes
__global__ void __launch_bounds__(GQ_COLS * GQ_LANES)
gate_theta_bwd_q_kernel(const float* __restrict__ gpre, const float* __restrict__ xn, const int* __restrict__ gptr,
                        int B, int D, int dbl, float* __restrict__ gq) {
  const int r = blockIdx.x;  // row of q: < B when dbl (q is [B,D]); < N otherwise (q is [N,D])
  int n0 = 0, n1 = 0;
  if (dbl) {
    const int g0 = min(gptr[r], B), g1 = min(gptr[r + 1], B);
    n0 = gptr[g0];
    n1 = gptr[g1];
  } else if (r < B) {  // single gather q[batch[n]]: only rows < B are referenced
    n0 = gptr[r];
    n1 = gptr[r + 1];
  }
  // With the double gather a handful of rows own ~Nmax graphs' worth of nodes each, so the node range is
  // split over GQ_LANES row lanes (float4 columns, 4 loads in flight) and folded in a fixed order.
  __shared__ float4 red[GQ_LANES][GQ_COLS];
  const int cx = threadIdx.x % GQ_COLS, ly = threadIdx.x / GQ_COLS;
  const int d4 = D >> 2;
  if ((D & 3) == 0 && d4 <= GQ_COLS && (((uintptr_t)xn | (uintptr_t)gq) & 15) == 0) {
    float4 acc = f4_zero();
    if (cx < d4) {
      int n = n0 + ly;
      for (; n + 3 * GQ_LANES < n1; n += 4 * GQ_LANES) {
        float4 v[4];
        float gp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          gp[u] = gpre[n + u * GQ_LANES];
          v[u] = Vec4<float>::ld(xn + (int64_t)(n + u * GQ_LANES) * D + 4 * cx);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          acc = make_float4(fmaf(gp[u], v[u].x, acc.x), fmaf(gp[u], v[u].y, acc.y), fmaf(gp[u], v[u].z, acc.z),
                            fmaf(gp[u], v[u].w, acc.w));
      }
      for (; n < n1; n += GQ_LANES) {
        const float gp = gpre[n];
        const float4 v = Vec4<float>::ld(xn + (int64_t)n * D + 4 * cx);
        acc = make_float4(fmaf(gp, v.x, acc.x), fmaf(gp, v.y, acc.y), fmaf(gp, v.z, acc.z), fmaf(gp, v.w, acc.w));
      }
    }
    red[ly][cx] = acc;
    __syncthreads();
    if (ly == 0 && cx < d4) {
      float4 t = red[0][cx];
#pragma unroll
      for (int y = 1; y < GQ_LANES; ++y) t = f4_add(t, red[y][cx]);
      Vec4<float>::st(gq + (int64_t)r * D + 4 * cx, t);
    }
    return;
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int n = n0; n < n1; ++n) acc = fmaf(gpre[n], xn[(int64_t)n * D + c], acc);
    gq[(int64_t)r * D + c] = acc;
  }
}

// ---------------------------------------------------------------- scatter-SDPA + GraphNorm + residual
constexpr int SG_THREADS = 320;  // 10 warps; thread c owns channel c (D = 300) in the per-channel phases

__global__ void __launch_bounds__(SG_THREADS)
sdpa_graphnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ ins,
                          const float* __restrict__ h_in, const float* __restrict__ weight,
                          const float* __restrict__ bias, const float* __restrict__ mean_scale,
                          const int* __restrict__ gptr, int D, float eps, float* __restrict__ h_out,
                          float* __restrict__ a_out, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out) {
  extern __shared__ float sm[];  // [nmax] attention weights
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1], cnt = n1 - n0;
  if (cnt <= 0) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      mean_out[(int64_t)b * D + c] = 0.f;
      rstd_out[(int64_t)b * D + c] = 0.f;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* ib = ins + (int64_t)b * D;
  const float rs = 1.0f / sqrtf((float)D);
  // logits: warp per node
  for (int i = warp; i < cnt; i += nw) {
    const float* vr = v + (int64_t)(n0 + i) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(ib[c], vr[c], s);
    s = warp_sum(s);
    if (lane == 0) sm[i] = s * rs;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) mx = fmaxf(mx, sm[i]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float e = expf(sm[i] - mx);
    sm[i] = e;
    se += e;
  }
  se = block_sum(se, red);
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float a = sm[i] / se;  // torch_scatter.scatter_softmax: no epsilon
    sm[i] = a;
    a_out[n0 + i] = a;
  }
  __syncthreads();
  const float inv_cnt = 1.0f / (float)cnt;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s1 = 0.f;
    for (int i = 0; i < cnt; ++i) s1 = fmaf(sm[i], v[(int64_t)(n0 + i) * D + c], s1);
    const float mean = s1 * inv_cnt;
    const float shift = mean * mean_scale[c];
    float s2 = 0.f;
    for (int i = 0; i < cnt; ++i) {
      const float o = sm[i] * v[(int64_t)(n0 + i) * D + c] - shift;
      s2 = fmaf(o, o, s2);
    }
    const float rstd = 1.0f / sqrtf(s2 * inv_cnt + eps);
    mean_out[(int64_t)b * D + c] = mean;
    rstd_out[(int64_t)b * D + c] = rstd;
    const float w = weight[c] * rstd, bb = bias[c];
    for (int i = 0; i < cnt; ++i) {
      const int64_t o_ = (int64_t)(n0 + i) * D + c;
      const float o = sm[i] * v[o_] - shift;
      h_out[o_] = fmaf(w, o, bb) + h_in[o_];
    }
  }
}

__global__ void __launch_bounds__(SG_THREADS)
sdpa_graphnorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                          const float* __restrict__ ins, const float* __restrict__ weight,
                          const float* __restrict__ mean_scale, const float* __restrict__ a,
                          const float* __restrict__ mean, const float* __restrict__ rstd,
                          const int* __restrict__ gptr, int D, float* __restrict__ g_v,
                          float* __restrict__ g_ins, float* __restrict__ gw_part,
                          float* __restrict__ gb_part, float* __restrict__ gms_part,
                          const float* __restrict__ zg) {
  // dynamic smem: coefA[D], coefB[D], coefC[D], shift[D], sa[nmax], sga[nmax]
  extern __shared__ float sm[];
  __shared__ float red[32];
  float* cA = sm;
  float* cB = cA + D;
  float* cC = cB + D;
  float* sh = cC + D;
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1], cnt = n1 - n0;
  float* sa = sh + D;
  float* sga = sa + max(cnt, 0);
  if (cnt <= 0) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const int64_t o = (int64_t)b * D + c;
      g_ins[o] = 0.f;
      gw_part[o] = 0.f;
      gb_part[o] = 0.f;
      gms_part[o] = 0.f;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv_cnt = 1.0f / (float)cnt;
  const float rs = 1.0f / sqrtf((float)D);
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sa[i] = a[n0 + i];
  __syncthreads();
  // P1: per-channel sums -> coefficients of g_y = A*G + B*o + C
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const int64_t bc = (int64_t)b * D + c;
    const float mu = mean[bc], r = rstd[bc], ms = mean_scale[c], w = weight[c];
    const float shift = mu * ms;
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < cnt; ++i) {
      const int64_t o_ = (int64_t)(n0 + i) * D + c;
      const float G = g[o_];
      const float o = sa[i] * v[o_] - shift;
      s1 += G;
      s2 = fmaf(G, o, s2);
    }
    gb_part[bc] = s1;
    gw_part[bc] = s2 * r;
    const float g_var = -0.5f * (w * s2) * r * r * r;
    // sum_n g_o = w r S1 + g_var*(2/cnt)*sum_n o ;  sum_n o = cnt*mean*(1-ms)
    const float sum_go = w * r * s1 + g_var * 2.0f * mu * (1.0f - ms);
    gms_part[bc] = -mu * sum_go;
    cA[c] = w * r;
    cB[c] = g_var * 2.0f * inv_cnt;
    cC[c] = -ms * sum_go * inv_cnt;
    sh[c] = shift;
  }
  __syncthreads();
  // P2: g_a[n] = sum_c g_y[n,c] * v[n,c]   (warp per node)
  for (int i = warp; i < cnt; i += nw) {
    const int64_t ro = (int64_t)(n0 + i) * D;
    const float ai = sa[i];
    float s = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float vv = v[ro + c];
      const float gy = fmaf(cA[c], g[ro + c], fmaf(cB[c], ai * vv - sh[c], cC[c]));
      s = fmaf(gy, vv, s);
    }
    s = warp_sum(s);
    if (lane == 0) sga[i] = s;
  }
  __syncthreads();
  // P3: softmax backward  g_logit = a * (g_a - sum a*g_a)
  float dp = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) dp = fmaf(sa[i], sga[i], dp);
  dp = block_sum(dp, red);
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sga[i] = sa[i] * (sga[i] - dp) * rs;  // includes 1/sqrt(D)
  __syncthreads();
  // P4: g_v and g_ins
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float ic = ins[(int64_t)b * D + c];
    float gi = 0.f;
    for (int i = 0; i < cnt; ++i) {
      const int64_t o_ = (int64_t)(n0 + i) * D + c;
      const float vv = v[o_];
      const float gy = fmaf(cA[c], g[o_], fmaf(cB[c], sa[i] * vv - sh[c], cC[c]));
      const float gvv = fmaf(gy, sa[i], sga[i] * ic);
      g_v[o_] = zg ? gvv * gelu_grad_f(zg[o_]) : gvv;  // (zg: the closing GELU of the producer of v, fused)
      gi = fmaf(sga[i], vv, gi);
    }
    g_ins[(int64_t)b * D + c] = gi;
  }
}

// ---- D % 4 == 0 variants: float4 columns x SG_LANES row lanes ------------------------------------------------
// The kernels above give every thread one channel and walk the graph's nodes serially in each of their per-channel
// passes (three in the forward, two in the backward): chains of ~20-40 dependent L2 round trips, 20 us per launch
// at the c3 size for 12 MB.  Here thread (cx, ly) owns the float4 column cx and the rows ly, ly + SG_LANES, ...;
// per-lane partial sums are folded through shared memory in a fixed order (deterministic).
constexpr int SG_COLS = 80, SG_LANES = 4;  // 320 threads (8 lanes measured slower: fwd 15.4 -> 19.9 us, bwd 20.9 -> 24.5); D <= 320 per column pass (looped beyond)

__device__ __forceinline__ float4 sg_fold(float4 (*red)[SG_COLS], int cx, int ly, float4 v) {
  __syncthreads();  // protect `red` from the previous use
  red[ly][cx] = v;
  __syncthreads();
  float4 t = red[0][cx];
#pragma unroll
  for (int y = 1; y < SG_LANES; ++y) t = f4_add(t, red[y][cx]);
  return t;
}

__global__ void __launch_bounds__(SG_COLS * SG_LANES)
sdpa_graphnorm_fwd_v4_kernel(const float* __restrict__ v, const float* __restrict__ ins, const float* __restrict__ h_in,
                             const float* __restrict__ weight, const float* __restrict__ bias,
                             const float* __restrict__ mean_scale, const int* __restrict__ gptr, int D, float eps,
                             float* __restrict__ h_out, float* __restrict__ a_out, float* __restrict__ mean_out,
                             float* __restrict__ rstd_out) {
  pdl_enter();
  extern __shared__ float sm[];  // [nmax] attention weights
  __shared__ float red1[32];
  __shared__ float4 red[SG_LANES][SG_COLS];
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1], cnt = n1 - n0;
  const int d4 = D >> 2;
  if (cnt <= 0) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      mean_out[(int64_t)b * D + c] = 0.f;
      rstd_out[(int64_t)b * D + c] = 0.f;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* ib = ins + (int64_t)b * D;
  const float rs = 1.0f / sqrtf((float)D);
  for (int i = warp; i < cnt; i += nw) {  // logits: warp per node, same arithmetic order as the scalar kernel
    const float* vr = v + (int64_t)(n0 + i) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(ib[c], vr[c], s);
    s = warp_sum(s);
    if (lane == 0) sm[i] = s * rs;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) mx = fmaxf(mx, sm[i]);
  mx = block_max(mx, red1);
  float se = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float e = expf(sm[i] - mx);
    sm[i] = e;
    se += e;
  }
  se = block_sum(se, red1);
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float a = sm[i] / se;  // torch_scatter.scatter_softmax: no epsilon
    sm[i] = a;
    a_out[n0 + i] = a;
  }
  __syncthreads();
  const float inv_cnt = 1.0f / (float)cnt;
  const int cx = threadIdx.x % SG_COLS, ly = threadIdx.x / SG_COLS;
  for (int c0 = 0; c0 < d4; c0 += SG_COLS) {
    const int c4 = c0 + cx;
    const bool on = c4 < d4;
    const int64_t cb = 4 * (int64_t)c4;
    float4 s1 = f4_zero();
    if (on)
      for (int i = ly; i < cnt; i += SG_LANES) s1 = f4_fma(Vec4<float>::ld(v + (int64_t)(n0 + i) * D + cb), sm[i], s1);
    s1 = sg_fold(red, cx, ly, s1);
    float4 mean = f4_zero(), shift = f4_zero(), s2 = f4_zero();
    if (on) {
      mean = f4_scale(s1, inv_cnt);
      const float4 ms = Vec4<float>::ld(mean_scale + cb);
      shift = make_float4(mean.x * ms.x, mean.y * ms.y, mean.z * ms.z, mean.w * ms.w);
      for (int i = ly; i < cnt; i += SG_LANES) {
        const float4 vv = Vec4<float>::ld(v + (int64_t)(n0 + i) * D + cb);
        const float a = sm[i];
        const float4 o = make_float4(a * vv.x - shift.x, a * vv.y - shift.y, a * vv.z - shift.z, a * vv.w - shift.w);
        s2 = make_float4(fmaf(o.x, o.x, s2.x), fmaf(o.y, o.y, s2.y), fmaf(o.z, o.z, s2.z), fmaf(o.w, o.w, s2.w));
      }
    }
    s2 = sg_fold(red, cx, ly, s2);
    if (on) {
      const float4 rstd = make_float4(1.0f / sqrtf(s2.x * inv_cnt + eps), 1.0f / sqrtf(s2.y * inv_cnt + eps),
                                      1.0f / sqrtf(s2.z * inv_cnt + eps), 1.0f / sqrtf(s2.w * inv_cnt + eps));
      if (ly == 0) {
        Vec4<float>::st(mean_out + (int64_t)b * D + cb, mean);
        Vec4<float>::st(rstd_out + (int64_t)b * D + cb, rstd);
      }
      const float4 wv = Vec4<float>::ld(weight + cb), bb = Vec4<float>::ld(bias + cb);
      const float4 w = make_float4(wv.x * rstd.x, wv.y * rstd.y, wv.z * rstd.z, wv.w * rstd.w);
      for (int i = ly; i < cnt; i += SG_LANES) {
        const int64_t o_ = (int64_t)(n0 + i) * D + cb;
        const float4 vv = Vec4<float>::ld(v + o_), hi = Vec4<float>::ld(h_in + o_);
        const float a = sm[i];
        Vec4<float>::st(h_out + o_, make_float4(fmaf(w.x, a * vv.x - shift.x, bb.x) + hi.x, fmaf(w.y, a * vv.y - shift.y, bb.y) + hi.y,
                                                fmaf(w.z, a * vv.z - shift.z, bb.z) + hi.z, fmaf(w.w, a * vv.w - shift.w, bb.w) + hi.w));
      }
    }
  }
}

__global__ void __launch_bounds__(SG_COLS * SG_LANES)
sdpa_graphnorm_bwd_v4_kernel(const float* __restrict__ g, const float* __restrict__ v, const float* __restrict__ ins,
                             const float* __restrict__ weight, const float* __restrict__ mean_scale,
                             const float* __restrict__ a, const float* __restrict__ mean, const float* __restrict__ rstd,
                             const int* __restrict__ gptr, int D, float* __restrict__ g_v, float* __restrict__ g_ins,
                             float* __restrict__ gw_part, float* __restrict__ gb_part, float* __restrict__ gms_part,
                             const float* __restrict__ zg) {
  pdl_enter();
  // dynamic smem: coefA[D], coefB[D], coefC[D], shift[D], sa[cnt], sga[cnt]
  extern __shared__ __align__(16) float sm4[];
  __shared__ float red1[32];
  __shared__ float4 red[SG_LANES][SG_COLS];
  float* cA = sm4;
  float* cB = cA + D;
  float* cC = cB + D;
  float* sh = cC + D;
  const int b = blockIdx.x;
  const int n0 = gptr[b], n1 = gptr[b + 1], cnt = n1 - n0;
  const int d4 = D >> 2;
  float* sa = sh + D;
  float* sga = sa + max(cnt, 0);
  if (cnt <= 0) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const int64_t o = (int64_t)b * D + c;
      g_ins[o] = 0.f;
      gw_part[o] = 0.f;
      gb_part[o] = 0.f;
      gms_part[o] = 0.f;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv_cnt = 1.0f / (float)cnt;
  const float rs = 1.0f / sqrtf((float)D);
  const int cx = threadIdx.x % SG_COLS, ly = threadIdx.x / SG_COLS;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sa[i] = a[n0 + i];
  __syncthreads();
  // P1: per-channel sums -> coefficients of g_y = A*G + B*o + C
  for (int c0 = 0; c0 < d4; c0 += SG_COLS) {
    const int c4 = c0 + cx;
    const bool on = c4 < d4;
    const int64_t cb = 4 * (int64_t)c4, bc = (int64_t)b * D + cb;
    float4 mu = f4_zero(), r = f4_zero(), ms = f4_zero(), w = f4_zero(), shift = f4_zero(), s1 = f4_zero(), s2 = f4_zero();
    if (on) {
      mu = Vec4<float>::ld(mean + bc);
      r = Vec4<float>::ld(rstd + bc);
      ms = Vec4<float>::ld(mean_scale + cb);
      w = Vec4<float>::ld(weight + cb);
      shift = make_float4(mu.x * ms.x, mu.y * ms.y, mu.z * ms.z, mu.w * ms.w);
      for (int i = ly; i < cnt; i += SG_LANES) {
        const int64_t o_ = (int64_t)(n0 + i) * D + cb;
        const float4 G = Vec4<float>::ld(g + o_), vv = Vec4<float>::ld(v + o_);
        const float ai = sa[i];
        s1 = f4_add(s1, G);
        s2 = make_float4(fmaf(G.x, ai * vv.x - shift.x, s2.x), fmaf(G.y, ai * vv.y - shift.y, s2.y),
                         fmaf(G.z, ai * vv.z - shift.z, s2.z), fmaf(G.w, ai * vv.w - shift.w, s2.w));
      }
    }
    s1 = sg_fold(red, cx, ly, s1);
    s2 = sg_fold(red, cx, ly, s2);
    if (on && ly == 0) {
      const float s1v[4] = {s1.x, s1.y, s1.z, s1.w}, s2v[4] = {s2.x, s2.y, s2.z, s2.w};
      const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rv[4] = {r.x, r.y, r.z, r.w}, msv[4] = {ms.x, ms.y, ms.z, ms.w};
      const float wv[4] = {w.x, w.y, w.z, w.w}, shv[4] = {shift.x, shift.y, shift.z, shift.w};
      float gbv[4], gwv[4], gmsv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gbv[j] = s1v[j];
        gwv[j] = s2v[j] * rv[j];
        const float g_var = -0.5f * (wv[j] * s2v[j]) * rv[j] * rv[j] * rv[j];
        // sum_n g_o = w r S1 + g_var*(2/cnt)*sum_n o ;  sum_n o = cnt*mean*(1-ms)
        const float sum_go = wv[j] * rv[j] * s1v[j] + g_var * 2.0f * muv[j] * (1.0f - msv[j]);
        gmsv[j] = -muv[j] * sum_go;
        const int c = 4 * c4 + j;
        cA[c] = wv[j] * rv[j];
        cB[c] = g_var * 2.0f * inv_cnt;
        cC[c] = -msv[j] * sum_go * inv_cnt;
        sh[c] = shv[j];
      }
      Vec4<float>::st(gb_part + bc, make_float4(gbv[0], gbv[1], gbv[2], gbv[3]));
      Vec4<float>::st(gw_part + bc, make_float4(gwv[0], gwv[1], gwv[2], gwv[3]));
      Vec4<float>::st(gms_part + bc, make_float4(gmsv[0], gmsv[1], gmsv[2], gmsv[3]));
    }
  }
  __syncthreads();
  // P2: g_a[n] = sum_c g_y[n,c] * v[n,c]   (warp per node)
  for (int i = warp; i < cnt; i += nw) {
    const int64_t ro = (int64_t)(n0 + i) * D;
    const float ai = sa[i];
    float s = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float vv = v[ro + c];
      const float gy = fmaf(cA[c], g[ro + c], fmaf(cB[c], ai * vv - sh[c], cC[c]));
      s = fmaf(gy, vv, s);
    }
    s = warp_sum(s);
    if (lane == 0) sga[i] = s;
  }
  __syncthreads();
  // P3: softmax backward  g_logit = a * (g_a - sum a*g_a)
  float dp = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) dp = fmaf(sa[i], sga[i], dp);
  dp = block_sum(dp, red1);
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sga[i] = sa[i] * (sga[i] - dp) * rs;  // includes 1/sqrt(D)
  __syncthreads();
  // P4: g_v and g_ins
  for (int c0 = 0; c0 < d4; c0 += SG_COLS) {
    const int c4 = c0 + cx;
    const bool on = c4 < d4;
    const int64_t cb = 4 * (int64_t)c4;
    float4 gi = f4_zero();
    if (on) {
      const float4 ic = Vec4<float>::ld(ins + (int64_t)b * D + cb);
      const float4 A = *reinterpret_cast<const float4*>(cA + cb), Bc = *reinterpret_cast<const float4*>(cB + cb);
      const float4 Cc = *reinterpret_cast<const float4*>(cC + cb), S = *reinterpret_cast<const float4*>(sh + cb);
      for (int i = ly; i < cnt; i += SG_LANES) {
        const int64_t o_ = (int64_t)(n0 + i) * D + cb;
        const float4 vv = Vec4<float>::ld(v + o_), G = Vec4<float>::ld(g + o_);
        const float ai = sa[i], gl = sga[i];
        const float4 gy = make_float4(fmaf(A.x, G.x, fmaf(Bc.x, ai * vv.x - S.x, Cc.x)), fmaf(A.y, G.y, fmaf(Bc.y, ai * vv.y - S.y, Cc.y)),
                                      fmaf(A.z, G.z, fmaf(Bc.z, ai * vv.z - S.z, Cc.z)), fmaf(A.w, G.w, fmaf(Bc.w, ai * vv.w - S.w, Cc.w)));
        float4 gvv = make_float4(fmaf(gy.x, ai, gl * ic.x), fmaf(gy.y, ai, gl * ic.y), fmaf(gy.z, ai, gl * ic.z),
                                 fmaf(gy.w, ai, gl * ic.w));
        if (zg) {  // the closing GELU of the projection that produced v (x_proj[2]), fused: g_z = g_v * gelu'(z)
          const float4 z = Vec4<float>::ld(zg + o_);
          gvv = make_float4(gvv.x * gelu_grad_f(z.x), gvv.y * gelu_grad_f(z.y), gvv.z * gelu_grad_f(z.z), gvv.w * gelu_grad_f(z.w));
        }
        Vec4<float>::st(g_v + o_, gvv);
        gi = f4_fma(vv, gl, gi);
      }
    }
    gi = sg_fold(red, cx, ly, gi);
    if (on && ly == 0) Vec4<float>::st(g_ins + (int64_t)b * D + cb, gi);
  }
}

// ---------------------------------------------------------------- misc

// ---------------------------------------------------------------- masked attention pooling (§8 f1)
// GlobalAttention.forward (reference models/att_pooling.py:57-77) after its node_nn / ques_nn projections:
//   xm = x * node_mask;  l_n = <xm_n, q[b]> / sqrt(D);  a = softmax over the nodes of graph b
//   (torch_geometric.utils.softmax: exp(l - max) / (sum + 1e-16));  out[b] = sum_n a_n * xm_n.
// One CTA per graph (batch is sorted, nodes of a graph are contiguous); logits live in shared memory.
constexpr int AP_THREADS = 256;

__global__ void __launch_bounds__(AP_THREADS)
attn_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ q,
                     const int* __restrict__ gptr, int D, float* __restrict__ out, float* __restrict__ gate) {
  extern __shared__ float ap_smem[];
  __shared__ float red[32];
  float* lg = ap_smem;  // [n_b] logits, then attention weights
  const int b = blockIdx.x;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = AP_THREADS / 32;
  const int d4 = D >> 2;
  const float inv_sqrt_d = 1.0f / sqrtf((float)D);
  const float* qb = q + (size_t)b * D;
  for (int n = warp; n < nb; n += nwarps) {
    const float* xr = x + (size_t)(n0 + n) * D;
    const float m = mask ? mask[n0 + n] : 1.f;
    float acc = 0.f;
    for (int v = lane; v < d4; v += 32) {
      const float4 xv = Vec4<float>::ld(xr + 4 * v), qv = Vec4<float>::ld(qb + 4 * v);
      acc += (xv.x * m) * qv.x + (xv.y * m) * qv.y + (xv.z * m) * qv.z + (xv.w * m) * qv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) lg[n] = acc * inv_sqrt_d;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < nb; n += AP_THREADS) mx = fmaxf(mx, lg[n]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int n = threadIdx.x; n < nb; n += AP_THREADS) {
    const float e = expf(lg[n] - mx);
    lg[n] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / (sum + 1e-16f);
  __syncthreads();
  for (int n = threadIdx.x; n < nb; n += AP_THREADS) {
    const float a = lg[n] * inv;
    lg[n] = a;
    gate[n0 + n] = a;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < d4; v += AP_THREADS) {
    float4 o = f4_zero();
    for (int n = 0; n < nb; ++n) {
      const float w = lg[n] * (mask ? mask[n0 + n] : 1.f);
      o = f4_fma(Vec4<float>::ld(x + (size_t)(n0 + n) * D + 4 * v), w, o);
    }
    Vec4<float>::st(out + (size_t)b * D + 4 * v, o);
  }
}

// backward: g_out [B,D], g_gate [N] or NULL -> g_x [N,D], g_mask [N] (or NULL), g_q [B,D]
__global__ void __launch_bounds__(AP_THREADS)
attn_pool_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ g_gate,
                     const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ q,
                     const float* __restrict__ gate, const int* __restrict__ gptr, int D,
                     float* __restrict__ g_x, float* __restrict__ g_mask, float* __restrict__ g_q) {
  extern __shared__ float ap_smem[];
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int n0 = gptr[b], nb = gptr[b + 1] - n0;
  float* ga = ap_smem;       // [n_b] dL/da_n, then dL/dl_n
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = AP_THREADS / 32;
  const int d4 = D >> 2;
  const float inv_sqrt_d = 1.0f / sqrtf((float)D);
  const float* qb = q + (size_t)b * D;
  const float* gob = g_out + (size_t)b * D;
  for (int n = warp; n < nb; n += nwarps) {
    const float* xr = x + (size_t)(n0 + n) * D;
    const float m = mask ? mask[n0 + n] : 1.f;
    float acc = 0.f;
    for (int v = lane; v < d4; v += 32) {
      const float4 xv = Vec4<float>::ld(xr + 4 * v), gv = Vec4<float>::ld(gob + 4 * v);
      acc += (xv.x * m) * gv.x + (xv.y * m) * gv.y + (xv.z * m) * gv.z + (xv.w * m) * gv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) ga[n] = acc + (g_gate ? g_gate[n0 + n] : 0.f);
  }
  __syncthreads();
  float dot = 0.f;
  for (int n = threadIdx.x; n < nb; n += AP_THREADS) dot += gate[n0 + n] * ga[n];
  dot = block_sum(dot, red);
  __syncthreads();
  for (int n = threadIdx.x; n < nb; n += AP_THREADS) ga[n] = gate[n0 + n] * (ga[n] - dot);  // dL/dl_n
  __syncthreads();
  for (int n = warp; n < nb; n += nwarps) {
    const float* xr = x + (size_t)(n0 + n) * D;
    const float m = mask ? mask[n0 + n] : 1.f;
    const float a = gate[n0 + n], gl = ga[n] * inv_sqrt_d;
    float gm = 0.f;
    for (int v = lane; v < d4; v += 32) {
      const float4 xv = Vec4<float>::ld(xr + 4 * v), gv = Vec4<float>::ld(gob + 4 * v), qv = Vec4<float>::ld(qb + 4 * v);
      const float4 gxm = make_float4(fmaf(gl, qv.x, a * gv.x), fmaf(gl, qv.y, a * gv.y), fmaf(gl, qv.z, a * gv.z),
                                     fmaf(gl, qv.w, a * gv.w));
      gm += gxm.x * xv.x + gxm.y * xv.y + gxm.z * xv.z + gxm.w * xv.w;
      Vec4<float>::st(g_x + (size_t)(n0 + n) * D + 4 * v, f4_scale(gxm, m));
    }
    gm = warp_sum(gm);
    if (g_mask && lane == 0) g_mask[n0 + n] = gm;
  }
  for (int v = threadIdx.x; v < d4; v += AP_THREADS) {
    float4 o = f4_zero();
    for (int n = 0; n < nb; ++n) {
      const float w = ga[n] * inv_sqrt_d * (mask ? mask[n0 + n] : 1.f);
      o = f4_fma(Vec4<float>::ld(x + (size_t)(n0 + n) * D + 4 * v), w, o);
    }
    Vec4<float>::st(g_q + (size_t)b * D + 4 * v, o);
  }
}

// gz may alias gy (the layer executor runs it in place): no __restrict__ on those two
__global__ void gelu_bwd_kernel(const float* gy, const float* __restrict__ z, float* gz, int64_t n) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) gz[i] = gy[i] * gelu_grad_f(z[i]);
}

constexpr int CS_ROWS = 64;  // rows per partial
// Each thread owns 4 adjacent columns (one float4) of a CS_ROWS-row slab: 16-byte coalesced loads,
// CS_ROWS independent loads per thread, grid = (cols/4/128) x (rows/CS_ROWS) CTAs.
__global__ void colsum_partial_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int cols,
                                      float* __restrict__ part) {
  pdl_enter();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t r1 = min(rows, r0 + CS_ROWS);
  float4 s = f4_zero();
  int64_t r = r0;
  for (; r + 8 <= r1; r += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = Vec4<float>::ld(in + (r + u) * ld + c);
#pragma unroll
    for (int u = 0; u < 8; ++u) s = f4_add(s, v[u]);
  }
  for (; r < r1; ++r) s = f4_add(s, Vec4<float>::ld(in + r * ld + c));
  Vec4<float>::st(part + (int64_t)blockIdx.y * cols + c, s);
}
// Second stage: block (64 column groups, CS_LANES row lanes); lane y sums partial rows y, y + CS_LANES, ...
// (4 independent loads in flight), then a fixed-order shared-memory fold over the lanes -> deterministic.
// (The first version walked all `nparts` rows serially in one thread per column: 15 us for the 622 partial
// rows of an [E, 1200] bias gradient, more than the first stage.)
constexpr int CS_LANES = 16;
__global__ void __launch_bounds__(64 * CS_LANES)
colsum_final_kernel(const float* __restrict__ part, int nparts, int cols, float* __restrict__ out) {
  pdl_enter();
  __shared__ float4 red[CS_LANES][64];
  const int c = (blockIdx.x * 64 + threadIdx.x) * 4;
  float4 s = f4_zero();
  if (c < cols) {
    int p = threadIdx.y;
    for (; p + 3 * CS_LANES < nparts; p += 4 * CS_LANES) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = Vec4<float>::ld(part + (int64_t)(p + u * CS_LANES) * cols + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) s = f4_add(s, v[u]);
    }
    for (; p < nparts; p += CS_LANES) s = f4_add(s, Vec4<float>::ld(part + (int64_t)p * cols + c));
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < CS_LANES; ++y) t = f4_add(t, red[y][threadIdx.x]);
    Vec4<float>::st(out + c, t);
  }
}

// ---- batched column sums: the bias / affine gradients of one layer's backward in two launches ----------------
// (the layer executor used to issue up to nine isg_colsum calls = 18 latency-bound launches per layer, 0.43 ms of
// the 5.3 ms c3 step).  Same slabs, same fold order as colsum_partial/final, so results are bit-identical.
constexpr int CS_MAX_JOBS = 12;
struct ColsumJob {
  const void* in;  // fp32, or bf16 when `bf16` is set (4 columns = one 8-byte load)
  float* out;
  float* part;
  int64_t ld;
  int64_t rows;
  int cols, parts, cblocks, block0, fblock0, bf16;
};
struct ColsumBatch {
  ColsumJob job[CS_MAX_JOBS];
  int n;
};

__global__ void __launch_bounds__(64) colsum_multi_partial_kernel(const __grid_constant__ ColsumBatch batch) {
  pdl_enter();
  int j = 0;
  while (j + 1 < batch.n && (int)blockIdx.x >= batch.job[j + 1].block0) ++j;
  const ColsumJob& J = batch.job[j];
  const int local = blockIdx.x - J.block0;
  const int cb = local % J.cblocks, slab = local / J.cblocks;
  const int c = (cb * 64 + threadIdx.x) * 4;
  if (c >= J.cols) return;
  const int64_t r0 = (int64_t)slab * CS_ROWS;
  const int64_t r1 = min(J.rows, r0 + CS_ROWS);
  const int64_t ld = J.ld;
  float4 s = f4_zero();
  int64_t r = r0;
  if (J.bf16) {
    const __nv_bfloat16* in = (const __nv_bfloat16*)J.in;
    for (; r + 8 <= r1; r += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = Vec4<__nv_bfloat16>::ld(in + (r + u) * ld + c);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = f4_add(s, v[u]);
    }
    for (; r < r1; ++r) s = f4_add(s, Vec4<__nv_bfloat16>::ld(in + r * ld + c));
  } else {
    const float* in = (const float*)J.in;
    for (; r + 8 <= r1; r += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = Vec4<float>::ld(in + (r + u) * ld + c);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = f4_add(s, v[u]);
    }
    for (; r < r1; ++r) s = f4_add(s, Vec4<float>::ld(in + r * ld + c));
  }
  Vec4<float>::st(J.part + (int64_t)slab * J.cols + c, s);
}

__global__ void __launch_bounds__(64 * CS_LANES) colsum_multi_final_kernel(const __grid_constant__ ColsumBatch batch) {
  pdl_enter();
  __shared__ float4 red[CS_LANES][64];
  int j = 0;
  while (j + 1 < batch.n && (int)blockIdx.x >= batch.job[j + 1].fblock0) ++j;
  const ColsumJob& J = batch.job[j];
  const int c = ((blockIdx.x - J.fblock0) * 64 + threadIdx.x) * 4;
  const int cols = J.cols, nparts = J.parts;
  const float* part = J.part;
  float4 s = f4_zero();
  if (c < cols) {
    int p = threadIdx.y;
    for (; p + 3 * CS_LANES < nparts; p += 4 * CS_LANES) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = Vec4<float>::ld(part + (int64_t)(p + u * CS_LANES) * cols + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) s = f4_add(s, v[u]);
    }
    for (; p < nparts; p += CS_LANES) s = f4_add(s, Vec4<float>::ld(part + (int64_t)p * cols + c));
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < CS_LANES; ++y) t = f4_add(t, red[y][threadIdx.x]);
    Vec4<float>::st(J.out + c, t);
  }
}

}  // namespace

extern "C" int isg_instr_gate_fwd(const float* x, const float* ins, const int32_t* batch32, int64_t N, int D,
                                  float* y, void* stream_) {
  if (N < 0 || D <= 0) return ISG_EINVAL;
  if (D % 4 != 0) return ISG_EUNSUPPORTED;
  if (N == 0) return ISG_OK;
  if (!x || !ins || !batch32 || !y) return ISG_EINVAL;
  const int64_t total = N * (D / 4);
  cudaError_t le = isg::launch_pdl(instr_gate_fwd_kernel, dim3((unsigned)isg::ceil_div(total, 256)), dim3(256), 0,
                                   (cudaStream_t)stream_, x, ins, batch32, N, D / 4, y);
  if (le != cudaSuccess) return (int)le;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_instr_gate_bwd(const float* g_y, const float* x, const float* ins, const int32_t* gptr,
                                  int64_t B, int D, const float* g_residual, int accumulate_ins, float* g_x,
                                  float* g_ins, void* stream_) {
  if (B < 0 || D <= 0) return ISG_EINVAL;
  if (B == 0) return ISG_OK;
  if (!g_y || !x || !ins || !gptr || !g_x || !g_ins) return ISG_EINVAL;
  const bool v4 = D % 4 == 0 && !(((uintptr_t)g_y | (uintptr_t)x | (uintptr_t)ins | (uintptr_t)g_x | (uintptr_t)g_ins |
                                   (uintptr_t)g_residual) & 15);
  if (v4) {
    cudaError_t le = isg::launch_pdl(instr_gate_bwd_v4_kernel, dim3((unsigned)B), dim3(IG_COLS * IG_LANES), 0,
                                     (cudaStream_t)stream_, g_y, x, ins, gptr, D, g_residual, accumulate_ins, g_x, g_ins);
    if (le != cudaSuccess) return (int)le;
  } else
    instr_gate_bwd_kernel<<<(unsigned)B, 320, 0, (cudaStream_t)stream_>>>(g_y, x, ins, gptr, D, g_residual,
                                                                          accumulate_ins, g_x, g_ins);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_concat_instr_fwd(const float* x, const float* instruction, const int32_t* batch32, int64_t N, int D,
                                    float* y, void* stream_) {
  if (N < 0 || D <= 0) return ISG_EINVAL;
  if (N == 0) return ISG_OK;
  if (!x || !instruction || !batch32 || !y) return ISG_EINVAL;
  concat_instr_fwd_kernel<<<(unsigned)isg::ceil_div(N * 2 * D, 256), 256, 0, (cudaStream_t)stream_>>>(x, instruction, batch32,
                                                                                                  N, D, y);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_concat_instr_bwd(const float* g_y, const int32_t* gptr, int64_t B, int D, const float* g_residual,
                                    int accumulate_ins, float* g_x, float* g_ins, void* stream_) {
  if (B < 0 || D <= 0) return ISG_EINVAL;
  if (B == 0) return ISG_OK;
  if (!g_y || !gptr || !g_x || !g_ins) return ISG_EINVAL;
  concat_instr_bwd_kernel<<<(unsigned)B, 320, 0, (cudaStream_t)stream_>>>(g_y, gptr, D, g_residual, accumulate_ins, g_x, g_ins);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_gate_theta_fwd(const float* xn, const float* q, const int32_t* batch32, int64_t N, int D,
                                  int double_gather, const float* keep, float* theta, void* stream_) {
  if (N < 0 || D <= 0) return ISG_EINVAL;
  if (N == 0) return ISG_OK;
  if (!xn || !q || !batch32 || !theta) return ISG_EINVAL;
  gate_theta_fwd_kernel<<<isg::ceil_div(N * 32, 128), 128, 0, (cudaStream_t)stream_>>>(xn, q, batch32, N, D, double_gather, keep, theta);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_gate_theta_bwd(const float* g_theta, const float* xn, const float* q, const int32_t* batch32,
                                  const int32_t* gptr, int64_t N, int64_t B, int D, int double_gather,
                                  const float* keep, float* g_xn, float* g_q, float* scratch, void* stream_) {
  if (N < 0 || B < 0 || D <= 0) return ISG_EINVAL;
  if (N == 0 || B == 0) return ISG_OK;
  if (!g_theta || !xn || !q || !batch32 || !gptr || !g_xn || !g_q || !scratch) return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  gate_theta_bwd_xn_kernel<<<isg::ceil_div(N * 32, 128), 128, 0, stream>>>(g_theta, xn, q, batch32, N, D, double_gather, keep, g_xn, scratch);
  ISG_CHECK_LAUNCH();
  gate_theta_bwd_q_kernel<<<(unsigned)(double_gather ? B : N), GQ_COLS * GQ_LANES, 0, stream>>>(scratch, xn, gptr, (int)B, D, double_gather, g_q);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_sdpa_graphnorm_fwd(const float* v, const float* ins, const float* h_in, const float* weight,
                                      const float* bias, const float* mean_scale, const int32_t* gptr, int64_t B,
                                      int D, int nmax, float eps, float* h_out, float* a, float* mean,
                                      float* rstd, void* stream_) {
  if (B < 0 || D <= 0 || nmax < 0) return ISG_EINVAL;
  if (B == 0) return ISG_OK;
  if (!v || !ins || !h_in || !weight || !bias || !mean_scale || !gptr || !h_out || !a || !mean || !rstd)
    return ISG_EINVAL;
  const size_t smem = (size_t)(nmax > 0 ? nmax : 1) * sizeof(float);
  if (smem > 200 * 1024) return ISG_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sdpa_graphnorm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const bool v4 = D % 4 == 0 && !(((uintptr_t)v | (uintptr_t)ins | (uintptr_t)h_in | (uintptr_t)weight | (uintptr_t)bias |
                                   (uintptr_t)mean_scale | (uintptr_t)h_out | (uintptr_t)mean | (uintptr_t)rstd) & 15);
  if (v4) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(sdpa_graphnorm_fwd_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    cudaError_t le = isg::launch_pdl(sdpa_graphnorm_fwd_v4_kernel, dim3((unsigned)B), dim3(SG_COLS * SG_LANES), smem,
                                     (cudaStream_t)stream_, v, ins, h_in, weight, bias, mean_scale, gptr, D, eps, h_out, a,
                                     mean, rstd);
    if (le != cudaSuccess) return (int)le;
    ISG_CHECK_LAUNCH();
    return ISG_OK;
  }
  sdpa_graphnorm_fwd_kernel<<<(unsigned)B, SG_THREADS, smem, (cudaStream_t)stream_>>>(
      v, ins, h_in, weight, bias, mean_scale, gptr, D, eps, h_out, a, mean, rstd);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_sdpa_graphnorm_bwd(const float* g_out, const float* v, const float* ins, const float* weight,
                                      const float* mean_scale, const float* a, const float* mean,
                                      const float* rstd, const int32_t* gptr, int64_t B, int D, int nmax,
                                      float* g_v, float* g_ins, float* gw_part, float* gb_part, float* gms_part,
                                      const float* z_gelu, void* stream_) {
  if (B < 0 || D <= 0 || nmax < 0) return ISG_EINVAL;
  if (B == 0) return ISG_OK;
  if (!g_out || !v || !ins || !weight || !mean_scale || !a || !mean || !rstd || !gptr || !g_v || !g_ins ||
      !gw_part || !gb_part || !gms_part)
    return ISG_EINVAL;
  const size_t smem = ((size_t)4 * D + 2 * (size_t)(nmax > 0 ? nmax : 1)) * sizeof(float);
  if (smem > 200 * 1024) return ISG_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sdpa_graphnorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const bool v4 = D % 4 == 0 && !(((uintptr_t)g_out | (uintptr_t)v | (uintptr_t)ins | (uintptr_t)weight | (uintptr_t)mean_scale |
                                   (uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)g_v | (uintptr_t)g_ins | (uintptr_t)gw_part |
                                   (uintptr_t)gb_part | (uintptr_t)gms_part | (uintptr_t)z_gelu) & 15);
  if (v4) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(sdpa_graphnorm_bwd_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    cudaError_t le = isg::launch_pdl(sdpa_graphnorm_bwd_v4_kernel, dim3((unsigned)B), dim3(SG_COLS * SG_LANES), smem,
                                     (cudaStream_t)stream_, g_out, v, ins, weight, mean_scale, a, mean, rstd, gptr, D, g_v,
                                     g_ins, gw_part, gb_part, gms_part, z_gelu);
    if (le != cudaSuccess) return (int)le;
    ISG_CHECK_LAUNCH();
    return ISG_OK;
  }
  sdpa_graphnorm_bwd_kernel<<<(unsigned)B, SG_THREADS, smem, (cudaStream_t)stream_>>>(
      g_out, v, ins, weight, mean_scale, a, mean, rstd, gptr, D, g_v, g_ins, gw_part, gb_part, gms_part, z_gelu);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_attn_pool_fwd(const float* x, const float* node_mask, const float* q, const int32_t* gptr,
                                 int64_t B, int D, int nmax, float* out, float* gate, void* stream_) {
  if (B < 0 || D <= 0 || nmax < 0) return ISG_EINVAL;
  if (D % 4 != 0) return ISG_EUNSUPPORTED;
  if (B == 0) return ISG_OK;
  if (!x || !q || !gptr || !out || !gate) return ISG_EINVAL;
  const size_t smem = (size_t)(nmax > 0 ? nmax : 1) * sizeof(float);
  if (smem > 200 * 1024) return ISG_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attn_pool_fwd_kernel<<<(unsigned)B, AP_THREADS, smem, (cudaStream_t)stream_>>>(x, node_mask, q, gptr, D, out, gate);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_attn_pool_bwd(const float* g_out, const float* g_gate, const float* x, const float* node_mask,
                                 const float* q, const float* gate, const int32_t* gptr, int64_t B, int D, int nmax,
                                 float* g_x, float* g_mask, float* g_q, void* stream_) {
  if (B < 0 || D <= 0 || nmax < 0) return ISG_EINVAL;
  if (D % 4 != 0) return ISG_EUNSUPPORTED;
  if (B == 0) return ISG_OK;
  if (!g_out || !x || !q || !gate || !gptr || !g_x || !g_q) return ISG_EINVAL;
  if ((node_mask == nullptr) != (g_mask == nullptr)) return ISG_EINVAL;
  const size_t smem = (size_t)(nmax > 0 ? nmax : 1) * sizeof(float);
  if (smem > 200 * 1024) return ISG_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attn_pool_bwd_kernel<<<(unsigned)B, AP_THREADS, smem, (cudaStream_t)stream_>>>(g_out, g_gate, x, node_mask, q, gate,
                                                                               gptr, D, g_x, g_mask, g_q);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_gelu_bwd(const float* g_y, const float* z, float* g_z, int64_t n, void* stream_) {
  if (n < 0) return ISG_EINVAL;
  if (n == 0) return ISG_OK;
  if (!g_y || !z || !g_z) return ISG_EINVAL;
  cudaError_t le = isg::launch_pdl(gelu_bwd_kernel, dim3((unsigned)isg::ceil_div(n, 256)), dim3(256), 0,
                                   (cudaStream_t)stream_, g_y, z, g_z, n);
  if (le != cudaSuccess) return (int)le;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" size_t isg_colsum_workspace_bytes(int64_t rows, int cols) {
  if (cols <= 0) return 0;  // (isg_colsum rejects it; no size wraps around here)
  const int64_t parts = (rows + CS_ROWS - 1) / CS_ROWS;
  return (size_t)(parts > 0 ? parts : 1) * (size_t)cols * sizeof(float);
}

extern "C" int isg_colsum(const float* in, int64_t ld, int64_t rows, int cols, float* out, void* workspace,
                          size_t ws_bytes, void* stream_) {
  if (rows < 0 || cols <= 0 || !out) return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows == 0) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), stream);
    return e == cudaSuccess ? ISG_OK : (int)e;
  }
  if (!in) return ISG_EINVAL;
  if (cols % 4 != 0 || ld % 4 != 0 || ((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return ISG_EUNSUPPORTED;
  if (ws_bytes < isg_colsum_workspace_bytes(rows, cols) || !workspace) return ISG_EWORKSPACE;
  const int parts = (int)((rows + CS_ROWS - 1) / CS_ROWS);
  cudaError_t le = isg::launch_pdl(colsum_partial_kernel, dim3(isg::ceil_div(cols / 4, 64), parts), dim3(64), 0, stream, in,
                                   ld, rows, cols, (float*)workspace);
  if (le != cudaSuccess) return (int)le;
  ISG_CHECK_LAUNCH();
  le = isg::launch_pdl(colsum_final_kernel, dim3(isg::ceil_div(cols / 4, 64)), dim3(64, CS_LANES), 0, stream,
                       (const float*)workspace, parts, cols, out);
  if (le != cudaSuccess) return (int)le;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" size_t isg_colsum_multi_workspace_bytes(int n, const int64_t* rows, const int* cols) {
  if (n <= 0 || !rows || !cols) return 0;
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += (isg_colsum_workspace_bytes(rows[i], cols[i]) + 255) & ~(size_t)255;
  return total;
}

extern "C" int isg_colsum_multi(int n, const void* const* in, const int* dtypes, const int64_t* ld, const int64_t* rows,
                                const int* cols, float* const* out, void* workspace, size_t ws_bytes, void* stream_) {
  if (n < 0 || n > CS_MAX_JOBS) return ISG_EINVAL;
  if (n == 0) return ISG_OK;
  if (!in || !ld || !rows || !cols || !out) return ISG_EINVAL;
  if (ws_bytes < isg_colsum_multi_workspace_bytes(n, rows, cols) || !workspace) return ISG_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  ColsumBatch batch;
  batch.n = 0;
  int blocks = 0, fblocks = 0;
  char* ws = (char*)workspace;
  for (int i = 0; i < n; ++i) {
    if (rows[i] < 0 || cols[i] <= 0 || !out[i]) return ISG_EINVAL;
    if (rows[i] == 0) {
      cudaError_t e = cudaMemsetAsync(out[i], 0, (size_t)cols[i] * sizeof(float), stream);
      if (e != cudaSuccess) return (int)e;
      continue;
    }
    if (!in[i]) return ISG_EINVAL;
    const int dt = dtypes ? dtypes[i] : ISG_F32;
    if (dt != ISG_F32 && dt != ISG_BF16) return ISG_EUNSUPPORTED;
    if (cols[i] % 4 != 0 || ld[i] % 4 != 0 || ((uintptr_t)in[i] & (dt == ISG_BF16 ? 7 : 15)) || ((uintptr_t)out[i] & 15))
      return ISG_EUNSUPPORTED;
    ColsumJob& J = batch.job[batch.n++];
    J.bf16 = dt == ISG_BF16 ? 1 : 0;
    J.in = in[i];
    J.out = out[i];
    J.part = (float*)ws;
    ws += (isg_colsum_workspace_bytes(rows[i], cols[i]) + 255) & ~(size_t)255;
    J.ld = ld[i];
    J.rows = rows[i];
    J.cols = cols[i];
    J.parts = (int)((rows[i] + CS_ROWS - 1) / CS_ROWS);
    J.cblocks = isg::ceil_div(cols[i] / 4, 64);
    J.block0 = blocks;
    J.fblock0 = fblocks;
    blocks += J.cblocks * J.parts;
    fblocks += J.cblocks;
  }
  if (batch.n == 0) return ISG_OK;
  cudaError_t le = isg::launch_pdl(colsum_multi_partial_kernel, dim3(blocks), dim3(64), 0, stream, batch);
  if (le != cudaSuccess) return (int)le;
  ISG_CHECK_LAUNCH();
  le = isg::launch_pdl(colsum_multi_final_kernel, dim3(fblocks), dim3(64, CS_LANES), 0, stream, batch);
  if (le != cudaSuccess) return (int)le;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}


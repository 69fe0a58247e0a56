// edge.cu — kernel (b): fused GATv2 edge attention (forward + backward) and NodeMaskToEdgeMask.
//
// Restates MaskingGATv2Conv.message + PyG propagate/softmax/aggregate
// (reference models/mgat_v2_conv.py:243-279, :215, :231-232) as ONE pass over a dst-sorted CSR:
// a warp owns one (destination node, head) pair, keeps x_r[dst,h,:], att[h,:] and the output
// accumulator in registers, streams x_l[src,h,:] (gathered, L2-resident) and e_proj[e,h,:]
// (read-once, HBM) with 16-byte loads, reduces the per-edge logit with warp shuffles and folds
// it into an online segment softmax.  PyG materialises ~8 [E,H,C] tensors for the same work.
//
// HBM-bound: algorithmic bytes per layer (fp32) = 4*HC*(E + 3N) + 4*E*H (+4E mask) + CSR ints
// (SURVEY.md §8d).  No tensor-core work here by design.
#include <stdlib.h>

#include "common.cuh"

namespace {

using namespace isg;

constexpr int EDGE_WARPS = 4;  // warps per CTA (128 threads)

// lane `lane` owns float4 slots v = lane + 32*k (k < VPL) of a C-wide head row, valid if v < C/4.
template <typename T, int VPL>
__device__ __forceinline__ void load_row(const T* __restrict__ row, int lane, int c4, float4 (&r)[VPL]) {
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = lane + 32 * k;
    r[k] = (v < c4) ? Vec4<T>::ld(row + 4 * v) : f4_zero();
  }
}
template <typename T, int VPL>
__device__ __forceinline__ void load_row_stream(const T* __restrict__ row, int lane, int c4, float4 (&r)[VPL]) {
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = lane + 32 * k;
    r[k] = (v < c4) ? Vec4<T>::ld_stream(row + 4 * v) : f4_zero();
  }
}
template <int VPL>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ row, int lane, int c4, float4 (&r)[VPL]) {
  load_row<float, VPL>(row, lane, c4, r);
}

__device__ __forceinline__ float leaky(float u, float slope) { return u > 0.f ? u : slope * u; }

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename T, int VPL>
__global__ void __launch_bounds__(EDGE_WARPS * 32)
gat_edge_fwd_kernel(const T* __restrict__ xl, const T* __restrict__ xr, int64_t ld_x,
                    const T* __restrict__ ep, const float* __restrict__ att,
                    const float* __restrict__ bias, const float* __restrict__ emask,
                    const int* __restrict__ rowptr, const int* __restrict__ nbr,
                    const int* __restrict__ eid, T* __restrict__ out, int64_t ld_out,
                    float* __restrict__ alpha, int64_t NH, int H, int C, float slope) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * EDGE_WARPS + (threadIdx.x >> 5);
  if (wid >= NH) return;
  const int64_t node = wid / H;
  const int head = (int)(wid - node * H);
  const int c4 = C >> 2;
  const int64_t HC = (int64_t)H * C;
  const int hoff = head * C;

  float4 xr_v[VPL], att_v[VPL], acc[VPL];
  load_row<T, VPL>(xr + node * ld_x + hoff, lane, c4, xr_v);
  load_row_f32<VPL>(att + hoff, lane, c4, att_v);
#pragma unroll
  for (int k = 0; k < VPL; ++k) acc[k] = f4_zero();

  const int beg = rowptr[node], end = rowptr[node + 1];
  float m_run = -INFINITY, s_run = 0.f;

  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int my_src = 0, my_eid = 0;
    float my_m = 1.f, my_logit = 0.f;
    if (lane < cnt) {
      my_src = nbr[base + lane];
      my_eid = eid[base + lane];
      if (emask != nullptr) my_m = emask[my_eid];
    }
    // two edges per iteration: both rows' loads are issued before either reduction
    for (int t = 0; t < cnt; t += 2) {
      const bool has2 = (t + 1) < cnt;
      const int j0 = __shfl_sync(ISG_FULL_MASK, my_src, t);
      const int e0 = __shfl_sync(ISG_FULL_MASK, my_eid, t);
      const float m0 = __shfl_sync(ISG_FULL_MASK, my_m, t);
      const int j1 = __shfl_sync(ISG_FULL_MASK, my_src, has2 ? t + 1 : t);
      const int e1 = __shfl_sync(ISG_FULL_MASK, my_eid, has2 ? t + 1 : t);
      const float m1 = __shfl_sync(ISG_FULL_MASK, my_m, has2 ? t + 1 : t);
      float4 x0[VPL], p0[VPL], x1[VPL], p1[VPL];
      load_row<T, VPL>(xl + (int64_t)j0 * ld_x + hoff, lane, c4, x0);
      load_row_stream<T, VPL>(ep + (int64_t)e0 * HC + hoff, lane, c4, p0);
      if (has2) {
        load_row<T, VPL>(xl + (int64_t)j1 * ld_x + hoff, lane, c4, x1);
        load_row_stream<T, VPL>(ep + (int64_t)e1 * HC + hoff, lane, c4, p1);
      }
      float part0 = 0.f, part1 = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        float4 s = f4_add(f4_add(xr_v[k], x0[k]), p0[k]);
        float4 w;
        w.x = leaky(s.x * m0, slope) * m0;
        w.y = leaky(s.y * m0, slope) * m0;
        w.z = leaky(s.z * m0, slope) * m0;
        w.w = leaky(s.w * m0, slope) * m0;
        part0 += f4_dot(w, att_v[k]);
      }
      if (has2) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          float4 s = f4_add(f4_add(xr_v[k], x1[k]), p1[k]);
          float4 w;
          w.x = leaky(s.x * m1, slope) * m1;
          w.y = leaky(s.y * m1, slope) * m1;
          w.z = leaky(s.z * m1, slope) * m1;
          w.w = leaky(s.w * m1, slope) * m1;
          part1 += f4_dot(w, att_v[k]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        part0 += __shfl_xor_sync(ISG_FULL_MASK, part0, o);
        part1 += __shfl_xor_sync(ISG_FULL_MASK, part1, o);
      }
      {  // online softmax update, edge t
        const float m_new = fmaxf(m_run, part0);
        const float sc = expf(m_run - m_new);
        const float p = expf(part0 - m_new);
        s_run = s_run * sc + p;
        const float pm = p * m0;
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = f4_fma(x0[k], pm, f4_scale(acc[k], sc));
        m_run = m_new;
        if (lane == t) my_logit = part0;
      }
      if (has2) {
        const float m_new = fmaxf(m_run, part1);
        const float sc = expf(m_run - m_new);
        const float p = expf(part1 - m_new);
        s_run = s_run * sc + p;
        const float pm = p * m1;
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = f4_fma(x1[k], pm, f4_scale(acc[k], sc));
        m_run = m_new;
        if (lane == t + 1) my_logit = part1;
      }
    }
    if (lane < cnt) alpha[(int64_t)my_eid * H + head] = my_logit;  // raw logit, normalised below
  }

  const float inv = 1.f / (s_run + 1e-16f);  // PyG softmax epsilon (mgat_v2_conv.py:272)
  T* orow = out + node * ld_out + hoff;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = lane + 32 * k;
    if (v < c4) {
      float4 o = f4_scale(acc[k], inv);
      if (bias != nullptr) o = f4_add(o, Vec4<float>::ld(bias + hoff + 4 * v));
      Vec4<T>::st(orow + 4 * v, o);
    }
  }
  // same lane wrote the raw logit above -> program order makes it visible here
  for (int base = beg; base < end; base += 32) {
    if (base + lane < end) {
      const int64_t idx = (int64_t)eid[base + lane] * H + head;
      alpha[idx] = expf(alpha[idx] - m_run) * inv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward, pass 1 (dst-major): g_eproj, g_xr, g_att partials, per-head g_edge_mask terms.
// Two sweeps over the in-edges of a (node, head): the first accumulates the softmax-backward dot
// product sum_e a_e m_e t_e, the second produces the per-edge gradients.
// Persistent over (node, head) pairs with the head fixed per warp so the g_att partial lives
// in registers; partials are reduced by gat_att_reduce_kernel in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------
template <typename T, int VPL, bool MASKED>
__global__ void __launch_bounds__(EDGE_WARPS * 32)
gat_edge_bwd_dst_kernel(const T* __restrict__ gout, int64_t ld_g, const T* __restrict__ xl,
                        const T* __restrict__ xr, int64_t ld_x, const T* __restrict__ ep,
                        const float* __restrict__ att, const float* __restrict__ bias,
                        const float* __restrict__ emask, const float* __restrict__ alpha,
                        const T* __restrict__ out, int64_t ld_out,
                        const int* __restrict__ rowptr, const int* __restrict__ nbr,
                        const int* __restrict__ eid, T* __restrict__ g_xr, int64_t ld_gx,
                        T* __restrict__ g_ep, float* __restrict__ gatt_part,
                        float* __restrict__ gm_h, int64_t N, int H, int C, float slope) {
  const int lane = threadIdx.x & 31;
  const int64_t wid0 = (int64_t)blockIdx.x * EDGE_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * EDGE_WARPS;  // multiple of H by construction
  const int head = (int)(wid0 % H);
  const int c4 = C >> 2;
  const int64_t HC = (int64_t)H * C;
  const int hoff = head * C;

  float4 att_v[VPL], gatt[VPL];
  load_row_f32<VPL>(att + hoff, lane, c4, att_v);
  (void)bias; (void)out; (void)ld_out;  // kept in the ABI; the direct softmax backward does not need them
#pragma unroll
  for (int k = 0; k < VPL; ++k) gatt[k] = f4_zero();

  for (int64_t node = wid0 / H; node < N; node += nwarps / H) {
    float4 G[VPL], xr_v[VPL], gxr[VPL];
    load_row<T, VPL>(gout + node * ld_g + hoff, lane, c4, G);
    load_row<T, VPL>(xr + node * ld_x + hoff, lane, c4, xr_v);
#pragma unroll
    for (int k = 0; k < VPL; ++k) gxr[k] = f4_zero();
    const int beg = rowptr[node], end = rowptr[node + 1];

    // sweep 1: dot = sum_e a_e * m_e * t_e with t_e = <G, x_l[src_e]>.  (The flash-attention
    // shortcut dot = <G, out - bias> is NOT used: once the segment softmax saturates, t_e - dot
    // cancels catastrophically unless both come from the very same t_e roundings; measured 6e-2
    // relative error on lin_r.weight gradients with the shortcut.)  Sweep 2 recomputes the
    // bit-identical t_e, so nothing is stored; the gathered rows are re-read from L1/L2.
    float dot = 0.f;
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_src = 0;
      float my_am = 0.f;
      if (lane < cnt) {
        my_src = nbr[base + lane];
        const int e = eid[base + lane];
        my_am = alpha[(int64_t)e * H + head];
        if (MASKED) my_am *= emask[e];
      }
      for (int t = 0; t < cnt; t += 2) {
        const bool has2 = (t + 1) < cnt;
        const int j0 = __shfl_sync(ISG_FULL_MASK, my_src, t);
        const int j1 = __shfl_sync(ISG_FULL_MASK, my_src, has2 ? t + 1 : t);
        const float am0 = __shfl_sync(ISG_FULL_MASK, my_am, t);
        const float am1 = has2 ? __shfl_sync(ISG_FULL_MASK, my_am, t + 1) : 0.f;
        float4 x0[VPL], x1[VPL];
        load_row<T, VPL>(xl + (int64_t)j0 * ld_x + hoff, lane, c4, x0);
        load_row<T, VPL>(xl + (int64_t)j1 * ld_x + hoff, lane, c4, x1);
        float p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          p0 = f4_dot_acc(G[k], x0[k], p0);
          p1 = f4_dot_acc(G[k], x1[k], p1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          p0 += __shfl_xor_sync(ISG_FULL_MASK, p0, o);
          p1 += __shfl_xor_sync(ISG_FULL_MASK, p1, o);
        }
        dot = fmaf(am0, p0, dot);
        dot = fmaf(am1, p1, dot);
      }
    }

    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_src = 0, my_eid = 0;
      float my_m = 1.f, my_a = 0.f, my_gm = 0.f;
      if (lane < cnt) {
        my_src = nbr[base + lane];
        my_eid = eid[base + lane];
        if (MASKED) my_m = emask[my_eid];
        my_a = alpha[(int64_t)my_eid * H + head];
      }
      for (int t = 0; t < cnt; ++t) {
        const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        const float m = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
        const float a = __shfl_sync(ISG_FULL_MASK, my_a, t);
        float4 xv[VPL], pv[VPL];
        load_row<T, VPL>(xl + (int64_t)j * ld_x + hoff, lane, c4, xv);
        load_row_stream<T, VPL>(ep + (int64_t)e * HC + hoff, lane, c4, pv);
        float tpart = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) tpart = f4_dot_acc(G[k], xv[k], tpart);
        const float tt = warp_sum(tpart);
        const float gl = a * (m * tt - dot);  // d loss / d logit[e,h]
        float gmpart = 0.f;
        T* gerow = g_ep + (int64_t)e * HC + hoff;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const float sx[4] = {xr_v[k].x + xv[k].x + pv[k].x, xr_v[k].y + xv[k].y + pv[k].y,
                               xr_v[k].z + xv[k].z + pv[k].z, xr_v[k].w + xv[k].w + pv[k].w};
          const float at[4] = {att_v[k].x, att_v[k].y, att_v[k].z, att_v[k].w};
          float gs[4], ga[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float u = sx[q] * m;
            const float lk = u > 0.f ? 1.f : slope;
            const float v = u * lk;
            const float gw = gl * at[q];
            const float gu = gw * m * lk;
            gs[q] = gu * m;
            ga[q] = gl * (v * m);
            if (MASKED) gmpart += gw * v + gu * sx[q];
          }
          gatt[k].x += ga[0]; gatt[k].y += ga[1]; gatt[k].z += ga[2]; gatt[k].w += ga[3];
          gxr[k].x += gs[0]; gxr[k].y += gs[1]; gxr[k].z += gs[2]; gxr[k].w += gs[3];
          const int v4 = lane + 32 * k;
          if (v4 < c4) Vec4<T>::st_stream(gerow + 4 * v4, make_float4(gs[0], gs[1], gs[2], gs[3]));
        }
        if (MASKED) {
          const float gm = warp_sum(gmpart) + tt * a;
          if (lane == t) my_gm = gm;
        }
      }
      if (MASKED && lane < cnt) gm_h[(int64_t)my_eid * H + head] = my_gm;
    }
    T* grow = g_xr + node * ld_gx + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v4 = lane + 32 * k;
      if (v4 < c4) Vec4<T>::st(grow + 4 * v4, gxr[k]);
    }
  }
  float* prow = gatt_part + wid0 * C;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v4 = lane + 32 * k;
    if (v4 < c4) Vec4<float>::st(prow + 4 * v4, gatt[k]);
  }
}

// g_att[h*C + c] = sum over warps w with (w % H == h) of part[w, c], fixed order.
// grid (ceil(C/32), H), block (32, 8): 8 row lanes stride through the partials, then a fixed-order
// shared-memory fold -> deterministic.
__global__ void gat_att_reduce_kernel(const float* __restrict__ part, int64_t nwarps, int H, int C,
                                      float* __restrict__ g_att) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, h = blockIdx.y;
  float s = 0.f;
  if (c < C)
    for (int64_t w = h + (int64_t)H * threadIdx.y; w < nwarps; w += (int64_t)H * 8) s += part[w * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    g_att[h * C + c] = t;
  }
}

// ------------------------------------------------------------------------------------------
// backward, pass 2 (src-major): g_xl[j] = sum_{e: src = j} ( g_eproj[e] + G[dst_e] * alpha*m ).
// ------------------------------------------------------------------------------------------
template <typename T, int VPL, bool MASKED>
__global__ void __launch_bounds__(EDGE_WARPS * 32)
gat_edge_bwd_src_kernel(const T* __restrict__ gout, int64_t ld_g, const T* __restrict__ g_ep,
                        const float* __restrict__ emask, const float* __restrict__ alpha,
                        const int* __restrict__ colptr, const int* __restrict__ nbr,
                        const int* __restrict__ eid, T* __restrict__ g_xl, int64_t ld_gx,
                        int64_t NH, int H, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * EDGE_WARPS + (threadIdx.x >> 5);
  if (wid >= NH) return;
  const int64_t node = wid / H;
  const int head = (int)(wid - node * H);
  const int c4 = C >> 2;
  const int64_t HC = (int64_t)H * C;
  const int hoff = head * C;
  float4 acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) acc[k] = f4_zero();
  const int beg = colptr[node], end = colptr[node + 1];
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int my_dst = 0, my_eid = 0;
    float my_am = 0.f;
    if (lane < cnt) {
      my_dst = nbr[base + lane];
      my_eid = eid[base + lane];
      my_am = alpha[(int64_t)my_eid * H + head];
      if (MASKED) my_am *= emask[my_eid];
    }
    for (int t = 0; t < cnt; ++t) {
      const int i = __shfl_sync(ISG_FULL_MASK, my_dst, t);
      const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
      const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
      float4 gv[VPL], Gv[VPL];
      load_row_stream<T, VPL>(g_ep + (int64_t)e * HC + hoff, lane, c4, gv);
      load_row<T, VPL>(gout + (int64_t)i * ld_g + hoff, lane, c4, Gv);
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = f4_add(acc[k], f4_fma(Gv[k], am, gv[k]));
    }
  }
  T* grow = g_xl + node * ld_gx + hoff;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v4 = lane + 32 * k;
    if (v4 < c4) Vec4<T>::st(grow + 4 * v4, acc[k]);
  }
}


// ==========================================================================================
// fp32 kernels with shared-memory row staging (the ones the fp32 path runs).
// Every gathered (x_l[src], G[dst]) or streamed (e_proj[e], g_eproj[e]) head-row slice — C floats =
// 1200 B at C = 300 — is brought in by one cp.async.bulk (UBLKCP) issued by an elected lane into a
// per-warp ring of RING slots; the copy of edge t+RING is issued as soon as edge t has been consumed.
// Streamed rows carry an L2 evict_first policy, gathered rows evict_last, so the node-feature rows of
// the batch stay L2-resident.
//
// Register diet (r1f): the first ring kernels kept x_r[dst], att[h] and the current edge's rows in
// registers (118-128 regs -> 16 warps/SM) and ran a persistent grid with a static stride over the
// (node, head) tasks.  ncu showed them latency-bound at 20 % warps active, and the static stride left
// a long tail at the c3 size (5 tasks of 1..75 edges per warp).  Now the per-task constants (x_r[dst,h,:],
// att[h,:]) live in per-warp shared memory, the edge rows are read from their ring slot where they are
// used, the ring is 2 deep, and the grid is one CTA per EDGE_WARPS tasks (hardware block scheduling
// balances the tail): 72-80 regs -> 24-28 warps/SM.  Forward at B=256: 106 -> 79 us (L2 flushed).
// ==========================================================================================
constexpr int RING = 2;  // slots per warp (power of two)

struct WarpRing {
  uint32_t data, bars, slot_bytes, seq;
  __device__ __forceinline__ void init(uint32_t data_, uint32_t bars_, uint32_t slot_bytes_, int lane, int extra_bars) {
    data = data_; bars = bars_; slot_bytes = slot_bytes_; seq = 0;
    if (lane == 0) {
      for (int r = 0; r < RING + extra_bars; ++r) ring_bar_init(bars + 8u * r, 1);
      ring_fence_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ uint32_t slot(uint32_t s) const { return data + (s & (RING - 1)) * slot_bytes; }
  __device__ __forceinline__ uint32_t bar(uint32_t s) const { return bars + 8u * (s & (RING - 1)); }
  __device__ __forceinline__ uint32_t extra_bar(int i) const { return bars + 8u * (RING + i); }
  // elected lane only
  __device__ __forceinline__ void issue2(uint32_t s, const void* a, uint64_t pol_a, const void* b, uint64_t pol_b,
                                         uint32_t bytes_each) const {
    ring_expect(bar(s), 2 * bytes_each);
    ring_copy(slot(s), a, bytes_each, bar(s), pol_a);
    ring_copy(slot(s) + bytes_each, b, bytes_each, bar(s), pol_b);
  }
  __device__ __forceinline__ void issue1(uint32_t s, const void* a, uint64_t pol_a, uint32_t bytes) const {
    ring_expect(bar(s), bytes);
    ring_copy(slot(s), a, bytes, bar(s), pol_a);
  }
  __device__ __forceinline__ void wait(uint32_t s) const { ring_wait(bar(s), (s / RING) & 1u); }
};

// Per-warp shared-memory layout of the three kernels: RING slots of 2 rows, then `extra_rows` rows of
// per-task constants; the mbarriers of all warps follow the data of all warps.
__host__ __device__ inline uint32_t ring_warp_bytes(int C, int extra_rows) {
  return (uint32_t)(2 * RING + extra_rows) * (uint32_t)C * 4u;
}
inline size_t ring_smem_bytes(int C, int extra_rows, int extra_bars) {
  return (size_t)EDGE_WARPS * (ring_warp_bytes(C, extra_rows) + 8 * (RING + extra_bars));
}

__device__ __forceinline__ void st_f4_policy(float* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 a) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
}

// ------------------------------------------------------------------------------------------
// forward: one warp per (node, head); online segment softmax over the in-edges.
// smem per warp: [slot0: x_l | e_proj][slot1][x_r row][att row]; barriers: RING + 1 (x_r).
// CT/HT > 0 bake C and H in as compile-time constants (the reference shape C=300, H=4): all row
// offsets become immediates, which removes ~40 % of the integer instructions of the generic build.
// ------------------------------------------------------------------------------------------
template <int VPL, int CT>
__device__ __forceinline__ bool col_active(int k, int lane, int c4) {
  if (k < VPL - 1) return true;  // vpl_for(): c4 > 32*(VPL-1), only the last 32-column group is ragged
  return lane + 32 * k < (CT > 0 ? CT / 4 : c4);
}

template <int VPL, bool MASKED, int CT, int HT>
__global__ void __launch_bounds__(EDGE_WARPS * 32, 7)
gat_edge_fwd_ring_kernel(const float* __restrict__ xl, const float* __restrict__ xr, int64_t ld_x,
                         const float* __restrict__ ep, const float* __restrict__ att,
                         const float* __restrict__ bias, const float* __restrict__ emask,
                         const int* __restrict__ rowptr, const int* __restrict__ nbr,
                         const int* __restrict__ eid, const int* __restrict__ order, float* __restrict__ out,
                         int64_t ld_out, float* __restrict__ alpha, int64_t NH, int H_, int C_, float slope) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  const int C = CT > 0 ? CT : C_, H = HT > 0 ? HT : H_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = C >> 2;
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  const uint32_t wbytes = ring_warp_bytes(C, 2);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes,
            smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * (RING + 1), 2 * row_bytes, lane, 1);
  const uint32_t xr_s = ring.data + 2 * RING * row_bytes, att_s = xr_s + row_bytes;
  const uint32_t xbar = ring.extra_bar(0);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();  // issues the bulk copies (see warp_elect_one)
  uint32_t xseq = 0;
  int cur_head = -1;

  for (int64_t wid = (int64_t)blockIdx.x * EDGE_WARPS + warp; wid < NH; wid += (int64_t)gridDim.x * EDGE_WARPS) {
    const int64_t slot = wid / H;
    const int head = (int)(wid - slot * H);
    const int64_t node = order ? order[slot] : slot;  // longest segments first (isg_degree_order)
    const int hoff = head * C;
    if (head != cur_head) {
      for (int v = lane; v < c4; v += 32) sts_f4(att_s + 16u * v, Vec4<float>::ld(att + hoff + 4 * v));
      cur_head = head;
      __syncwarp();
    }
    if (leader) {
      ring_expect(xbar, row_bytes);
      ring_copy(xr_s, xr + node * ld_x + hoff, row_bytes, xbar, pol_keep);
    }
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = f4_zero();
    const int beg = rowptr[node], end = rowptr[node + 1];
    const bool single = end - beg <= 32;  // whole segment in one batch: alpha is written once, normalised
    float m_run = -INFINITY, s_run = 0.f;
    bool xr_ready = false;
    int keep_eid = 0;
    float keep_logit = 0.f;

    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_src = 0, my_eid = 0;
      float my_m = 1.f, my_logit = 0.f;
      if (lane < cnt) {
        my_src = nbr[base + lane];
        my_eid = eid[base + lane];
        if (MASKED) my_m = emask[my_eid];
      }
      const int npre = min(RING, cnt);
      for (int t = 0; t < npre; ++t) {
        const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader)
          ring.issue2(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff, pol_stream,
                      row_bytes);
      }
      for (int t = 0; t < cnt; ++t) {
        const uint32_t q = ring.seq + t;
        const float m0 = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
        const int tn = t + RING < cnt ? t + RING : t;
        const int jn = __shfl_sync(ISG_FULL_MASK, my_src, tn);
        const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
        if (!xr_ready) {
          ring_wait(xbar, xseq & 1u);
          xr_ready = true;
        }
        ring.wait(q);
        const uint32_t sx = ring.slot(q) + l16, sp = sx + row_bytes;
        // logit = att . (leaky(s*m)*m),  s = x_r[i] + x_l[j] + e_proj[e]   (mgat_v2_conv.py:262-270)
        float2 part2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          if (col_active<VPL, CT>(k, lane, c4)) {
            float4 sv = p4_add(p4_add(lds_f4(xr_s + l16 + 512u * k), lds_f4(sx + 512u * k)), lds_f4(sp + 512u * k));
            if (MASKED) sv = p4_scale(sv, m0);
            float4 w = p4_leaky(sv, slope);
            if (MASKED) w = p4_scale(w, m0);
            part2 = p4_dot_acc(w, lds_f4(att_s + l16 + 512u * k), part2);
          }
        }
        const float part0 = warp_sum(p2_sum(part2));
        if (part0 > m_run) {
          const float sc = expf(m_run - part0);
          s_run *= sc;
#pragma unroll
          for (int k = 0; k < VPL; ++k) acc[k] = p4_scale(acc[k], sc);
          m_run = part0;
        }
        const float pe = expf(part0 - m_run);
        s_run += pe;
        const float pm = MASKED ? pe * m0 : pe;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (col_active<VPL, CT>(k, lane, c4)) acc[k] = p4_fma_s(lds_f4(sx + 512u * k), pm, acc[k]);
        if (lane == t) my_logit = part0;
        __syncwarp();  // every lane is done with the slot
        if (leader && t + RING < cnt)
          ring.issue2(q + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, ep + (int64_t)en * HC + hoff, pol_stream,
                      row_bytes);
      }
      ring.seq += cnt;
      if (single) {
        keep_eid = my_eid;
        keep_logit = my_logit;
      } else if (lane < cnt) {
        alpha[(int64_t)my_eid * H + head] = my_logit;
      }
    }
    if (!xr_ready) ring_wait(xbar, xseq & 1u);  // node without in-edges: still consume the x_r copy
    ++xseq;
    __syncwarp();  // the x_r row is free for the next task's copy

    const float inv = 1.f / (s_run + 1e-16f);
    float* orow = out + node * ld_out + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (col_active<VPL, CT>(k, lane, c4)) {
        const int v = lane + 32 * k;
        float4 o = f4_scale(acc[k], inv);
        if (bias != nullptr) o = f4_add(o, Vec4<float>::ld(bias + hoff + 4 * v));
        Vec4<float>::st(orow + 4 * v, o);
      }
    }
    if (single) {
      if (lane < end - beg) alpha[(int64_t)keep_eid * H + head] = expf(keep_logit - m_run) * inv;
    } else {
      for (int base = beg; base < end; base += 32) {
        if (base + lane < end) {
          const int64_t idx = (int64_t)eid[base + lane] * H + head;
          alpha[idx] = expf(alpha[idx] - m_run) * inv;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward, pass 1 (dst-major).  Warp w: head = w % H, column = w / H; column c owns the nodes
// [c*K, (c+1)*K).  The g_att partial of a warp stays in registers over its K nodes and is written to
// gatt_part[w, :]; gat_att_reduce* folds the partials in a fixed order (deterministic).
// smem per warp: [slot0][slot1][x_r row][att row]; barriers: RING + 1 (x_r).
//
// Per (node, head): sweep 1 accumulates dot = sum_e a_e m_e t_e with t_e = <G, x_l[src_e]>, sweep 2
// produces the per-edge gradients from gl_e = a_e (m_e t_e - dot) (see the note in
// gat_edge_bwd_dst_kernel on why both must use the very same t_e roundings).  Segments of <= 32 edges
// (the usual case) keep t_e in lane e's register and run both sweeps as ONE stream of 2*deg ring items,
// so the x_l|e_proj copies of sweep 2 are already in flight while sweep 1 drains; longer segments take
// the two-loop path that recomputes the bit-identical t_e.
// ------------------------------------------------------------------------------------------
template <int VPL, bool MASKED, int CT, int HT>
__global__ void __launch_bounds__(EDGE_WARPS * 32, 6)
gat_edge_bwd_dst_ring_kernel(const float* __restrict__ gout, int64_t ld_g, const float* __restrict__ xl,
                             const float* __restrict__ xr, int64_t ld_x, const float* __restrict__ ep,
                             const float* __restrict__ att, const float* __restrict__ emask,
                             const float* __restrict__ alpha, const int* __restrict__ rowptr,
                             const int* __restrict__ nbr, const int* __restrict__ eid,
                             const int* __restrict__ order, float* __restrict__ g_xr, int64_t ld_gx,
                             float* __restrict__ g_ep, float* __restrict__ gatt_part, float* __restrict__ gm_h,
                             int64_t N, int H_, int C_, float slope, int K, int ep_keep, int* __restrict__ red_cnt,
                             int red_n) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  const int C = CT > 0 ? CT : C_, H = HT > 0 ? HT : H_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // arrival counters of the g_att fold that rides in the src kernel (launched after this one completes)
  if (blockIdx.x == 0)
    for (int v = threadIdx.x; v < red_n; v += blockDim.x) red_cnt[v] = 0;
  const int64_t wid0 = (int64_t)blockIdx.x * EDGE_WARPS + warp;  // gridDim.x*EDGE_WARPS is a multiple of H
  const int head = (int)(wid0 % H);
  int64_t col = wid0 / H;
  // K == 1 (one node per column): columns are taken in the order of isg_degree_order, longest segment first; the
  // g_att partial row stays keyed by the NODE, so the fixed-order fold below does not depend on that order
  if (order != nullptr && K == 1 && col < N) col = order[col];
  const int c4 = C >> 2;
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  const int hoff = head * C;
  const uint32_t wbytes = ring_warp_bytes(C, 2);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes,
            smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * (RING + 1), 2 * row_bytes, lane, 1);
  const uint32_t xr_s = ring.data + 2 * RING * row_bytes, att_s = xr_s + row_bytes;
  const uint32_t xbar = ring.extra_bar(0);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();
  uint32_t xseq = 0;

  for (int v = lane; v < c4; v += 32) sts_f4(att_s + 16u * v, Vec4<float>::ld(att + hoff + 4 * v));
  __syncwarp();
  float4 gatt[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) gatt[k] = f4_zero();

  const int64_t node_end = min(N, (col + 1) * (int64_t)K);
  for (int64_t node = col * K; node < node_end; ++node) {
    if (leader) {
      ring_expect(xbar, row_bytes);
      ring_copy(xr_s, xr + node * ld_x + hoff, row_bytes, xbar, pol_keep);
    }
    float4 G[VPL], gxr[VPL];
    load_row<float, VPL>(gout + node * ld_g + hoff, lane, c4, G);
#pragma unroll
    for (int k = 0; k < VPL; ++k) gxr[k] = f4_zero();
    const int beg = rowptr[node], end = rowptr[node + 1];
    const int deg = end - beg;
    bool xr_ready = false;

    // t_e = <G, x_l row in the slot>; one fixed instruction sequence for every use
    auto row_dot = [&](uint32_t sx) -> float {
      float2 pp2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (col_active<VPL, CT>(k, lane, c4)) pp2 = p4_dot_acc(G[k], lds_f4(sx + 512u * k), pp2);
      return warp_sum(p2_sum(pp2));
    };
    // per-edge gradients; returns the per-head edge-mask gradient (MASKED only)
    auto edge_grad = [&](uint32_t sx, uint32_t sp, float tt, float a, float m, float dot, int e) -> float {
      if (!xr_ready) {
        ring_wait(xbar, xseq & 1u);
        xr_ready = true;
      }
      const float gl = a * (m * tt - dot);  // d loss / d logit[e,h]
      const float glmm = MASKED ? gl * m * m : gl;
      float2 av2 = make_float2(0.f, 0.f);  // sum_c att*v (edge-mask gradient, see below)
      float* gerow = g_ep + (int64_t)e * HC + hoff + 4 * lane;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (col_active<VPL, CT>(k, lane, c4)) {
          const float4 sxv =
              p4_add(p4_add(lds_f4(xr_s + l16 + 512u * k), lds_f4(sx + 512u * k)), lds_f4(sp + 512u * k));
          const float4 u = MASKED ? p4_scale(sxv, m) : sxv;
          const float4 v = p4_leaky(u, slope);
          const float4 at = lds_f4(att_s + l16 + 512u * k);
          // g_att += gl * (v*m);  g_s = gl*m*m * att * leaky'(u)
          gatt[k] = p4_fma_s(MASKED ? p4_scale(v, m) : v, gl, gatt[k]);
          const float4 lk = make_float4(u.x > 0.f ? 1.f : slope, u.y > 0.f ? 1.f : slope, u.z > 0.f ? 1.f : slope,
                                        u.w > 0.f ? 1.f : slope);
          const float4 gs = p4_scale(p4_mul(at, lk), glmm);
          gxr[k] = p4_add(gxr[k], gs);
          if (MASKED) av2 = p4_dot_acc(at, v, av2);
          if (ep_keep) st_f4_policy(gerow + 128 * k, gs, pol_keep);  // g_eproj is re-read by the src pass
          else Vec4<float>::st_stream(gerow + 128 * k, gs);
        }
      }
      // d/dm of logit = sum_c att*(dw/dm): w = leaky(s*m)*m  =>  gw*v + gu*s = 2*gl*att*v per channel
      // (gu*s = gw*m*leaky'(u)*s = gw*v), so the per-head edge-mask gradient is 2*gl*sum_c att*v + t*a.
      return MASKED ? warp_sum(2.f * gl * p2_sum(av2)) + tt * a : 0.f;
    };

    if (deg <= 32) {
      int my_src = 0, my_eid = 0;
      float my_m = 1.f, my_a = 0.f, my_t = 0.f, my_gm = 0.f;
      if (lane < deg) {
        my_src = nbr[beg + lane];
        my_eid = eid[beg + lane];
        if (MASKED) my_m = emask[my_eid];
        my_a = alpha[(int64_t)my_eid * H + head];
      }
      const float my_am = MASKED ? my_a * my_m : my_a;
      const int items = 2 * deg;  // item i < deg: sweep-1 edge i (x_l only); i >= deg: sweep-2 edge i-deg
      auto issue_item = [&](int i) {  // all lanes call (shuffles); the elected lane issues
        const int t = i < deg ? i : i - deg;
        const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader) {
          if (i < deg)
            ring.issue1(ring.seq + i, xl + (int64_t)j * ld_x + hoff, pol_keep, row_bytes);
          else
            ring.issue2(ring.seq + i, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff,
                        pol_stream, row_bytes);
        }
      };
      for (int i = 0; i < min(RING, items); ++i) issue_item(i);
      float dot = 0.f;
      for (int t = 0; t < deg; ++t) {
        const uint32_t s = ring.seq + t;
        const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
        ring.wait(s);
        const float pp = row_dot(ring.slot(s) + l16);  // warp_sum inside: every lane is past its slot reads
        if (t + RING < items) issue_item(t + RING);
        if (lane == t) my_t = pp;
        dot = fmaf(am, pp, dot);
      }
      for (int t = 0; t < deg; ++t) {
        const uint32_t s = ring.seq + deg + t;
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        const float m = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
        const float a = __shfl_sync(ISG_FULL_MASK, my_a, t);
        const float tt = __shfl_sync(ISG_FULL_MASK, my_t, t);
        ring.wait(s);
        const uint32_t sx = ring.slot(s) + l16;
        const float gm = edge_grad(sx, sx + row_bytes, tt, a, m, dot, e);
        __syncwarp();
        if (deg + t + RING < items) issue_item(deg + t + RING);
        if (MASKED && lane == t) my_gm = gm;
      }
      ring.seq += items;
      if (MASKED && lane < deg) gm_h[(int64_t)my_eid * H + head] = my_gm;
    } else {
      // sweep 1
      float dot = 0.f;
      for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int my_src = 0;
        float my_am = 0.f;
        if (lane < cnt) {
          my_src = nbr[base + lane];
          const int e = eid[base + lane];
          my_am = alpha[(int64_t)e * H + head];
          if (MASKED) my_am *= emask[e];
        }
        const int npre = min(RING, cnt);
        for (int t = 0; t < npre; ++t) {
          const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
          if (leader) ring.issue1(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, row_bytes);
        }
        for (int t = 0; t < cnt; ++t) {
          const uint32_t s = ring.seq + t;
          const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
          const int jn = __shfl_sync(ISG_FULL_MASK, my_src, t + RING < cnt ? t + RING : t);
          ring.wait(s);
          const float pp = row_dot(ring.slot(s) + l16);
          if (leader && t + RING < cnt) ring.issue1(s + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, row_bytes);
          dot = fmaf(am, pp, dot);
        }
        ring.seq += cnt;
      }
      // sweep 2 (recomputes the bit-identical t_e)
      for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int my_src = 0, my_eid = 0;
        float my_m = 1.f, my_a = 0.f, my_gm = 0.f;
        if (lane < cnt) {
          my_src = nbr[base + lane];
          my_eid = eid[base + lane];
          if (MASKED) my_m = emask[my_eid];
          my_a = alpha[(int64_t)my_eid * H + head];
        }
        const int npre = min(RING, cnt);
        for (int t = 0; t < npre; ++t) {
          const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
          const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
          if (leader)
            ring.issue2(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff,
                        pol_stream, row_bytes);
        }
        for (int t = 0; t < cnt; ++t) {
          const uint32_t s = ring.seq + t;
          const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
          const float m = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
          const float a = __shfl_sync(ISG_FULL_MASK, my_a, t);
          const int tn = t + RING < cnt ? t + RING : t;
          const int jn = __shfl_sync(ISG_FULL_MASK, my_src, tn);
          const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
          ring.wait(s);
          const uint32_t sx = ring.slot(s) + l16;
          const float tt = row_dot(sx);
          const float gm = edge_grad(sx, sx + row_bytes, tt, a, m, dot, e);
          __syncwarp();
          if (leader && t + RING < cnt)
            ring.issue2(s + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, ep + (int64_t)en * HC + hoff, pol_stream,
                        row_bytes);
          if (MASKED && lane == t) my_gm = gm;
        }
        ring.seq += cnt;
        if (MASKED && lane < cnt) gm_h[(int64_t)my_eid * H + head] = my_gm;
      }
    }
    if (!xr_ready) ring_wait(xbar, xseq & 1u);
    ++xseq;
    __syncwarp();
    float* grow = g_xr + node * ld_gx + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (col_active<VPL, CT>(k, lane, c4)) Vec4<float>::st(grow + 4 * (lane + 32 * k), gxr[k]);
  }
  float* prow = gatt_part + (col * H + head) * C;
#pragma unroll
  for (int k = 0; k < VPL; ++k)
    if (col_active<VPL, CT>(k, lane, c4)) Vec4<float>::st(prow + 4 * (lane + 32 * k), gatt[k]);
}

// g_att = column sums of gatt_part viewed as [rows, H*C] (row = column index of the dst kernel), two
// fixed-order stages: slabs of GR_ROWS rows -> part2 [parts2, H*C], then part2 -> g_att with GR2_LANES
// row lanes per column and a fixed-order shared-memory fold.
constexpr int GR_ROWS = 64;   // rows per slab (64 keeps the second stage's serial chain short: N/64/16 steps)
constexpr int GR_BATCH = 16;  // loads in flight per thread
constexpr int GR2_LANES = 16;
// rows [r0, r1) of column group c summed in ascending row order, starting from row r0's value
__device__ __forceinline__ float4 slab_fold(const float* __restrict__ part, int64_t r0, int64_t r1, int cols, int c) {
  float4 s = f4_zero();
  for (int64_t b = r0; b < r1; b += GR_BATCH) {
    float4 v[GR_BATCH];
#pragma unroll
    for (int u = 0; u < GR_BATCH; ++u) v[u] = (b + u < r1) ? Vec4<float>::ld(part + (b + u) * cols + c) : f4_zero();
#pragma unroll
    for (int u = 0; u < GR_BATCH; ++u)
      if (b + u < r1) s = (b + u == r0) ? v[u] : f4_add(s, v[u]);
  }
  return s;
}
__global__ void __launch_bounds__(64)
gat_att_reduce1_kernel(const float* __restrict__ part, int64_t rows, int cols, float* __restrict__ part2) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * GR_ROWS, r1 = min(rows, r0 + GR_ROWS);
  Vec4<float>::st(part2 + (int64_t)blockIdx.y * cols + c, slab_fold(part, r0, r1, cols, c));
}
__global__ void __launch_bounds__(64 * GR2_LANES)
gat_att_reduce2_kernel(const float* __restrict__ part2, int nparts, int cols, float* __restrict__ g_att) {
  __shared__ float4 red[GR2_LANES][64];
  const int c = (blockIdx.x * 64 + threadIdx.x) * 4;
  float4 s = f4_zero();
  if (c < cols) {
    int p = threadIdx.y;
    for (; p + 3 * GR2_LANES < nparts; p += 4 * GR2_LANES) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = Vec4<float>::ld(part2 + (int64_t)(p + u * GR2_LANES) * cols + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) s = f4_add(s, v[u]);
    }
    for (; p < nparts; p += GR2_LANES) s = f4_add(s, Vec4<float>::ld(part2 + (int64_t)p * cols + c));
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < GR2_LANES; ++y) t = f4_add(t, red[y][threadIdx.x]);
    Vec4<float>::st(g_att + c, t);
  }
}

// The same two-stage fold as block roles of the src kernel (fp32 ring path).  Block b < red_blocks folds slab
// b / cb of column block b % cb exactly as gat_att_reduce1_kernel does, then bumps the column block's arrival
// counter; the block that arrives last folds part2 over the slabs in gat_att_reduce2_kernel's order (row p goes to
// lane sum p % GR2_LANES, ascending p; lane sums folded 0..15), so g_att is bit-identical to the two-kernel fold and
// independent of which block arrives last.  Counters are zeroed by the dst kernel.
struct AttFold {
  const float* part;   // [rows, H*C] g_att partials of the dst pass
  float* part2;        // [parts2, H*C]
  float* g_att;        // [H*C]
  int* cnt;            // [cb] arrival counters
  const float* gm_h;   // [E, H] per-head edge-mask gradients (masked only)
  float* g_emask;      // [E]
  int64_t rows, E;
  int parts2, red_blocks, gm_blocks;
};
constexpr int FOLD_THREADS = EDGE_WARPS * 32;
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __noinline__ void att_fold_role(const AttFold& f, int cols) {
  __shared__ int s_last;
  const int cb = (cols / 4 + FOLD_THREADS - 1) / FOLD_THREADS;
  const int y = (int)blockIdx.x / cb, x = (int)blockIdx.x - y * cb;
  const int c = (x * FOLD_THREADS + (int)threadIdx.x) * 4;
  if (c < cols) {
    const int64_t r0 = (int64_t)y * GR_ROWS, r1 = min(f.rows, r0 + GR_ROWS);
    Vec4<float>::st(f.part2 + (int64_t)y * cols + c, slab_fold(f.part, r0, r1, cols, c));
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(f.cnt + x, 1) == f.parts2 - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (c >= cols) return;
  float4 s[GR2_LANES];
#pragma unroll
  for (int u = 0; u < GR2_LANES; ++u) s[u] = f4_zero();
  for (int p0 = 0; p0 < f.parts2; p0 += GR2_LANES) {
    float4 v[GR2_LANES];
#pragma unroll
    for (int u = 0; u < GR2_LANES; ++u)
      v[u] = p0 + u < f.parts2 ? ldcg_f4(f.part2 + (int64_t)(p0 + u) * cols + c) : f4_zero();
#pragma unroll
    for (int u = 0; u < GR2_LANES; ++u)
      if (p0 + u < f.parts2) s[u] = f4_add(s[u], v[u]);
  }
  float4 t = s[0];
#pragma unroll
  for (int u = 1; u < GR2_LANES; ++u) t = f4_add(t, s[u]);
  Vec4<float>::st(f.g_att + c, t);
}

// g_edge_mask[e] = sum_h gm_h[e, h] (gm_head_sum_kernel's arithmetic), grid-strided over the gm blocks
__device__ __forceinline__ void gm_fold_role(const AttFold& f, int H) {
  const int64_t stride = (int64_t)f.gm_blocks * FOLD_THREADS;
  for (int64_t e = (int64_t)((int)blockIdx.x - f.red_blocks) * FOLD_THREADS + threadIdx.x; e < f.E; e += stride) {
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += f.gm_h[e * H + h];
    f.g_emask[e] = s;
  }
}

// ------------------------------------------------------------------------------------------
// backward, pass 2 (src-major): g_xl[j] = sum_{e: src = j} ( g_eproj[e] + G[dst_e] * alpha*m ).
// ------------------------------------------------------------------------------------------
template <int VPL, bool MASKED, int CT, int HT>
__global__ void __launch_bounds__(EDGE_WARPS * 32, 7)
gat_edge_bwd_src_ring_kernel(const float* __restrict__ gout, int64_t ld_g, const float* __restrict__ g_ep,
                             const float* __restrict__ emask, const float* __restrict__ alpha,
                             const int* __restrict__ colptr, const int* __restrict__ nbr,
                             const int* __restrict__ eid, const int* __restrict__ order, float* __restrict__ g_xl,
                             int64_t ld_gx, int64_t NH, int H_, int C_, AttFold fold) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  const int C = CT > 0 ? CT : C_, H = HT > 0 ? HT : H_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = C >> 2;
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  // the leading blocks of the grid fold the dst pass's g_att partials (and the per-head edge-mask gradients) while
  // the remaining blocks stream g_eproj: three launches fewer, and the fold is hidden under the src pass
  if ((int)blockIdx.x < fold.red_blocks + fold.gm_blocks) {
    if ((int)blockIdx.x < fold.red_blocks) att_fold_role(fold, (int)HC);
    else if (MASKED) gm_fold_role(fold, H);
    return;
  }
  const int64_t first = (int64_t)(blockIdx.x - fold.red_blocks - fold.gm_blocks);
  const int64_t nblk = (int64_t)(gridDim.x - fold.red_blocks - fold.gm_blocks);
  const uint32_t wbytes = ring_warp_bytes(C, 0);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes, smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * RING,
            2 * row_bytes, lane, 0);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();

  for (int64_t wid = first * EDGE_WARPS + warp; wid < NH; wid += nblk * EDGE_WARPS) {
    const int64_t slot = wid / H;
    const int head = (int)(wid - slot * H);
    const int64_t node = order ? order[slot] : slot;  // longest segments first (isg_degree_order)
    const int hoff = head * C;
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = f4_zero();
    const int beg = colptr[node], end = colptr[node + 1];
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_dst = 0, my_eid = 0;
      float my_am = 0.f;
      if (lane < cnt) {
        my_dst = nbr[base + lane];
        my_eid = eid[base + lane];
        my_am = alpha[(int64_t)my_eid * H + head];
        if (MASKED) my_am *= emask[my_eid];
      }
      const int npre = min(RING, cnt);
      for (int t = 0; t < npre; ++t) {
        const int i = __shfl_sync(ISG_FULL_MASK, my_dst, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader)
          ring.issue2(ring.seq + t, g_ep + (int64_t)e * HC + hoff, pol_stream, gout + (int64_t)i * ld_g + hoff,
                      pol_keep, row_bytes);
      }
      for (int t = 0; t < cnt; ++t) {
        const uint32_t s = ring.seq + t;
        const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
        const int tn = t + RING < cnt ? t + RING : t;
        const int in_ = __shfl_sync(ISG_FULL_MASK, my_dst, tn);
        const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
        ring.wait(s);
        const uint32_t sg = ring.slot(s) + l16, sG = sg + row_bytes;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (col_active<VPL, CT>(k, lane, c4))
            acc[k] = p4_add(acc[k], p4_fma_s(lds_f4(sG + 512u * k), am, lds_f4(sg + 512u * k)));
        __syncwarp();
        if (leader && t + RING < cnt)
          ring.issue2(s + RING, g_ep + (int64_t)en * HC + hoff, pol_stream, gout + (int64_t)in_ * ld_g + hoff,
                      pol_keep, row_bytes);
      }
      ring.seq += cnt;
    }
    float* grow = g_xl + node * ld_gx + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (col_active<VPL, CT>(k, lane, c4)) Vec4<float>::st(grow + 4 * (lane + 32 * k), acc[k]);
  }
}

// ------------------------------------------------------------------------------------------
// backward, ONE launch (fp32 ring path, graph-closed edge sets): the dst-major and the src-major pass above as
// two block ROLES of the same kernel, interleaved so that g_eproj is consumed out of L2.
//
// Run as two launches, the src pass re-reads all of g_eproj (E x H*C floats: 191 MB at the c3 size, more than
// the 126 MB L2) from DRAM — 216 of the 700 MB the two-pass backward moves for 501 MB of algorithmic traffic.
// Here every block draws a ticket t from a global counter (tickets follow the order in which blocks actually
// start, which blockIdx does not guarantee):   t even -> dst block t/2,   t odd -> src block (t-1)/2 - LAG,
// so the src block of a node starts ~2*LAG tickets after the dst block of the same node.  A src task needs
// the g_eproj rows of all out-edges of its node, which are written by the dst tasks of the nodes of the SAME
// graph (edges never leave a graph — checked when the GraphIndex is built); each dst task bumps its graph's
// completion counter with release semantics and the src task acquires it (spinning only if the lag was too
// short).  LAG >= Nmax guarantees that every dst task a src task can wait for has already been ticketed, i.e.
// is running or finished: no deadlock.  The g_eproj rows in flight between the two roles are
// ~(LAG + resident blocks / 2) nodes' worth (~40 MB at the default lag), so the src role finds them in L2.
// Arithmetic, task granularity ((node, head) per warp) and summation orders are exactly those of the two-pass
// kernels: results are bit-identical to them and independent of the schedule.
// sync[0] = ticket counter, sync[1 + g] = finished dst tasks of graph g; zeroed by a memset before the launch.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int VPL, bool MASKED, int CT, int HT>
__global__ void __launch_bounds__(EDGE_WARPS * 32, 6)
gat_edge_bwd_fused_ring_kernel(const float* __restrict__ gout, int64_t ld_g, const float* __restrict__ xl,
                               const float* __restrict__ xr, int64_t ld_x, const float* __restrict__ ep,
                               const float* __restrict__ att, const float* __restrict__ emask,
                               const float* __restrict__ alpha, const int* __restrict__ rowptr,
                               const int* __restrict__ nbr, const int* __restrict__ eid,
                               const int* __restrict__ colptr, const int* __restrict__ snbr,
                               const int* __restrict__ seid, const int* __restrict__ batch32,
                               const int* __restrict__ graph_ptr, float* __restrict__ g_xl,
                               float* __restrict__ g_xr, int64_t ld_gx, float* __restrict__ g_ep,
                               float* __restrict__ gatt_part, float* __restrict__ gm_h, int* __restrict__ sync,
                               int64_t NH, int64_t nblocks_role, int lag_blocks, int H_, int C_, float slope) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  __shared__ int s_ticket;
  const int C = CT > 0 ? CT : C_, H = HT > 0 ? HT : H_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&sync[0], 1);
  __syncthreads();
  const int ticket = s_ticket;
  const bool src_role = (ticket & 1) != 0;
  const int64_t blk = src_role ? (int64_t)(ticket >> 1) - lag_blocks : (int64_t)(ticket >> 1);
  if (blk < 0 || blk >= nblocks_role) return;
  const int64_t wid = blk * EDGE_WARPS + warp;
  if (wid >= NH) return;  // warp-uniform; nothing below synchronises across warps
  const int64_t node = wid / H;
  const int head = (int)(wid - node * H);
  const int c4 = C >> 2;
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  const int hoff = head * C;
  const uint32_t wbytes = ring_warp_bytes(C, 2);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes,
            smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * (RING + 1), 2 * row_bytes, lane, 1);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();
  const int graph = batch32[node];

  if (src_role) {
    // ---- wait until every dst task of this node's graph has published its g_eproj rows
    const int need = (graph_ptr[graph + 1] - graph_ptr[graph]) * H;
    if (lane == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(sync + 1 + graph) < need) {
        __nanosleep(100);
        if (++spins > (1u << 23)) __trap();  // a lost signal must fail loudly, not hang the GPU
      }
    }
    __syncwarp();
    __threadfence();          // order the acquire before this warp's reads (all lanes)
    fence_proxy_async_all();  // ... including the bulk copies (async proxy) of data written by generic stores
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = f4_zero();
    const int beg = colptr[node], end = colptr[node + 1];
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_dst = 0, my_eid = 0;
      float my_am = 0.f;
      if (lane < cnt) {
        my_dst = snbr[base + lane];
        my_eid = seid[base + lane];
        my_am = alpha[(int64_t)my_eid * H + head];
        if (MASKED) my_am *= emask[my_eid];
      }
      const int npre = min(RING, cnt);
      for (int t = 0; t < npre; ++t) {
        const int i = __shfl_sync(ISG_FULL_MASK, my_dst, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader)
          ring.issue2(ring.seq + t, g_ep + (int64_t)e * HC + hoff, pol_stream, gout + (int64_t)i * ld_g + hoff,
                      pol_keep, row_bytes);
      }
      for (int t = 0; t < cnt; ++t) {
        const uint32_t s = ring.seq + t;
        const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
        const int tn = t + RING < cnt ? t + RING : t;
        const int in_ = __shfl_sync(ISG_FULL_MASK, my_dst, tn);
        const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
        ring.wait(s);
        const uint32_t sg = ring.slot(s) + l16, sG = sg + row_bytes;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (col_active<VPL, CT>(k, lane, c4))
            acc[k] = p4_add(acc[k], p4_fma_s(lds_f4(sG + 512u * k), am, lds_f4(sg + 512u * k)));
        __syncwarp();
        if (leader && t + RING < cnt)
          ring.issue2(s + RING, g_ep + (int64_t)en * HC + hoff, pol_stream, gout + (int64_t)in_ * ld_g + hoff,
                      pol_keep, row_bytes);
      }
      ring.seq += cnt;
    }
    float* grow = g_xl + node * ld_gx + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (col_active<VPL, CT>(k, lane, c4)) Vec4<float>::st(grow + 4 * (lane + 32 * k), acc[k]);
    return;
  }

  // ---- dst role: one (node, head) task — the body of gat_edge_bwd_dst_ring_kernel with K = 1
  const uint32_t xr_s = ring.data + 2 * RING * row_bytes, att_s = xr_s + row_bytes;
  const uint32_t xbar = ring.extra_bar(0);
  for (int v = lane; v < c4; v += 32) sts_f4(att_s + 16u * v, Vec4<float>::ld(att + hoff + 4 * v));
  __syncwarp();
  float4 gatt[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) gatt[k] = f4_zero();
  {
    if (leader) {
      ring_expect(xbar, row_bytes);
      ring_copy(xr_s, xr + node * ld_x + hoff, row_bytes, xbar, pol_keep);
    }
    float4 G[VPL], gxr[VPL];
    load_row<float, VPL>(gout + node * ld_g + hoff, lane, c4, G);
#pragma unroll
    for (int k = 0; k < VPL; ++k) gxr[k] = f4_zero();
    const int beg = rowptr[node], end = rowptr[node + 1];
    const int deg = end - beg;
    bool xr_ready = false;

    auto row_dot = [&](uint32_t sx) -> float {
      float2 pp2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (col_active<VPL, CT>(k, lane, c4)) pp2 = p4_dot_acc(G[k], lds_f4(sx + 512u * k), pp2);
      return warp_sum(p2_sum(pp2));
    };
    auto edge_grad = [&](uint32_t sx, uint32_t sp, float tt, float a, float m, float dot, int e) -> float {
      if (!xr_ready) {
        ring_wait(xbar, 0u);
        xr_ready = true;
      }
      const float gl = a * (m * tt - dot);
      const float glmm = MASKED ? gl * m * m : gl;
      float2 av2 = make_float2(0.f, 0.f);
      float* gerow = g_ep + (int64_t)e * HC + hoff + 4 * lane;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (col_active<VPL, CT>(k, lane, c4)) {
          const float4 sxv =
              p4_add(p4_add(lds_f4(xr_s + l16 + 512u * k), lds_f4(sx + 512u * k)), lds_f4(sp + 512u * k));
          const float4 u = MASKED ? p4_scale(sxv, m) : sxv;
          const float4 v = p4_leaky(u, slope);
          const float4 at = lds_f4(att_s + l16 + 512u * k);
          gatt[k] = p4_fma_s(MASKED ? p4_scale(v, m) : v, gl, gatt[k]);
          const float4 lk = make_float4(u.x > 0.f ? 1.f : slope, u.y > 0.f ? 1.f : slope, u.z > 0.f ? 1.f : slope,
                                        u.w > 0.f ? 1.f : slope);
          const float4 gs = p4_scale(p4_mul(at, lk), glmm);
          gxr[k] = p4_add(gxr[k], gs);
          if (MASKED) av2 = p4_dot_acc(at, v, av2);
          Vec4<float>::st_stream(gerow + 128 * k, gs);  // L1 no-allocate only: the row stays in L2 for the src role
        }
      }
      return MASKED ? warp_sum(2.f * gl * p2_sum(av2)) + tt * a : 0.f;
    };

    if (deg <= 32) {
      int my_src = 0, my_eid = 0;
      float my_m = 1.f, my_a = 0.f, my_t = 0.f, my_gm = 0.f;
      if (lane < deg) {
        my_src = nbr[beg + lane];
        my_eid = eid[beg + lane];
        if (MASKED) my_m = emask[my_eid];
        my_a = alpha[(int64_t)my_eid * H + head];
      }
      const float my_am = MASKED ? my_a * my_m : my_a;
      const int items = 2 * deg;
      auto issue_item = [&](int i) {
        const int t = i < deg ? i : i - deg;
        const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader) {
          if (i < deg)
            ring.issue1(ring.seq + i, xl + (int64_t)j * ld_x + hoff, pol_keep, row_bytes);
          else
            ring.issue2(ring.seq + i, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff,
                        pol_stream, row_bytes);
        }
      };
      for (int i = 0; i < min(RING, items); ++i) issue_item(i);
      float dot = 0.f;
      for (int t = 0; t < deg; ++t) {
        const uint32_t s = ring.seq + t;
        const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
        ring.wait(s);
        const float pp = row_dot(ring.slot(s) + l16);
        if (t + RING < items) issue_item(t + RING);
        if (lane == t) my_t = pp;
        dot = fmaf(am, pp, dot);
      }
      for (int t = 0; t < deg; ++t) {
        const uint32_t s = ring.seq + deg + t;
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        const float m = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
        const float a = __shfl_sync(ISG_FULL_MASK, my_a, t);
        const float tt = __shfl_sync(ISG_FULL_MASK, my_t, t);
        ring.wait(s);
        const uint32_t sx = ring.slot(s) + l16;
        const float gm = edge_grad(sx, sx + row_bytes, tt, a, m, dot, e);
        __syncwarp();
        if (deg + t + RING < items) issue_item(deg + t + RING);
        if (MASKED && lane == t) my_gm = gm;
      }
      ring.seq += items;
      if (MASKED && lane < deg) gm_h[(int64_t)my_eid * H + head] = my_gm;
    } else {
      float dot = 0.f;
      for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int my_src = 0;
        float my_am = 0.f;
        if (lane < cnt) {
          my_src = nbr[base + lane];
          const int e = eid[base + lane];
          my_am = alpha[(int64_t)e * H + head];
          if (MASKED) my_am *= emask[e];
        }
        const int npre = min(RING, cnt);
        for (int t = 0; t < npre; ++t) {
          const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
          if (leader) ring.issue1(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, row_bytes);
        }
        for (int t = 0; t < cnt; ++t) {
          const uint32_t s = ring.seq + t;
          const float am = __shfl_sync(ISG_FULL_MASK, my_am, t);
          const int jn = __shfl_sync(ISG_FULL_MASK, my_src, t + RING < cnt ? t + RING : t);
          ring.wait(s);
          const float pp = row_dot(ring.slot(s) + l16);
          if (leader && t + RING < cnt) ring.issue1(s + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, row_bytes);
          dot = fmaf(am, pp, dot);
        }
        ring.seq += cnt;
      }
      for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int my_src = 0, my_eid = 0;
        float my_m = 1.f, my_a = 0.f, my_gm = 0.f;
        if (lane < cnt) {
          my_src = nbr[base + lane];
          my_eid = eid[base + lane];
          if (MASKED) my_m = emask[my_eid];
          my_a = alpha[(int64_t)my_eid * H + head];
        }
        const int npre = min(RING, cnt);
        for (int t = 0; t < npre; ++t) {
          const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
          const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
          if (leader)
            ring.issue2(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff,
                        pol_stream, row_bytes);
        }
        for (int t = 0; t < cnt; ++t) {
          const uint32_t s = ring.seq + t;
          const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
          const float m = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
          const float a = __shfl_sync(ISG_FULL_MASK, my_a, t);
          const int tn = t + RING < cnt ? t + RING : t;
          const int jn = __shfl_sync(ISG_FULL_MASK, my_src, tn);
          const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
          ring.wait(s);
          const uint32_t sx = ring.slot(s) + l16;
          const float tt = row_dot(sx);
          const float gm = edge_grad(sx, sx + row_bytes, tt, a, m, dot, e);
          __syncwarp();
          if (leader && t + RING < cnt)
            ring.issue2(s + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, ep + (int64_t)en * HC + hoff, pol_stream,
                        row_bytes);
          if (MASKED && lane == t) my_gm = gm;
        }
        ring.seq += cnt;
        if (MASKED && lane < cnt) gm_h[(int64_t)my_eid * H + head] = my_gm;
      }
    }
    if (!xr_ready) ring_wait(xbar, 0u);
    __syncwarp();
    float* grow = g_xr + node * ld_gx + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (col_active<VPL, CT>(k, lane, c4)) Vec4<float>::st(grow + 4 * (lane + 32 * k), gxr[k]);
  }
  float* prow = gatt_part + wid * C;
#pragma unroll
  for (int k = 0; k < VPL; ++k)
    if (col_active<VPL, CT>(k, lane, c4)) Vec4<float>::st(prow + 4 * (lane + 32 * k), gatt[k]);
  // publish: every lane orders its own g_eproj stores (also towards the async proxy that will read them),
  // then one lane signals the graph's counter
  __threadfence();
  fence_proxy_async_all();
  __syncwarp();
  if (lane == 0) red_release_gpu_add(sync + 1 + graph, 1);
}

// Columns of the dst-major backward: K nodes per column, sized for ~6 waves of resident CTAs so that the
// hardware block scheduler evens out the degree imbalance; K and the grid depend on (N, H) only.
// ==========================================================================================
// bf16 storage, ring kernels (r2): the bf16 configuration's edge kernels with 16-byte accesses.
// A head row of bf16 is C*2 = 600 bytes — not a legal bulk-copy size and only 8-byte aligned for odd heads, which
// is why the first bf16 kernels (above) used 8-byte register loads.  Here a warp owns a (node, HEAD PAIR): the rows
// of heads (2p, 2p+1) are contiguous, 4C = 1200 bytes, 16-byte aligned — byte for byte the shape the fp32 ring
// moves — so the same per-warp ring of cp.async.bulk copies applies.  Lane `lane` owns the 16-byte chunks
// c = lane + 32k (k < VPL, c < C/4) of a pair row = 8 bf16 = two float4 groups; group (k, half) holds elements
// 8c + 4*half .. +3 of the pair and belongs to the second head iff that offset is >= C (C % 4 == 0: no group
// straddles the heads).  Every per-head scalar (logit, softmax state, alpha, t_e, gl) exists twice.
// Arithmetic is fp32 throughout; only loads / stores convert.
// ==========================================================================================
__device__ __forceinline__ void bf8_to_f4(uint4 raw, float4& lo, float4& hi) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.z));
  const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.w));
  lo = make_float4(a.x, a.y, b.x, b.y);
  hi = make_float4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ uint4 f4_to_bf8(float4 lo, float4 hi) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(lo.x, lo.y), b = __floats2bfloat162_rn(lo.z, lo.w);
  const __nv_bfloat162 c = __floats2bfloat162_rn(hi.x, hi.y), d = __floats2bfloat162_rn(hi.z, hi.w);
  uint4 r;
  r.x = *reinterpret_cast<const uint32_t*>(&a);
  r.y = *reinterpret_cast<const uint32_t*>(&b);
  r.z = *reinterpret_cast<const uint32_t*>(&c);
  r.w = *reinterpret_cast<const uint32_t*>(&d);
  return r;
}
// acc_h += d for the head the group belongs to (selects, no dynamic register indexing)
__device__ __forceinline__ void sel_add(float2& acc0, float2& acc1, bool second, float2 d) {
  const float2 z = make_float2(0.f, 0.f);
  acc0 = __fadd2_rn(acc0, second ? z : d);
  acc1 = __fadd2_rn(acc1, second ? d : z);
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// smem per warp: RING slots of 2 pair rows (4C bytes each), then x_r pair row (4C bytes) and the pair's att (8C bytes)
__host__ __device__ inline uint32_t pair_warp_bytes(int C, bool with_consts) {
  return (uint32_t)(2 * RING) * (uint32_t)C * 4u + (with_consts ? (uint32_t)C * 12u : 0u);
}
inline size_t pair_smem_bytes(int C, bool with_consts, int extra_bars) {
  return (size_t)EDGE_WARPS * (pair_warp_bytes(C, with_consts) + 8 * (RING + extra_bars));
}

// forward: one warp per (node, head pair)
#ifndef ISG_PAIR_FWD_CTAS
#define ISG_PAIR_FWD_CTAS 4  // 128 registers, no spills: 0.086 vs 0.101 ms at c3 with 5 (96 registers, spilling)
#endif
#ifndef ISG_PAIR_DST_CTAS
#define ISG_PAIR_DST_CTAS 3  // c3: 0.189 vs 0.208 ms with 4 (batch 1024: 0.603 vs 0.575)
#endif
template <int VPL, bool MASKED>
__global__ void __launch_bounds__(EDGE_WARPS * 32, ISG_PAIR_FWD_CTAS)
gat_edge_fwd_pair_kernel(const __nv_bfloat16* __restrict__ xl, const __nv_bfloat16* __restrict__ xr, int64_t ld_x,
                         const __nv_bfloat16* __restrict__ ep, const float* __restrict__ att,
                         const float* __restrict__ bias, const float* __restrict__ emask,
                         const int* __restrict__ rowptr, const int* __restrict__ nbr, const int* __restrict__ eid,
                         const int* __restrict__ order, __nv_bfloat16* __restrict__ out, int64_t ld_out,
                         float* __restrict__ alpha, int64_t NP, int H, int C, float slope) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = C >> 2, HP = H >> 1;  // 16-byte chunks per pair row; head pairs
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  const uint32_t wbytes = pair_warp_bytes(C, true);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes,
            smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * (RING + 1), 2 * row_bytes, lane, 1);
  const uint32_t xr_s = ring.data + 2 * RING * row_bytes, att_s = xr_s + row_bytes;
  const uint32_t xbar = ring.extra_bar(0);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();
  uint32_t xseq = 0;
  int cur_pair = -1;
  // per-lane group -> head selector: group (k, half) belongs to the pair's second head iff 8c + 4*half >= C
  bool hb[2 * VPL], ok[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = lane + 32 * k;
    ok[k] = c < c4;
    hb[2 * k] = 8 * c >= C;
    hb[2 * k + 1] = 8 * c + 4 >= C;
  }

  for (int64_t wid = (int64_t)blockIdx.x * EDGE_WARPS + warp; wid < NP; wid += (int64_t)gridDim.x * EDGE_WARPS) {
    const int64_t slot = wid / HP;
    const int pair = (int)(wid - slot * HP);
    const int64_t node = order ? order[slot] : slot;
    const int hoff = pair * 2 * C;
    if (pair != cur_pair) {
      for (int v = lane; v < 2 * c4; v += 32) sts_f4(att_s + 16u * v, Vec4<float>::ld(att + hoff + 4 * v));
      cur_pair = pair;
      __syncwarp();
    }
    if (leader) {
      ring_expect(xbar, row_bytes);
      ring_copy(xr_s, xr + node * ld_x + hoff, row_bytes, xbar, pol_keep);
    }
    float4 acc[2 * VPL];
#pragma unroll
    for (int g = 0; g < 2 * VPL; ++g) acc[g] = f4_zero();
    const int beg = rowptr[node], end = rowptr[node + 1];
    const bool single = end - beg <= 32;
    float m_run0 = -INFINITY, m_run1 = -INFINITY, s_run0 = 0.f, s_run1 = 0.f;
    bool xr_ready = false;
    int keep_eid = 0;
    float keep_logit0 = 0.f, keep_logit1 = 0.f;

    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_src = 0, my_eid = 0;
      float my_m = 1.f, my_logit0 = 0.f, my_logit1 = 0.f;
      if (lane < cnt) {
        my_src = nbr[base + lane];
        my_eid = eid[base + lane];
        if (MASKED) my_m = emask[my_eid];
      }
      const int npre = min(RING, cnt);
      for (int t = 0; t < npre; ++t) {
        const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader)
          ring.issue2(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff, pol_stream,
                      row_bytes);
      }
      for (int t = 0; t < cnt; ++t) {
        const uint32_t q = ring.seq + t;
        const float m0 = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
        const int tn = t + RING < cnt ? t + RING : t;
        const int jn = __shfl_sync(ISG_FULL_MASK, my_src, tn);
        const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
        if (!xr_ready) {
          ring_wait(xbar, xseq & 1u);
          xr_ready = true;
        }
        ring.wait(q);
        const uint32_t sx = ring.slot(q) + l16, sp = sx + row_bytes;
        float2 part0 = make_float2(0.f, 0.f), part1 = make_float2(0.f, 0.f);
        float4 xlv[2 * VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          if (ok[k]) {
            float4 r0, r1, e0, e1;
            bf8_to_f4(lds_u4(sx + 512u * k), xlv[2 * k], xlv[2 * k + 1]);
            bf8_to_f4(lds_u4(sp + 512u * k), e0, e1);
            bf8_to_f4(lds_u4(xr_s + l16 + 512u * k), r0, r1);
            float4 s0 = p4_add(p4_add(r0, xlv[2 * k]), e0), s1 = p4_add(p4_add(r1, xlv[2 * k + 1]), e1);
            if (MASKED) { s0 = p4_scale(s0, m0); s1 = p4_scale(s1, m0); }
            float4 w0 = p4_leaky(s0, slope), w1 = p4_leaky(s1, slope);
            if (MASKED) { w0 = p4_scale(w0, m0); w1 = p4_scale(w1, m0); }
            const uint32_t as = att_s + 32u * (lane + 32 * k);
            const float2 d0 = p4_dot_acc(w0, lds_f4(as), make_float2(0.f, 0.f));
            const float2 d1 = p4_dot_acc(w1, lds_f4(as + 16u), make_float2(0.f, 0.f));
            sel_add(part0, part1, hb[2 * k], d0);
            sel_add(part0, part1, hb[2 * k + 1], d1);
          } else {
            xlv[2 * k] = f4_zero();
            xlv[2 * k + 1] = f4_zero();
          }
        }
        float pm0, pm1, sc0 = 1.f, sc1 = 1.f;
        // both heads' butterfly reductions interleaved (independent shuffle chains)
        float lg0 = p2_sum(part0), lg1 = p2_sum(part1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lg0 += __shfl_xor_sync(ISG_FULL_MASK, lg0, o);
          lg1 += __shfl_xor_sync(ISG_FULL_MASK, lg1, o);
        }
        {
          const float lg = lg0;
          if (lg > m_run0) {
            sc0 = expf(m_run0 - lg);
            s_run0 *= sc0;
            m_run0 = lg;
          }
          const float pe = expf(lg - m_run0);
          s_run0 += pe;
          pm0 = MASKED ? pe * m0 : pe;
          if (lane == t) my_logit0 = lg;
        }
        {
          const float lg = lg1;
          if (lg > m_run1) {
            sc1 = expf(m_run1 - lg);
            s_run1 *= sc1;
            m_run1 = lg;
          }
          const float pe = expf(lg - m_run1);
          s_run1 += pe;
          pm1 = MASKED ? pe * m0 : pe;
          if (lane == t) my_logit1 = lg;
        }
        if (sc0 != 1.f || sc1 != 1.f) {  // warp-uniform
#pragma unroll
          for (int g = 0; g < 2 * VPL; ++g) acc[g] = p4_scale(acc[g], hb[g] ? sc1 : sc0);
        }
#pragma unroll
        for (int g = 0; g < 2 * VPL; ++g) acc[g] = p4_fma_s(xlv[g], hb[g] ? pm1 : pm0, acc[g]);
        __syncwarp();
        if (leader && t + RING < cnt)
          ring.issue2(q + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, ep + (int64_t)en * HC + hoff, pol_stream,
                      row_bytes);
      }
      ring.seq += cnt;
      if (single) {
        keep_eid = my_eid;
        keep_logit0 = my_logit0;
        keep_logit1 = my_logit1;
      } else if (lane < cnt) {
        alpha[(int64_t)my_eid * H + 2 * pair] = my_logit0;
        alpha[(int64_t)my_eid * H + 2 * pair + 1] = my_logit1;
      }
    }
    if (!xr_ready) ring_wait(xbar, xseq & 1u);
    ++xseq;
    __syncwarp();

    const float inv0 = 1.f / (s_run0 + 1e-16f), inv1 = 1.f / (s_run1 + 1e-16f);
    __nv_bfloat16* orow = out + node * ld_out + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (ok[k]) {
        const int c = lane + 32 * k;
        float4 o0 = f4_scale(acc[2 * k], hb[2 * k] ? inv1 : inv0), o1 = f4_scale(acc[2 * k + 1], hb[2 * k + 1] ? inv1 : inv0);
        if (bias != nullptr) {
          o0 = f4_add(o0, Vec4<float>::ld(bias + hoff + 8 * c));
          o1 = f4_add(o1, Vec4<float>::ld(bias + hoff + 8 * c + 4));
        }
        *reinterpret_cast<uint4*>(orow + 8 * c) = f4_to_bf8(o0, o1);
      }
    }
    if (single) {
      if (lane < end - beg) {
        alpha[(int64_t)keep_eid * H + 2 * pair] = expf(keep_logit0 - m_run0) * inv0;
        alpha[(int64_t)keep_eid * H + 2 * pair + 1] = expf(keep_logit1 - m_run1) * inv1;
      }
    } else {
      for (int base = beg; base < end; base += 32) {
        if (base + lane < end) {
          const int64_t idx = (int64_t)eid[base + lane] * H + 2 * pair;
          alpha[idx] = expf(alpha[idx] - m_run0) * inv0;
          alpha[idx + 1] = expf(alpha[idx + 1] - m_run1) * inv1;
        }
      }
    }
  }
}

// backward, src-major: g_xl[j, pair] = sum_{e: src = j} ( g_eproj[e, pair] + G[dst_e, pair] * alpha_h*m )
template <int VPL, bool MASKED>
__global__ void __launch_bounds__(EDGE_WARPS * 32, 6)
gat_edge_bwd_src_pair_kernel(const __nv_bfloat16* __restrict__ gout, int64_t ld_g, const __nv_bfloat16* __restrict__ g_ep,
                             const float* __restrict__ emask, const float* __restrict__ alpha,
                             const int* __restrict__ colptr, const int* __restrict__ nbr, const int* __restrict__ eid,
                             const int* __restrict__ order, __nv_bfloat16* __restrict__ g_xl, int64_t ld_gx, int64_t NP,
                             int H, int C, AttFold fold) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = C >> 2, HP = H >> 1;
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  // leading blocks: the g_att / edge-mask folds (see att_fold_role)
  if ((int)blockIdx.x < fold.red_blocks + fold.gm_blocks) {
    if ((int)blockIdx.x < fold.red_blocks) att_fold_role(fold, (int)HC);
    else if (MASKED) gm_fold_role(fold, H);
    return;
  }
  const int64_t first = (int64_t)(blockIdx.x - fold.red_blocks - fold.gm_blocks);
  const int64_t nblk = (int64_t)(gridDim.x - fold.red_blocks - fold.gm_blocks);
  const uint32_t wbytes = pair_warp_bytes(C, false);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes, smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * RING,
            2 * row_bytes, lane, 0);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();
  bool hb[2 * VPL], ok[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = lane + 32 * k;
    ok[k] = c < c4;
    hb[2 * k] = 8 * c >= C;
    hb[2 * k + 1] = 8 * c + 4 >= C;
  }
  for (int64_t wid = first * EDGE_WARPS + warp; wid < NP; wid += nblk * EDGE_WARPS) {
    const int64_t slot = wid / HP;
    const int pair = (int)(wid - slot * HP);
    const int64_t node = order ? order[slot] : slot;
    const int hoff = pair * 2 * C;
    float4 acc[2 * VPL];
#pragma unroll
    for (int g = 0; g < 2 * VPL; ++g) acc[g] = f4_zero();
    const int beg = colptr[node], end = colptr[node + 1];
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      int my_dst = 0, my_eid = 0;
      float my_am0 = 0.f, my_am1 = 0.f;
      if (lane < cnt) {
        my_dst = nbr[base + lane];
        my_eid = eid[base + lane];
        const float2 a2 = *reinterpret_cast<const float2*>(alpha + (int64_t)my_eid * H + 2 * pair);
        const float m = MASKED ? emask[my_eid] : 1.f;
        my_am0 = a2.x * m;
        my_am1 = a2.y * m;
      }
      const int npre = min(RING, cnt);
      for (int t = 0; t < npre; ++t) {
        const int i = __shfl_sync(ISG_FULL_MASK, my_dst, t);
        const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
        if (leader)
          ring.issue2(ring.seq + t, g_ep + (int64_t)e * HC + hoff, pol_stream, gout + (int64_t)i * ld_g + hoff, pol_keep,
                      row_bytes);
      }
      for (int t = 0; t < cnt; ++t) {
        const uint32_t s = ring.seq + t;
        const float am0 = __shfl_sync(ISG_FULL_MASK, my_am0, t), am1 = __shfl_sync(ISG_FULL_MASK, my_am1, t);
        const int tn = t + RING < cnt ? t + RING : t;
        const int in_ = __shfl_sync(ISG_FULL_MASK, my_dst, tn);
        const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
        ring.wait(s);
        const uint32_t sg = ring.slot(s) + l16, sG = sg + row_bytes;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          if (ok[k]) {
            float4 g0, g1, G0, G1;
            bf8_to_f4(lds_u4(sg + 512u * k), g0, g1);
            bf8_to_f4(lds_u4(sG + 512u * k), G0, G1);
            acc[2 * k] = p4_add(acc[2 * k], p4_fma_s(G0, hb[2 * k] ? am1 : am0, g0));
            acc[2 * k + 1] = p4_add(acc[2 * k + 1], p4_fma_s(G1, hb[2 * k + 1] ? am1 : am0, g1));
          }
        }
        __syncwarp();
        if (leader && t + RING < cnt)
          ring.issue2(s + RING, g_ep + (int64_t)en * HC + hoff, pol_stream, gout + (int64_t)in_ * ld_g + hoff, pol_keep,
                      row_bytes);
      }
      ring.seq += cnt;
    }
    __nv_bfloat16* grow = g_xl + node * ld_gx + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (ok[k]) *reinterpret_cast<uint4*>(grow + 8 * (lane + 32 * k)) = f4_to_bf8(acc[2 * k], acc[2 * k + 1]);
  }
}

// backward, dst-major: one warp per (node, head pair); two sweeps per node as in gat_edge_bwd_dst_ring_kernel (sweep 1:
// dot_h = sum_e a m t_h with t_h = <G_h, x_l[src]_h>; sweep 2: per-edge gradients from gl_h = a_h (m t_h - dot_h)).
// The two-loop form only (sweep 2 recomputes the bit-identical t): g_att partial row (node*H + 2p) covers both heads.
template <int VPL, bool MASKED>
__global__ void __launch_bounds__(EDGE_WARPS * 32, ISG_PAIR_DST_CTAS)
gat_edge_bwd_dst_pair_kernel(const __nv_bfloat16* __restrict__ gout, int64_t ld_g, const __nv_bfloat16* __restrict__ xl,
                             const __nv_bfloat16* __restrict__ xr, int64_t ld_x, const __nv_bfloat16* __restrict__ ep,
                             const float* __restrict__ att, const float* __restrict__ emask,
                             const float* __restrict__ alpha, const int* __restrict__ rowptr,
                             const int* __restrict__ nbr, const int* __restrict__ eid, const int* __restrict__ order,
                             __nv_bfloat16* __restrict__ g_xr, int64_t ld_gx, __nv_bfloat16* __restrict__ g_ep,
                             float* __restrict__ gatt_part, float* __restrict__ gm_h, int64_t N, int H, int C,
                             float slope, int* __restrict__ red_cnt, int red_n) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  pdl_enter();  // first global access below (no-op unless launched through launch_pdl)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = C >> 2, HP = H >> 1;
  if (blockIdx.x == 0)  // arrival counters of the fold roles in the src kernel
    for (int v = threadIdx.x; v < red_n; v += blockDim.x) red_cnt[v] = 0;
  const int64_t wid = (int64_t)blockIdx.x * EDGE_WARPS + warp;
  const int64_t slot = wid / HP;
  const int pair = (int)(wid - slot * HP);
  if (slot >= N) return;  // (whole warps; no block-level synchronisation below)
  const int64_t node = order ? order[slot] : slot;
  const uint32_t row_bytes = (uint32_t)C * 4u, l16 = 16u * lane;
  const int64_t HC = (int64_t)H * C;
  const int hoff = pair * 2 * C;
  const uint32_t wbytes = pair_warp_bytes(C, true);
  WarpRing ring;
  ring.init(smem_addr_u32(ring_smem) + warp * wbytes,
            smem_addr_u32(ring_smem) + EDGE_WARPS * wbytes + warp * 8 * (RING + 1), 2 * row_bytes, lane, 1);
  const uint32_t xr_s = ring.data + 2 * RING * row_bytes, att_s = xr_s + row_bytes;
  const uint32_t xbar = ring.extra_bar(0);
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool leader = warp_elect_one();
  bool hb[2 * VPL], ok[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = lane + 32 * k;
    ok[k] = c < c4;
    hb[2 * k] = 8 * c >= C;
    hb[2 * k + 1] = 8 * c + 4 >= C;
  }
  for (int v = lane; v < 2 * c4; v += 32) sts_f4(att_s + 16u * v, Vec4<float>::ld(att + hoff + 4 * v));
  __syncwarp();
  if (leader) {
    ring_expect(xbar, row_bytes);
    ring_copy(xr_s, xr + node * ld_x + hoff, row_bytes, xbar, pol_keep);
  }
  float4 G[2 * VPL], gxr[2 * VPL], gatt[2 * VPL];
  {
    const __nv_bfloat16* grow = gout + node * ld_g + hoff;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (ok[k]) bf8_to_f4(*reinterpret_cast<const uint4*>(grow + 8 * (lane + 32 * k)), G[2 * k], G[2 * k + 1]);
      else G[2 * k] = G[2 * k + 1] = f4_zero();
    }
  }
#pragma unroll
  for (int g = 0; g < 2 * VPL; ++g) { gxr[g] = f4_zero(); gatt[g] = f4_zero(); }
  const int beg = rowptr[node], end = rowptr[node + 1];
  bool xr_ready = false;

  // t_h = <G_h, x_l row in the slot>, one fixed instruction sequence for both sweeps (bit-identical t)
  auto row_dot = [&](uint32_t sx, float& t0, float& t1) {
    float2 pp0 = make_float2(0.f, 0.f), pp1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (ok[k]) {
        float4 x0, x1;
        bf8_to_f4(lds_u4(sx + 512u * k), x0, x1);
        const float2 d0 = p4_dot_acc(G[2 * k], x0, make_float2(0.f, 0.f));
        const float2 d1 = p4_dot_acc(G[2 * k + 1], x1, make_float2(0.f, 0.f));
        sel_add(pp0, pp1, hb[2 * k], d0);
        sel_add(pp0, pp1, hb[2 * k + 1], d1);
      }
    }
    t0 = warp_sum(p2_sum(pp0));
    t1 = warp_sum(p2_sum(pp1));
  };

  // sweep 1
  float dot0 = 0.f, dot1 = 0.f;
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int my_src = 0;
    float my_am0 = 0.f, my_am1 = 0.f;
    if (lane < cnt) {
      my_src = nbr[base + lane];
      const int e = eid[base + lane];
      const float2 a2 = *reinterpret_cast<const float2*>(alpha + (int64_t)e * H + 2 * pair);
      const float m = MASKED ? emask[e] : 1.f;
      my_am0 = a2.x * m;
      my_am1 = a2.y * m;
    }
    const int npre = min(RING, cnt);
    for (int t = 0; t < npre; ++t) {
      const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
      if (leader) ring.issue1(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, row_bytes);
    }
    for (int t = 0; t < cnt; ++t) {
      const uint32_t s = ring.seq + t;
      const float am0 = __shfl_sync(ISG_FULL_MASK, my_am0, t), am1 = __shfl_sync(ISG_FULL_MASK, my_am1, t);
      const int jn = __shfl_sync(ISG_FULL_MASK, my_src, t + RING < cnt ? t + RING : t);
      ring.wait(s);
      float t0, t1;
      row_dot(ring.slot(s) + l16, t0, t1);  // warp_sum inside: every lane is past its slot reads
      if (leader && t + RING < cnt) ring.issue1(s + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, row_bytes);
      dot0 = fmaf(am0, t0, dot0);
      dot1 = fmaf(am1, t1, dot1);
    }
    ring.seq += cnt;
  }
  // sweep 2
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int my_src = 0, my_eid = 0;
    float my_m = 1.f, my_a0 = 0.f, my_a1 = 0.f, my_gm0 = 0.f, my_gm1 = 0.f;
    if (lane < cnt) {
      my_src = nbr[base + lane];
      my_eid = eid[base + lane];
      if (MASKED) my_m = emask[my_eid];
      const float2 a2 = *reinterpret_cast<const float2*>(alpha + (int64_t)my_eid * H + 2 * pair);
      my_a0 = a2.x;
      my_a1 = a2.y;
    }
    const int npre = min(RING, cnt);
    for (int t = 0; t < npre; ++t) {
      const int j = __shfl_sync(ISG_FULL_MASK, my_src, t);
      const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
      if (leader)
        ring.issue2(ring.seq + t, xl + (int64_t)j * ld_x + hoff, pol_keep, ep + (int64_t)e * HC + hoff, pol_stream,
                    row_bytes);
    }
    for (int t = 0; t < cnt; ++t) {
      const uint32_t s = ring.seq + t;
      const int e = __shfl_sync(ISG_FULL_MASK, my_eid, t);
      const float m = MASKED ? __shfl_sync(ISG_FULL_MASK, my_m, t) : 1.f;
      const float a0 = __shfl_sync(ISG_FULL_MASK, my_a0, t), a1 = __shfl_sync(ISG_FULL_MASK, my_a1, t);
      const int tn = t + RING < cnt ? t + RING : t;
      const int jn = __shfl_sync(ISG_FULL_MASK, my_src, tn);
      const int en = __shfl_sync(ISG_FULL_MASK, my_eid, tn);
      ring.wait(s);
      const uint32_t sx = ring.slot(s) + l16, sp = sx + row_bytes;
      float t0, t1;
      row_dot(sx, t0, t1);
      if (!xr_ready) {
        ring_wait(xbar, 0u);
        xr_ready = true;
      }
      const float gl0 = a0 * (m * t0 - dot0), gl1 = a1 * (m * t1 - dot1);  // d loss / d logit[e, h]
      const float glmm0 = MASKED ? gl0 * m * m : gl0, glmm1 = MASKED ? gl1 * m * m : gl1;
      float2 av0 = make_float2(0.f, 0.f), av1 = make_float2(0.f, 0.f);
      __nv_bfloat16* gerow = g_ep + (int64_t)e * HC + hoff;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (ok[k]) {
          float4 x0, x1, e0, e1, r0, r1;
          bf8_to_f4(lds_u4(sx + 512u * k), x0, x1);
          bf8_to_f4(lds_u4(sp + 512u * k), e0, e1);
          bf8_to_f4(lds_u4(xr_s + l16 + 512u * k), r0, r1);
          const uint32_t as = att_s + 32u * (lane + 32 * k);
          float4 gs[2];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int g = 2 * k + hf;
            const bool h1 = hb[g];
            const float4 sv = p4_add(p4_add(hf ? r1 : r0, hf ? x1 : x0), hf ? e1 : e0);
            const float4 u = MASKED ? p4_scale(sv, m) : sv;
            const float4 v = p4_leaky(u, slope);
            const float4 at = lds_f4(as + 16u * hf);
            gatt[g] = p4_fma_s(MASKED ? p4_scale(v, m) : v, h1 ? gl1 : gl0, gatt[g]);
            const float4 lk = make_float4(u.x > 0.f ? 1.f : slope, u.y > 0.f ? 1.f : slope, u.z > 0.f ? 1.f : slope,
                                          u.w > 0.f ? 1.f : slope);
            gs[hf] = p4_scale(p4_mul(at, lk), h1 ? glmm1 : glmm0);
            gxr[g] = p4_add(gxr[g], gs[hf]);
            if (MASKED) sel_add(av0, av1, h1, p4_dot_acc(at, v, make_float2(0.f, 0.f)));
          }
          const uint4 packed = f4_to_bf8(gs[0], gs[1]);
          asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(gerow + 8 * (lane + 32 * k)),
                       "r"(packed.x), "r"(packed.y), "r"(packed.z), "r"(packed.w)
                       : "memory");
        }
      }
      if (MASKED) {
        const float gm0 = warp_sum(2.f * gl0 * p2_sum(av0)) + t0 * a0;
        const float gm1 = warp_sum(2.f * gl1 * p2_sum(av1)) + t1 * a1;
        if (lane == t) { my_gm0 = gm0; my_gm1 = gm1; }
      }
      __syncwarp();
      if (leader && t + RING < cnt)
        ring.issue2(s + RING, xl + (int64_t)jn * ld_x + hoff, pol_keep, ep + (int64_t)en * HC + hoff, pol_stream,
                    row_bytes);
    }
    ring.seq += cnt;
    if (MASKED && lane < cnt) {
      gm_h[(int64_t)my_eid * H + 2 * pair] = my_gm0;
      gm_h[(int64_t)my_eid * H + 2 * pair + 1] = my_gm1;
    }
  }
  if (!xr_ready) ring_wait(xbar, 0u);
  __syncwarp();
  __nv_bfloat16* grow = g_xr + node * ld_gx + hoff;
  float* prow = gatt_part + (node * H + 2 * pair) * C;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (ok[k]) {
      const int c = lane + 32 * k;
      *reinterpret_cast<uint4*>(grow + 8 * c) = f4_to_bf8(gxr[2 * k], gxr[2 * k + 1]);
      Vec4<float>::st(prow + 8 * c, gatt[2 * k]);
      Vec4<float>::st(prow + 8 * c + 4, gatt[2 * k + 1]);
    }
  }
}


struct BwdRingPlan {
  int K;
  int64_t cols, blocks, warps, parts2;
};
inline BwdRingPlan bwd_ring_plan(int64_t N, int H) {
  BwdRingPlan p;
  if (N < 1) N = 1;
  if (H < 1) H = 1;
  const int64_t target_cols = (int64_t)ISG_NUM_SMS * 6 * 6;
  p.K = (int)(N / target_cols);
  if (p.K < 1) p.K = 1;
  p.cols = (N + p.K - 1) / p.K;
  p.blocks = (p.cols * H + EDGE_WARPS - 1) / EDGE_WARPS;
  {  // next block count whose warps divide by H (closed form, see bwd_grid_blocks)
    int64_t g = H, r = EDGE_WARPS;
    while (r) { const int64_t t = g % r; g = r; r = t; }
    const int64_t step = H / g;
    p.blocks = (p.blocks + step - 1) / step * step;
  }
  p.warps = p.blocks * EDGE_WARPS;
  p.parts2 = (p.warps / H + GR_ROWS - 1) / GR_ROWS;
  return p;
}

// ------------------------------------------------------------------------------------------
// NodeMaskToEdgeMask (reference sampling/node_edge_masks.py:5-19)
// ------------------------------------------------------------------------------------------
__global__ void node_edge_mask_fwd_kernel(const float* __restrict__ m, const int64_t* __restrict__ ei,
                                          int64_t E, float* __restrict__ em) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < E) em[e] = __fmul_rn(m[ei[e]], m[ei[E + e]]);
}

// g_mask[i] = sum_{e -> i} g_em[e]  (custom backward: dst only).  g_em [E] in original order.
__global__ void node_edge_mask_bwd_kernel(const float* __restrict__ g_em, const int* __restrict__ rowptr,
                                          const int* __restrict__ eid, int64_t N, float* __restrict__ g_m) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = 0.f;
  for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) s += g_em[eid[p]];
  g_m[i] = s;
}

// g_edge_mask[e] = sum_h gm_h[e,h]
__global__ void gm_head_sum_kernel(const float* __restrict__ gm_h, int64_t E, int H, float* __restrict__ g_em) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  float s = 0.f;
  for (int h = 0; h < H; ++h) s += gm_h[e * H + h];
  g_em[e] = s;
}

inline int bwd_grid_blocks(int64_t N, int H) {
  // persistent grid: CTAs hold EDGE_WARPS warps; total warps must be a multiple of H
  int64_t want = (N * H + EDGE_WARPS - 1) / EDGE_WARPS;
  int64_t cap = (int64_t)ISG_NUM_SMS * 4;  // 4 CTAs of 128 threads per SM at ~100 regs
  int64_t blocks = want < cap ? want : cap;
  if (blocks < 1) blocks = 1;
  // round up to the next count whose warps divide by H: a multiple of H / gcd(H, EDGE_WARPS) (closed form — the
  // sizing helper must not spin on an absurd H)
  int64_t g = H, r = EDGE_WARPS;
  while (r) { const int64_t t = g % r; g = r; r = t; }
  const int64_t step = H / g;
  blocks = (blocks + step - 1) / step * step;
  return (int)blocks;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int fold_col_blocks(int H, int C) { return (H * C / 4 + FOLD_THREADS - 1) / FOLD_THREADS; }

template <int VPL, bool MASKED>
int launch_fwd_ring(const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj, const float* att,
                    const float* bias, const float* emask, const int* dst_ptr, const int* dst_nbr,
                    const int* dst_eid, const int* dst_order, void* out, int64_t ld_out, float* alpha, int64_t N,
                    int H, int C, float slope, cudaStream_t stream) {
  const int64_t NH = N * H;
  const size_t smem = ring_smem_bytes(C, 2, 1);
  // the reference shape (C = 300, H = 4) runs the build with C and H as compile-time constants
  const bool ref_shape = VPL == 3 && C == 300 && H == 4;
  auto kern = ref_shape ? gat_edge_fwd_ring_kernel<VPL, MASKED, VPL == 3 ? 300 : 0, VPL == 3 ? 4 : 0>
                        : gat_edge_fwd_ring_kernel<VPL, MASKED, 0, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // one CTA per EDGE_WARPS (node, head) tasks: the block scheduler balances the degree spread
  const int64_t blocks = ceil_div(NH, (int64_t)EDGE_WARPS);
  e = launch_pdl(kern, dim3((unsigned)blocks), dim3(EDGE_WARPS * 32), smem, stream, (const float*)x_l,
                 (const float*)x_r, ld_x, (const float*)e_proj, att, bias, emask, dst_ptr, dst_nbr, dst_eid, dst_order,
                 (float*)out, ld_out, alpha, NH, H, C, slope);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

// g_eproj stores of the dst pass carry an L2 evict_last hint: the src pass and the two lin_edge backward products
// re-read them (191 MB at c3, more than L2 keeps, so this only raises the share that survives).  Measured in-step at
// c3: backward 0.189 -> 0.186 ms.  ISG_EDGE_EP_KEEP=0 restores the plain streaming stores.
inline int ep_keep_mode() {
  static const int v = getenv("ISG_EDGE_EP_KEEP") ? atoi(getenv("ISG_EDGE_EP_KEEP")) : 1;
  return v;
}

// ISG_EDGE_FOLD=0: g_att / edge-mask folds as separate launches instead of block roles of the src kernel
inline int fold_mode() {
  static const int v = getenv("ISG_EDGE_FOLD") ? atoi(getenv("ISG_EDGE_FOLD")) : 1;
  return v;
}

template <int VPL, bool MASKED>
int launch_bwd_ring(const void* g_out, int64_t ld_g, const void* x_l, const void* x_r, int64_t ld_x,
                    const void* e_proj, const float* att, const float* emask, const float* alpha,
                    const int* dst_ptr, const int* dst_nbr, const int* dst_eid, const int* dst_order,
                    const int* src_ptr, const int* src_nbr, const int* src_eid, const int* src_order, void* g_xl,
                    void* g_xr, int64_t ld_gx, void* g_eproj, float* g_att, float* g_emask, int64_t N, int64_t E,
                    int H, int C, float slope, float* gatt_part, float* gatt_part2, float* gm_h, int* red_cnt,
                    const BwdRingPlan& plan, cudaStream_t stream) {
  const size_t smem_d = ring_smem_bytes(C, 2, 1), smem_s = ring_smem_bytes(C, 0, 0);
  const int fold_cb = fold_col_blocks(H, C);
  const bool ref_shape = VPL == 3 && C == 300 && H == 4;
  auto kd = ref_shape ? gat_edge_bwd_dst_ring_kernel<VPL, MASKED, VPL == 3 ? 300 : 0, VPL == 3 ? 4 : 0>
                      : gat_edge_bwd_dst_ring_kernel<VPL, MASKED, 0, 0>;
  auto ks = ref_shape ? gat_edge_bwd_src_ring_kernel<VPL, MASKED, VPL == 3 ? 300 : 0, VPL == 3 ? 4 : 0>
                      : gat_edge_bwd_src_ring_kernel<VPL, MASKED, 0, 0>;
  cudaError_t e = cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
  if (e != cudaSuccess) return (int)e;
  e = launch_pdl(kd, dim3((unsigned)plan.blocks), dim3(EDGE_WARPS * 32), smem_d, stream, (const float*)g_out, ld_g,
                 (const float*)x_l, (const float*)x_r, ld_x, (const float*)e_proj, att, emask, alpha, dst_ptr, dst_nbr,
                 dst_eid, dst_order, (float*)g_xr, ld_gx, (float*)g_eproj, gatt_part, gm_h, N, H, C, slope, plan.K,
                 ep_keep_mode(), red_cnt, fold_cb);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  const int64_t NH = N * H;
  AttFold fold;
  fold.part = gatt_part;
  fold.part2 = gatt_part2;
  fold.g_att = g_att;
  fold.cnt = red_cnt;
  fold.gm_h = gm_h;
  fold.g_emask = g_emask;
  fold.rows = plan.warps / H;
  fold.E = E;
  fold.parts2 = (int)plan.parts2;
  fold.red_blocks = (int)plan.parts2 * fold_cb;
  fold.gm_blocks = (MASKED && E > 0) ? (int)std::min<int64_t>(ceil_div(E, (int64_t)FOLD_THREADS), 2 * ISG_NUM_SMS) : 0;
  const bool folded = fold_mode() != 0;
  if (!folded) {  // ISG_EDGE_FOLD=0: the fold as launches of its own, ahead of the src pass
    fold.red_blocks = fold.gm_blocks = 0;
    const int HC = H * C;
    gat_att_reduce1_kernel<<<dim3(ceil_div(HC / 4, 64), (unsigned)plan.parts2), 64, 0, stream>>>(gatt_part, fold.rows, HC,
                                                                                             gatt_part2);
    ISG_CHECK_LAUNCH();
    gat_att_reduce2_kernel<<<ceil_div(HC / 4, 64), dim3(64, GR2_LANES), 0, stream>>>(gatt_part2, (int)plan.parts2, HC,
                                                                                     g_att);
    ISG_CHECK_LAUNCH();
  }
  e = launch_pdl(ks, dim3((unsigned)(fold.red_blocks + fold.gm_blocks + ceil_div(NH, (int64_t)EDGE_WARPS))),
                 dim3(EDGE_WARPS * 32), smem_s, stream, (const float*)g_out, ld_g, (const float*)g_eproj, emask, alpha,
                 src_ptr, src_nbr, src_eid, src_order, (float*)g_xl, ld_gx, NH, H, C, fold);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  if (!folded && MASKED && E > 0) {
    gm_head_sum_kernel<<<ceil_div(E, 256), 256, 0, stream>>>(gm_h, E, H, g_emask);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

// Lag between the dst role and the src role of a node, in blocks (see gat_edge_bwd_fused_ring_kernel): at least the
// blocks that hold Nmax nodes (deadlock freedom) and by default 640, which lets the ~888 resident blocks of dst
// work drain before the matching src blocks start while keeping the in-flight g_eproj rows (~40 MB) inside L2.
inline int fused_lag_blocks(int nmax, int H) {
  static const int env = [] {
    const char* v = getenv("ISG_EDGE_LAG_BLOCKS");
    return v ? atoi(v) : 0;
  }();
  const int need = (int)(((int64_t)nmax * H + EDGE_WARPS - 1) / EDGE_WARPS) + 1;
  const int want = env > 0 ? env : 640;
  return want > need ? want : need;
}

template <int VPL, bool MASKED>
int launch_bwd_fused(const void* g_out, int64_t ld_g, const void* x_l, const void* x_r, int64_t ld_x,
                     const void* e_proj, const float* att, const float* emask, const float* alpha,
                     const int* dst_ptr, const int* dst_nbr, const int* dst_eid, const int* src_ptr,
                     const int* src_nbr, const int* src_eid, const int* batch32, const int* graph_ptr, int64_t B,
                     int nmax, void* g_xl, void* g_xr, int64_t ld_gx, void* g_eproj, float* g_att, float* g_emask,
                     int64_t N, int64_t E, int H, int C, float slope, float* gatt_part, float* gatt_part2,
                     float* gm_h, int* sync, cudaStream_t stream) {
  const size_t smem = ring_smem_bytes(C, 2, 1);
  const bool ref_shape = VPL == 3 && C == 300 && H == 4;
  auto kf = ref_shape ? gat_edge_bwd_fused_ring_kernel<VPL, MASKED, VPL == 3 ? 300 : 0, VPL == 3 ? 4 : 0>
                      : gat_edge_bwd_fused_ring_kernel<VPL, MASKED, 0, 0>;
  cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(sync, 0, (size_t)(1 + B) * sizeof(int), stream)) != cudaSuccess) return (int)e;
  const int64_t NH = N * H;
  const int64_t nb = ceil_div(NH, (int64_t)EDGE_WARPS);
  const int lag = fused_lag_blocks(nmax, H);
  const int64_t grid = 2 * (nb + lag);
  if (grid >= (int64_t)INT32_MAX) return ISG_EUNSUPPORTED;
  kf<<<(unsigned)grid, EDGE_WARPS * 32, smem, stream>>>(
      (const float*)g_out, ld_g, (const float*)x_l, (const float*)x_r, ld_x, (const float*)e_proj, att, emask, alpha,
      dst_ptr, dst_nbr, dst_eid, src_ptr, src_nbr, src_eid, batch32, graph_ptr, (float*)g_xl, (float*)g_xr, ld_gx,
      (float*)g_eproj, gatt_part, gm_h, sync, NH, nb, lag, H, C, slope);
  ISG_CHECK_LAUNCH();
  const int HC = H * C;
  const int64_t parts2 = (N + GR_ROWS - 1) / GR_ROWS;  // gatt_part viewed as [N rows, H*C]
  gat_att_reduce1_kernel<<<dim3(ceil_div(HC / 4, 64), (unsigned)parts2), 64, 0, stream>>>(gatt_part, N, HC, gatt_part2);
  ISG_CHECK_LAUNCH();
  gat_att_reduce2_kernel<<<ceil_div(HC / 4, 64), dim3(64, GR2_LANES), 0, stream>>>(gatt_part2, (int)parts2, HC, g_att);
  ISG_CHECK_LAUNCH();
  if (MASKED && E > 0) {
    gm_head_sum_kernel<<<ceil_div(E, 256), 256, 0, stream>>>(gm_h, E, H, g_emask);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

template <typename T, int VPL>
int launch_fwd(const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj, const float* att,
               const float* bias, const float* emask, const int* dst_ptr, const int* dst_nbr,
               const int* dst_eid, void* out, int64_t ld_out, float* alpha, int64_t N, int H, int C,
               float slope, cudaStream_t stream) {
  const int64_t NH = N * H;
  const int blocks = ceil_div(NH, EDGE_WARPS);
  gat_edge_fwd_kernel<T, VPL><<<blocks, EDGE_WARPS * 32, 0, stream>>>(
      (const T*)x_l, (const T*)x_r, ld_x, (const T*)e_proj, att, bias, emask, dst_ptr, dst_nbr, dst_eid,
      (T*)out, ld_out, alpha, NH, H, C, slope);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

template <typename T, int VPL, bool MASKED>
int launch_bwd(const void* g_out, int64_t ld_g, const void* x_l, const void* x_r, int64_t ld_x,
               const void* e_proj, const float* att, const float* bias, const float* emask,
               const float* alpha, const void* out, int64_t ld_out, const int* dst_ptr,
               const int* dst_nbr, const int* dst_eid, const int* src_ptr, const int* src_nbr,
               const int* src_eid, void* g_xl, void* g_xr, int64_t ld_gx, void* g_eproj, float* g_att,
               float* g_emask, int64_t N, int64_t E, int H, int C, float slope, float* gatt_part,
               float* gm_h, cudaStream_t stream) {
  const int blocks = bwd_grid_blocks(N, H);
  gat_edge_bwd_dst_kernel<T, VPL, MASKED><<<blocks, EDGE_WARPS * 32, 0, stream>>>(
      (const T*)g_out, ld_g, (const T*)x_l, (const T*)x_r, ld_x, (const T*)e_proj, att, bias, emask, alpha,
      (const T*)out, ld_out, dst_ptr, dst_nbr, dst_eid, (T*)g_xr, ld_gx, (T*)g_eproj, gatt_part, gm_h, N, H,
      C, slope);
  ISG_CHECK_LAUNCH();
  gat_att_reduce_kernel<<<dim3(ceil_div(C, 32), H), dim3(32, 8), 0, stream>>>(
      gatt_part, (int64_t)blocks * EDGE_WARPS, H, C, g_att);
  ISG_CHECK_LAUNCH();
  const int64_t NH = N * H;
  gat_edge_bwd_src_kernel<T, VPL, MASKED><<<ceil_div(NH, EDGE_WARPS), EDGE_WARPS * 32, 0, stream>>>(
      (const T*)g_out, ld_g, (const T*)g_eproj, emask, alpha, src_ptr, src_nbr, src_eid, (T*)g_xl, ld_gx, NH,
      H, C);
  ISG_CHECK_LAUNCH();
  if (MASKED && E > 0) {
    gm_head_sum_kernel<<<ceil_div(E, 256), 256, 0, stream>>>(gm_h, E, H, g_emask);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}


template <int VPL, bool MASKED>
int launch_fwd_pair(const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj, const float* att, const float* bias,
                    const float* emask, const int* dst_ptr, const int* dst_nbr, const int* dst_eid, const int* dst_order,
                    void* out, int64_t ld_out, float* alpha, int64_t N, int H, int C, float slope, cudaStream_t stream) {
  const int64_t NP = N * (H / 2);
  const size_t smem = pair_smem_bytes(C, true, 1);
  auto kern = gat_edge_fwd_pair_kernel<VPL, MASKED>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = launch_pdl(kern, dim3((unsigned)ceil_div(NP, (int64_t)EDGE_WARPS)), dim3(EDGE_WARPS * 32), smem, stream,
                 (const __nv_bfloat16*)x_l, (const __nv_bfloat16*)x_r, ld_x, (const __nv_bfloat16*)e_proj, att, bias, emask,
                 dst_ptr, dst_nbr, dst_eid, dst_order, (__nv_bfloat16*)out, ld_out, alpha, NP, H, C, slope);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

template <int VPL, bool MASKED>
int launch_bwd_pair(const void* g_out, int64_t ld_g, const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj,
                    const float* att, const float* emask, const float* alpha, const int* dst_ptr, const int* dst_nbr,
                    const int* dst_eid, const int* dst_order, const int* src_ptr, const int* src_nbr, const int* src_eid,
                    const int* src_order, void* g_xl, void* g_xr, int64_t ld_gx, void* g_eproj, float* g_att,
                    float* g_emask, int64_t N, int64_t E, int H, int C, float slope, float* gatt_part, float* gatt_part2,
                    float* gm_h, int* red_cnt, cudaStream_t stream) {
  const size_t smem_d = pair_smem_bytes(C, true, 1), smem_s = pair_smem_bytes(C, false, 0);
  const int fold_cb = fold_col_blocks(H, C);
  auto kd = gat_edge_bwd_dst_pair_kernel<VPL, MASKED>;
  auto ks = gat_edge_bwd_src_pair_kernel<VPL, MASKED>;
  cudaError_t e = cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
  if (e != cudaSuccess) return (int)e;
  const int64_t NP = N * (H / 2);
  const unsigned blocks = (unsigned)ceil_div(NP, (int64_t)EDGE_WARPS);
  e = launch_pdl(kd, dim3(blocks), dim3(EDGE_WARPS * 32), smem_d, stream, (const __nv_bfloat16*)g_out, ld_g,
                 (const __nv_bfloat16*)x_l, (const __nv_bfloat16*)x_r, ld_x, (const __nv_bfloat16*)e_proj, att, emask, alpha,
                 dst_ptr, dst_nbr, dst_eid, dst_order, (__nv_bfloat16*)g_xr, ld_gx, (__nv_bfloat16*)g_eproj, gatt_part, gm_h,
                 N, H, C, slope, red_cnt, fold_cb);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  const int HC = H * C;
  const int64_t parts2 = (N + GR_ROWS - 1) / GR_ROWS;  // gatt_part viewed as [N rows, H*C], keyed by the node
  AttFold fold;
  fold.part = gatt_part;
  fold.part2 = gatt_part2;
  fold.g_att = g_att;
  fold.cnt = red_cnt;
  fold.gm_h = gm_h;
  fold.g_emask = g_emask;
  fold.rows = N;
  fold.E = E;
  fold.parts2 = (int)parts2;
  fold.red_blocks = (int)parts2 * fold_cb;
  fold.gm_blocks = (MASKED && E > 0) ? (int)std::min<int64_t>(ceil_div(E, (int64_t)FOLD_THREADS), 2 * ISG_NUM_SMS) : 0;
  const bool folded = fold_mode() != 0;
  if (!folded) {
    fold.red_blocks = fold.gm_blocks = 0;
    gat_att_reduce1_kernel<<<dim3(ceil_div(HC / 4, 64), (unsigned)parts2), 64, 0, stream>>>(gatt_part, N, HC, gatt_part2);
    ISG_CHECK_LAUNCH();
    gat_att_reduce2_kernel<<<ceil_div(HC / 4, 64), dim3(64, GR2_LANES), 0, stream>>>(gatt_part2, (int)parts2, HC, g_att);
    ISG_CHECK_LAUNCH();
  }
  e = launch_pdl(ks, dim3(blocks + (unsigned)(fold.red_blocks + fold.gm_blocks)), dim3(EDGE_WARPS * 32), smem_s, stream,
                 (const __nv_bfloat16*)g_out, ld_g, (const __nv_bfloat16*)g_eproj, emask, alpha, src_ptr, src_nbr, src_eid,
                 src_order, (__nv_bfloat16*)g_xl, ld_gx, NP, H, C, fold);
  if (e != cudaSuccess) return (int)e;
  ISG_CHECK_LAUNCH();
  if (!folded && MASKED && E > 0) {
    gm_head_sum_kernel<<<ceil_div(E, 256), 256, 0, stream>>>(gm_h, E, H, g_emask);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

// the pair kernels need an even head count and 16-byte aligned pair rows (ISG_EDGE_BF16_PAIR=0: the 8-byte kernels)
inline bool pair_ok(int H, int C, int64_t ld0, int64_t ld1, int64_t ld2, const void* a, const void* b, const void* c,
                    const void* d, const void* e) {
  static const bool off = getenv("ISG_EDGE_BF16_PAIR") != nullptr && atoi(getenv("ISG_EDGE_BF16_PAIR")) == 0;
  if (off || H % 2 || C % 4 || ld0 % 8 || ld1 % 8 || ld2 % 8) return false;
  return !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e) & 15);
}

inline int vpl_for(int C) { return (C / 4 + 31) / 32; }

}  // namespace

extern "C" int isg_gat_edge_fwd(const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj,
                                const float* att, const float* bias, const float* edge_mask,
                                const int32_t* dst_ptr, const int32_t* dst_nbr, const int32_t* dst_eid,
                                const int32_t* dst_order, void* out, int64_t ld_out, float* alpha, int64_t N,
                                int64_t E, int H, int C, float slope, int dtype, void* stream_) {
  if (N < 0 || E < 0 || H <= 0 || C <= 0) return ISG_EINVAL;
  if (C % 4 != 0 || C > 512 || ld_x % 4 != 0 || ld_out % 4 != 0) return ISG_EUNSUPPORTED;
  if (!(slope >= 0.f && slope <= 1.f)) return ISG_EUNSUPPORTED;  // leaky_relu evaluated as max(u, slope*u)
  if (N == 0) return ISG_OK;
  if (!x_l || !x_r || !att || !dst_ptr || !out || (E > 0 && (!e_proj || !dst_nbr || !dst_eid || !alpha)))
    return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
#define ISG_FWD_CASE(T, V)                                                                              \
  return launch_fwd<T, V>(x_l, x_r, ld_x, e_proj, att, bias, edge_mask, dst_ptr, dst_nbr, dst_eid, out, \
                          ld_out, alpha, N, H, C, slope, stream)
  const int vpl = vpl_for(C);
  if (dtype == ISG_F32) {
#define ISG_FWD_RING(V)                                                                                  \
  return edge_mask ? launch_fwd_ring<V, true>(x_l, x_r, ld_x, e_proj, att, bias, edge_mask, dst_ptr, dst_nbr, \
                                              dst_eid, dst_order, out, ld_out, alpha, N, H, C, slope, stream) \
                   : launch_fwd_ring<V, false>(x_l, x_r, ld_x, e_proj, att, bias, edge_mask, dst_ptr, dst_nbr, \
                                               dst_eid, dst_order, out, ld_out, alpha, N, H, C, slope, stream)
    if (((uintptr_t)x_l & 15) || ((uintptr_t)x_r & 15) || ((uintptr_t)e_proj & 15)) return ISG_EUNSUPPORTED;
    switch (vpl) {
      case 1: ISG_FWD_RING(1);
      case 2: ISG_FWD_RING(2);
      case 3: ISG_FWD_RING(3);
      case 4: ISG_FWD_RING(4);
    }
#undef ISG_FWD_RING
  } else if (dtype == ISG_BF16) {
    if (pair_ok(H, C, ld_x, ld_out, 8, x_l, x_r, e_proj, out, nullptr)) {  // 16-byte ring kernels, one warp per head pair
#define ISG_FWD_PAIR(V)                                                                                        \
  return edge_mask ? launch_fwd_pair<V, true>(x_l, x_r, ld_x, e_proj, att, bias, edge_mask, dst_ptr, dst_nbr, dst_eid, \
                                              dst_order, out, ld_out, alpha, N, H, C, slope, stream)               \
                   : launch_fwd_pair<V, false>(x_l, x_r, ld_x, e_proj, att, bias, edge_mask, dst_ptr, dst_nbr,     \
                                               dst_eid, dst_order, out, ld_out, alpha, N, H, C, slope, stream)
      switch (vpl) {
        case 1: ISG_FWD_PAIR(1);
        case 2: ISG_FWD_PAIR(2);
        case 3: ISG_FWD_PAIR(3);
        case 4: ISG_FWD_PAIR(4);
      }
#undef ISG_FWD_PAIR
    }
    switch (vpl) {
      case 1: ISG_FWD_CASE(__nv_bfloat16, 1);
      case 2: ISG_FWD_CASE(__nv_bfloat16, 2);
      case 3: ISG_FWD_CASE(__nv_bfloat16, 3);
      case 4: ISG_FWD_CASE(__nv_bfloat16, 4);
    }
  }
#undef ISG_FWD_CASE
  return ISG_EUNSUPPORTED;
}

extern "C" size_t isg_gat_edge_bwd_workspace_bytes(int64_t N, int64_t E, int64_t B, int H, int C) {
  // max over the three layouts (register-load kernels: persistent grid; two-launch ring kernels: BwdRingPlan;
  // single-launch ring kernel: one g_att partial per (node, head) + the ticket / per-graph counters)
  if (N < 0) N = 0;
  if (B < 0) B = 0;
  const int h = H > 0 ? H : 1;
  const int blocks = bwd_grid_blocks(N > 0 ? N : 1, h);
  const BwdRingPlan plan = bwd_ring_plan(N, h);
  const size_t a = align256((size_t)blocks * EDGE_WARPS * (size_t)C * sizeof(float));
  const size_t b = align256((size_t)plan.warps * (size_t)C * sizeof(float)) +
                   align256((size_t)plan.parts2 * (size_t)h * (size_t)C * sizeof(float));
  const size_t c = align256((size_t)(N + EDGE_WARPS) * (size_t)h * (size_t)C * sizeof(float)) +
                   align256((size_t)((N + GR_ROWS - 1) / GR_ROWS) * (size_t)h * (size_t)C * sizeof(float)) +
                   align256((size_t)(1 + B) * sizeof(int));
  size_t m = a > b ? a : b;
  m = m > c ? m : c;
  return m + align256((size_t)E * (size_t)h * sizeof(float)) + align256((size_t)fold_col_blocks(h, C) * sizeof(int));
}

extern "C" int isg_gat_edge_bwd(const void* g_out, int64_t ld_g, const void* x_l, const void* x_r,
                                int64_t ld_x, const void* e_proj, const float* att, const float* bias,
                                const float* edge_mask, const float* alpha, const void* out, int64_t ld_out,
                                const int32_t* dst_ptr, const int32_t* dst_nbr, const int32_t* dst_eid,
                                const int32_t* dst_order, const int32_t* src_ptr, const int32_t* src_nbr,
                                const int32_t* src_eid, const int32_t* src_order,
                                void* g_xl, void* g_xr, int64_t ld_gx, void* g_eproj, float* g_att,
                                float* g_edge_mask, int64_t N, int64_t E, int H, int C, float slope, int dtype,
                                const int32_t* batch32, const int32_t* graph_ptr, int64_t B, int nmax,
                                void* workspace, size_t ws_bytes, void* stream_) {
  if (N < 0 || E < 0 || H <= 0 || C <= 0) return ISG_EINVAL;
  if (C % 4 != 0 || C > 512 || ld_x % 4 != 0 || ld_out % 4 != 0 || ld_g % 4 != 0 || ld_gx % 4 != 0)
    return ISG_EUNSUPPORTED;
  if (!(slope >= 0.f && slope <= 1.f)) return ISG_EUNSUPPORTED;
  if (!g_att) return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N == 0) {
    cudaError_t err = cudaMemsetAsync(g_att, 0, (size_t)H * C * sizeof(float), stream);
    return err == cudaSuccess ? ISG_OK : (int)err;
  }
  if (!g_out || !x_l || !x_r || !att || !out || !dst_ptr || !src_ptr || !g_xl || !g_xr ||
      (E > 0 && (!e_proj || !alpha || !g_eproj || !dst_nbr || !dst_eid || !src_nbr || !src_eid)))
    return ISG_EINVAL;
  if ((edge_mask != nullptr) != (g_edge_mask != nullptr)) return ISG_EINVAL;
  if (B < 0) return ISG_EINVAL;
  if (ws_bytes < isg_gat_edge_bwd_workspace_bytes(N, E, B, H, C) || !workspace) return ISG_EWORKSPACE;
  const int blocks = bwd_grid_blocks(N, H);
  float* gatt_part = (float*)workspace;
  float* gm_h = (float*)((char*)workspace + align256((size_t)blocks * EDGE_WARPS * (size_t)C * sizeof(float)));
#define ISG_BWD_CASE(T, V)                                                                                   \
  return edge_mask                                                                                           \
             ? launch_bwd<T, V, true>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, bias, edge_mask, alpha, out, \
                                      ld_out, dst_ptr, dst_nbr, dst_eid, src_ptr, src_nbr, src_eid, g_xl,    \
                                      g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, H, C, slope, gatt_part, \
                                      gm_h, stream)                                                          \
             : launch_bwd<T, V, false>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, bias, edge_mask, alpha, out, \
                                       ld_out, dst_ptr, dst_nbr, dst_eid, src_ptr, src_nbr, src_eid, g_xl,   \
                                       g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, H, C, slope,          \
                                       gatt_part, gm_h, stream)
  const int vpl = vpl_for(C);
  if (dtype == ISG_F32) {
#define ISG_BWD_RING(V)                                                                                        \
  return edge_mask ? launch_bwd_ring<V, true>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, edge_mask, alpha,     \
                                              dst_ptr, dst_nbr, dst_eid, dst_order, src_ptr, src_nbr, src_eid, \
                                              src_order, g_xl, g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, \
                                              H, C, slope, rp_part, rp_part2, rp_gmh, rp_cnt, rplan, stream)          \
                   : launch_bwd_ring<V, false>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, edge_mask, alpha,    \
                                               dst_ptr, dst_nbr, dst_eid, dst_order, src_ptr, src_nbr, src_eid, \
                                               src_order, g_xl, g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, \
                                               H, C, slope, rp_part, rp_part2, rp_gmh, rp_cnt, rplan, stream)
    if (((uintptr_t)x_l & 15) || ((uintptr_t)g_out & 15) || ((uintptr_t)e_proj & 15) || ((uintptr_t)g_eproj & 15))
      return ISG_EUNSUPPORTED;
    if (batch32 && graph_ptr && nmax > 0 && B > 0) {
      float* fp_part = (float*)workspace;
      float* fp_part2 = (float*)((char*)workspace + align256((size_t)(N + EDGE_WARPS) * (size_t)H * (size_t)C * sizeof(float)));
      int* fp_sync = (int*)((char*)fp_part2 +
                            align256((size_t)((N + GR_ROWS - 1) / GR_ROWS) * (size_t)H * (size_t)C * sizeof(float)));
      float* fp_gmh = (float*)((char*)fp_sync + align256((size_t)(1 + B) * sizeof(int)));
#define ISG_BWD_FUSED(V)                                                                                         \
  return edge_mask ? launch_bwd_fused<V, true>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, edge_mask, alpha, dst_ptr, \
                                               dst_nbr, dst_eid, src_ptr, src_nbr, src_eid, batch32, graph_ptr, B,  \
                                               nmax, g_xl, g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, H, C,    \
                                               slope, fp_part, fp_part2, fp_gmh, fp_sync, stream)                   \
                   : launch_bwd_fused<V, false>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, edge_mask, alpha, dst_ptr, \
                                                dst_nbr, dst_eid, src_ptr, src_nbr, src_eid, batch32, graph_ptr, B, \
                                                nmax, g_xl, g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, H, C,   \
                                                slope, fp_part, fp_part2, fp_gmh, fp_sync, stream)
      switch (vpl) {
        case 1: ISG_BWD_FUSED(1);
        case 2: ISG_BWD_FUSED(2);
        case 3: ISG_BWD_FUSED(3);
        case 4: ISG_BWD_FUSED(4);
      }
#undef ISG_BWD_FUSED
    }
    const BwdRingPlan rplan = bwd_ring_plan(N, H);
    float* rp_part = (float*)workspace;
    float* rp_part2 = (float*)((char*)workspace + align256((size_t)rplan.warps * (size_t)C * sizeof(float)));
    float* rp_gmh = (float*)((char*)rp_part2 + align256((size_t)rplan.parts2 * (size_t)H * (size_t)C * sizeof(float)));
    int* rp_cnt = (int*)((char*)rp_gmh + align256((size_t)E * (size_t)H * sizeof(float)));
    switch (vpl) {
      case 1: ISG_BWD_RING(1);
      case 2: ISG_BWD_RING(2);
      case 3: ISG_BWD_RING(3);
      case 4: ISG_BWD_RING(4);
    }
#undef ISG_BWD_RING
  } else if (dtype == ISG_BF16) {
    if (pair_ok(H, C, ld_x, ld_g, ld_gx, x_l, x_r, e_proj, g_out, g_eproj) && !(((uintptr_t)g_xl | (uintptr_t)g_xr) & 15)) {
      float* pp_part = (float*)workspace;
      float* pp_part2 = (float*)((char*)workspace + align256((size_t)(N + EDGE_WARPS) * (size_t)H * (size_t)C * sizeof(float)));
      float* pp_gmh = (float*)((char*)pp_part2 +
                               align256((size_t)((N + GR_ROWS - 1) / GR_ROWS) * (size_t)H * (size_t)C * sizeof(float)) +
                               align256((size_t)(1 + B) * sizeof(int)));
      int* pp_cnt = (int*)((char*)pp_gmh + align256((size_t)E * (size_t)H * sizeof(float)));
#define ISG_BWD_PAIR(V)                                                                                          \
  return edge_mask ? launch_bwd_pair<V, true>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, edge_mask, alpha, dst_ptr,   \
                                              dst_nbr, dst_eid, dst_order, src_ptr, src_nbr, src_eid, src_order,  \
                                              g_xl, g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, H, C, slope,   \
                                              pp_part, pp_part2, pp_gmh, pp_cnt, stream)                                 \
                   : launch_bwd_pair<V, false>(g_out, ld_g, x_l, x_r, ld_x, e_proj, att, edge_mask, alpha, dst_ptr,  \
                                               dst_nbr, dst_eid, dst_order, src_ptr, src_nbr, src_eid, src_order, \
                                               g_xl, g_xr, ld_gx, g_eproj, g_att, g_edge_mask, N, E, H, C, slope,  \
                                               pp_part, pp_part2, pp_gmh, pp_cnt, stream)
      switch (vpl) {
        case 1: ISG_BWD_PAIR(1);
        case 2: ISG_BWD_PAIR(2);
        case 3: ISG_BWD_PAIR(3);
        case 4: ISG_BWD_PAIR(4);
      }
#undef ISG_BWD_PAIR
    }
    switch (vpl) {
      case 1: ISG_BWD_CASE(__nv_bfloat16, 1);
      case 2: ISG_BWD_CASE(__nv_bfloat16, 2);
      case 3: ISG_BWD_CASE(__nv_bfloat16, 3);
      case 4: ISG_BWD_CASE(__nv_bfloat16, 4);
    }
  }
#undef ISG_BWD_CASE
  return ISG_EUNSUPPORTED;
}

extern "C" int isg_node_edge_mask_fwd(const float* node_mask, const int64_t* edge_index, int64_t E,
                                      float* edge_mask, void* stream_) {
  if (E < 0) return ISG_EINVAL;
  if (E == 0) return ISG_OK;
  if (!node_mask || !edge_index || !edge_mask) return ISG_EINVAL;
  node_edge_mask_fwd_kernel<<<isg::ceil_div(E, 256), 256, 0, (cudaStream_t)stream_>>>(node_mask, edge_index, E,
                                                                                      edge_mask);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_node_edge_mask_bwd(const float* g_edge_mask, const int32_t* dst_ptr, const int32_t* dst_eid,
                                      int64_t N, float* g_node_mask, void* stream_) {
  if (N < 0) return ISG_EINVAL;
  if (N == 0) return ISG_OK;
  if (!g_edge_mask || !dst_ptr || !dst_eid || !g_node_mask) return ISG_EINVAL;
  node_edge_mask_bwd_kernel<<<isg::ceil_div(N, 128), 128, 0, (cudaStream_t)stream_>>>(g_edge_mask, dst_ptr,
                                                                                      dst_eid, N, g_node_mask);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

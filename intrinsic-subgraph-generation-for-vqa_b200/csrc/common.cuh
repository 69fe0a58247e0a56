// common.cuh — shared device helpers for libisg.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/isg.h"

#define ISG_FULL_MASK 0xffffffffu
#define ISG_NUM_SMS 148  // B200: 2 dies x 74 SMs

#define ISG_CHECK_LAUNCH()                         \
  do {                                             \
    cudaError_t _e = cudaGetLastError();           \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

namespace isg {

// Programmatic dependent launch (griddepcontrol): a kernel launched through launch_pdl() may become resident while
// the previous kernel of its stream is still draining; it must call pdl_enter() BEFORE its first global-memory access
// (the wait returns once every prerequisite grid has completed and its writes are visible, so the data flow is the
// plain stream order), and only on-chip set-up (barrier init, tensor-memory allocation, descriptor prefetch) may
// precede it.  What overlaps is the next grid's launch latency and prologue with this grid's tail.  Both
// instructions are no-ops in a kernel launched without the attribute.  ISG_PDL=0 drops the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_launch_dependents();
}
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ISG_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}
inline void pdl_attribute(cudaLaunchAttribute* a) {
  a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a->val.programmaticStreamSerializationAllowed = 1;
}
// kern<<<grid, block, smem, stream>>>(args...) with the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  pdl_attribute(&attr[0]);
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ISG_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(ISG_FULL_MASK, v, o));
  return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// exact (erf) GELU and its derivative — torch.nn.functional.gelu default.
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// 4-element vector access for fp32 (16 B) and bf16 (8 B) rows.
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
  }
  // streaming load: read-once data (e_proj / g_eproj rows) should not displace gathered rows in L1
  static __device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
  }
  static __device__ __forceinline__ void st(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
  }
  static __device__ __forceinline__ void st_stream(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 cvt(uint2 raw) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ uint2 pack(float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    const __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<const uint32_t*>(&a);
    raw.y = *reinterpret_cast<const uint32_t*>(&b);
    return raw;
  }
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    return cvt(__ldg(reinterpret_cast<const uint2*>(p)));
  }
  static __device__ __forceinline__ float4 ld_stream(const __nv_bfloat16* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return cvt(r);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = pack(v);
  }
  static __device__ __forceinline__ void st_stream(__nv_bfloat16* p, float4 v) {
    const uint2 r = pack(v);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(r.x), "r"(r.y)
                 : "memory");
  }
};

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 f4_fma(float4 a, float s, float4 c) {  // a*s + c
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}
// explicit fma chain: two call sites with the same inputs produce bit-identical results (the edge
// backward relies on that; `a.x*b.x + ...` may be contracted differently per site)
__device__ __forceinline__ float f4_dot_acc(float4 a, float4 b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

// ---- packed fp32 (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE-rn fp32 lanes per instruction).  The edge
// kernels need ~10 fp32 ops per streamed element; at the HBM roofline that is ~0.9 issued
// instructions per cycle per scheduler with scalar ops, so halving the FP instruction count is what
// lets them approach the memory bound.
__device__ __forceinline__ float4 p4_add(float4 a, float4 b) {
  const float2 l = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 h = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(l.x, l.y, h.x, h.y);
}
__device__ __forceinline__ float4 p4_mul(float4 a, float4 b) {
  const float2 l = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 h = __fmul2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(l.x, l.y, h.x, h.y);
}
__device__ __forceinline__ float4 p4_scale(float4 a, float s) {
  const float2 ss = make_float2(s, s);
  const float2 l = __fmul2_rn(make_float2(a.x, a.y), ss);
  const float2 h = __fmul2_rn(make_float2(a.z, a.w), ss);
  return make_float4(l.x, l.y, h.x, h.y);
}
__device__ __forceinline__ float4 p4_fma(float4 a, float4 b, float4 c) {  // a*b + c
  const float2 l = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  const float2 h = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), make_float2(c.z, c.w));
  return make_float4(l.x, l.y, h.x, h.y);
}
__device__ __forceinline__ float4 p4_fma_s(float4 a, float s, float4 c) {  // a*s + c
  const float2 ss = make_float2(s, s);
  const float2 l = __ffma2_rn(make_float2(a.x, a.y), ss, make_float2(c.x, c.y));
  const float2 h = __ffma2_rn(make_float2(a.z, a.w), ss, make_float2(c.z, c.w));
  return make_float4(l.x, l.y, h.x, h.y);
}
// leaky_relu for 0 < slope < 1: max(u, slope*u) has the same value as (u > 0 ? u : slope*u)
__device__ __forceinline__ float4 p4_leaky(float4 u, float slope) {
  const float4 l = p4_scale(u, slope);
  return make_float4(fmaxf(u.x, l.x), fmaxf(u.y, l.y), fmaxf(u.z, l.z), fmaxf(u.w, l.w));
}
// two-lane dot-product accumulator: acc2 += a * b (pairwise); reduce with p2_sum
__device__ __forceinline__ float2 p4_dot_acc(float4 a, float4 b, float2 acc) {
  acc = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), acc);
  return __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), acc);
}
__device__ __forceinline__ float p2_sum(float2 a) { return a.x + a.y; }

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Row ring: per-warp shared-memory staging of gathered / streamed feature rows with bulk async
// copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers.  One elected lane issues the copies
// of the next edges while the warp computes on the current one, so the bytes in flight per SM are
// set by the ring depth and not by the register file (a row slice of C fp32 = 3 float4 per lane).
// ---------------------------------------------------------------------------------------------
// One lane of a converged warp, chosen by elect.sync.  Bulk-copy / TMA / tcgen05 instructions take uniform
// registers: inside a branch ptxas can prove single-lane (an elect.sync predicate) they are emitted directly,
// whereas under `lane == 0` every one is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop.
__device__ __forceinline__ bool warp_elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(e));
  return e != 0;
}
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ring_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ring_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void ring_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                          uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void ring_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();  // a lost copy must fail loudly, not hang the GPU
  }
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

}  // namespace isg

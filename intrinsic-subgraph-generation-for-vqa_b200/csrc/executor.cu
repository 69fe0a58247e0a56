// executor.cu — layer-level entry points of the hot path: ONE C call runs every kernel of one MGAT layer
// (reference: the body of the loop in MGAT.forward, models/mgat.py:131-177 = MaskingGATv2Conv.forward
// models/mgat_v2_conv.py:138-241 with MaskingModel.forward models/masking.py:132-199, then x_proj,
// scatter-SDPA, GraphNorm and the residual), and one more its backward.
//
// Why: through the per-operator entry points the eager drop-in issues ~155 C-ABI calls, ~220 tensor
// allocations and 33 autograd Functions per training step — about as much host time as the GPU needs for the
// kernels (r1: 7.6 ms/step end to end against 5.1 ms of kernels).  Here the host does 4 + 4 calls per step; the
// launches are enqueued back to back from C, activations live in one caller-provided arena and the backward's
// temporaries in one workspace.  No new arithmetic: every step below is one of the isg_* entry points of this
// library, called in the order of the reference's forward (and the reverse for the backward).
//
// Interface: three flat arrays indexed by named slots — dims (int64), scalars (double), ptrs (device
// pointers).  isg_layer_slot("P_XLR") etc. returns a slot's index, so the host binding never hard-codes the
// numbering (lib.py queries it at load time).
#include <string.h>

#include "common.cuh"

// clang-format off
#define ISG_LAYER_DIMS(X)                                                                                      \
  X(D_N) X(D_E) X(D_B) X(D_D) X(D_H) X(D_HID) X(D_NMAX)                                                          \
  X(D_MASKED)          /* 1: this layer draws a node mask (masking_threshold != 1) */                            \
  X(D_SAMPLER)         /* 1 imle, 2 aimle, 3 gumbel, 4 simple */                                                  \
  X(D_K) X(D_GEMM_MODE) X(D_GATE_MODE) X(D_CLOSED) X(D_EDGE_FUSED) X(D_AIMLE_ADAPTIVE)                            \
  X(D_ACC_EDGE_ATTR)   /* 1: g_edge_attr += (another layer already wrote it) */                                   \
  X(D_ACC_GLF)         /* 1: g_glf += */                                                                          \
  X(D_NEED_GEA)        /* 0: skip the edge_attr gradient */                                                       \
  X(D_BF16)            /* 1: bf16 configuration — projections on kind::f16, bf16 storage of xlr / e_proj / out /   \
                          y1 / z1 and of the gradients g_z1 / g_out / g_xlr / g_eproj (the P_*_BF slots are set) */ \
  X(D_SIDE_WGRAD)      /* 1: the weight-gradient products run on P_SIDE_STREAM, forked / joined with P_EV_FORK / P_EV_JOIN */ \
  X(D_DEFER_JOIN)      /* 1 (backward, with D_SIDE_WGRAD): the main stream does not wait for this layer's weight         \
                          gradients at the end of the call — P_EV_JOIN is only recorded on the side stream; the host     \
                          alternates two workspaces, passes the event of the layer that used this one last as           \
                          P_EV_WS_FREE (waited for before the first kernel) and joins after the last layer */            \
  X(D_EPROJ_READY)     /* 1 (forward): P_EPROJ is being written by a lin_edge product the host issued on another    \
                          stream (edge_attr does not depend on the layer); the edge kernel waits for P_EV_EPROJ */  \
  X(D_WS_BYTES)
#define ISG_LAYER_SCALARS(X) X(F_SLOPE) X(F_EPS) X(F_ALPHA) X(F_BETA) X(F_TAU_IN) X(F_TAU_TGT) X(F_GUMBEL_TAU)
#define ISG_LAYER_PTRS(X)                                                                                      \
  /* graph index */                                                                                            \
  X(P_DST_PTR) X(P_DST_NBR) X(P_DST_EID) X(P_SRC_PTR) X(P_SRC_NBR) X(P_SRC_EID) X(P_GRAPH_PTR) X(P_BATCH32)      \
  X(P_EDGE_INDEX) X(P_DST_ORDER) X(P_SRC_ORDER)                                                                  \
  /* inputs */                                                                                                 \
  X(P_X_IN) X(P_INS) X(P_GLF) X(P_EDGE_ATTR) X(P_NOISE) X(P_KEEP)                                                \
  /* parameters */                                                                                             \
  X(P_W_LR) X(P_B_LR) X(P_W_E) X(P_ATT) X(P_BIAS) X(P_WP0) X(P_BP0) X(P_WP2) X(P_BP2) X(P_BN_W) X(P_BN_B)       \
  X(P_BN_MS) X(P_WN) X(P_BNN) X(P_WQ) X(P_BQ) X(P_AIMLE_STATE)                                                   \
  /* bf16 configuration: edge_attr [E, pad8(D)] and the weights as bf16 W [Nout, pad8(K)] / W^T [K, pad8(Nout)] */ \
  X(P_EDGE_ATTR_BF) X(P_W_LR_BF) X(P_W_LR_T_BF) X(P_W_E_BF) X(P_W_E_T_BF) X(P_WP0_BF) X(P_WP0_T_BF) X(P_WP2_BF)   \
  X(P_WP2_T_BF)                                                                                                \
  /* activations (written by the forward, read by the backward) */                                             \
  X(P_XG) X(P_XG_BF) X(P_XLR) X(P_EPROJ) X(P_OUT) X(P_ALPHA) X(P_Z1) X(P_Y1) X(P_Z2) X(P_Y2) X(P_SA) X(P_MEAN) X(P_RSTD)    \
  X(P_H_OUT) X(P_XN_PRE) X(P_XN) X(P_Q_PRE) X(P_Q) X(P_THETA) X(P_MASK) X(P_ZD) X(P_MARG) X(P_EMASK)              \
  /* backward: incoming gradients */                                                                           \
  X(P_G_H_OUT) X(P_G_MASK_EXT)                                                                                   \
  /* backward: outgoing gradients */                                                                           \
  X(P_G_X_IN) X(P_G_INS) X(P_G_GLF) X(P_G_EDGE_ATTR)                                                             \
  X(P_G_W_LR) X(P_G_B_LR) X(P_G_W_E) X(P_G_ATT) X(P_G_BIAS) X(P_G_WP0) X(P_G_BP0) X(P_G_WP2) X(P_G_BP2)          \
  X(P_G_BN_W) X(P_G_BN_B) X(P_G_BN_MS) X(P_G_WN) X(P_G_BNN) X(P_G_WQ) X(P_G_BQ)                                   \
  X(P_SIDE_STREAM) X(P_EV_FORK) X(P_EV_JOIN) X(P_EV_EPROJ) X(P_EV_WS_FREE)                                                      \
  X(P_WS)
// clang-format on

namespace {

#define X(n) n,
enum LayerDim { ISG_LAYER_DIMS(X) NUM_DIMS };
enum LayerScalar { ISG_LAYER_SCALARS(X) NUM_SCALARS };
enum LayerPtr { ISG_LAYER_PTRS(X) NUM_PTRS };
#undef X

struct Slot {
  const char* name;
  int index;
};
#define X(n) {#n, n},
const Slot kSlots[] = {ISG_LAYER_DIMS(X) ISG_LAYER_SCALARS(X) ISG_LAYER_PTRS(X)};
#undef X

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int pad8(int x) { return (x + 7) & ~7; }
inline size_t max2(size_t a, size_t b) { return a > b ? a : b; }

enum { SAMP_IMLE = 1, SAMP_AIMLE = 2, SAMP_GUMBEL = 3, SAMP_SIMPLE = 4 };

// Backward workspace layout (bytes), shared by the size query and the executor.
struct BwdWs {
  size_t g_y2, g_y2_bf, parts, g_z1, g_out, g_xlr, g_ep, g_em, g_xg, dy, g_theta, g_xn, g_q, scratch, colsum, colsum_bytes,
      wgrad, wgrad_bytes, edge, edge_bytes, aimle, aimle_bytes, total;
};
BwdWs bwd_ws(const int64_t* d) {
  const int64_t N = d[D_N], E = d[D_E], B = d[D_B];
  const int D = (int)d[D_D], H = (int)d[D_H], HID = (int)d[D_HID];
  const int HC = H * D;
  BwdWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += al256(bytes);
    return o;
  };
  w.g_y2 = take((size_t)N * D * 4);
  w.g_y2_bf = take(d[D_BF16] ? (size_t)N * pad8(D) * 2 : 0);
  w.parts = take((size_t)3 * B * D * 4);
  w.g_z1 = take((size_t)N * HID * 4);
  w.g_out = take((size_t)N * HC * 4);
  w.g_xlr = take((size_t)N * 2 * HC * 4);
  w.g_ep = take((size_t)E * HC * 4);
  w.g_em = take((size_t)E * 4);
  w.g_xg = take((size_t)N * D * 4);
  w.dy = take((size_t)N * 4);
  w.g_theta = take((size_t)N * 4);
  w.g_xn = take((size_t)N * D * 4);
  w.g_q = take((size_t)B * D * 4);
  w.scratch = take((size_t)N * 4);
  {  // all column sums of the layer run as ONE batched call at the end of the backward (isg_colsum_multi)
    const int64_t rows[9] = {B, B, B, N, N, N, N, N, B};
    const int cols[9] = {D, D, D, D, HID, HC, 2 * HC, D, D};
    w.colsum_bytes = isg_colsum_multi_workspace_bytes(9, rows, cols);
  }
  w.colsum = take(w.colsum_bytes);
  w.wgrad_bytes = max2(max2(isg_linear_wgrad_workspace_bytes(E, HC, D), isg_linear_wgrad_workspace_bytes(N, 2 * HC, D)),
                       max2(max2(isg_linear_wgrad_workspace_bytes(N, HID, HC), isg_linear_wgrad_workspace_bytes(N, D, HID)),
                            max2(isg_linear_wgrad_workspace_bytes(N, D, D), isg_linear_wgrad_workspace_bytes(B, D, D))));
  if (d[D_BF16])
    w.wgrad_bytes = max2(w.wgrad_bytes,
                         max2(max2(isg_linear_bf16_wgrad_workspace_bytes(E, HC, D), isg_linear_bf16_wgrad_workspace_bytes(N, 2 * HC, D)),
                              max2(isg_linear_bf16_wgrad_workspace_bytes(N, HID, HC), isg_linear_bf16_wgrad_workspace_bytes(N, D, HID))));
  w.wgrad = take(w.wgrad_bytes);
  w.edge_bytes = isg_gat_edge_bwd_workspace_bytes(N, E, B, H, D);
  w.edge = take(w.edge_bytes);
  w.aimle_bytes = isg_aimle_workspace_bytes();
  w.aimle = take(w.aimle_bytes);
  w.total = off;
  return w;
}

__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = __fadd_rn(a[i], b[i]);
}

#define CK(call)             \
  do {                       \
    const int rc_ = (call);  \
    if (rc_ != 0) return rc_; \
  } while (0)

template <typename T>
inline T* ptr(void* const* p, int slot) {
  return reinterpret_cast<T*>(p[slot]);
}

}  // namespace

extern "C" int isg_layer_slot(const char* name) {
  if (!name) return -1;
  for (const Slot& s : kSlots)
    if (strcmp(s.name, name) == 0) return s.index;
  return -1;
}

extern "C" int isg_layer_slot_count(int which) {
  return which == 0 ? (int)NUM_DIMS : which == 1 ? (int)NUM_SCALARS : which == 2 ? (int)NUM_PTRS : -1;
}

extern "C" size_t isg_mgat_layer_bwd_workspace_bytes(const int64_t* dims) {
  if (!dims) return 0;
  if (dims[D_N] < 0 || dims[D_E] < 0 || dims[D_B] < 0 || dims[D_D] <= 0 || dims[D_H] <= 0 || dims[D_HID] <= 0)
    return 0;  // the layer calls reject these with ISG_EINVAL; no size wraps around here
  return bwd_ws(dims).total;
}

// ---------------------------------------------------------------------------------------------------------------
// forward of one layer:  gating -> [node mask + edge mask] -> lin_l|lin_r, lin_edge -> edge attention ->
// x_proj (Linear GELU Linear GELU) -> scatter-SDPA + GraphNorm + residual
// ---------------------------------------------------------------------------------------------------------------
extern "C" int isg_mgat_layer_fwd(const int64_t* d, const double* f, void* const* p, void* stream) {
  if (!d || !f || !p) return ISG_EINVAL;
  const int64_t N = d[D_N], E = d[D_E], B = d[D_B];
  const int D = (int)d[D_D], H = (int)d[D_H], HID = (int)d[D_HID], nmax = (int)d[D_NMAX];
  if (N < 0 || E < 0 || B < 0 || D <= 0 || H <= 0 || HID <= 0) return ISG_EINVAL;
  if (N == 0) return ISG_OK;
  const int HC = H * D, mode = (int)d[D_GEMM_MODE], gate_mode = (int)d[D_GATE_MODE], k = (int)d[D_K];
  const int sampler = (int)d[D_SAMPLER];
  const float* x_in = ptr<const float>(p, P_X_IN);
  const float* ins = ptr<const float>(p, P_INS);
  const int32_t* gptr = ptr<const int32_t>(p, P_GRAPH_PTR);
  const int32_t* batch32 = ptr<const int32_t>(p, P_BATCH32);
  float* xg = ptr<float>(p, P_XG);

  CK(isg_instr_gate_fwd(x_in, ins, batch32, N, D, xg, stream));  // mgat_v2_conv.py:156-157

  float* emask = nullptr;
  if (d[D_MASKED]) {  // mgat_v2_conv.py:161-171 -> masking.py:132-199
    float *xn = ptr<float>(p, P_XN), *q = ptr<float>(p, P_Q), *theta = ptr<float>(p, P_THETA);
    float *mask = ptr<float>(p, P_MASK), *zd = ptr<float>(p, P_ZD);
    const float *keep = ptr<const float>(p, P_KEEP), *noise = ptr<const float>(p, P_NOISE);
    emask = ptr<float>(p, P_EMASK);
    CK(isg_linear_fwd(xg, D, p[P_WN], nullptr, nullptr, ptr<const float>(p, P_BNN), xn, D, p[P_XN_PRE], D, N, D, D,
                      ISG_ACT_GELU, gate_mode, ISG_F32, stream));
    CK(isg_linear_fwd(p[P_GLF], D, p[P_WQ], nullptr, nullptr, ptr<const float>(p, P_BQ), q, D, p[P_Q_PRE], D, B, D, D,
                      ISG_ACT_GELU, gate_mode, ISG_F32, stream));
    if ((sampler == SAMP_IMLE || sampler == SAMP_AIMLE) && d[D_CLOSED]) {
      CK(isg_sampler_fused_fwd(xn, q, keep, noise, batch32, gptr, ptr<const int32_t>(p, P_DST_PTR),
                               ptr<const int32_t>(p, P_DST_NBR), ptr<const int32_t>(p, P_DST_EID), B, D, 1, nmax, k,
                               (float)f[F_TAU_IN], theta, mask, zd, emask, stream));
    } else {
      CK(isg_gate_theta_fwd(xn, q, batch32, N, D, 1, keep, theta, stream));
      switch (sampler) {
        case SAMP_IMLE:
        case SAMP_AIMLE:
          CK(isg_topk_mask_fwd(theta, noise, gptr, B, nmax, k, (float)f[F_TAU_IN], mask, zd, stream));
          break;
        case SAMP_GUMBEL:
          CK(isg_gumbel_topk_fwd(theta, noise, gptr, B, nmax, k, (float)f[F_GUMBEL_TAU], mask, zd, stream));
          break;
        case SAMP_SIMPLE:
          CK(isg_simple_marginals_fwd(theta, noise, gptr, B, nmax, k, mask, ptr<float>(p, P_MARG), stream));
          break;
        default:
          return ISG_EUNSUPPORTED;
      }
      CK(isg_node_edge_mask_fwd(mask, ptr<const int64_t>(p, P_EDGE_INDEX), E, emask, stream));
    }
  }

  if (d[D_BF16]) {
    // bf16 configuration: same sequence, projections on kind::f16 (bf16 operands, fp32 accumulation), bf16 storage
    // of the big activations; the gate logits above and scatter-SDPA / GraphNorm below stay fp32
    const int Dp = pad8(D);
    if (HC % 8 || HID % 8) return ISG_EUNSUPPORTED;
    CK(isg_to_bf16(xg, D, N, D, p[P_XG_BF], Dp, stream));
    CK(isg_linear_bf16_fwd(p[P_XG_BF], Dp, p[P_W_LR_BF], Dp, ptr<const float>(p, P_B_LR), p[P_XLR], 2 * HC, nullptr, 0, N,
                           2 * HC, D, ISG_ACT_NONE, ISG_BF16, stream));
    if (E > 0 && !d[D_EPROJ_READY])
      CK(isg_linear_bf16_fwd(p[P_EDGE_ATTR_BF], Dp, p[P_W_E_BF], Dp, nullptr, p[P_EPROJ], HC, nullptr, 0, E, HC, D,
                             ISG_ACT_NONE, ISG_BF16, stream));
    if (d[D_EPROJ_READY] && p[P_EV_EPROJ]) {
      cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)p[P_EV_EPROJ], 0);
      if (e != cudaSuccess) return (int)e;
    }
    const __nv_bfloat16* xlr16 = ptr<const __nv_bfloat16>(p, P_XLR);
    CK(isg_gat_edge_fwd(xlr16, xlr16 + HC, 2 * HC, p[P_EPROJ], ptr<const float>(p, P_ATT), ptr<const float>(p, P_BIAS),
                        emask, ptr<const int32_t>(p, P_DST_PTR), ptr<const int32_t>(p, P_DST_NBR),
                        ptr<const int32_t>(p, P_DST_EID), ptr<const int32_t>(p, P_DST_ORDER), p[P_OUT], HC,
                        ptr<float>(p, P_ALPHA), N, E, H, D, (float)f[F_SLOPE], ISG_BF16, stream));
    CK(isg_linear_bf16_fwd(p[P_OUT], HC, p[P_WP0_BF], HC, ptr<const float>(p, P_BP0), p[P_Y1], HID, p[P_Z1], HID, N, HID,
                           HC, ISG_ACT_GELU, ISG_BF16, stream));
    CK(isg_linear_bf16_fwd(p[P_Y1], HID, p[P_WP2_BF], HID, ptr<const float>(p, P_BP2), p[P_Y2], D, p[P_Z2], D, N, D, HID,
                           ISG_ACT_GELU, ISG_F32, stream));
    CK(isg_sdpa_graphnorm_fwd(ptr<const float>(p, P_Y2), ins, x_in, ptr<const float>(p, P_BN_W),
                              ptr<const float>(p, P_BN_B), ptr<const float>(p, P_BN_MS), gptr, B, D, nmax,
                              (float)f[F_EPS], ptr<float>(p, P_H_OUT), ptr<float>(p, P_SA), ptr<float>(p, P_MEAN),
                              ptr<float>(p, P_RSTD), stream));
    return ISG_OK;
  }
  float* xlr = ptr<float>(p, P_XLR);
  CK(isg_linear_fwd(xg, D, p[P_W_LR], nullptr, nullptr, ptr<const float>(p, P_B_LR), xlr, 2 * HC, nullptr, 0, N, 2 * HC,
                    D, ISG_ACT_NONE, mode, ISG_F32, stream));  // :177,181 (lin_l | lin_r, stacked weights)
  if (E > 0 && !d[D_EPROJ_READY])
    CK(isg_linear_fwd(p[P_EDGE_ATTR], D, p[P_W_E], nullptr, nullptr, nullptr, p[P_EPROJ], HC, nullptr, 0, E, HC, D,
                      ISG_ACT_NONE, mode, ISG_F32, stream));  // :259
  if (d[D_EPROJ_READY] && p[P_EV_EPROJ]) {
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)p[P_EV_EPROJ], 0);
    if (e != cudaSuccess) return (int)e;
  }
  CK(isg_gat_edge_fwd(xlr, xlr + HC, 2 * HC, p[P_EPROJ], ptr<const float>(p, P_ATT), ptr<const float>(p, P_BIAS), emask,
                      ptr<const int32_t>(p, P_DST_PTR), ptr<const int32_t>(p, P_DST_NBR),
                      ptr<const int32_t>(p, P_DST_EID), ptr<const int32_t>(p, P_DST_ORDER), p[P_OUT], HC,
                      ptr<float>(p, P_ALPHA), N, E, H, D,
                      (float)f[F_SLOPE], ISG_F32, stream));  // :215,243-279
  // mgat.py:156 x_proj = Linear(HC, HID) GELU Linear(HID, D) GELU
  CK(isg_linear_fwd(p[P_OUT], HC, p[P_WP0], nullptr, nullptr, ptr<const float>(p, P_BP0), p[P_Y1], HID, p[P_Z1], HID, N,
                    HID, HC, ISG_ACT_GELU, mode, ISG_F32, stream));
  CK(isg_linear_fwd(p[P_Y1], HID, p[P_WP2], nullptr, nullptr, ptr<const float>(p, P_BP2), p[P_Y2], D, p[P_Z2], D, N, D,
                    HID, ISG_ACT_GELU, mode, ISG_F32, stream));
  // mgat.py:168-172
  CK(isg_sdpa_graphnorm_fwd(ptr<const float>(p, P_Y2), ins, x_in, ptr<const float>(p, P_BN_W),
                            ptr<const float>(p, P_BN_B), ptr<const float>(p, P_BN_MS), gptr, B, D, nmax,
                            (float)f[F_EPS], ptr<float>(p, P_H_OUT), ptr<float>(p, P_SA), ptr<float>(p, P_MEAN),
                            ptr<float>(p, P_RSTD), stream));
  return ISG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// backward of one layer.  In: P_G_H_OUT (d loss / d h_out), P_G_MASK_EXT (gradient that reaches this layer's node
// mask from outside MGAT, or NULL).  Out: P_G_X_IN, P_G_INS, P_G_GLF / P_G_EDGE_ATTR (written or accumulated,
// D_ACC_*), and every parameter gradient of the layer.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int isg_mgat_layer_bwd(const int64_t* d, const double* f, void* const* p, void* stream_) {
  if (!d || !f || !p) return ISG_EINVAL;
  const int64_t N = d[D_N], E = d[D_E], B = d[D_B];
  const int D = (int)d[D_D], H = (int)d[D_H], HID = (int)d[D_HID], nmax = (int)d[D_NMAX];
  if (N < 0 || E < 0 || B < 0 || D <= 0 || H <= 0 || HID <= 0) return ISG_EINVAL;
  if (N == 0) return ISG_OK;
  const int HC = H * D, mode = (int)d[D_GEMM_MODE], k = (int)d[D_K], sampler = (int)d[D_SAMPLER];
  const BwdWs w = bwd_ws(d);
  if ((size_t)d[D_WS_BYTES] < w.total || !p[P_WS]) return ISG_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = ptr<char>(p, P_WS);
  float *g_y2 = (float*)(ws + w.g_y2), *parts = (float*)(ws + w.parts), *g_z1 = (float*)(ws + w.g_z1);
  float *g_out = (float*)(ws + w.g_out), *g_xlr = (float*)(ws + w.g_xlr), *g_ep = (float*)(ws + w.g_ep);
  float *g_em = (float*)(ws + w.g_em), *g_xg = (float*)(ws + w.g_xg), *dy = (float*)(ws + w.dy);
  float *g_theta = (float*)(ws + w.g_theta), *g_xn = (float*)(ws + w.g_xn), *g_q = (float*)(ws + w.g_q);
  float* scratch = (float*)(ws + w.scratch);
  void *cs = ws + w.colsum, *wgw = ws + w.wgrad;
  const int32_t* gptr = ptr<const int32_t>(p, P_GRAPH_PTR);
  const int32_t* batch32 = ptr<const int32_t>(p, P_BATCH32);
  const float* ins = ptr<const float>(p, P_INS);
  const float* g_h_out = ptr<const float>(p, P_G_H_OUT);
  const bool masked = d[D_MASKED] != 0;
  // bias / affine gradients are column sums of tensors this backward leaves in its workspace: they are collected
  // here and issued as one batched call (two launches) after the last producer
  const void* cs_in[9];
  float* cs_out[9];
  int64_t cs_ld[9], cs_rows[9];
  int cs_cols[9], cs_dt[9], cs_n = 0;
  auto colsum_t = [&](const void* t, int64_t rows, int cols, int slot, int dt) -> int {
    cs_in[cs_n] = t;
    cs_out[cs_n] = ptr<float>(p, slot);
    cs_ld[cs_n] = cols;
    cs_rows[cs_n] = rows;
    cs_cols[cs_n] = cols;
    cs_dt[cs_n] = dt;
    ++cs_n;
    return ISG_OK;
  };
  auto colsum = [&](const float* t, int64_t rows, int cols, int slot) -> int { return colsum_t(t, rows, cols, slot, ISG_F32); };
  const bool bf = d[D_BF16] != 0;
  const int Dp = pad8(D);
  // Weight gradients on a side stream.  Nothing downstream in this layer's backward reads a weight gradient, and the
  // projections are persistent kernels whose tile counts rarely fill whole waves of 148 CTAs (lin_l|lin_r: 5.007
  // waves): issued on a second stream, a wgrad's CTAs start on the SMs the main stream's kernel leaves idle in its tail
  // (and next to the memory-bound edge kernels).  fork(): the side stream waits for everything issued so far on the
  // main stream; the join at the end of the layer makes the main stream wait for the side stream (the next layer
  // re-uses this workspace).  Capturable: the event edges become graph dependencies.
  const bool side_on = d[D_SIDE_WGRAD] && p[P_SIDE_STREAM] && p[P_EV_FORK] && p[P_EV_JOIN];
  void* wstream = side_on ? p[P_SIDE_STREAM] : stream_;
  auto fork = [&]() -> int {
    if (!side_on) return ISG_OK;
    cudaError_t e = cudaEventRecord((cudaEvent_t)p[P_EV_FORK], stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)p[P_SIDE_STREAM], (cudaEvent_t)p[P_EV_FORK], 0);
    return e == cudaSuccess ? ISG_OK : (int)e;
  };

  if (p[P_EV_WS_FREE]) {  // deferred join: the weight gradients of the layer that used this workspace last
    cudaError_t e = cudaStreamWaitEvent(stream, (cudaEvent_t)p[P_EV_WS_FREE], 0);
    if (e != cudaSuccess) return (int)e;
  }
  // scatter-SDPA + GraphNorm + residual (mgat.py:168-172); the residual's share of g_h_in is added in the last step
  CK(isg_sdpa_graphnorm_bwd(g_h_out, ptr<const float>(p, P_Y2), ins, ptr<const float>(p, P_BN_W),
                            ptr<const float>(p, P_BN_MS), ptr<const float>(p, P_SA), ptr<const float>(p, P_MEAN),
                            ptr<const float>(p, P_RSTD), gptr, B, D, nmax, g_y2, ptr<float>(p, P_G_INS), parts,
                            parts + (size_t)B * D, parts + (size_t)2 * B * D, ptr<const float>(p, P_Z2), stream_));
  CK(colsum(parts, B, D, P_G_BN_W));
  CK(colsum(parts + (size_t)B * D, B, D, P_G_BN_B));
  CK(colsum(parts + (size_t)2 * B * D, B, D, P_G_BN_MS));
  // x_proj[2]: y2 = gelu(z2), z2 = y1 Wp2^T + bp2 — the GELU's backward is fused into the kernel above (z_gelu)
  if (bf) {
    void* gy2b = ws + w.g_y2_bf;
    CK(isg_to_bf16(g_y2, D, N, D, gy2b, Dp, stream_));
    CK(isg_linear_bf16_dgrad(gy2b, Dp, p[P_WP2_T_BF], Dp, p[P_Z1], HID, g_z1, HID, 0, N, D, HID, ISG_BF16, stream_));
    CK(fork());
  CK(isg_linear_bf16_wgrad(gy2b, Dp, p[P_Y1], HID, ptr<float>(p, P_G_WP2), N, D, HID, wgw, w.wgrad_bytes, wstream));
    CK(colsum(g_y2, N, D, P_G_BP2));
    CK(isg_linear_bf16_dgrad(g_z1, HID, p[P_WP0_T_BF], HID, nullptr, 0, g_out, HC, 0, N, HID, HC, ISG_BF16, stream_));
    CK(fork());
  CK(isg_linear_bf16_wgrad(g_z1, HID, p[P_OUT], HC, ptr<float>(p, P_G_WP0), N, HID, HC, wgw, w.wgrad_bytes, wstream));
    CK(colsum_t(g_z1, N, HID, P_G_BP0, ISG_BF16));
  } else {
  CK(isg_linear_dgrad(g_y2, D, p[P_WP2], nullptr, p[P_Z1], HID, g_z1, HID, 0, N, D, HID, mode, ISG_F32,
                      stream_));  // (g_z2 Wp2) * gelu'(z1): the first GELU's derivative in the epilogue
  CK(fork());
  CK(isg_linear_wgrad(g_y2, D, p[P_Y1], HID, ptr<float>(p, P_G_WP2), nullptr, N, D, HID, mode, ISG_F32, wgw,
                      w.wgrad_bytes, wstream));
  CK(colsum(g_y2, N, D, P_G_BP2));
  // x_proj[0]
  CK(isg_linear_dgrad(g_z1, HID, p[P_WP0], nullptr, nullptr, 0, g_out, HC, 0, N, HID, HC, mode, ISG_F32, stream_));
  CK(fork());
  CK(isg_linear_wgrad(g_z1, HID, p[P_OUT], HC, ptr<float>(p, P_G_WP0), nullptr, N, HID, HC, mode, ISG_F32, wgw,
                      w.wgrad_bytes, wstream));
  CK(colsum(g_z1, N, HID, P_G_BP0));
  }
  // edge attention
  const bool fused = d[D_EDGE_FUSED] && d[D_CLOSED] && !bf;
  const float* xlr = ptr<const float>(p, P_XLR);
  const void* xr_ptr = bf ? (const void*)(ptr<const __nv_bfloat16>(p, P_XLR) + HC) : (const void*)(xlr + HC);
  void* g_xr_ptr = bf ? (void*)((__nv_bfloat16*)g_xlr + HC) : (void*)(g_xlr + HC);
  CK(isg_gat_edge_bwd(g_out, HC, xlr, xr_ptr, 2 * HC, p[P_EPROJ], ptr<const float>(p, P_ATT),
                      ptr<const float>(p, P_BIAS), masked ? ptr<const float>(p, P_EMASK) : nullptr,
                      ptr<const float>(p, P_ALPHA), p[P_OUT], HC, ptr<const int32_t>(p, P_DST_PTR),
                      ptr<const int32_t>(p, P_DST_NBR), ptr<const int32_t>(p, P_DST_EID),
                      ptr<const int32_t>(p, P_DST_ORDER), ptr<const int32_t>(p, P_SRC_PTR),
                      ptr<const int32_t>(p, P_SRC_NBR), ptr<const int32_t>(p, P_SRC_EID),
                      ptr<const int32_t>(p, P_SRC_ORDER), g_xlr, g_xr_ptr, 2 * HC, g_ep, ptr<float>(p, P_G_ATT),
                      masked ? g_em : nullptr, N, E, H, D, (float)f[F_SLOPE], bf ? ISG_BF16 : ISG_F32,
                      fused ? batch32 : nullptr,
                      fused ? gptr : nullptr, B, fused ? nmax : 0, ws + w.edge, w.edge_bytes, stream_));
  CK(colsum_t(g_out, N, HC, P_G_BIAS, bf ? ISG_BF16 : ISG_F32));
  if (bf) {
    if (E > 0) {
      if (d[D_NEED_GEA])
        CK(isg_linear_bf16_dgrad(g_ep, HC, p[P_W_E_T_BF], HC, nullptr, 0, p[P_G_EDGE_ATTR], D, d[D_ACC_EDGE_ATTR] ? 1 : 0, E,
                                 HC, D, ISG_F32, stream_));
      CK(fork());
    CK(isg_linear_bf16_wgrad(g_ep, HC, p[P_EDGE_ATTR_BF], Dp, ptr<float>(p, P_G_W_E), E, HC, D, wgw, w.wgrad_bytes,
                               wstream));
    } else {
      cudaError_t e = cudaMemsetAsync(p[P_G_W_E], 0, (size_t)HC * D * 4, stream);
      if (e != cudaSuccess) return (int)e;
    }
    CK(isg_linear_bf16_dgrad(g_xlr, 2 * HC, p[P_W_LR_T_BF], 2 * HC, nullptr, 0, g_xg, D, 0, N, 2 * HC, D, ISG_F32, stream_));
    CK(fork());
  CK(isg_linear_bf16_wgrad(g_xlr, 2 * HC, p[P_XG_BF], Dp, ptr<float>(p, P_G_W_LR), N, 2 * HC, D, wgw, w.wgrad_bytes,
                             wstream));
    CK(colsum_t(g_xlr, N, 2 * HC, P_G_B_LR, ISG_BF16));
  } else {
  // lin_edge
  if (E > 0) {
    if (d[D_NEED_GEA])
      CK(isg_linear_dgrad(g_ep, HC, p[P_W_E], nullptr, nullptr, 0, p[P_G_EDGE_ATTR], D, d[D_ACC_EDGE_ATTR] ? 1 : 0, E, HC,
                          D, mode, ISG_F32, stream_));
    CK(fork());
  CK(isg_linear_wgrad(g_ep, HC, p[P_EDGE_ATTR], D, ptr<float>(p, P_G_W_E), nullptr, E, HC, D, mode, ISG_F32, wgw,
                        w.wgrad_bytes, wstream));
  } else {
    cudaError_t e = cudaMemsetAsync(p[P_G_W_E], 0, (size_t)HC * D * 4, stream);
    if (e != cudaSuccess) return (int)e;
  }
  // lin_l | lin_r
  CK(isg_linear_dgrad(g_xlr, 2 * HC, p[P_W_LR], nullptr, nullptr, 0, g_xg, D, 0, N, 2 * HC, D, mode, ISG_F32, stream_));
  CK(fork());
  CK(isg_linear_wgrad(g_xlr, 2 * HC, p[P_XG], D, ptr<float>(p, P_G_W_LR), nullptr, N, 2 * HC, D, mode, ISG_F32, wgw,
                      w.wgrad_bytes, wstream));
  CK(colsum(g_xlr, N, 2 * HC, P_G_B_LR));
  }
  // node mask: NodeMaskToEdgeMask's custom backward, the sampler's perturbation gradient, the gate projections
  if (masked) {
    CK(isg_node_edge_mask_bwd(g_em, ptr<const int32_t>(p, P_DST_PTR), ptr<const int32_t>(p, P_DST_EID), N, dy, stream_));
    if (p[P_G_MASK_EXT]) {
      add_inplace_kernel<<<isg::ceil_div(N, 256), 256, 0, stream>>>(dy, ptr<const float>(p, P_G_MASK_EXT), N);
      ISG_CHECK_LAUNCH();
    }
    const float *theta = ptr<const float>(p, P_THETA), *noise = ptr<const float>(p, P_NOISE);
    switch (sampler) {
      case SAMP_IMLE:
        CK(isg_imle_bwd(dy, theta, noise, ptr<const float>(p, P_ZD), gptr, B, nmax, k, (float)f[F_ALPHA],
                        (float)f[F_BETA], (float)f[F_TAU_TGT], g_theta, stream_));
        break;
      case SAMP_AIMLE:
        CK(isg_aimle_bwd(dy, theta, noise, gptr, N, B, nmax, k, (float)f[F_TAU_TGT], (int)d[D_AIMLE_ADAPTIVE],
                         ptr<double>(p, P_AIMLE_STATE), g_theta, ws + w.aimle, w.aimle_bytes, stream_));
        break;
      case SAMP_GUMBEL:
        CK(isg_gumbel_topk_bwd(dy, ptr<const float>(p, P_ZD), gptr, B, nmax, k, (float)f[F_GUMBEL_TAU], g_theta,
                               stream_));
        break;
      case SAMP_SIMPLE:
        CK(isg_simple_marginals_bwd(dy, nullptr, theta, gptr, B, nmax, k, g_theta, stream_));
        break;
      default:
        return ISG_EUNSUPPORTED;
    }
    CK(isg_gate_theta_bwd(g_theta, ptr<const float>(p, P_XN), ptr<const float>(p, P_Q), batch32, gptr, N, B, D, 1,
                          ptr<const float>(p, P_KEEP), g_xn, g_q, scratch, stream_));
    // node_nn: xn = gelu(xg Wn^T + bn)
    CK(isg_gelu_bwd(g_xn, ptr<const float>(p, P_XN_PRE), g_xn, N * (int64_t)D, stream_));
    CK(isg_linear_dgrad(g_xn, D, p[P_WN], nullptr, nullptr, 0, g_xg, D, 1, N, D, D, mode, ISG_F32, stream_));
    CK(fork());
  CK(isg_linear_wgrad(g_xn, D, p[P_XG], D, ptr<float>(p, P_G_WN), nullptr, N, D, D, mode, ISG_F32, wgw, w.wgrad_bytes,
                        wstream));
    CK(colsum(g_xn, N, D, P_G_BNN));
    // ques_nn: q = gelu(glf Wq^T + bq)
    CK(isg_gelu_bwd(g_q, ptr<const float>(p, P_Q_PRE), g_q, B * (int64_t)D, stream_));
    CK(isg_linear_dgrad(g_q, D, p[P_WQ], nullptr, nullptr, 0, p[P_G_GLF], D, d[D_ACC_GLF] ? 1 : 0, B, D, D, mode, ISG_F32,
                        stream_));
    CK(fork());
  CK(isg_linear_wgrad(g_q, D, p[P_GLF], D, ptr<float>(p, P_G_WQ), nullptr, B, D, D, mode, ISG_F32, wgw, w.wgrad_bytes,
                        wstream));
    CK(colsum(g_q, B, D, P_G_BQ));
  }
  if (side_on) {  // join: the next layer's backward re-uses this workspace (D_DEFER_JOIN: the host alternates two)
    cudaError_t e = cudaEventRecord((cudaEvent_t)p[P_EV_JOIN], (cudaStream_t)p[P_SIDE_STREAM]);
    if (e == cudaSuccess && !d[D_DEFER_JOIN]) e = cudaStreamWaitEvent(stream, (cudaEvent_t)p[P_EV_JOIN], 0);
    if (e != cudaSuccess) return (int)e;
  }
  CK(isg_colsum_multi(cs_n, cs_in, cs_dt, cs_ld, cs_rows, cs_cols, cs_out, cs, w.colsum_bytes, stream_));
  // gating + residual: g_x_in = g_xg * d gelu(x*ins)/dx + g_h_out;  g_ins += ...
  CK(isg_instr_gate_bwd(g_xg, ptr<const float>(p, P_X_IN), ins, gptr, B, D, g_h_out, 1, ptr<float>(p, P_G_X_IN),
                        ptr<float>(p, P_G_INS), stream_));
  return ISG_OK;
}

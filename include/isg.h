/* isg.h — C ABI of libisg.so: the B200 (sm_100a) kernels behind the ISubGVQA hot path.
 *
 * The reference (DigitalPhonetics/Intrinsic-Subgraph-Generation-for-VQA) has no FFI layer: its
 * hot path reaches native code only through torch_geometric / torch_scatter / ATen from the
 * Python modules cited at each entry point below (paths relative to the reference root).  This
 * header is what a maintainer binds with ctypes to replace those call sites.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all
 *     memory (outputs and workspaces are caller-allocated; the library never allocates,
 *     frees or retains device memory and keeps no global state);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and never
 *     synchronises the host;
 *   - return value: 0 = ok, negative = ISG_E* (bad argument / unsupported shape), positive =
 *     cudaError_t of a failed launch.  Nothing throws, nothing exits;
 *   - feature tensors are row-major; `dtype` selects their storage type: ISG_F32 (parity
 *     configuration) or ISG_BF16 (storage only, fp32 accumulation).  Indices are int32
 *     internally; the int64 COO of PyG is converted once by isg_csr_build;
 *   - H = heads, C = channels per head (C % 4 == 0, C <= 512), HC = H*C.
 */
#ifndef ISG_H_
#define ISG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISG_OK 0
#define ISG_EINVAL (-1)      /* null pointer / negative size / inconsistent argument */
#define ISG_EUNSUPPORTED (-2) /* shape or dtype outside what the kernels are built for */
#define ISG_EWORKSPACE (-3)  /* workspace too small */

#define ISG_F32 0
#define ISG_BF16 1

#define ISG_ACT_NONE 0
#define ISG_OPT_MAX_TENSORS 64 /* tensors per isg_grad_sq_partials / isg_adam_update call (passed by value) */
#define ISG_ACT_GELU 1 /* exact erf GELU, torch.nn.GELU() default */

int isg_version(void);
const char* isg_error_string(int code);

/* ---------------------------------------------------------------------------------------
 * (a) destination-sorted (and source-sorted) CSR rebuild of the COO edge_index.
 * No reference counterpart: the reference keeps COO and lets PyG's propagate() gather /
 * index_add_ (models/mgat_v2_conv.py:215; torch_geometric MessagePassing).  Result ==
 * stable sort of edge ids by key:  *_ptr [N+1], *_eid [E] (original edge id of the p-th
 * sorted edge), *_nbr [E] (the other endpoint: source for the dst ordering, destination for
 * the src ordering).  `status` (1 int32, device) is set to the number of out-of-range indices.
 * ------------------------------------------------------------------------------------- */
size_t isg_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int isg_csr_build(const int64_t* edge_index /* [2,E] */, int64_t num_edges, int64_t num_nodes,
                  int32_t* dst_ptr, int32_t* dst_nbr, int32_t* dst_eid,
                  int32_t* src_ptr, int32_t* src_nbr, int32_t* src_eid,
                  int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* graph_ptr [B+1] from the sorted `batch` vector [N] int64 (replaces the count/cumsum inside
 * torch_geometric.utils.to_dense_batch, models/masking.py:162, and GraphNorm's int(batch.max())).
 * Also writes batch32 [N] (int32 copy) and nmax (1 int32: max nodes per graph). */
int isg_graph_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs,
                  int32_t* graph_ptr, int32_t* batch32, int32_t* nmax, void* stream);

/* crossing (1 int32, device) = number of edges whose endpoints lie in different graphs or out of range.
 * PyG batches (datasets/gqa.py:237-272, Batch.from_data_list) never have such edges; 0 is the precondition
 * of the single-launch edge backward (isg_gat_edge_bwd with graph_ptr != NULL). */
/* Task order of the edge kernels: dst_order / src_order [N] = the nodes whose in- / out-degree is at least
 * max(2, ceil(2E/N)) (twice the mean) first, then all other nodes, each group in increasing node id (a stable
 * partition, deterministic).  The edge kernels run one warp per (node, head); starting the heavy segments first
 * bounds the end-of-launch tail by a mean-sized task while the other ~93 % of the nodes keep the natural order's
 * L1 / L2 re-use of gathered rows.  The order never changes any result.  No reference counterpart. */
size_t isg_degree_order_workspace_bytes(int64_t num_nodes);
int isg_degree_order(const int32_t* dst_ptr, const int32_t* src_ptr, int64_t num_nodes, int64_t num_edges,
                     int32_t* dst_order, int32_t* src_order, void* workspace, size_t workspace_bytes, void* stream);

int isg_graph_closure(const int64_t* edge_index /* [2,E] */, int64_t num_edges, const int64_t* batch,
                      int64_t num_nodes, int32_t* crossing, void* stream);

/* ---------------------------------------------------------------------------------------
 * (b) fused edge kernel = MaskingGATv2Conv.message (models/mgat_v2_conv.py:243-279) + PyG
 * propagate gathers + torch_geometric.utils.softmax (:272) + sum aggregation + bias (:231-232).
 *   s = x_r[dst] + x_l[src] + e_proj[e];  u = s*m;  v = leaky_relu(u, slope);  w = v*m
 *   logit[e,h] = sum_c w*att[h,c];  alpha = exp(l - max_dst) / (sum_dst + 1e-16)
 *   out[dst,h,:] = sum_e x_l[src,h,:] * alpha*m  + bias
 * x_l / x_r: [N, HC] with row pitch ld_x elements (they may be the two halves of one fused
 * projection); e_proj [E,HC] dense, ORIGINAL edge order; edge_mask [E] fp32 or NULL;
 * out [N,HC] pitch ld_out; alpha [E,H] fp32 in ORIGINAL edge order.
 * dst_order / src_order [N] (isg_degree_order) or NULL: the order in which the (node, head) tasks are scheduled.
 * ------------------------------------------------------------------------------------- */
int isg_gat_edge_fwd(const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj,
                     const float* att /* [H*C] */, const float* bias /* [H*C] or NULL */,
                     const float* edge_mask,
                     const int32_t* dst_ptr, const int32_t* dst_nbr, const int32_t* dst_eid,
                     const int32_t* dst_order /* or NULL */, void* out, int64_t ld_out, float* alpha,
                     int64_t num_nodes, int64_t num_edges, int heads, int channels,
                     float negative_slope, int dtype, void* stream);

/* Backward of the above (the reference gets it from autograd through message()/softmax).
 * Inputs: g_out [N,HC] pitch ld_g; saved x_l,x_r,e_proj,att,bias,edge_mask,alpha,out.
 * Outputs: g_xl,g_xr [N,HC] pitch ld_gx; g_eproj [E,HC]; g_att [H*C] fp32;
 *          g_edge_mask [E] fp32 (NULL iff edge_mask NULL).
 * Two deterministic passes: dst-major (g_eproj, g_xr, g_att, g_edge_mask) then src-major
 * (g_xl) — no floating-point atomics.  With batch32 / graph_ptr / nmax given (fp32, and every edge inside
 * one graph: isg_graph_closure == 0) the two passes run as block roles of ONE launch, ordered so that the
 * src role reads g_eproj out of L2 instead of DRAM; results are bit-identical to the two-launch form.
 * batch32 == NULL or graph_ptr == NULL or nmax <= 0 selects the two-launch form. */
size_t isg_gat_edge_bwd_workspace_bytes(int64_t num_nodes, int64_t num_edges, int64_t num_graphs, int heads,
                                        int channels);
int isg_gat_edge_bwd(const void* g_out, int64_t ld_g,
                     const void* x_l, const void* x_r, int64_t ld_x, const void* e_proj,
                     const float* att, const float* bias, const float* edge_mask,
                     const float* alpha, const void* out, int64_t ld_out,
                     const int32_t* dst_ptr, const int32_t* dst_nbr, const int32_t* dst_eid,
                     const int32_t* dst_order /* or NULL */,
                     const int32_t* src_ptr, const int32_t* src_nbr, const int32_t* src_eid,
                     const int32_t* src_order /* or NULL */,
                     void* g_xl, void* g_xr, int64_t ld_gx, void* g_eproj,
                     float* g_att, float* g_edge_mask,
                     int64_t num_nodes, int64_t num_edges, int heads, int channels,
                     float negative_slope, int dtype,
                     const int32_t* batch32 /* [N] or NULL */, const int32_t* graph_ptr /* [B+1] or NULL */,
                     int64_t num_graphs, int nmax,
                     void* workspace, size_t workspace_bytes, void* stream);

/* NodeMaskToEdgeMask (sampling/node_edge_masks.py:5-19).
 * fwd: edge_mask[e] = mask[src]*mask[dst] (fp32).  bwd (the reference's custom, non-true
 * gradient): g_mask[i] = sum over edges with dst == i of g_edge_mask[e]. */
int isg_node_edge_mask_fwd(const float* node_mask, const int64_t* edge_index, int64_t num_edges,
                           float* edge_mask, void* stream);
int isg_node_edge_mask_bwd(const float* g_edge_mask, const int32_t* dst_ptr, const int32_t* dst_eid,
                           int64_t num_nodes, float* g_node_mask, void* stream);

/* ---------------------------------------------------------------------------------------
 * (c) perturb-and-MAP top-k samplers.  Dense layout = torch_geometric.utils.to_dense_batch
 * (models/masking.py:162): graph b owns slots [b*Nmax, (b+1)*Nmax), real nodes first, pads 0.0
 * (pads COMPETE in top-k).  nb_samples S = 1.
 * MAP = select_from_edge_candidates (sampling/methods/deterministic_scheme.py:36-43):
 *   k >= Nmax -> all ones; else mask = score >= (k-th largest score), ties give > k ones.
 * ------------------------------------------------------------------------------------- */

/* IMLE/AIMLE forward (sampling/methods/wrapper.py:75-121, aimle.py:83-138):
 * z = MAP(theta_dense + noise*tau).  theta [N] ragged; noise [B,Nmax] or NULL (=0);
 * outputs: mask [N] ragged (z[valid]) and z_dense [B,Nmax] (saved for backward). */
int isg_topk_mask_fwd(const float* theta, const float* noise, const int32_t* graph_ptr,
                      int64_t num_graphs, int nmax, int k, float tau,
                      float* mask, float* z_dense, void* stream);

/* Fused sampler forward for IMLE / AIMLE — one launch for models/masking.py:151-176 + mgat_v2_conv.py:166-171:
 * theta[n] = gelu(<xn[n], q[...]>/sqrt(D)) * keep[n]  (isg_gate_theta_fwd semantics, bit-identical),
 * z = MAP(theta_dense + noise*tau), mask [N], z_dense [B,Nmax], and — when edge_mask != NULL — the
 * NodeMaskToEdgeMask forward edge_mask[e] = mask[src]*mask[dst] (sampling/node_edge_masks.py:5-12) over the
 * dst-sorted CSR.  The edge-mask part needs every edge inside one graph (isg_graph_closure == 0).
 * Backward: isg_imle_bwd / isg_aimle_bwd, isg_gate_theta_bwd, isg_node_edge_mask_bwd as before. */
int isg_sampler_fused_fwd(const float* xn, const float* q, const float* keep /* [N] or NULL */,
                          const float* noise /* [B,Nmax] or NULL */, const int32_t* batch32,
                          const int32_t* graph_ptr, const int32_t* dst_ptr, const int32_t* dst_nbr,
                          const int32_t* dst_eid, int64_t num_graphs, int dim, int double_gather, int nmax, int k,
                          float tau, float* theta, float* mask, float* z_dense, float* edge_mask /* [E] or NULL */,
                          void* stream);

/* IMLE backward (wrapper.py:124-172, target.py:44-48):
 * z' = MAP(alpha*theta - beta*dy + noise*tau_target);  g_theta = z - z'   (ragged [N]). */
int isg_imle_bwd(const float* dy, const float* theta, const float* noise, const float* z_dense,
                 const int32_t* graph_ptr, int64_t num_graphs, int nmax, int k,
                 float alpha, float beta, float tau_target, float* g_theta, void* stream);

/* AIMLE backward with the adaptive target (aimle.py:141-243, target_aimle.py:87-162),
 * symmetric perturbation.  `state` is 8 doubles on the device, updated in place with no host
 * sync: [0] beta, [1] grad_norm (fp32 value), [2] previous_beta_update, [3] alpha,
 * [4] beta_update_step, [5] grad_norm_decay_rate, [6] target_norm, [7] beta_update_momentum.
 * adaptive = 0 gives the fixed TargetDistribution(alpha,beta) of the validation scheme.
 * workspace: isg_aimle_workspace_bytes(). */
size_t isg_aimle_workspace_bytes(void);
int isg_aimle_bwd(const float* dy, const float* theta, const float* noise,
                  const int32_t* graph_ptr, int64_t num_nodes, int64_t num_graphs, int nmax, int k,
                  float tau_target, int adaptive, double* state, float* g_theta,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Gumbel relaxed top-k with straight-through (sampling/methods/gumbel_scheme.py:26-107,
 * policy edge_candid, tau 0.1, hard=True).  gumbel [B,Nmax] = Gumbel(0,1) noise.
 * fwd: mask [N] ragged = (hard - khot) + khot; saves khot-rounds in `saved`
 * [B, k, Nmax] (softmax one-hot approximations).  bwd: g_theta [N]. */
int isg_gumbel_topk_fwd(const float* theta, const float* gumbel, const int32_t* graph_ptr,
                        int64_t num_graphs, int nmax, int k, float tau,
                        float* mask, float* saved, void* stream);
int isg_gumbel_topk_bwd(const float* dy, const float* saved, const int32_t* graph_ptr,
                        int64_t num_graphs, int nmax, int k, float tau,
                        float* g_theta, void* stream);

/* SIMPLE: exact k-subset marginals + Gumbel top-k sample + straight-through
 * (sampling/methods/simple_scheme.py:44-162 'edge_candid'; sampling/methods/simple.py:113-252 evaluated over
 * the SDD of create_simple_constraint.py:34-73, including simple.py's -1000 dummy pads, which decide the
 * result whenever a logit is exactly 0 — to_dense_batch pads and dropped-out logits are).
 * n_pad = isg_simple_npad(nmax) = nmax rounded up to a power of two (simple_scheme.py:87, slots >= nmax are
 * -1e10).  gumbel [B, n_pad] = Gumbel(0,1) noise (-log(-log U), simple.py:91-96).
 * fwd: mask [N] ragged = (hot - marginals) + marginals, hot = one-hot of topk(theta_pad + gumbel, min(k,nmax));
 *      marginals [B, nmax] or NULL.
 * bwd: g_theta [N] from dy [N] (gradient of mask) and optionally d_marginals [B, nmax] (or NULL); flows
 *      through the positive literals only (simple.py:215-217 detaches the negative weight).
 * Supported: 2 <= n_pad <= 1024, k <= 7, circuit must fit 200 KB of shared memory per graph. */
int isg_simple_npad(int nmax);
int isg_simple_marginals_fwd(const float* theta, const float* gumbel, const int32_t* graph_ptr,
                             int64_t num_graphs, int nmax, int k, float* mask, float* marginals, void* stream);
int isg_simple_marginals_bwd(const float* dy, const float* d_marginals, const float* theta,
                             const int32_t* graph_ptr, int64_t num_graphs, int nmax, int k, float* g_theta,
                             void* stream);

/* ---------------------------------------------------------------------------------------
 * node-side fused segment ops
 * ------------------------------------------------------------------------------------- */

/* instruction gating  y = gelu(x * ins[batch])  (models/mgat_v2_conv.py:156-157).
 * bwd: g_x [N,D] and g_ins [B,D] (segmented sum per graph, deterministic).  g_residual [N,D] or NULL is added
 * to g_x (the gradient that reaches x through the layer's residual connection, models/mgat.py:172);
 * accumulate_ins != 0 adds to g_ins instead of overwriting it. */
int isg_instr_gate_fwd(const float* x, const float* ins, const int32_t* batch32,
                       int64_t num_nodes, int dim, float* y, void* stream);
int isg_instr_gate_bwd(const float* g_y, const float* x, const float* ins,
                       const int32_t* graph_ptr, int64_t num_graphs, int dim,
                       const float* g_residual, int accumulate_ins,
                       float* g_x, float* g_ins, void* stream);

/* gate logits (models/masking.py:151-155).  double_gather = 1: q is [B,D] (one row per graph) and
 * theta[n] = gelu(<xn[n], q[batch[batch[n]]]> / sqrt(D)) — the double gather that results from
 * models/mgat_v2_conv.py:166-168 passing imle_att[batch] into a forward that indexes [batch] again.
 * double_gather = 0: q is [N,D] and theta[n] = gelu(<xn[n], q[batch[n]]> / sqrt(D)) (MaskingModel.forward
 * called directly with a per-node u).  keep [N] or NULL: the dropout keep-mask of models/masking.py:159
 * (0 or 1/(1-p)), multiplied into theta (and into g_theta in the backward).
 * bwd: g_xn [N,D], g_q (same shape as q). */
/* concat_instr variant (models/mgat_v2_conv.py:153-154, `--concat_instr 1`, off by default): y[n] = [x[n],
 * instruction[batch[n]]] [N,2D]; the conv's lin_l / lin_r and the mask's node_nn then take 2D inputs.
 * bwd: g_x = g_y[:, :D] (+ g_residual), g_ins[b] (+)= sum of g_y[n, D:] over the graph's nodes. */
int isg_concat_instr_fwd(const float* x, const float* instruction, const int32_t* batch32, int64_t num_nodes, int dim,
                         float* y /* [N,2D] */, void* stream);
int isg_concat_instr_bwd(const float* g_y /* [N,2D] */, const int32_t* graph_ptr, int64_t num_graphs, int dim,
                         const float* g_residual /* [N,D] or NULL */, int accumulate_ins, float* g_x, float* g_ins,
                         void* stream);
int isg_gate_theta_fwd(const float* xn, const float* q, const int32_t* batch32,
                       int64_t num_nodes, int dim, int double_gather, const float* keep, float* theta,
                       void* stream);
int isg_gate_theta_bwd(const float* g_theta, const float* xn, const float* q,
                       const int32_t* batch32, const int32_t* graph_ptr,
                       int64_t num_nodes, int64_t num_graphs, int dim, int double_gather, const float* keep,
                       float* g_xn, float* g_q, float* scratch /* [N] */, void* stream);

/* scatter-SDPA + GraphNorm + residual (models/mgat.py:168-172; utils/scatter_scaled_dot_product.py:6-15;
 * torch_geometric GraphNorm eps 1e-5):
 *   a = softmax_graph(<ins[b], v_n>/sqrt(D)); y = a*v; o = y - mean*mean_scale;
 *   h_out = weight*o*rsqrt(mean(o^2)+eps) + bias + h_in.
 * saves a [N], mean [B,D], rstd [B,D].  bwd returns g_v [N,D], g_ins [B,D] and per-graph
 * partials gw_part/gb_part/gms_part [B,D] (column-summed by isg_colsum). g_h_in == g_out.
 * z_gelu [N,D] or NULL: when v = gelu(z) closes the projection that produced it (x_proj[2], models/mgat.py:156), the
 * backward of that GELU is fused: g_v is returned already multiplied by gelu'(z_gelu). */
int isg_sdpa_graphnorm_fwd(const float* v, const float* ins, const float* h_in,
                           const float* weight, const float* bias, const float* mean_scale,
                           const int32_t* graph_ptr, int64_t num_graphs, int dim, int nmax, float eps,
                           float* h_out, float* a, float* mean, float* rstd, void* stream);
int isg_sdpa_graphnorm_bwd(const float* g_out, const float* v, const float* ins,
                           const float* weight, const float* mean_scale,
                           const float* a, const float* mean, const float* rstd,
                           const int32_t* graph_ptr, int64_t num_graphs, int dim, int nmax,
                           float* g_v, float* g_ins, float* gw_part, float* gb_part, float* gms_part,
                           const float* z_gelu /* or NULL */, void* stream);

/* masked attention pooling — SURVEY.md section 8 row f1, GlobalAttention.forward
 * (models/att_pooling.py:57-77; called from models/isubgvqa.py:280-287) after its node_nn / ques_nn MLPs:
 *   xm = x * node_mask;  l_n = <xm_n, q[b]> / sqrt(D);  a = torch_geometric softmax over the nodes of graph b
 *   (exp(l - max) / (sum + 1e-16));  out[b] = sum_n a_n * xm_n;  gate [N] = a.
 * node_mask [N] or NULL.  bwd: g_gate [N] or NULL (gradient of the returned gate); g_mask NULL iff node_mask NULL. */
int isg_attn_pool_fwd(const float* x, const float* node_mask, const float* q, const int32_t* graph_ptr,
                      int64_t num_graphs, int dim, int nmax, float* out, float* gate, void* stream);
int isg_attn_pool_bwd(const float* g_out, const float* g_gate, const float* x, const float* node_mask,
                      const float* q, const float* gate, const int32_t* graph_ptr, int64_t num_graphs, int dim,
                      int nmax, float* g_x, float* g_mask, float* g_q, void* stream);

/* ---------------------------------------------------------------------------------------
 * (d) dense projections (lin_l / lin_r / lin_edge models/mgat_v2_conv.py:63-103,177,181,259;
 * x_proj models/mgat.py:79-89,156; node_nn / ques_nn models/masking.py:82-87,137,152).
 *   fwd:   y = act(x W^T + b);  optionally also writes the pre-activation z (for backward)
 *   dgrad: g_x = (g_y W) [* gelu'(z_prev) if z_prev != NULL]
 *   wgrad: g_W = g_y^T x (deterministic split over M; bias gradient = isg_colsum(g_y))
 * x [M,K] pitch ldx; W [Nout,K] dense; y [M,Nout] pitch ldy.
 * `mode`: 0 = fp32 FFMA (0.8-1.7e-6 vs fp64), 1 = tcgen05 3xTF32 split (fp32-grade: 1.4-2.0e-6; 6.5e-7 in the
 *         -DISG_TC_NARROW build), 2 = tcgen05 single-pass TF32 (~8e-4)
 * Optional pre-split weight planes (mode 1 only; NULL keeps the in-kernel split; results are bit-identical):
 *   dgrad  `w_lo`  = w - tf32_trunc(w), [Nout,K] dense, from isg_split_lo;
 *   fwd    `w_t`, `w_t_lo` = the transposed weight [K,Nout] dense and its lo plane, from isg_transpose_split
 *          (both or neither).  The kernel then fetches the lo plane by TMA instead of splitting the weight tile in
 *          its per-k-block loop, and the forward product reads the weight as an MN-major operand (128-byte TMA rows
 *          instead of 64-byte ones).  One small kernel per weight and step instead of work on the GEMM's critical chain.
 * ------------------------------------------------------------------------------------- */
int isg_split_lo(const float* w, int64_t n /* multiple of 4 */, float* w_lo, void* stream);
int isg_transpose_split(const float* w /* [Nout,K] */, int Nout, int K, float* w_t /* [K,Nout] */,
                        float* w_t_lo /* [K,Nout] */, void* stream);
int isg_linear_fwd(const void* x, int64_t ldx, const void* w, const float* w_t /* or NULL */,
                   const float* w_t_lo /* or NULL */, const float* bias,
                   void* y, int64_t ldy, void* z_pre /* or NULL */, int64_t ldz,
                   int64_t M, int Nout, int K, int act, int mode, int dtype, void* stream);
int isg_linear_dgrad(const void* g_y, int64_t ldg, const void* w, const float* w_lo /* or NULL */,
                     const void* z_prev /* or NULL */, int64_t ldz,
                     void* g_x, int64_t ldgx, int accumulate,
                     int64_t M, int Nout, int K, int mode, int dtype, void* stream);
size_t isg_linear_wgrad_workspace_bytes(int64_t M, int Nout, int K);
int isg_linear_wgrad(const void* g_y, int64_t ldg, const void* x, int64_t ldx,
                     float* g_w /* [Nout,K] */, float* g_b /* reserved: use isg_colsum(g_y) */,
                     int64_t M, int Nout, int K, int mode, int dtype,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (d), bf16 configuration (BASELINE.json config 3 "fp32 vs bf16"): the same projections with bf16 operands on
 * tcgen05 kind::f16, fp32 accumulation — what torch.autocast(bfloat16) would give the reference's Linear layers
 * (models/mgat_v2_conv.py:177,181,259, models/mgat.py:156; the reference imports autocast and never enters it,
 * training/train_epoch.py:7).  Operands are bf16 with row pitches that are multiples of 8 elements (16 bytes;
 * a 300-wide operand uses pitch 304); outputs are bf16 (out_dtype = ISG_BF16) or fp32 (ISG_F32).
 *   fwd:   y = act(x W^T + b), optional pre-activation z (same dtype as y)      x [M,K], w [Nout,K]
 *   dgrad: g_x = (g_y W) [* gelu'(z_prev)] [+= when accumulate, fp32 only]      g_y [M,Nout], w_t = W^T [K,Nout]
 *   wgrad: g_W = g_y^T x, fp32, deterministic split over M                      g_y [M,Nout], x [M,K]
 * isg_weights_to_bf16 produces the bf16 copies W [rows, ld_w] and W^T [cols, ld_t] of up to 24 fp32 weights
 * [rows, cols] in one launch (pad columns zero); isg_to_bf16 converts an activation (pad columns zero). */
int isg_to_bf16(const float* in, int64_t ld_in, int64_t rows, int cols, void* out_bf16, int64_t ld_out, void* stream);
int isg_weights_to_bf16(int n, const float* const* w, const int* rows, const int* cols, void* const* w_bf16,
                        const int* ld_w, void* const* w_t_bf16, const int* ld_t, void* stream);
int isg_linear_bf16_fwd(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias,
                        void* y, int64_t ldy, void* z_pre /* or NULL */, int64_t ldz,
                        int64_t M, int Nout, int K, int act, int out_dtype, void* stream);
int isg_linear_bf16_dgrad(const void* g_y, int64_t ldg, const void* w_t, int64_t ldwt,
                          const void* z_prev /* or NULL, dtype of g_x */, int64_t ldz,
                          void* g_x, int64_t ldgx, int accumulate,
                          int64_t M, int Nout, int K, int out_dtype, void* stream);
size_t isg_linear_bf16_wgrad_workspace_bytes(int64_t M, int Nout, int K);
int isg_linear_bf16_wgrad(const void* g_y, int64_t ldg, const void* x, int64_t ldx, float* g_w /* [Nout,K] */,
                          int64_t M, int Nout, int K, void* workspace, size_t workspace_bytes, void* stream);

/* g = g_y * gelu'(z)  elementwise (backward of a GELU that closes a projection). */
int isg_gelu_bwd(const float* g_y, const float* z, float* g_z, int64_t n, void* stream);

/* out[c] = sum_r in[r, c]  (deterministic two-stage column sum; rows x cols fp32, pitch ld). */
size_t isg_colsum_workspace_bytes(int64_t rows, int cols);
int isg_colsum(const float* in, int64_t ld, int64_t rows, int cols, float* out,
               void* workspace, size_t workspace_bytes, void* stream);
/* n (<= 12) independent column sums in two launches: out[i][c] = sum_r in[i][r, c].  Bit-identical to n isg_colsum
 * calls (same slabs, same fold order).  The arrays are host arrays; dtypes[i] = ISG_F32 / ISG_BF16 is the storage
 * type of in[i] (NULL = all fp32; sums are always fp32).  Used by isg_mgat_layer_bwd for the bias /
 * GraphNorm-affine gradients of one layer (models/mgat_v2_conv.py:63-103 biases, models/mgat.py:79-89,163). */
size_t isg_colsum_multi_workspace_bytes(int n, const int64_t* rows, const int* cols);
int isg_colsum_multi(int n, const void* const* in, const int* dtypes, const int64_t* ld, const int64_t* rows,
                     const int* cols, float* const* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * SURVEY.md section 8 row f2 — scene-graph encoding layer in front of MGAT: torch_geometric MetaLayer(EdgeModel,
 * NodeModel) (models/scene_graph_encoder.py:107-146) and the GraphNorm SceneGraphEncoder.forward evaluates in
 * float64 on the CPU (:99-102).  The [300, 900] / [300, 600] weights of the concatenated inputs are split by
 * column block; the node blocks are applied per NODE by isg_linear_fwd and gathered here.
 * ------------------------------------------------------------------------------------- */

/* z[e] = a[src[e]] + b[dst[e]] + q[e]  (a, b: [N,D] or NULL; q [E,D]; edge_index [2,E] int64);  y = act(z).
 * z_pre [E,D] or NULL receives z (for the GELU backward). */
int isg_gather_add_act_fwd(const float* a, const float* b, const float* q, const int64_t* edge_index,
                           int64_t num_edges, int dim, int act, float* z_pre, float* y, void* stream);

/* out[n] = sum over p in [ptr[n], ptr[n+1]) of in[eid[p]]  (x 1/max(deg,1) if mean != 0): torch_scatter.scatter_mean
 * by destination (scene_graph_encoder.py:141) over the dst-sorted CSR, and the transposes of the two gathers above
 * (over the dst- / src-sorted CSR) in the backward.  Deterministic (fixed CSR order). */
int isg_segment_sum(const float* in /* [E,D] */, const int32_t* ptr, const int32_t* eid, int64_t num_nodes, int dim,
                    int mean, float* out /* [N,D] */, void* stream);

/* out[e] = in[idx[e]] (x 1/max(deg(idx[e]),1) if ptr != NULL): backward of the segment mean. */
int isg_gather_rows(const float* in /* [N,D] */, const int64_t* idx /* [E] */, const int32_t* ptr /* [N+1] or NULL */,
                    int64_t num_edges, int dim, float* out /* [E,D] */, void* stream);

/* GraphNorm (eps 1e-5) with float64 statistics and arithmetic, result rounded to float once — what
 * scene_graph_encoder.py:99-102 computes through a CPU DoubleTensor round trip.  Saves mean, rstd [B,D] double.
 * bwd: g_x [N,D] and per-graph partials of g_weight / g_bias / g_mean_scale [B,D] (summed by isg_colsum). */
int isg_graphnorm64_fwd(const float* x, const float* weight, const float* bias, const float* mean_scale,
                        const int32_t* graph_ptr, int64_t num_graphs, int dim, double eps,
                        float* y, double* mean, double* rstd, void* stream);
int isg_graphnorm64_bwd(const float* g_y, const float* x, const float* weight, const float* mean_scale,
                        const double* mean, const double* rstd, const int32_t* graph_ptr, int64_t num_graphs, int dim,
                        float* g_x, float* gw_part, float* gb_part, float* gms_part, void* stream);

/* ---------------------------------------------------------------------------------------
 * SURVEY.md section 8 row f4 — optimizer tail of the training step (training/train_epoch.py:111-118:
 * GradScaler.unscale_ + clip_grad_norm_(max_norm 2.0) + GradScaler.step(torch.optim.Adam) + update) without a
 * host synchronisation.  Tensors are fp32, passed as arrays of <= ISG_OPT_MAX_TENSORS device pointers per call.
 *   1. isg_grad_sq_partials  (per group)  partial[off + b] = sum over block b's 4096-element chunk of
 *                                         (g * inv_scale)^2; *nonfinite |= any inf/nan      (isg_opt_blocks = #b)
 *   2. isg_clip_finalize     (once)       state[0] = total norm, state[1] = min(1, max_norm / (norm + 1e-6)),
 *                                         state[2] = found_inf, state[3] += 1 unless found_inf (step counter);
 *                                         found_inf_out (or NULL) receives state[2] (GradScaler's per-device flag)
 *   3. isg_adam_update       (per group)  torch.optim.Adam step on g * inv_scale * state[1]; no-op if found_inf.
 * inv_scale / lr_dev: device scalars or NULL (1.0 / the host `lr`).  `nonfinite` must be zeroed by the caller.
 * ------------------------------------------------------------------------------------- */
int isg_opt_max_tensors(void);
int64_t isg_opt_blocks(const int64_t* numel, int count);
int isg_grad_sq_partials(const void* const* grads, const int64_t* numel, int count, const float* inv_scale,
                         double* partial, int partial_off, int32_t* nonfinite, void* stream);
int isg_clip_finalize(const double* partial, int nparts, const int32_t* nonfinite, float max_norm,
                      float* state /* [4] */, float* found_inf_out, void* stream);
int isg_adam_update(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                    const int64_t* numel, int count, const float* inv_scale, const float* state,
                    const float* lr_dev, float lr, float beta1, float beta2, float eps, void* stream);

/* ---------------------------------------------------------------------------------------
 * Layer executor: one call runs every kernel of one MGAT layer — the body of the loop in MGAT.forward
 * (models/mgat.py:131-177: MaskingGATv2Conv.forward models/mgat_v2_conv.py:138-241 incl. MaskingModel.forward
 * models/masking.py:132-199, x_proj, scatter-SDPA, GraphNorm, residual) — and one more its backward, in the
 * order of the entry points above.  Arguments are three flat arrays indexed by NAMED slots (csrc/executor.cu
 * lists them: D_* dims int64, F_* scalars double, P_* device pointers); isg_layer_slot(name) returns a slot's
 * index (-1 if unknown) and isg_layer_slot_count(0|1|2) the length of the dims | scalars | ptrs array.
 * Activations are written to caller-provided buffers (P_XG ... P_EMASK) by the forward and read back by the
 * backward; the backward's temporaries live in P_WS (isg_mgat_layer_bwd_workspace_bytes(dims) bytes, passed in
 * D_WS_BYTES).  Unused pointer slots are NULL.
 * ------------------------------------------------------------------------------------- */
int isg_layer_slot(const char* name);
int isg_layer_slot_count(int which);
size_t isg_mgat_layer_bwd_workspace_bytes(const int64_t* dims);
int isg_mgat_layer_fwd(const int64_t* dims, const double* scalars, void* const* ptrs, void* stream);
int isg_mgat_layer_bwd(const int64_t* dims, const double* scalars, void* const* ptrs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISG_H_ */

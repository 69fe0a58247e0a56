// csr.cu — kernel (a): destination-sorted and source-sorted CSR rebuild of the COO edge_index,
// plus graph_ptr from the sorted batch vector.  Integer-only, HBM/latency-bound work.
//
// Definition (oracle/isg_oracle.py::csr_build): stable sort of edge ids by key.  Implemented as
//   histogram (int atomics) -> exclusive scan -> unordered placement (int atomic cursor)
//   -> per-segment ascending sort of the edge ids.
// Edge ids are unique, so sorting each segment ascending yields exactly the stable order and
// the result is deterministic although the placement itself is not.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void csr_hist_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                int* __restrict__ deg_dst, int* __restrict__ deg_src,
                                int* __restrict__ status) {
  int bad = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) {
      ++bad;
      continue;
    }
    atomicAdd(&deg_dst[d], 1);
    atomicAdd(&deg_src[s], 1);
  }
  if (bad) atomicAdd(status, bad);
}

// grid = (tiles, 2): y selects the array.  Phase 1: per-tile totals.
__global__ void scan_tile_sums_kernel(const int* __restrict__ a0, const int* __restrict__ a1,
                                      int64_t N, int* __restrict__ tile_sums, int tiles) {
  __shared__ float red_unused[1];
  (void)red_unused;
  __shared__ int sred[32];
  const int* a = blockIdx.y ? a1 : a0;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
  for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS) {
    const int64_t idx = base + i;
    if (idx < N) s += a[idx];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(ISG_FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int r = threadIdx.x < (SCAN_THREADS / 32) ? sred[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(ISG_FULL_MASK, r, o);
    if (threadIdx.x == 0) tile_sums[blockIdx.y * tiles + blockIdx.x] = r;
  }
}

// Phase 2: in-place exclusive scan of a[0..N) -> a[0..N], a[N] = total.
__global__ void scan_apply_kernel(int* __restrict__ a0, int* __restrict__ a1, int64_t N,
                                  const int* __restrict__ tile_sums, int tiles) {
  __shared__ int sred[32];
  __shared__ int s_tile_off;
  __shared__ int s_warp_off[SCAN_THREADS / 32];
  int* a = blockIdx.y ? a1 : a0;
  const int* ts = tile_sums + blockIdx.y * tiles;
  // offset of this tile = sum of previous tile totals
  int part = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += SCAN_THREADS) part += ts[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(ISG_FULL_MASK, part, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int r = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) r += sred[w];
    s_tile_off = r;
  }
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t idx = base + i;
    v[i] = (idx < N) ? a[idx] : 0;
    tsum += v[i];
  }
  // exclusive scan of per-thread sums across the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(ISG_FULL_MASK, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_off[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
      const int t = s_warp_off[w];
      s_warp_off[w] = run;
      run += t;
    }
  }
  __syncthreads();
  int off = s_tile_off + s_warp_off[warp] + (incl - tsum);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t idx = base + i;
    if (idx < N) a[idx] = off;
    off += v[i];
    if (idx == N - 1) a[N] = off;
  }
  if (N == 0 && blockIdx.x == 0 && threadIdx.x == 0) a[0] = 0;
}

__global__ void csr_place_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                 const int* __restrict__ dst_ptr, const int* __restrict__ src_ptr,
                                 int* __restrict__ cur_dst, int* __restrict__ cur_src,
                                 int* __restrict__ dst_eid, int* __restrict__ src_eid) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) continue;
    dst_eid[dst_ptr[d] + atomicAdd(&cur_dst[d], 1)] = (int)e;
    src_eid[src_ptr[s] + atomicAdd(&cur_src[s], 1)] = (int)e;
  }
}

// One WARP per segment: ascending sort of the edge ids the atomics of csr_place_kernel left in arbitrary order
// (=> the stable order of the reference's COO), then emit the neighbour.  Segments of <= 32 edges (every node of
// a GQA graph at ~8 edges per node) are ranked in registers: lane i holds one id and counts the smaller ones with
// 32 shuffles.  Longer segments (hub nodes, the 200-object sweep points) run a normalised bitonic network in
// place — every compare-exchange ascending, so the virtual +inf padding beyond the segment end never moves —
// O(d log^2 d / 32) per warp instead of the O(d^2) single-thread insertion sort this replaces.
__global__ void csr_sort_segments_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                         const int* __restrict__ dst_ptr, const int* __restrict__ src_ptr,
                                         int* __restrict__ dst_eid, int* __restrict__ src_eid,
                                         int* __restrict__ dst_nbr, int* __restrict__ src_nbr) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (t >= 2 * N) return;
  const bool by_src = t >= N;
  const int64_t node = by_src ? t - N : t;
  const int* ptr = by_src ? src_ptr : dst_ptr;
  int* eid = by_src ? src_eid : dst_eid;
  int* nbr = by_src ? src_nbr : dst_nbr;
  const int64_t* other = by_src ? ei + E : ei;  // src ordering stores dst, dst ordering stores src
  const int beg = ptr[node], deg = ptr[node + 1] - beg;
  if (deg <= 0) return;
  int* a = eid + beg;
  if (deg <= 32) {
    const int key = lane < deg ? a[lane] : 0x7fffffff;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) rank += (__shfl_sync(ISG_FULL_MASK, key, j) < key) ? 1 : 0;  // ids are distinct
    __syncwarp();
    if (lane < deg) {
      a[rank] = key;
      nbr[beg + rank] = (int)other[key];
    }
    return;
  }
  int n2 = 64;
  while (n2 < deg) n2 <<= 1;
  auto cmpswap = [&](int i, int partner) {
    if (partner > i && partner < deg) {
      const int x = a[i], y = a[partner];
      if (x > y) {
        a[i] = y;
        a[partner] = x;
      }
    }
  };
  for (int k = 2; k <= n2; k <<= 1) {
    for (int i = lane; i < deg; i += 32) cmpswap(i, i ^ (k - 1));
    __syncwarp();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = lane; i < deg; i += 32) cmpswap(i, i ^ j);
      __syncwarp();
    }
  }
  for (int i = lane; i < deg; i += 32) nbr[beg + i] = (int)other[a[i]];
}

__global__ void graph_ptr_kernel(const int64_t* __restrict__ batch, int64_t N, int64_t B,
                                 int* __restrict__ graph_ptr, int* __restrict__ batch32,
                                 int* __restrict__ nmax) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n == 0) *nmax = 0;
  if (n > N) return;
  int64_t prev = (n > 0) ? batch[n - 1] : -1;
  int64_t cur = (n < N) ? batch[n] : B;
  if (n < N) batch32[n] = (int)cur;
  if (prev < -1) prev = -1;
  if (cur > B) cur = B;
  for (int64_t b = prev + 1; b <= cur; ++b) graph_ptr[b] = (int)n;
}

__global__ void graph_nmax_kernel(const int* __restrict__ graph_ptr, int64_t B, int* __restrict__ nmax) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int v = (b < B) ? graph_ptr[b + 1] - graph_ptr[b] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(ISG_FULL_MASK, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(nmax, v);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t isg_csr_workspace_bytes(int64_t N, int64_t E) {
  (void)E;
  if (N < 0) N = 0;  // (isg_csr_build rejects it; no size wraps around here)
  const int tiles = isg::ceil_div(N > 0 ? N : 1, SCAN_TILE);
  return align256((size_t)2 * (size_t)N * sizeof(int)) + align256((size_t)2 * tiles * sizeof(int));
}

extern "C" int isg_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* dst_ptr,
                             int32_t* dst_nbr, int32_t* dst_eid, int32_t* src_ptr, int32_t* src_nbr,
                             int32_t* src_eid, int32_t* status, void* workspace, size_t ws_bytes,
                             void* stream_) {
  if (E < 0 || N < 0 || E >= (int64_t)INT32_MAX || N >= (int64_t)INT32_MAX) return ISG_EINVAL;
  if (!dst_ptr || !src_ptr || !status || (E > 0 && (!edge_index || !dst_nbr || !dst_eid || !src_nbr || !src_eid)))
    return ISG_EINVAL;
  if (ws_bytes < isg_csr_workspace_bytes(N, E) || (!workspace && N > 0)) return ISG_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int tiles = isg::ceil_div(N > 0 ? N : 1, SCAN_TILE);
  int* cur_dst = (int*)workspace;
  int* cur_src = cur_dst + N;
  int* tile_sums = (int*)((char*)workspace + align256((size_t)2 * (size_t)N * sizeof(int)));

  cudaError_t err;
  if ((err = cudaMemsetAsync(dst_ptr, 0, (size_t)(N + 1) * sizeof(int), stream)) != cudaSuccess) return (int)err;
  if ((err = cudaMemsetAsync(src_ptr, 0, (size_t)(N + 1) * sizeof(int), stream)) != cudaSuccess) return (int)err;
  if ((err = cudaMemsetAsync(status, 0, sizeof(int), stream)) != cudaSuccess) return (int)err;
  if (N > 0 && (err = cudaMemsetAsync(workspace, 0, (size_t)2 * N * sizeof(int), stream)) != cudaSuccess)
    return (int)err;
  if (N == 0) return ISG_OK;

  const int eblocks = (int)min((int64_t)ISG_NUM_SMS * 8, (int64_t)isg::ceil_div(E > 0 ? E : 1, 256));
  if (E > 0) {
    csr_hist_kernel<<<eblocks, 256, 0, stream>>>(edge_index, E, N, dst_ptr, src_ptr, status);
    ISG_CHECK_LAUNCH();
  }
  scan_tile_sums_kernel<<<dim3(tiles, 2), SCAN_THREADS, 0, stream>>>(dst_ptr, src_ptr, N, tile_sums, tiles);
  ISG_CHECK_LAUNCH();
  scan_apply_kernel<<<dim3(tiles, 2), SCAN_THREADS, 0, stream>>>(dst_ptr, src_ptr, N, tile_sums, tiles);
  ISG_CHECK_LAUNCH();
  if (E > 0) {
    csr_place_kernel<<<eblocks, 256, 0, stream>>>(edge_index, E, N, dst_ptr, src_ptr, cur_dst, cur_src,
                                                  dst_eid, src_eid);
    ISG_CHECK_LAUNCH();
    csr_sort_segments_kernel<<<isg::ceil_div(2 * N * 32, 256), 256, 0, stream>>>(
        edge_index, E, N, dst_ptr, src_ptr, dst_eid, src_eid, dst_nbr, src_nbr);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

// crossing += #edges whose endpoints lie in different graphs (or out of range).  `crossing` must be zeroed by the
// caller's memset below.
__global__ void graph_closure_kernel(const int64_t* __restrict__ ei, int64_t E, const int64_t* __restrict__ batch,
                                     int64_t N, int* __restrict__ crossing) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool bad = false;
  if (e < E) {
    const int64_t s = ei[e], d = ei[E + e];
    bad = s < 0 || s >= N || d < 0 || d >= N || batch[s] != batch[d];
  }
  const unsigned m = __ballot_sync(ISG_FULL_MASK, bad);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(crossing, __popc(m));
}

extern "C" int isg_graph_closure(const int64_t* edge_index, int64_t E, const int64_t* batch, int64_t N,
                                 int32_t* crossing, void* stream_) {
  if (E < 0 || N < 0 || !crossing || (E > 0 && (!edge_index || !batch))) return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t err = cudaMemsetAsync(crossing, 0, sizeof(int), stream);
  if (err != cudaSuccess) return (int)err;
  if (E == 0) return ISG_OK;
  graph_closure_kernel<<<isg::ceil_div(E, 256), 256, 0, stream>>>(edge_index, E, batch, N, crossing);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

extern "C" int isg_graph_ptr(const int64_t* batch, int64_t N, int64_t B, int32_t* graph_ptr,
                             int32_t* batch32, int32_t* nmax, void* stream_) {
  if (N < 0 || B < 0 || !graph_ptr || !nmax || (N > 0 && (!batch || !batch32))) return ISG_EINVAL;
  if (N >= (int64_t)INT32_MAX) return ISG_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  graph_ptr_kernel<<<isg::ceil_div(N + 1, 256), 256, 0, stream>>>(batch, N, B, graph_ptr, batch32, nmax);
  ISG_CHECK_LAUNCH();
  if (B > 0) {
    graph_nmax_kernel<<<isg::ceil_div(B, 256), 256, 0, stream>>>(graph_ptr, B, nmax);
    ISG_CHECK_LAUNCH();
  }
  return ISG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Task order for the edge kernels: HEAVY NODES FIRST, everything else in its natural order.  The edge kernels run one
// warp per (node, head) and the hardware dispatches CTAs in index order; in the natural node order a 40-edge node
// that happens to sit near the end of the batch finishes ~15 us after everything else at the c3 size (4.7 waves of
// CTAs: a simulated makespan 1.47x the balanced one).  A full sort by degree fixes the tail but scatters the
// gathers: the natural order walks a graph's nodes together and re-uses their x_l rows in L1 / L2, and at batch 4096
// (x_l|x_r = 390 MB, beyond L2) the fully sorted schedule cost the forward kernel 22 % (0.78 -> 0.61 of the HBM peak).
// So only the nodes whose segment is at least twice the mean length (~7 % of them) are moved to the front — a STABLE
// partition, deterministic — which bounds the tail by a mean-sized task and keeps the locality of the other 93 %.
// The order only decides WHEN a task runs, never what it computes or where it writes.
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr int ORD_THREADS = 256;
// heavy(n) = deg(n) >= thr, thr = max(2, 2 * E / N) computed by the host from sizes it already knows
__global__ void __launch_bounds__(ORD_THREADS)
order_count_kernel(const int* __restrict__ ptr_a, const int* __restrict__ ptr_b, int64_t N, int thr,
                   int* __restrict__ blk /* [2][nblocks] */, int nblocks) {
  __shared__ int ca, cb;
  if (threadIdx.x == 0) ca = cb = 0;
  __syncthreads();
  const int64_t n = blockIdx.x * (int64_t)ORD_THREADS + threadIdx.x;
  const bool ha = n < N && ptr_a[n + 1] - ptr_a[n] >= thr, hb = n < N && ptr_b[n + 1] - ptr_b[n] >= thr;
  const unsigned ma = __ballot_sync(ISG_FULL_MASK, ha), mb = __ballot_sync(ISG_FULL_MASK, hb);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ca, __popc(ma));  // integer counts: the order of the adds does not matter
    atomicAdd(&cb, __popc(mb));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    blk[blockIdx.x] = ca;
    blk[nblocks + blockIdx.x] = cb;
  }
}
// exclusive scan of the per-block heavy counts (one block per ordering); blk[which][nblocks] receives the total
__global__ void __launch_bounds__(1024) order_scan_kernel(int* __restrict__ blk, int nblocks) {
  int* b = blk + (int64_t)blockIdx.x * (nblocks + 1);
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? b[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(ISG_FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(ISG_FULL_MASK, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int before = carry + (warp ? warp_tot[warp - 1] : 0) + incl - v;
    if (i < nblocks) b[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) b[nblocks] = carry;
}
__global__ void __launch_bounds__(ORD_THREADS)
order_place_kernel(const int* __restrict__ ptr_a, const int* __restrict__ ptr_b, int64_t N, int thr,
                   const int* __restrict__ blk, int nblocks, int* __restrict__ order_a, int* __restrict__ order_b,
                   int reverse_b) {
  __shared__ int wcount[2][ORD_THREADS / 32];
  const int64_t n = blockIdx.x * (int64_t)ORD_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool ha = n < N && ptr_a[n + 1] - ptr_a[n] >= thr, hb = n < N && ptr_b[n + 1] - ptr_b[n] >= thr;
  const unsigned ma = __ballot_sync(ISG_FULL_MASK, ha), mb = __ballot_sync(ISG_FULL_MASK, hb);
  if (lane == 0) {
    wcount[0][warp] = __popc(ma);
    wcount[1][warp] = __popc(mb);
  }
  __syncthreads();
  if (n >= N) return;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const unsigned m = which ? mb : ma;
    const bool heavy = which ? hb : ha;
    int before = __popc(m & lt);  // heavy nodes of this block that precede n
    for (int w = 0; w < warp; ++w) before += wcount[which][w];
    const int* b = blk + (int64_t)which * (nblocks + 1);
    const int heavy_before = b[blockIdx.x] + before, total_heavy = b[nblocks];
    int64_t pos = heavy ? heavy_before : total_heavy + (n - heavy_before);
    // ordering b (the src pass of the edge backward) walks the light nodes in DESCENDING order: it starts with the
    // nodes whose g_eproj rows the dst pass wrote last, which are the ones still in L2
    if (which && reverse_b && !heavy) pos = total_heavy + (N - 1 - pos);
    (which ? order_b : order_a)[pos] = (int)n;
  }
}
}  // namespace

extern "C" size_t isg_degree_order_workspace_bytes(int64_t N) {
  const int64_t nblocks = ((N > 0 ? N : 1) + ORD_THREADS - 1) / ORD_THREADS;
  return (size_t)2 * (size_t)(nblocks + 1) * sizeof(int);
}

extern "C" int isg_degree_order(const int32_t* dst_ptr, const int32_t* src_ptr, int64_t N, int64_t E, int32_t* dst_order,
                                int32_t* src_order, void* workspace, size_t ws_bytes, void* stream_) {
  if (N < 0 || E < 0 || N >= (int64_t)INT32_MAX) return ISG_EINVAL;
  if (N == 0) return ISG_OK;
  if (!dst_ptr || !src_ptr || !dst_order || !src_order) return ISG_EINVAL;
  if (ws_bytes < isg_degree_order_workspace_bytes(N) || !workspace) return ISG_EWORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nblocks = (int)((N + ORD_THREADS - 1) / ORD_THREADS);
  int thr = (int)((2 * E + N - 1) / N);
  if (thr < 2) thr = 2;
  // layout: [heavy counts of dst ordering: nblocks + 1][of src ordering: nblocks + 1] — the count kernel writes
  // with stride nblocks, so it gets a compact staging view and the scan kernel the padded one
  int* blk = (int*)workspace;
  order_count_kernel<<<nblocks, ORD_THREADS, 0, stream>>>(dst_ptr, src_ptr, N, thr, blk, nblocks + 1);
  ISG_CHECK_LAUNCH();
  order_scan_kernel<<<2, 1024, 0, stream>>>(blk, nblocks);
  ISG_CHECK_LAUNCH();
  static const int reverse_src = getenv("ISG_SRC_ORDER_REV") ? atoi(getenv("ISG_SRC_ORDER_REV")) : 1;
  order_place_kernel<<<nblocks, ORD_THREADS, 0, stream>>>(dst_ptr, src_ptr, N, thr, blk, nblocks, dst_order, src_order,
                                                          reverse_src);
  ISG_CHECK_LAUNCH();
  return ISG_OK;
}

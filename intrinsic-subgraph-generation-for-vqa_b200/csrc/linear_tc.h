// linear_tc.h — internal interface between linear.cu (C-ABI entry points, mode dispatch) and
// linear_tc.cu (tcgen05 / TMEM / TMA GEMM).  Not part of the public ABI.
#pragma once
#include <stdint.h>

namespace isg {

struct TcGemm {
  const float* A;  // a_mn == 0: [rows, R] R-contiguous;  a_mn == 1: [R, rows] rows-contiguous
  int64_t lda;
  const float* B;  // b_mn == 0: [cols, R] R-contiguous;  b_mn == 1: [R, cols] cols-contiguous
  int64_t ldb;
  const float* B_lo;  // optional (mode 1): B - tf32_trunc(B), same layout and pitch as B, fetched by TMA instead of
                      // being produced by the splitter warps (isg_split_lo); NULL = split B inside the kernel
  float* C;        // [rows, cols] (+ split * c_split_stride)
  int64_t ldc;
  int64_t rows;
  int cols;
  int64_t R;
  int a_mn, b_mn;
  int epi;         // 0 fwd (bias / Z / act), 1 dgrad (Zprev / accumulate), 2 plain
  int splits;      // split of the reduction range (wgrad); 1 otherwise
  int64_t r_chunk; // reduction elements per split, multiple of 32; == R rounded up when splits == 1
  int64_t c_split_stride;
  const float* bias;
  float* Z;
  int64_t ldz;
  const float* Zprev;
  int act;
  int accumulate;
  int no_cluster;  // 1: launch single CTAs (no B-tile multicast pairs); debugging / A-B comparison
  int split3;      // 1 = 3xTF32 (fp32-grade), 0 = single-pass TF32
};

int tc_gemm(const TcGemm& p, void* stream);
// number of reduction splits (and the chunk length) the tensor-core wgrad uses for [M] x [Nout, K]
int tc_wgrad_splits(int64_t M, int Nout, int K, int64_t* r_chunk);

}  // namespace isg

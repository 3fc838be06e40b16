// gemm.cuh — parameter block shared by the dense-product kernels (gemm_simt.cu, gemm_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gps {

// C[m,n] (+)= epilogue( sum_k Aop[m,k] * Bop[k,n] )
//   ta == 0: Aop[m,k] = A[m*lda + k]      ta == 1: Aop[m,k] = A[k*lda + m]
//   tb == 0: Bop[k,n] = B[n*ldb + k]      tb == 1: Bop[k,n] = B[k*ldb + n]
// Epilogue order: +bias[n] -> (store pre-activation) -> act -> *act'(mask_src) -> dropout ->
//                 +R1 +R2 -> store -> column statistics (sum, sum of squares, double atomics).
// With splitk > 1 the partial products are atomically added into a pre-zeroed C and only the
// plain product is supported.
struct GemmParams {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int lda = 0; int ta = 0;
  const float* B = nullptr; int ldb = 0; int tb = 0;
  float* C = nullptr; int ldc = 0;
  const float* bias = nullptr;
  float* C_pre = nullptr; int ldpre = 0;   // optional copy of the pre-activation value
  int act = -1;                            // -1 none, GPS_ACT_*
  const float* mask_src = nullptr; int ldmask = 0; int mask_act = -1;  // multiply by act'(mask_src)
  int mask_is_post = 0;                    // relu only: mask_src holds the post-activation value
  float p_drop = 0.f; uint64_t seed = 0, offset = 0; int site = 0;    // dropout on the result
  const unsigned long long* offset_dev = nullptr;                     // optional device-resident addend to offset
  float p_drop2 = 0.f; int site2 = 0;      // optional inner dropout applied before the one above (Performer: attn_dropout
                                           // on to_out(O) inside SelfAttention, then GPSLayer.dropout_attn)
  const float* R1 = nullptr; int ldr1 = 0;
  const float* R2 = nullptr; int ldr2 = 0;
  double* stats = nullptr;                 // [2][N] column sum / sum of squares of the stored C
  int splitk = 1;
  float* colsum_a = nullptr;               // ta==1 only: += sum_k Aop[m,k]  (bias gradient), [M]
  int precision = GPS_PREC_FP32;
  // optional pre-packed image of B (tb == 0 only; see prepack_weights): bf16 hi plane at bpk, lo plane at
  // bpk + bpk_lo_off, bpk_groups 8-row groups per 64-wide k-block, this GEMM's B starts at packed row bpk_row0
  // bpk_mn: the image is MN-major (tb == 1: B is [K x N]; bpk_groups = 64-column blocks per k-block)
  const void* bpk = nullptr; int64_t bpk_lo_off = 0; int bpk_groups = 0; int bpk_row0 = 0; int bpk_mn = 0;
  int bpk_kb0 = 0;   // first 64-deep k-block of this GEMM inside the packed planes (reduction sub-range of a packed matrix)
  // bf16 hi/lo planes of the operands as stored (A: [M,K] or, ta == 1, [K,M]; B: [N,K] or, tb == 1, [K,N]): when both
  // are given the TMA-fed kernel (gemm_tma.cu) runs and A/B (fp32) are not read.  Cp: optional plane copy of the
  // result for the next GEMM; C may then be null (planes-only output).
  Planes Ap, Bp, Cp;
  // Cp column remap for the attention operands: output column c lands at (c / cp_hd) * cp_hd_pad + c % cp_hd and the
  // cp_hd_pad - cp_hd pad columns of every head are written as zeros (per-head layout padded to a multiple of 16)
  int cp_hd = 0, cp_hd_pad = 0, cp_col0 = 0;
  // Backward-reduce of up to two BatchNorms that consume this GEMM's output C as their upstream gradient (TMA kernel
  // only): with zhat_k = (z_k - mean_k) * invstd_k, sums_k[0][n] += sum_m C[m,n], sums_k[1][n] += sum_m C[m,n] * zhat_k[m,n]
  // (what bn_bwd_reduce(C, z_k, ...) computes; norm1_local / norm1_attn both read g_s, gps_layer.py:194,217,222)
  struct BnRed { const float* z = nullptr; int ldz = 0; const float* mean = nullptr; const float* invstd = nullptr; double* sums = nullptr; };
  BnRed bnred[2];   // cp_col0: first output column of the remapped block
};

struct ToPlanesItem { const float* src; int64_t ld; int rows; int cols; Planes dst; };
// fp32 [rows, cols] (pitch ld) -> bf16 hi/lo planes, up to 16 matrices per launch
int to_planes(const ToPlanesItem* items, int n, cudaStream_t stream);
// TMA-fed tcgen05 product on plane operands; GPS_ERR_UNSUPPORTED when the planes are missing / misaligned
int gemm_tma(const GemmParams& p, cudaStream_t stream);
// rank-3 tensor map {cols, rows, planes} over a plane pair with a {64, box_rows, 1} SWIZZLE_128B box (cached)
int make_tensor_map(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int planes, int64_t rows, int64_t cols, int64_t ld,
                    int box_rows, CUtensorMap* out);
void gemm_tma_set_force_bn(int bn);
void gemm_tma_set_trace(unsigned long long* buf);   // bring-up: per-CTA phase timestamps (tools/gemm_trace.py)

// Pre-packs up to 8 weight matrices (fp32 [rows, K] row-major) into the tcgen05 kernel's shared-memory tile image.
// K-major (mn = 0): W is [rows x K], dst sized by prepack_bytes(rows, K).
// MN-major (mn = 1): W is [K x rows] (rows = GEMM output columns), dst sized by prepack_bytes_mn(rows, K).
struct PrepackItem { const float* W; int rows; int K; int ld; void* dst; int mn; };
int64_t prepack_bytes(int rows, int K);          // both planes
int64_t prepack_plane_bytes(int rows, int K);    // offset of the lo plane
int prepack_groups(int rows);
int64_t prepack_bytes_mn(int cols, int K);
int64_t prepack_plane_bytes_mn(int cols, int K);
int prepack_groups_mn(int cols);
int prepack_weights(const PrepackItem* items, int n, cudaStream_t stream);

// exact fp32 CUDA-core product (validation path and shapes the tensor-core kernel does not take)
int gemm_simt(const GemmParams& p, cudaStream_t stream);
// tcgen05 tensor-core product; returns GPS_ERR_UNSUPPORTED for shapes it does not take
int gemm_tc(const GemmParams& p, cudaStream_t stream);
void gemm_tc_set_debug(int v);
// dispatcher used by the layer
int gemm(const GemmParams& p, cudaStream_t stream);

}  // namespace gps

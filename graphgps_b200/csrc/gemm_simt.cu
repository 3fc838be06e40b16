// gemm_simt.cu — exact-fp32 CUDA-core dense product with the layer's fused epilogues.
//
// Role: bring-up/validation path for every Linear of the hot path (A,B,C,D,E of GatedGCN
// gatedgcn_layer.py:57-61; in_proj/out_proj of nn.MultiheadAttention gps_layer.py:104-106; FFN
// gps_layer.py:253-257) and their gradients, and the fallback for shapes gemm_tc.cu rejects.
// 64x64x16 tile, 256 threads, 4x4 register micro-tile, float4 shared-memory reads.
#include "gemm.cuh"

namespace gps {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, NT = 256;

template <bool TRANS>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int ld, int rows_total,
                                          int k_end, int r0, int k0, float (*S)[BM + PAD], bool vec_ok) {
  // Fills S[k][r] = op(P)[r0 + r, k0 + k] for r < 64, k < 16 (zero outside rows_total / k_end).
  const int t = threadIdx.x;
  if (!TRANS) {  // element (r,k) at P[r*ld + k]: k contiguous
    const int r = t >> 2, kq = (t & 3) * 4;
    const int gr = r0 + r, gk = k0 + kq;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < rows_total) {
      const float* src = P + (int64_t)gr * ld + gk;
      if (vec_ok && gk + 3 < k_end) {
        float4 q = ld4(src);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (gk + i < k_end) v[i] = src[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) S[kq + i][r] = v[i];
  } else {  // element (r,k) at P[k*ld + r]: r contiguous
    const int k = t >> 4, rq = (t & 15) * 4;
    const int gk = k0 + k, gr = r0 + rq;
    float4 q = f4zero();
    if (gk < k_end) {
      const float* src = P + (int64_t)gk * ld + gr;
      if (vec_ok && gr + 3 < rows_total) {
        q = ld4(src);
      } else {
        if (gr + 0 < rows_total) q.x = src[0];
        if (gr + 1 < rows_total) q.y = src[1];
        if (gr + 2 < rows_total) q.z = src[2];
        if (gr + 3 < rows_total) q.w = src[3];
      }
    }
    *reinterpret_cast<float4*>(&S[k][rq]) = q;
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) k_gemm_simt(GemmParams p, int kchunk, bool vecA, bool vecB, bool vecC) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  __shared__ float red[2][16][BN];

  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * kchunk;
  const int ke = min(p.K, kb + kchunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;  // bias-gradient partial (threads < 64 own column m0+threadIdx.x of A^T)
  const bool do_colsum = p.colsum_a != nullptr && blockIdx.x == 0;

  for (int k0 = kb; k0 < ke; k0 += BK) {
    load_tile<TA>(p.A, p.lda, p.M, ke, m0, k0, As, vecA);
    load_tile<TB>(p.B, p.ldb, p.N, ke, n0, k0, Bs, vecB);
    __syncthreads();
    if (do_colsum && threadIdx.x < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) csum += As[k][threadIdx.x];
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (do_colsum && threadIdx.x < BM && m0 + (int)threadIdx.x < p.M)
    atomicAdd(&p.colsum_a[m0 + threadIdx.x], csum);

  const int gn = n0 + tx * 4;
  if (p.splitk > 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int gm = m0 + ty * 4 + i;
      if (gm >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < p.N) {
          float w = acc[i][j];
          if (blockIdx.z == 0) {   // the first split also carries the residual terms
            if (p.R1) w += p.R1[(int64_t)gm * p.ldr1 + gn + j];
            if (p.R2) w += p.R2[(int64_t)gm * p.ldr2 + gn + j];
          }
          atomicAdd(&p.C[(int64_t)gm * p.ldc + gn + j], w);
        }
    }
    return;
  }

  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (gn + j < p.N) bias[j] = p.bias[gn + j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (p.C_pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < p.N) p.C_pre[(int64_t)gm * p.ldpre + gn + j] = v[j];
    }
    if (p.act >= 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = act_fwd_rt(p.act, v[j]);
    }
    if (p.mask_src) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < p.N) {
          float ms = p.mask_src[(int64_t)gm * p.ldmask + gn + j];
          v[j] *= p.mask_is_post ? (ms > 0.f ? 1.f : 0.f) : act_bwd_rt(p.mask_act, ms);
        }
    }
    if (p.p_drop2 > 0.f) {
      float4 sc = dropout_scale4(p.p_drop2, p.seed, p.offset + (p.offset_dev ? *p.offset_dev : 0ull), p.site2,
                                 ((uint64_t)gm * (uint64_t)p.N + gn) >> 2);
      v[0] *= sc.x; v[1] *= sc.y; v[2] *= sc.z; v[3] *= sc.w;
    }
    if (p.p_drop > 0.f) {
      // gn is a multiple of 4 and ldc-independent: flat index over a dense [M, N] grid
      float4 sc = dropout_scale4(p.p_drop, p.seed, p.offset + (p.offset_dev ? *p.offset_dev : 0ull), p.site,
                                 ((uint64_t)gm * (uint64_t)p.N + gn) >> 2);
      v[0] *= sc.x; v[1] *= sc.y; v[2] *= sc.z; v[3] *= sc.w;
    }
    if (p.R1) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < p.N) v[j] += p.R1[(int64_t)gm * p.ldr1 + gn + j];
    }
    if (p.R2) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < p.N) v[j] += p.R2[(int64_t)gm * p.ldr2 + gn + j];
    }
    float* dst = p.C + (int64_t)gm * p.ldc + gn;
    if (vecC && gn + 3 < p.N) {
      st4(dst, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < p.N) dst[j] = v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s1[j] += v[j];
      s2[j] += v[j] * v[j];
    }
  }
  if (p.stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[0][ty][tx * 4 + j] = s1[j];
      red[1][ty][tx * 4 + j] = s2[j];
    }
    __syncthreads();
    if (threadIdx.x < 2 * BN) {
      const int which = threadIdx.x / BN, c = threadIdx.x % BN;
      if (n0 + c < p.N) {
        double tot = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) tot += (double)red[which][r][c];
        atomic_add_f64(&p.stats[(int64_t)which * p.N + n0 + c], tot);
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int gemm_simt(const GemmParams& p, cudaStream_t stream) {
  GPS_REQUIRE(p.M >= 0 && p.N >= 0 && p.K >= 0 && p.A && p.B && p.C, GPS_ERR_ARG, "gemm: bad argument");
  if (p.M == 0 || p.N == 0) return GPS_OK;
  GPS_REQUIRE(p.splitk >= 1, GPS_ERR_ARG, "gemm: splitk < 1");
  if (p.splitk > 1)
    GPS_REQUIRE(!p.bias && p.act < 0 && !p.mask_src && !p.stats && !p.C_pre && p.p_drop == 0.f && p.p_drop2 == 0.f, GPS_ERR_ARG,
                "gemm: split-K supports the plain product (+ residuals) only");
  GPS_REQUIRE(p.colsum_a == nullptr || p.ta == 1, GPS_ERR_ARG, "gemm: colsum_a needs ta == 1");
  GPS_REQUIRE((p.p_drop == 0.f && p.p_drop2 == 0.f) || (p.N % 4 == 0), GPS_ERR_ARG, "gemm: dropout epilogue needs N %% 4 == 0");
  int splitk = p.splitk;
  int kchunk = (int)round_up(ceil_div(p.K > 0 ? p.K : 1, splitk), BK);
  splitk = (int)ceil_div(p.K > 0 ? p.K : 1, kchunk);
  GemmParams q = p;
  q.splitk = p.splitk > 1 ? 2 : 1;  // only "is split" matters inside the kernel
  const bool vecA = aligned16(p.A) && p.lda % 4 == 0;
  const bool vecB = aligned16(p.B) && p.ldb % 4 == 0;
  const bool vecC = aligned16(p.C) && p.ldc % 4 == 0;
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM), (unsigned)splitk);
  if (!p.ta && !p.tb) k_gemm_simt<false, false><<<grid, NT, 0, stream>>>(q, kchunk, vecA, vecB, vecC);
  else if (!p.ta && p.tb) k_gemm_simt<false, true><<<grid, NT, 0, stream>>>(q, kchunk, vecA, vecB, vecC);
  else if (p.ta && !p.tb) k_gemm_simt<true, false><<<grid, NT, 0, stream>>>(q, kchunk, vecA, vecB, vecC);
  else k_gemm_simt<true, true><<<grid, NT, 0, stream>>>(q, kchunk, vecA, vecB, vecC);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

}  // namespace gps

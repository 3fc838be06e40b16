// gemm_tma.cu — TMA-fed tcgen05 dense product: the round-2 operand path of every Linear of the GPS layer.
//
//   C[M,N] (+)= epi( Aop[M,K] * Bop[K,N] )
//
// Both operands live in HBM as bf16 "planes": plain row-major bf16 matrices holding the hi part of the fp32 value
// and (fp32-grade mode) its bf16 residual lo, written ONCE by the kernel that produced the tensor (GEMM epilogues,
// the row-wise BatchNorm kernels, the gather-reduce kernels, attention) or by k_to_planes for layer inputs and
// weights.  The consumer therefore never converts anything: one elected thread issues tensor-map TMA
// (cp.async.bulk.tensor, SASS UTMALDG) boxes that land in shared memory already in the canonical UMMA
// SWIZZLE_128B image -- a {64 x rows} box is a K-major tile, a {64 x 64} box is one MN-major block -- so the same
// planes serve y = x W^T (K-major), g_x = g_y W (B MN-major) and dW = G^T X (both MN-major, reduction over rows)
// without transposes or per-layout copies.  Out-of-range rows/columns are zero-filled by the TMA unit.
//
// Warp roles (320 threads): warps 0-7 epilogue (tcgen05.ld -> bias / act / act' / dropout / residuals / fp32 store
// / bf16 hi-lo plane store / BatchNorm column sums / split-K atomics), warp 8 MMA issuer (one lane,
// tcgen05.mma kind::f16 128 x BN x 16, fp32 accumulation in TMEM; fp32-grade mode issues lo*hi + hi*lo + hi*hi),
// warp 9 TMA producer (one lane).  Stages form an mbarrier ring: full = expect_tx bytes, empty = tcgen05.commit
// (+ one arrival per epilogue warp in CTAs that also reduce the bias gradient from the staged A tiles).
// Narrow tiles (BN <= 64) are launched two CTAs per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>
#include <unordered_map>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace gps {

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr int kThreads = 320;
constexpr int kATile = BM * BK * 2;     // 16 KB per plane
constexpr int kBlock = 64 * BK * 2;     // 8 KB: one 64-column MN-major block / 64 K-major rows

struct TmaArgs {
  GemmParams p;
  int BN, nb_blocks, stages, kb_per_split, tmem_cols, planes;
};

template <bool A_MN, bool B_MN, bool NARROW>
__global__ void __launch_bounds__(kThreads, NARROW ? 2 : 1)
k_gemm_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TmaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const GemmParams& p = a.p;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int planes = a.planes;
  const int b_tile = a.nb_blocks * kBlock;
  const int stage_bytes = planes * (kATile + b_tile);
  const int S = a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);   // full[S], empty[S], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);   // 16 x 16 x 8 floats (bias-gradient partials)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * a.BN;
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * a.kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + a.kb_per_split);
  const int nkb = kb_end - kb_begin;
  const bool do_colsum = A_MN && p.colsum_a != nullptr && blockIdx.x == 0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[S + s]), 1 + (do_colsum ? kEpiWarps : 0));
    }
    mbar_init(smem_u32(&bars[2 * S]), 1);
    fence_barrier_init();
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    // =========================================================== TMA producer
    if (lane == 0 && nkb > 0) {
      // blocks of an MN-major tile that lie completely outside the matrix are not fetched (their smem content only
      // reaches accumulator rows / columns that are never stored)
      int a_blocks = 2, b_blocks = a.nb_blocks;
      if (A_MN) a_blocks = min(2, (p.M - m0 + 63) / 64);
      if (B_MN) b_blocks = min(a.nb_blocks, (p.N - n0 + 63) / 64);
      const uint32_t tx = (uint32_t)planes * ((A_MN ? a_blocks * kBlock : kATile) +
                                              (B_MN ? b_blocks * kBlock : a.BN * 128));
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        if (i >= S) mbar_wait(smem_u32(&bars[S + s]), (uint32_t)((i / S) - 1) & 1u);
        const uint32_t full = smem_u32(&bars[s]);
        mbar_arrive_expect_tx(full, tx);
        const int k0 = (kb_begin + i) * BK;
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb = sa + planes * kATile;
        for (int pl = 0; pl < planes; ++pl) {
          if (!A_MN) {
            tma_tile_3d(sa + pl * kATile, &tmA, k0, m0, pl, full);
          } else {
            for (int b = 0; b < a_blocks; ++b) tma_tile_3d(sa + pl * kATile + b * kBlock, &tmA, m0 + 64 * b, k0, pl, full);
          }
          if (!B_MN) {
            tma_tile_3d(sb + pl * b_tile, &tmB, k0, n0, pl, full);
          } else {
            for (int b = 0; b < b_blocks; ++b) tma_tile_3d(sb + pl * b_tile + b * kBlock, &tmB, n0 + 64 * b, k0, pl, full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // =========================================================== MMA issuer
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc(BM, a.BN, A_MN, B_MN);
      const uint32_t a_lbo = A_MN ? kBlock : 16, b_lbo = B_MN ? kBlock : 16;
      const uint32_t a_kstep = A_MN ? 2048 : 32, b_kstep = B_MN ? 2048 : 32;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        mbar_wait(smem_u32(&bars[s]), (uint32_t)(i / S) & 1u);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb_hi = sa_hi + planes * kATile;
        const uint32_t sa_lo = sa_hi + kATile;
        const uint32_t sb_lo = sb_hi + b_tile;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t da_hi = make_desc(sa_hi + kk * a_kstep, a_lbo, 1024);
          const uint64_t db_hi = make_desc(sb_hi + kk * b_kstep, b_lbo, 1024);
          if (planes == 2) {
            const uint64_t da_lo = make_desc(sa_lo + kk * a_kstep, a_lbo, 1024);
            const uint64_t db_lo = make_desc(sb_lo + kk * b_kstep, b_lbo, 1024);
            umma_bf16(tmem_base, da_lo, db_hi, idesc, (i | kk) != 0);
            umma_bf16(tmem_base, da_hi, db_lo, idesc, 1u);
            umma_bf16(tmem_base, da_hi, db_hi, idesc, 1u);
          } else {
            umma_bf16(tmem_base, da_hi, db_hi, idesc, (i | kk) != 0);
          }
        }
        umma_commit(smem_u32(&bars[S + s]));   // frees the smem stage once these MMAs retire
      }
      umma_commit(smem_u32(&bars[2 * S]));     // accumulator complete
    }
    __syncwarp();
  } else {
    // =========================================================== epilogue warps
    // bias gradient db[m] = sum_k Aop[m,k]: the n-tile-0 CTAs sum the staged (MN-major) A tiles while the tensor
    // core works on them.  Thread t owns the 8-column chunk (t & 15) and k-rows 4 (t >> 4) .. +3 of every k-block.
    if (do_colsum) {
      float csum[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) csum[e] = 0.f;
      const int cm = tid & 15, kq = tid >> 4;
      const uint32_t blk_off = (uint32_t)(cm >> 3) * kBlock;
      const int cc = cm & 7;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        if (lane == 0) mbar_wait(smem_u32(&bars[s]), (uint32_t)(i / S) & 1u);
        __syncwarp();
        const uint8_t* sa = smem + (size_t)s * stage_bytes;
        for (int pl = 0; pl < planes; ++pl) {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int r = kq * 4 + rr;
            const uint4 q = *reinterpret_cast<const uint4*>(sa + pl * kATile + blk_off + (r >> 3) * 1024 + (r & 7) * 128 +
                                                            ((cc ^ (r & 7)) << 4));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              csum[2 * e] += __uint_as_float(w[e] << 16);
              csum[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[S + s]));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) red[(kq * 16 + cm) * 8 + e] = csum[e];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (do_colsum && tid < 128) {
      const int cm = tid >> 3, e = tid & 7;
      float tot = 0.f;
#pragma unroll
      for (int o = 0; o < 16; ++o) tot += red[(o * 16 + cm) * 8 + e];
      const int gm = m0 + cm * 8 + e;
      if (gm < p.M) atomicAdd(&p.colsum_a[gm], tot);
    }

    if (nkb > 0) {
      if (lane == 0) mbar_wait(smem_u32(&bars[2 * S]), 0u);
      __syncwarp();
      tc_fence_after();
    }
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const int nchunks = a.BN >> 4;
    const uint64_t drop_off = p.offset + ((p.p_drop > 0.f || p.p_drop2 > 0.f) && p.offset_dev ? *p.offset_dev : 0ull);
    for (int c = half; c < nchunks; c += 2) {
      const int gn = n0 + c * 16;
      if (gn >= p.N) break;
      float v[16];
      if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), v);
      else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0.f;
      }
      if (p.splitk > 1) {
        if (row_ok) {
          float* dst = p.C + (int64_t)row * p.ldc + gn;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {   // N % 4 == 0: whole 16-byte groups; red.global.add.v4.f32
            if (gn + e >= p.N) continue;
            float4 w4 = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            if (blockIdx.z == 0) {            // the first split also carries the residual terms
              if (p.R1) w4 = f4add(w4, ld4(p.R1 + (int64_t)row * p.ldr1 + gn + e));
              if (p.R2) w4 = f4add(w4, ld4(p.R2 + (int64_t)row * p.ldr2 + gn + e));
            }
            atomicAdd(reinterpret_cast<float4*>(dst + e), w4);
          }
        }
        continue;
      }
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const int col = gn + g4 * 4;
        const bool ok = row_ok && col < p.N;
        float* w = v + g4 * 4;
        if (p.bias && col < p.N) {
          float4 bb = ld4(p.bias + col);
          w[0] += bb.x; w[1] += bb.y; w[2] += bb.z; w[3] += bb.w;
        }
        if (ok && p.C_pre) st4(p.C_pre + (int64_t)row * p.ldpre + col, make_float4(w[0], w[1], w[2], w[3]));
        if (p.act >= 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = act_fwd_rt(p.act, w[e]);
        }
        if (ok && p.mask_src) {
          float4 ms = ld4(p.mask_src + (int64_t)row * p.ldmask + col);
          float mv[4] = {ms.x, ms.y, ms.z, ms.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] *= p.mask_is_post ? (mv[e] > 0.f ? 1.f : 0.f) : act_bwd_rt(p.mask_act, mv[e]);
        }
        if (ok && p.p_drop2 > 0.f) {
          float4 sc = dropout_scale4(p.p_drop2, p.seed, drop_off, p.site2, ((uint64_t)row * (uint64_t)p.N + col) >> 2);
          w[0] *= sc.x; w[1] *= sc.y; w[2] *= sc.z; w[3] *= sc.w;
        }
        if (ok && p.p_drop > 0.f) {
          float4 sc = dropout_scale4(p.p_drop, p.seed, drop_off, p.site, ((uint64_t)row * (uint64_t)p.N + col) >> 2);
          w[0] *= sc.x; w[1] *= sc.y; w[2] *= sc.z; w[3] *= sc.w;
        }
        if (ok && p.R1) {
          float4 r = ld4(p.R1 + (int64_t)row * p.ldr1 + col);
          w[0] += r.x; w[1] += r.y; w[2] += r.z; w[3] += r.w;
        }
        if (ok && p.R2) {
          float4 r = ld4(p.R2 + (int64_t)row * p.ldr2 + col);
          w[0] += r.x; w[1] += r.y; w[2] += r.z; w[3] += r.w;
        }
        if (ok && p.C) st4(p.C + (int64_t)row * p.ldc + col, make_float4(w[0], w[1], w[2], w[3]));
        if (ok && p.Cp.hi) planes_store4(p.Cp, row, col, make_float4(w[0], w[1], w[2], w[3]));
        if (!ok) { w[0] = w[1] = w[2] = w[3] = 0.f; }
      }
      if (p.stats) {
        // column sums over the warp's 32 rows: butterfly reduce-scatter, 16 columns x {sum, sumsq}
        float s1[16], s2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { s1[e] = v[e]; s2[e] = v[e] * v[e]; }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool up = (lane & 16) != 0;
          float send1 = up ? s1[e] : s1[e + 8], send2 = up ? s2[e] : s2[e + 8];
          float keep1 = up ? s1[e + 8] : s1[e], keep2 = up ? s2[e + 8] : s2[e];
          s1[e] = keep1 + __shfl_xor_sync(0xffffffffu, send1, 16);
          s2[e] = keep2 + __shfl_xor_sync(0xffffffffu, send2, 16);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool up = (lane & 8) != 0;
          float send1 = up ? s1[e] : s1[e + 4], send2 = up ? s2[e] : s2[e + 4];
          float keep1 = up ? s1[e + 4] : s1[e], keep2 = up ? s2[e + 4] : s2[e];
          s1[e] = keep1 + __shfl_xor_sync(0xffffffffu, send1, 8);
          s2[e] = keep2 + __shfl_xor_sync(0xffffffffu, send2, 8);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool up = (lane & 4) != 0;
          float send1 = up ? s1[e] : s1[e + 2], send2 = up ? s2[e] : s2[e + 2];
          float keep1 = up ? s1[e + 2] : s1[e], keep2 = up ? s2[e + 2] : s2[e];
          s1[e] = keep1 + __shfl_xor_sync(0xffffffffu, send1, 4);
          s2[e] = keep2 + __shfl_xor_sync(0xffffffffu, send2, 4);
        }
        {
          const bool up = (lane & 2) != 0;
          float send1 = up ? s1[0] : s1[1], send2 = up ? s2[0] : s2[1];
          float keep1 = up ? s1[1] : s1[0], keep2 = up ? s2[1] : s2[0];
          s1[0] = keep1 + __shfl_xor_sync(0xffffffffu, send1, 2);
          s2[0] = keep2 + __shfl_xor_sync(0xffffffffu, send2, 2);
        }
        s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
        s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
        const int colj = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        if ((lane & 1) == 0 && gn + colj < p.N) {
          atomic_add_f64(&p.stats[gn + colj], (double)s1[0]);
          atomic_add_f64(&p.stats[(int64_t)p.N + gn + colj], (double)s2[0]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

struct MapKey {
  const void* base; int64_t cols, rows, ld, plane_stride; int planes, box_rows;
  bool operator==(const MapKey& o) const {
    return base == o.base && cols == o.cols && rows == o.rows && ld == o.ld && plane_stride == o.plane_stride &&
           planes == o.planes && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    auto mix = [&](int64_t v) { h ^= (size_t)v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.cols); mix(k.rows); mix(k.ld); mix(k.plane_stride); mix(k.planes); mix(k.box_rows);
    return h;
  }
};

// rank-3 map over {cols (contiguous), rows, planes} of bf16 with a {64, box_rows, 1} SWIZZLE_128B box
int tensor_map(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int planes, int64_t rows, int64_t cols, int64_t ld,
               int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const int64_t plane_stride = planes == 2 ? (int64_t)(lo - hi) : rows * ld;
  GPS_REQUIRE(planes == 1 || plane_stride > 0, GPS_ERR_ARG, "gemm_tma: the lo plane must follow the hi plane in memory");
  MapKey key{hi, cols, rows, ld, plane_stride, planes, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return GPS_OK;
    }
  }
  EncodeTiledFn enc = encode_fn();
  GPS_REQUIRE(enc, GPS_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(plane_stride > 0 ? plane_stride : 8) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(hi), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GPS_REQUIRE(r == CUDA_SUCCESS, GPS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows %lld cols %lld ld %lld box %d",
              (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return GPS_OK;
}

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

template <bool A_MN, bool B_MN, bool NARROW>
int launch(const CUtensorMap& tA, const CUtensorMap& tB, const TmaArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    GPS_CUDA(cudaFuncSetAttribute(k_gemm_tma<A_MN, B_MN, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  k_gemm_tma<A_MN, B_MN, NARROW><<<grid, kThreads, smem, stream>>>(tA, tB, a);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

int g_tma_force_bn = 0;

}  // namespace

void gemm_tma_set_force_bn(int bn) { g_tma_force_bn = bn; }

int gemm_tma(const GemmParams& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return GPS_OK;
  if (p.K <= 0 || !p.Ap.hi || !p.Bp.hi) return GPS_ERR_UNSUPPORTED;
  const bool split = p.precision == GPS_PREC_FP32;
  const int planes = split ? 2 : 1;
  if (split && (!p.Ap.lo || !p.Bp.lo)) return GPS_ERR_UNSUPPORTED;
  // TMA: 16-byte aligned bases and row pitches; the epilogue's 128-bit paths as in the register-staged kernel
  if (!aligned16(p.Ap.hi) || !aligned16(p.Bp.hi) || p.Ap.ld % 8 || p.Bp.ld % 8 || p.N % 4) return GPS_ERR_UNSUPPORTED;
  if (split && (!aligned16(p.Ap.lo) || !aligned16(p.Bp.lo))) return GPS_ERR_UNSUPPORTED;
  if ((p.C && (!aligned16(p.C) || p.ldc % 4)) || (!p.C && !p.Cp.hi)) return GPS_ERR_UNSUPPORTED;
  if (p.Cp.hi && (p.Cp.ld % 4 || (reinterpret_cast<uintptr_t>(p.Cp.hi) & 7) || (p.Cp.lo && (reinterpret_cast<uintptr_t>(p.Cp.lo) & 7))))
    return GPS_ERR_UNSUPPORTED;
  if ((p.bias && !aligned16(p.bias)) || (p.R1 && (!aligned16(p.R1) || p.ldr1 % 4)) ||
      (p.R2 && (!aligned16(p.R2) || p.ldr2 % 4)) || (p.mask_src && (!aligned16(p.mask_src) || p.ldmask % 4)) ||
      (p.C_pre && (!aligned16(p.C_pre) || p.ldpre % 4)))
    return GPS_ERR_UNSUPPORTED;
  if (p.splitk > 1 && (p.bias || p.act >= 0 || p.mask_src || p.stats || p.C_pre || p.p_drop != 0.f || p.p_drop2 != 0.f ||
                       p.Cp.hi || !p.C)) {
    set_error("gemm: split-K supports the plain fp32 product (+ residuals) only");
    return GPS_ERR_ARG;
  }
  if (p.colsum_a && !p.ta) {
    set_error("gemm: colsum_a needs ta == 1");
    return GPS_ERR_ARG;
  }
  const int mt = (int)ceil_div(p.M, BM);
  const int nkb = (int)ceil_div(p.K, BK);
  const int splits_hint = p.splitk > 1 ? (p.splitk < nkb ? p.splitk : nkb) : 1;
  // tile width: minimise waves x staged bytes per CTA, wider on ties (L2 -> SM operand traffic bounds the kernel)
  int bestBN = 128;
  long bestCost = -1;
  for (int nt = (int)ceil_div(p.N, 256); nt <= (int)ceil_div(p.N, 48) + 1; ++nt) {
    int bn = (int)round_up(ceil_div(p.N, nt), 16);
    if (bn > 256) continue;
    if (bn < 16) bn = 16;
    const int nb = (bn + 63) / 64;
    const long tiles = (long)mt * ceil_div(p.N, bn) * splits_hint;
    const long waves = ceil_div(tiles, nb == 1 ? 2L * kNumSMs : (long)kNumSMs);   // narrow tiles: two CTAs per SM
    const long cost = waves * (BM + nb * 64L);
    if (bestCost < 0 || cost < bestCost) { bestCost = cost; bestBN = bn; }
  }
  if (g_tma_force_bn > 0) bestBN = g_tma_force_bn;
  TmaArgs a;
  a.p = p;
  a.BN = bestBN;
  a.nb_blocks = (a.BN + 63) / 64;
  a.planes = planes;
  const bool narrow = a.nb_blocks == 1;
  const int stage_bytes = planes * (kATile + a.nb_blocks * kBlock);
  const int fixed = 1024 /*align*/ + 1024 /*barriers*/ + 16 * 16 * 8 * 4;
  int stages = ((narrow ? 112 : 226) * 1024 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return GPS_ERR_UNSUPPORTED;
  a.stages = stages;
  int splitk = p.splitk > 1 ? p.splitk : 1;
  if (splitk > nkb) splitk = nkb;
  a.kb_per_split = (int)ceil_div(nkb, splitk);
  splitk = (int)ceil_div(nkb, a.kb_per_split);
  a.p.splitk = p.splitk > 1 ? 2 : 1;   // "accumulate atomically" flag
  a.tmem_cols = a.BN <= 32 ? 32 : a.BN <= 64 ? 64 : a.BN <= 128 ? 128 : 256;
  const bool amn = p.ta != 0, bmn = p.tb != 0;
  // planes are addressed as stored: Aop[m,k] = A[m, k] (ta = 0: rows = M, cols = K) or A[k, m] (ta = 1: rows = K, cols = M)
  CUtensorMap tA, tB;
  GPS_TRY(tensor_map(p.Ap.hi, p.Ap.lo, planes, amn ? p.K : p.M, amn ? p.M : p.K, p.Ap.ld, amn ? 64 : BM, &tA));
  GPS_TRY(tensor_map(p.Bp.hi, p.Bp.lo, planes, bmn ? p.K : p.N, bmn ? p.N : p.K, p.Bp.ld, bmn ? 64 : a.BN, &tB));
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  dim3 grid((unsigned)ceil_div(p.N, a.BN), (unsigned)mt, (unsigned)splitk);
#define GPS_TMA_CASE(AM, BMN)                                                                           \
  if (amn == AM && bmn == BMN)                                                                          \
    return narrow ? launch<AM, BMN, true>(tA, tB, a, grid, smem, stream) : launch<AM, BMN, false>(tA, tB, a, grid, smem, stream);
  GPS_TMA_CASE(false, false)
  GPS_TMA_CASE(false, true)
  GPS_TMA_CASE(true, false)
  GPS_TMA_CASE(true, true)
#undef GPS_TMA_CASE
  return GPS_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------ fp32 -> planes
namespace {
struct ToPlanesDesc {
  ToPlanesItem it[16];
  int start[17];   // first 8-element chunk row-block of each item in the 1-D grid
  int n;
};
// one thread per 8 consecutive elements of a row: 2 x 128-bit loads, one 128-bit store per plane
__global__ void k_to_planes(ToPlanesDesc d) {
  int item = 0;
  while (item + 1 < d.n && (int)blockIdx.x >= d.start[item + 1]) ++item;
  const ToPlanesItem& it = d.it[item];
  const int cpr = (it.cols + 7) >> 3;                               // chunks per row
  const int64_t total = (int64_t)it.rows * cpr;
  const int64_t idx = ((int64_t)blockIdx.x - d.start[item]) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t r = idx / cpr;
  const int c = (int)(idx - r * cpr) * 8;
  const float* src = it.src + r * it.ld + c;
  float v[8];
  if (c + 8 <= it.cols) {
    const float4 x = ld4(src), y = ld4(src + 4);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = c + i < it.cols ? src[i] : 0.f;
  }
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  *reinterpret_cast<uint4*>(it.dst.hi + r * it.dst.ld + c) = hi;
  if (it.dst.lo) *reinterpret_cast<uint4*>(it.dst.lo + r * it.dst.ld + c) = lo;
}
}  // namespace

int to_planes(const ToPlanesItem* items, int n, cudaStream_t stream) {
  if (n <= 0) return GPS_OK;
  GPS_REQUIRE(n <= 16, GPS_ERR_ARG, "to_planes: at most 16 matrices per call");
  ToPlanesDesc d;
  d.n = 0;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    const ToPlanesItem& it = items[i];
    if (it.rows <= 0 || it.cols <= 0) continue;
    GPS_REQUIRE(it.src && it.dst.hi && it.ld % 4 == 0 && it.dst.ld % 8 == 0 && it.dst.ld >= round_up(it.cols, 8) &&
                    (reinterpret_cast<uintptr_t>(it.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(it.dst.hi) & 15) == 0 &&
                    (!it.dst.lo || (reinterpret_cast<uintptr_t>(it.dst.lo) & 15) == 0),
                GPS_ERR_ARG, "to_planes: operands must be 16-byte aligned, ld %% 4 == 0, plane ld %% 8 == 0");
    d.it[d.n] = it;
    d.start[d.n] = total;
    total += (int)ceil_div((int64_t)it.rows * ((it.cols + 7) >> 3), 256);
    ++d.n;
  }
  d.start[d.n] = total;
  if (total == 0) return GPS_OK;
  k_to_planes<<<(unsigned)total, 256, 0, stream>>>(d);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}

}  // namespace gps

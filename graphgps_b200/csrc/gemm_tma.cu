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
  int debug;                   // bring-up: 1 no global stores in the epilogue, 2 no TMEM loads
  unsigned long long* trace;   // bring-up: 16 globaltimer stamps per CTA (first 256 CTAs), see tools/gemm_trace.py
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GPS_TRACE(slot)                                                                         \
  do {                                                                                          \
    if (a.trace && cta_lin < 256) a.trace[cta_lin * 16 + (slot)] = gtimer();                    \
  } while (0)

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
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

  if (tid == 0) {
    GPS_TRACE(0);
    if (a.trace && cta_lin < 256) {
      unsigned smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      a.trace[cta_lin * 16 + 8] = smid;
    }
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
  if (tid == 0) GPS_TRACE(1);

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
        if (i == 0) GPS_TRACE(2);
      }
      GPS_TRACE(9);
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
        if (i == 0) GPS_TRACE(3);
        if (i == nkb - 1) GPS_TRACE(10);
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
      GPS_TRACE(4);
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
    if (tid == 0) GPS_TRACE(5);
    // ---- phase 1: accumulator TMEM -> registers -> shared staging tile [128][BN + 4] fp32.  All MMAs have retired, so
    // the operand stages are free and double as the staging buffer.  The pad keeps both the row-per-lane writes here and
    // the row-contiguous reads of phase 2 bank-conflict free.
    float* stage = reinterpret_cast<float*>(smem);
    const int sld = a.BN + 4;
    {
      const int q = warp & 3, half = warp >> 2;
      const int r = q * 32 + lane;
      const int nchunks = a.BN >> 4;
      for (int c = half; c < nchunks; c += 2) {
        float v[16];
        if (nkb > 0 && !(a.debug & 2)) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), v);
        else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        }
        float4* dst = reinterpret_cast<float4*>(stage + r * sld + c * 16);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- phase 2: G = BN/4 threads per row, each owning 4 consecutive columns for all of its rows: every global access
    // (bias, residuals, act' mask, fp32 / plane stores, split-K atomics) is a contiguous row segment, the per-column
    // constants live in registers and the BatchNorm column sums are accumulated per thread.
    const int G = a.BN >> 2;
    const int rpp = kEpiWarps * 32 / G;            // rows per pass
    const int rip = tid / G, cg = tid - rip * G;
    const int col = n0 + cg * 4;
    const bool col_ok = rip < rpp && col < p.N;
    float4 s1 = f4zero(), s2 = f4zero(), q0 = f4zero(), q1 = f4zero();   // sum w, sum w^2, sum w*zhat_0, sum w*zhat_1
    if (col_ok) {
      // rows of this thread: r = rip + k * rpp, k < nrows.  Everything is addressed through per-thread base pointers
      // advanced by a constant stride, and the loop is unrolled by 4 rows so that the shared/global loads of a group are
      // in flight together: with 2 epilogue warps per scheduler the pass is instruction-latency bound otherwise
      // (measured: 8.5 us for a 128 x 256 tile with neither the TMEM loads nor the global stores on the critical path).
      const int rows_here = min(BM, p.M - m0);
      const int nrows = rip < rows_here ? (rows_here - rip + rpp - 1) / rpp : 0;
      const int64_t row0 = m0 + rip;
      const float* sp = stage + rip * sld + cg * 4;
      const int s_st = rpp * sld;
      float* cp = p.C ? p.C + row0 * p.ldc + col : nullptr;
      const int64_t c_st = (int64_t)rpp * p.ldc;
      const float* r1 = p.R1 ? p.R1 + row0 * p.ldr1 + col : nullptr;
      const int64_t r1_st = (int64_t)rpp * p.ldr1;
      const float* r2 = p.R2 ? p.R2 + row0 * p.ldr2 + col : nullptr;
      const int64_t r2_st = (int64_t)rpp * p.ldr2;
      if (p.splitk > 1) {
        const bool res = blockIdx.z == 0;            // the first split also carries the residual terms
        for (int k = 0; k < nrows; k += 4) {
          float4 w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) w[u] = k + u < nrows ? *reinterpret_cast<const float4*>(sp + (k + u) * s_st) : f4zero();
          if (res && r1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) if (k + u < nrows) w[u] = f4add(w[u], ld4(r1 + (k + u) * r1_st));
          }
          if (res && r2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) if (k + u < nrows) w[u] = f4add(w[u], ld4(r2 + (k + u) * r2_st));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k + u < nrows) atomicAdd(reinterpret_cast<float4*>(cp + (k + u) * c_st), w[u]);   // red.global.add.v4.f32
        }
      } else {
        const bool fast = !p.C_pre && !(p.mask_src && !p.mask_is_post) && p.p_drop == 0.f && p.p_drop2 == 0.f &&
                          (p.act < 0 || p.act == GPS_ACT_RELU);
        const float4 bb = p.bias ? ld4(p.bias + col) : f4zero();
        const float* mk = p.mask_src ? p.mask_src + row0 * p.ldmask + col : nullptr;
        const int64_t mk_st = (int64_t)rpp * p.ldmask;
        // plane column: identity, or the per-head padded layout of the attention operands (cp_hd)
        int pcol = col, npad = 0;
        if (p.cp_hd > 0) {
          const int cl = col - (p.cp_col0), hq = cl / p.cp_hd, cw = cl - hq * p.cp_hd;
          pcol = hq * p.cp_hd_pad + cw;
          npad = (cw + 4 == p.cp_hd) ? p.cp_hd_pad - p.cp_hd : 0;   // this thread also zeroes the head's pad columns
        }
        __nv_bfloat16* ph = p.Cp.hi ? p.Cp.hi + row0 * p.Cp.ld + pcol : nullptr;
        __nv_bfloat16* pl = p.Cp.lo ? p.Cp.lo + row0 * p.Cp.ld + pcol : nullptr;
        const int64_t p_st = (int64_t)rpp * p.Cp.ld;
        const bool relu = p.act == GPS_ACT_RELU;
        // fused BatchNorm-backward reductions (bnred): per-column constants -mean and invstd, row pointers into z_k
        const float* bz0 = p.bnred[0].sums ? p.bnred[0].z + row0 * p.bnred[0].ldz + col : nullptr;
        const float* bz1 = p.bnred[1].sums ? p.bnred[1].z + row0 * p.bnred[1].ldz + col : nullptr;
        const int64_t bz0_st = (int64_t)rpp * p.bnred[0].ldz, bz1_st = (int64_t)rpp * p.bnred[1].ldz;
        float4 bm0 = f4zero(), bi0 = f4zero(), bm1 = f4zero(), bi1 = f4zero();
        if (bz0) { bm0 = f4scale(ld4(p.bnred[0].mean + col), -1.f); bi0 = ld4(p.bnred[0].invstd + col); }
        if (bz1) { bm1 = f4scale(ld4(p.bnred[1].mean + col), -1.f); bi1 = ld4(p.bnred[1].invstd + col); }
        if (fast) {
          for (int k = 0; k < nrows; k += 4) {
            float4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              w[u] = f4add(k + u < nrows ? *reinterpret_cast<const float4*>(sp + (k + u) * s_st) : f4zero(), bb);
            if (relu) {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                w[u] = make_float4(fmaxf(w[u].x, 0.f), fmaxf(w[u].y, 0.f), fmaxf(w[u].z, 0.f), fmaxf(w[u].w, 0.f));
            }
            if (mk) {   // relu': the saved post-activation value is positive
              float4 m[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) m[u] = k + u < nrows ? ld4(mk + (k + u) * mk_st) : f4zero();
#pragma unroll
              for (int u = 0; u < 4; ++u)
                w[u] = make_float4(m[u].x > 0.f ? w[u].x : 0.f, m[u].y > 0.f ? w[u].y : 0.f, m[u].z > 0.f ? w[u].z : 0.f,
                                   m[u].w > 0.f ? w[u].w : 0.f);
            }
            if (r1) {
#pragma unroll
              for (int u = 0; u < 4; ++u) if (k + u < nrows) w[u] = f4add(w[u], ld4(r1 + (k + u) * r1_st));
            }
            if (r2) {
#pragma unroll
              for (int u = 0; u < 4; ++u) if (k + u < nrows) w[u] = f4add(w[u], ld4(r2 + (k + u) * r2_st));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (k + u >= nrows) continue;
              if (cp && !(a.debug & 1)) st4(cp + (k + u) * c_st, w[u]);
              if (ph) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(w[u].x, w[u].y), h1 = __floats2bfloat162_rn(w[u].z, w[u].w);
                *reinterpret_cast<uint2*>(ph + (k + u) * p_st) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                if (pl) {
                  __nv_bfloat162 l0 = __floats2bfloat162_rn(w[u].x - __low2float(h0), w[u].y - __high2float(h0));
                  __nv_bfloat162 l1 = __floats2bfloat162_rn(w[u].z - __low2float(h1), w[u].w - __high2float(h1));
                  *reinterpret_cast<uint2*>(pl + (k + u) * p_st) =
                      make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
                }
                for (int z = 0; z < npad; z += 4) {
                  *reinterpret_cast<uint2*>(ph + (k + u) * p_st + 4 + z) = make_uint2(0u, 0u);
                  if (pl) *reinterpret_cast<uint2*>(pl + (k + u) * p_st + 4 + z) = make_uint2(0u, 0u);
                }
              }
              s1 = f4add(s1, w[u]);
              s2 = f4fma(w[u], w[u], s2);
              if (bz0) q0 = f4fma(w[u], f4mul(f4add(ld4(bz0 + (k + u) * bz0_st), bm0), bi0), q0);
              if (bz1) q1 = f4fma(w[u], f4mul(f4add(ld4(bz1 + (k + u) * bz1_st), bm1), bi1), q1);
            }
          }
        } else {
          // general path: pre-activation copy, GELU and its derivative, dropout (two chained sites for the Performer)
          const uint64_t drop_off = p.offset + ((p.p_drop > 0.f || p.p_drop2 > 0.f) && p.offset_dev ? *p.offset_dev : 0ull);
          for (int k = 0; k < nrows; ++k) {
            const int64_t row = row0 + (int64_t)k * rpp;
            float4 w = f4add(*reinterpret_cast<const float4*>(sp + k * s_st), bb);
            if (p.C_pre) st4(p.C_pre + row * p.ldpre + col, w);
            if (p.act >= 0) w = make_float4(act_fwd_rt(p.act, w.x), act_fwd_rt(p.act, w.y), act_fwd_rt(p.act, w.z), act_fwd_rt(p.act, w.w));
            if (mk) {
              const float4 ms = ld4(mk + k * mk_st);
              if (p.mask_is_post) {
                w.x = ms.x > 0.f ? w.x : 0.f; w.y = ms.y > 0.f ? w.y : 0.f; w.z = ms.z > 0.f ? w.z : 0.f; w.w = ms.w > 0.f ? w.w : 0.f;
              } else {
                w.x *= act_bwd_rt(p.mask_act, ms.x); w.y *= act_bwd_rt(p.mask_act, ms.y);
                w.z *= act_bwd_rt(p.mask_act, ms.z); w.w *= act_bwd_rt(p.mask_act, ms.w);
              }
            }
            if (p.p_drop2 > 0.f) w = f4mul(w, dropout_scale4(p.p_drop2, p.seed, drop_off, p.site2, ((uint64_t)row * (uint64_t)p.N + col) >> 2));
            if (p.p_drop > 0.f) w = f4mul(w, dropout_scale4(p.p_drop, p.seed, drop_off, p.site, ((uint64_t)row * (uint64_t)p.N + col) >> 2));
            if (r1) w = f4add(w, ld4(r1 + k * r1_st));
            if (r2) w = f4add(w, ld4(r2 + k * r2_st));
            if (cp) st4(cp + k * c_st, w);
            if (ph) {
              planes_store4(p.Cp, row, pcol, w);
              for (int z = 0; z < npad; z += 4) planes_store4(p.Cp, row, pcol + 4 + z, f4zero());
            }
            s1 = f4add(s1, w);
            s2 = f4fma(w, w, s2);
          }
        }
      }
    }
    if (p.stats || p.bnred[0].sums || p.bnred[1].sums) {
      // column sums: the rpp threads that share a column group meet in shared memory (the staging tile is free again
      // after the barrier), one double atomic per column per CTA and accumulator
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float4* rs = reinterpret_cast<float4*>(stage);
      if (rip < rpp) {
        rs[(rip * G + cg) * 4] = s1;
        rs[(rip * G + cg) * 4 + 1] = s2;
        rs[(rip * G + cg) * 4 + 2] = q0;
        rs[(rip * G + cg) * 4 + 3] = q1;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < G && n0 + tid * 4 < p.N) {
        float4 t[4] = {f4zero(), f4zero(), f4zero(), f4zero()};
        for (int y = 0; y < rpp; ++y)
#pragma unroll
          for (int k = 0; k < 4; ++k) t[k] = f4add(t[k], rs[(y * G + tid) * 4 + k]);
        const int c0 = n0 + tid * 4;
        auto add4 = [&](double* dst, float4 v) {
          atomic_add_f64(dst + 0, (double)v.x); atomic_add_f64(dst + 1, (double)v.y);
          atomic_add_f64(dst + 2, (double)v.z); atomic_add_f64(dst + 3, (double)v.w);
        };
        if (p.stats) {
          add4(p.stats + c0, t[0]);
          add4(p.stats + (int64_t)p.N + c0, t[1]);
        }
        if (p.bnred[0].sums) {
          add4(p.bnred[0].sums + c0, t[0]);
          add4(p.bnred[0].sums + (int64_t)p.N + c0, t[2]);
        }
        if (p.bnred[1].sums) {
          add4(p.bnred[1].sums + c0, t[0]);
          add4(p.bnred[1].sums + (int64_t)p.N + c0, t[3]);
        }
      }
    }
  }

  if (tid == 0) GPS_TRACE(6);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) GPS_TRACE(7);
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

}  // namespace

int make_tensor_map(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int planes, int64_t rows, int64_t cols, int64_t ld,
                    int box_rows, CUtensorMap* out) {
  return tensor_map(hi, lo, planes, rows, cols, ld, box_rows, out);
}

namespace {

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
unsigned long long* g_tma_trace = nullptr;
int g_tma_debug = 0;

}  // namespace

void gemm_tma_set_force_bn(int bn) { g_tma_force_bn = bn & 0xFFFF; g_tma_debug = bn >> 16; }
void gemm_tma_set_trace(unsigned long long* buf) { g_tma_trace = buf; }

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
  if ((p.bnred[0].sums || p.bnred[1].sums) &&
      (p.splitk > 1 || p.C_pre || (p.mask_src && !p.mask_is_post) || p.p_drop != 0.f || p.p_drop2 != 0.f ||
       (p.act >= 0 && p.act != GPS_ACT_RELU))) {
    set_error("gemm: fused BatchNorm-backward reductions need the plain epilogue (no split-K / dropout / GELU)");
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
  a.trace = g_tma_trace;
  a.debug = g_tma_debug;
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

// gemm_tc.cu — tcgen05 (5th-gen tensor core) dense product for every Linear of the GPS layer and
// their data/weight gradients, with the layer's fused epilogues.
//
//   C[M,N] (+)= epi( Aop[M,K] * Bop[K,N] ),  fp32 in HBM, bf16 operands on the tensor cores,
//   fp32 accumulation in TMEM.  precision FP32: split-bf16 x3 (hi*hi + hi*lo + lo*hi, ~2^-16
//   relative), precision BF16: a single bf16 pass.
//
// One 128 x BN output tile per CTA (BN a runtime multiple of 16 up to 256; wide tiles matter: the kernel
// is bound by L2->SM operand traffic, M*N*K*4*(1/BM + 1/BN) bytes):
//   * warps 0-7 stage operands: coalesced 128-bit global loads of the fp32 tiles -> bf16 hi/lo split in
//     registers -> 16-byte st.shared into the canonical UMMA SWIZZLE_128B layout (K-major when the
//     reduction dim is contiguous in HBM, MN-major when it is the row dim, e.g. weight gradients
//     dW = G^T X) -> fence.proxy.async -> one mbarrier arrive per warp.  All per-chunk addresses are
//     (value for chunk 0) + q * constant, computed once per CTA.  No transposes, no separate conversion
//     pass; the fp32->bf16 split costs no extra HBM bytes.
//   * warp 8: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16) per
//     16-wide K step; tcgen05.commit releases the smem stage / signals the epilogue.
//   * epilogue (warps 0-7): tcgen05.ld 32x32b.x16 -> bias / activation / act' mask / dropout /
//     residuals / 128-bit stores; BatchNorm column sums by a warp butterfly reduce-scatter + double
//     atomics; split-K partials by fp32 atomics; bias-gradient column sums from the staged A tile.
// Smem stages form an mbarrier ring (full: one arrival per producer warp; empty: tcgen05.commit).
// Measured lessons kept in the code: per-thread mbarrier arrivals (256 per stage) serialise on the
// barrier unit (~1 us per k-block) -> per-warp arrivals; per-chunk integer divisions made the producers
// issue-bound -> precomputed addressing (tools/gemm_triage.py has the switches and the clock64 trace).
#include <cuda_bf16.h>

#include <algorithm>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace gps {

namespace {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // k-block: one 128-byte swizzle row of bf16
constexpr int kProducerWarps = 8;           // operand staging + epilogue
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kATileBytes = BM * BK * 2;       // 16 KB
constexpr int kBBlockBytes = 64 * BK * 2;      // 8 KB per 64 columns of B
using namespace tc;   // PTX wrappers: tc_ptx.cuh

// Byte offset of chunk `c` inside an operand tile stored in the canonical SWIZZLE_128B layout.
//  K-major : rows of 128 B (64 bf16 of K), 8-row groups of 1024 B.
//  MN-major: 64-column blocks of kBBlockBytes; inside a block 8-k-row groups of 1024 B, each k-row 128 B.
template <bool MN>
__device__ __forceinline__ uint32_t chunk_offset(int c, int tile_rows) {
  if (!MN) {
    const int r = c >> 3, ck = c & 7;
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((ck ^ (r & 7)) << 4));
  }
  const int rc = tile_rows >> 3;
  const int k = c / rc, cm = c % rc;
  const int blk = cm >> 3, cc = cm & 7;
  return (uint32_t)(blk * kBBlockBytes + (k >> 3) * 1024 + (k & 7) * 128 + ((cc ^ (k & 7)) << 4));
}

struct TcArgs {
  GemmParams p;
  int BN;          // tile width (multiple of 16, <= 256)
  int nb_blocks;   // ceil(BN / 64)
  int stages;
  int kb_per_split;
  int tmem_cols;
  const uint8_t* bpk_hi;   // pre-packed K-major B planes (BPRE variants), else null
  const uint8_t* bpk_lo;
  int bpk_groups;          // 8-row groups per k-block in the packed planes
  int bpk_row0;            // first row of this GEMM's B inside the packed matrix (multiple of 8)
  int bpk_kb0;             // k-block offset of this GEMM's reduction range inside the planes
  int bpk_shift;           // 3: K-major planes (1 KB per 8 rows), 6: MN-major planes (8 KB per 64 columns)
  int debug;       // perf-triage switches (gps_debug_set): 1 no global loads, 2 no convert/store, 4 no MMA, 8 no epilogue
};

// NBC = B chunks per producer thread per k-block (2: tiles up to 64 columns, 8: up to 256).  The narrow variant
// fits in 112 registers and ~100 KB of shared memory, so two CTAs share an SM and one CTA's load/convert phase
// overlaps the other's MMA/epilogue phase.
// BPRE: the B operand (an nn.Linear weight) was pre-packed once per step by k_prepack_weights into bf16 hi/lo
// planes that already have the shared-memory image of a tile (SWIZZLE_128B; K-major: 8-row groups, MN-major:
// 64-column blocks; one 64-deep k-block after the other), so a stage's B tile is ONE contiguous range: a single
// elected thread fetches it with bulk TMA (cp.async.bulk ... mbarrier::complete_tx) and the producer warps only
// stage A.  K-major planes serve y = x W^T (forward), MN-major planes serve g_x = g_y W (data gradients).
template <bool A_MN, bool B_MN, bool SPLIT, int NBC, bool BPRE>
__global__ void __launch_bounds__(kThreads, NBC == 2 ? 2 : 1) k_gemm_tc(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const GemmParams& p = a.p;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = a.nb_blocks * kBBlockBytes;
  const int plane = SPLIT ? 2 : 1;
  const int stage_bytes = plane * (kATileBytes + b_tile_bytes);
  const int S = a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);  // 16 x 16 x 8 floats (bias-gradient partials)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * a.BN;
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * a.kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + a.kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bars[s]), kProducerWarps + (BPRE ? 1 : 0));
      mbar_init(smem_u32(&bars[S + s]), 1);
    }
    mbar_init(smem_u32(&bars[2 * S]), 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kMmaWarp) {
    // =========================================================== MMA issuer
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = A_MN ? kBBlockBytes : 16, b_lbo = B_MN ? kBBlockBytes : 16;
      const uint32_t a_kstep = A_MN ? 2048 : 32, b_kstep = B_MN ? 2048 : 32;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        mbar_wait(smem_u32(&bars[s]), (uint32_t)(i / S) & 1u);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb_hi = sa_hi + plane * kATileBytes;
        const uint32_t sa_lo = sa_hi + kATileBytes;
        const uint32_t sb_lo = sb_hi + b_tile_bytes;
        if (!(a.debug & 4)) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da_hi = make_desc(sa_hi + kk * a_kstep, a_lbo, 1024);
            const uint64_t db_hi = make_desc(sb_hi + kk * b_kstep, b_lbo, 1024);
            if (SPLIT) {
              const uint64_t da_lo = make_desc(sa_lo + kk * a_kstep, a_lbo, 1024);
              const uint64_t db_lo = make_desc(sb_lo + kk * b_kstep, b_lbo, 1024);
              umma_bf16(tmem_base, da_lo, db_hi, idesc, (i | kk) != 0);
              umma_bf16(tmem_base, da_hi, db_lo, idesc, 1u);
              umma_bf16(tmem_base, da_hi, db_hi, idesc, 1u);
            } else {
              umma_bf16(tmem_base, da_hi, db_hi, idesc, (i | kk) != 0);
            }
          }
        }
        umma_commit(smem_u32(&bars[S + s]));  // frees the smem stage once these MMAs retire
      }
      umma_commit(smem_u32(&bars[2 * S]));    // accumulator complete
    }
    __syncwarp();
  } else {
    // =========================================================== operand producers
    // chunk c = tid + 256 q: every per-q quantity is (value at q = 0) + q * constant
    //   K-major : row (tid>>3) + 32 q, k-chunk (tid&7)            -> global += 32 rows,   smem += 4096 B
    //   MN-major: k-row (tid>>rcs) + (256>>rcs) q, row chunk tid & (rc-1) -> global += (256>>rcs) k-rows
    const int b_rows = a.nb_blocks * 64;
    const int nb_chunks = a.nb_blocks * 2;  // per thread: (nb_blocks*64 rows * 8 chunks) / 256
    const int a_rcs = 4;
    const int b_rcs = a.nb_blocks == 1 ? 3 : a.nb_blocks == 2 ? 4 : 5;   // log2(b_rows / 8); nb_blocks in {1,2,4}
    const int k_end = min(p.K, kb_end * BK);
    const int64_t a_kstride = A_MN ? (int64_t)BK * p.lda : BK;   // floats per k-block
    const int64_t b_kstride = B_MN ? (int64_t)BK * p.ldb : BK;
    const int a_row0 = A_MN ? m0 + (tid & 15) * 8 : m0 + (tid >> 3);
    const int a_k0 = A_MN ? (tid >> a_rcs) : (tid & 7) * 8;
    const float* a_g0 = A_MN ? p.A + ((int64_t)kb_begin * BK + a_k0) * p.lda + a_row0
                             : p.A + (int64_t)a_row0 * p.lda + (int64_t)kb_begin * BK + a_k0;
    const int64_t a_gq = A_MN ? (int64_t)(kProducerThreads >> a_rcs) * p.lda : (int64_t)32 * p.lda;
    const uint32_t a_s0 = chunk_offset<A_MN>(tid, BM);
    const uint32_t a_sq = A_MN ? (uint32_t)((kProducerThreads >> a_rcs) / 8) * 1024u : 4096u;
    const int b_row0 = B_MN ? n0 + (tid & ((1 << b_rcs) - 1)) * 8 : n0 + (tid >> 3);
    const int b_k0 = B_MN ? (tid >> b_rcs) : (tid & 7) * 8;
    const float* b_g0 = B_MN ? p.B + ((int64_t)kb_begin * BK + b_k0) * p.ldb + b_row0
                             : p.B + (int64_t)b_row0 * p.ldb + (int64_t)kb_begin * BK + b_k0;
    const int64_t b_gq = B_MN ? (int64_t)(kProducerThreads >> b_rcs) * p.ldb : (int64_t)32 * p.ldb;
    const uint32_t b_s0 = chunk_offset<B_MN>(tid, b_rows);
    const uint32_t b_sq = B_MN ? (uint32_t)((kProducerThreads >> b_rcs) / 8) * 1024u : 4096u;
    float csum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) csum[e] = 0.f;
    const bool do_colsum = A_MN && p.colsum_a != nullptr && blockIdx.x == 0;

    for (int i = 0; i < nkb; ++i) {
      const int s = i % S;
      const uint32_t ph = (uint32_t)(i / S) & 1u;
      const int krem = k_end - (kb_begin + i) * BK;     // valid k extent of this k-block (<= 64 on the tail)
      float4 va[4][2], vb[NBC][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool ok = !(a.debug & 1) && (A_MN ? (a_row0 < p.M && a_k0 + q * (kProducerThreads >> a_rcs) < krem)
                                                : (a_row0 + 32 * q < p.M && a_k0 < krem));
        const float* src = a_g0 + (int64_t)i * a_kstride + q * a_gq;
        va[q][0] = ok ? ld4(src) : f4zero();
        va[q][1] = ok ? ld4(src + 4) : f4zero();
      }
#pragma unroll
      for (int q = 0; q < (BPRE ? 0 : NBC); ++q) {
        const bool ok = !(a.debug & 1) && q < nb_chunks &&
                        (B_MN ? (b_row0 < p.N && b_k0 + q * (kProducerThreads >> b_rcs) < krem)
                              : (b_row0 + 32 * q < p.N && b_k0 < krem));
        const float* src = b_g0 + (int64_t)i * b_kstride + q * b_gq;
        vb[q][0] = ok ? ld4(src) : f4zero();
        vb[q][1] = ok ? ld4(src + 4) : f4zero();
      }
      if (do_colsum) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          csum[0] += va[q][0].x; csum[1] += va[q][0].y; csum[2] += va[q][0].z; csum[3] += va[q][0].w;
          csum[4] += va[q][1].x; csum[5] += va[q][1].y; csum[6] += va[q][1].z; csum[7] += va[q][1].w;
        }
      }
      // one lane per warp polls / arrives: per-thread mbarrier traffic serialises on the barrier unit
      if (lane == 0) mbar_wait(smem_u32(&bars[S + s]), ph ^ 1u);   // slot free (first pass returns at once)
      __syncwarp();
      uint8_t* st = smem + (size_t)s * stage_bytes;
      uint8_t* sa_hi = st;
      uint8_t* sa_lo = st + kATileBytes;
      uint8_t* sb_hi = st + plane * kATileBytes;
      uint8_t* sb_lo = sb_hi + b_tile_bytes;
      if (BPRE && tid == 0) {   // bulk TMA of the pre-packed weight tile(s) of this k-block
        const int64_t off = ((int64_t)(kb_begin + i + a.bpk_kb0) * a.bpk_groups + ((n0 + a.bpk_row0) >> a.bpk_shift))
                            << (7 + a.bpk_shift);
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_arrive_expect_tx(bar, (uint32_t)(plane * b_tile_bytes));
        tma_bulk_g2s(smem_u32(sb_hi), a.bpk_hi + off, (uint32_t)b_tile_bytes, bar);
        if (SPLIT) tma_bulk_g2s(smem_u32(sb_lo), a.bpk_lo + off, (uint32_t)b_tile_bytes, bar);
      }
      if (!(a.debug & 2)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 hi, lo;
          split8(reinterpret_cast<const float*>(va[q]), hi, lo);
          *reinterpret_cast<uint4*>(sa_hi + a_s0 + q * a_sq) = hi;
          if (SPLIT) *reinterpret_cast<uint4*>(sa_lo + a_s0 + q * a_sq) = lo;
        }
#pragma unroll
        for (int q = 0; q < (BPRE ? 0 : NBC); ++q)
          if (q < nb_chunks) {
            uint4 hi, lo;
            split8(reinterpret_cast<const float*>(vb[q]), hi, lo);
            *reinterpret_cast<uint4*>(sb_hi + b_s0 + q * b_sq) = hi;
            if (SPLIT) *reinterpret_cast<uint4*>(sb_lo + b_s0 + q * b_sq) = lo;
          }
      } else if (va[0][0].x == 123.456f) { sa_hi[0] = (uint8_t)va[1][0].x; }   // keep the loads alive
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[s]));
    }

    // bias gradient: thread t always owns MN chunk (t % 16) of A^T -> reduce the 16 owners in smem
    if (do_colsum) {
      const int cm = tid & 15, owner = tid >> 4;
#pragma unroll
      for (int e = 0; e < 8; ++e) red[(owner * 16 + cm) * 8 + e] = csum[e];
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

    // =========================================================== epilogue
    if (nkb > 0) {
      if (lane == 0) mbar_wait(smem_u32(&bars[2 * S]), 0u);
      __syncwarp();
      tc_fence_after();
    }
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const int nchunks = (a.debug & 8) ? 0 : (a.BN >> 4);
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
            for (int e = 0; e < 16; e += 4) {  // N % 4 == 0: whole 16-byte groups; red.global.add.v4.f32
              if (gn + e >= p.N) continue;
              float4 w4 = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
              if (blockIdx.z == 0) {           // the first split also carries the residual terms
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
            float4 sc = dropout_scale4(p.p_drop2, p.seed, p.offset + (p.offset_dev ? *p.offset_dev : 0ull), p.site2,
                                     ((uint64_t)row * (uint64_t)p.N + col) >> 2);
            w[0] *= sc.x; w[1] *= sc.y; w[2] *= sc.z; w[3] *= sc.w;
          }
          if (ok && p.p_drop > 0.f) {
            float4 sc = dropout_scale4(p.p_drop, p.seed, p.offset + (p.offset_dev ? *p.offset_dev : 0ull), p.site,
                                     ((uint64_t)row * (uint64_t)p.N + col) >> 2);
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
          if (ok) st4(p.C + (int64_t)row * p.ldc + col, make_float4(w[0], w[1], w[2], w[3]));
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

int g_tc_debug = 0;
int g_tc_force_bn = 0;

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

template <bool A_MN, bool B_MN, bool SPLIT, int NBC, bool BPRE>
int launch1(const TcArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    GPS_CUDA(cudaFuncSetAttribute(k_gemm_tc<A_MN, B_MN, SPLIT, NBC, BPRE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    attr_done = true;
  }
  k_gemm_tc<A_MN, B_MN, SPLIT, NBC, BPRE><<<grid, kThreads, smem, stream>>>(a);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}
template <bool A_MN, bool B_MN, bool SPLIT>
int launch(const TcArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  if constexpr (!A_MN) {
    if (a.bpk_hi)
      return a.nb_blocks == 1 ? launch1<false, B_MN, SPLIT, 2, true>(a, grid, smem, stream)
                              : launch1<false, B_MN, SPLIT, 8, true>(a, grid, smem, stream);
  }
  return a.nb_blocks == 1 ? launch1<A_MN, B_MN, SPLIT, 2, false>(a, grid, smem, stream)
                          : launch1<A_MN, B_MN, SPLIT, 8, false>(a, grid, smem, stream);
}

}  // namespace

void gemm_tc_set_debug(int v) {   // low 4 bits: triage switches; bits 8.. : forced tile width (0 = heuristic)
  g_tc_debug = v & 0xFF;
  g_tc_force_bn = (v >> 8) & 0x1FF;
}

int gemm_tc(const GemmParams& p, cudaStream_t stream) {
  static const int mode = [] {
    const char* e = getenv("GPS_B200_TC");
    return e ? atoi(e) : 3;   // bit0: K-major operands, bit1: MN-major operands
  }();
  if (p.M <= 0 || p.N <= 0) return GPS_OK;
  if (p.K <= 0) return GPS_ERR_UNSUPPORTED;
  const bool any_mn = p.ta || p.tb;
  if (!(mode & 1)) return GPS_ERR_UNSUPPORTED;
  if (any_mn && !(mode & 2)) return GPS_ERR_UNSUPPORTED;
  // 128-bit paths: aligned bases, leading dimensions and N multiples of 4
  if (!aligned16(p.A) || !aligned16(p.B) || !aligned16(p.C) || p.lda % 4 || p.ldb % 4 || p.ldc % 4 || p.N % 4)
    return GPS_ERR_UNSUPPORTED;
  if ((p.bias && !aligned16(p.bias)) || (p.R1 && (!aligned16(p.R1) || p.ldr1 % 4)) ||
      (p.R2 && (!aligned16(p.R2) || p.ldr2 % 4)) || (p.mask_src && (!aligned16(p.mask_src) || p.ldmask % 4)) ||
      (p.C_pre && (!aligned16(p.C_pre) || p.ldpre % 4)))
    return GPS_ERR_UNSUPPORTED;
  // lean producer loop: whole 8-element chunks are either inside or outside the operand
  if ((!p.ta && p.K % 8) || (!p.tb && p.K % 8) || (p.ta && p.M % 8) || (p.tb && p.N % 8)) return GPS_ERR_UNSUPPORTED;
  if (p.splitk > 1 && (p.bias || p.act >= 0 || p.mask_src || p.stats || p.C_pre || p.p_drop != 0.f || p.p_drop2 != 0.f)) {
    set_error("gemm: split-K supports the plain product (+ residuals) only");
    return GPS_ERR_ARG;
  }
  if (p.colsum_a && !p.ta) {
    set_error("gemm: colsum_a needs ta == 1");
    return GPS_ERR_ARG;
  }
  const bool split = p.precision == GPS_PREC_FP32;
  const int plane = split ? 2 : 1;
  const int mt = (int)ceil_div(p.M, BM);
  const int nkb = (int)ceil_div(p.K, BK);
  const int splits_hint = p.splitk > 1 ? (p.splitk < nkb ? p.splitk : nkb) : 1;

  // tile width: BN in {64,128,256}-block granularity (1, 2 or 4 staged 64-column blocks); the kernel is bound
  // by operand traffic ~ tiles x (128 + staged B rows), so minimise waves x staged rows, wider on ties
  const bool pre_k = p.bpk && !p.bpk_mn && !p.ta && !p.tb && p.bpk_row0 % 8 == 0 && (split ? p.bpk_lo_off > 0 : true);
  const bool pre_mn = p.bpk && p.bpk_mn && !p.ta && p.tb && p.bpk_row0 % 64 == 0 && (split ? p.bpk_lo_off > 0 : true);
  int bestBN = 128;
  long bestCost = -1;
  for (int nt = (int)ceil_div(p.N, 256); nt <= (int)ceil_div(p.N, 48) + 1; ++nt) {
    int bn = (int)round_up(ceil_div(p.N, nt), pre_mn ? 64 : 16);   // MN-major planes are cut at 64-column blocks
    if (bn > 256) continue;
    if (bn < 16) bn = 16;
    int nb = bn <= 64 ? 1 : bn <= 128 ? 2 : 4;
    long tiles = (long)mt * ceil_div(p.N, bn) * splits_hint;
    long waves = ceil_div(tiles, nb == 1 ? 2L * kNumSMs : (long)kNumSMs);   // narrow tiles: two CTAs per SM
    long cost = waves * (BM + nb * 64L);
    if (bestCost < 0 || cost < bestCost) { bestCost = cost; bestBN = bn; }
  }
  if (g_tc_force_bn > 0 && !(pre_mn && g_tc_force_bn % 64)) bestBN = g_tc_force_bn;   // tuning hook (tools/gemm_tune.py)
  TcArgs a;
  a.p = p;
  a.BN = bestBN;
  a.nb_blocks = a.BN <= 64 ? 1 : a.BN <= 128 ? 2 : 4;
  const int stage_bytes = plane * (kATileBytes + a.nb_blocks * kBBlockBytes);
  int stages = ((a.nb_blocks == 1 ? 100 : 200) * 1024) / stage_bytes;   // narrow tiles leave room for a second CTA
  if (stages > 4) stages = 4;
  if (stages < 2) return GPS_ERR_UNSUPPORTED;
  a.stages = stages;
  int splitk = p.splitk > 1 ? p.splitk : 1;
  if (splitk > nkb) splitk = nkb;
  a.kb_per_split = (int)ceil_div(nkb, splitk);
  splitk = (int)ceil_div(nkb, a.kb_per_split);
  a.p.splitk = p.splitk > 1 ? 2 : 1;   // "accumulate atomically" flag
  a.tmem_cols = a.BN <= 32 ? 32 : a.BN <= 64 ? 64 : a.BN <= 128 ? 128 : 256;
  a.debug = g_tc_debug;
  a.bpk_hi = a.bpk_lo = nullptr; a.bpk_groups = 0; a.bpk_row0 = 0; a.bpk_shift = 3; a.bpk_kb0 = p.bpk_kb0;
  if (pre_k || pre_mn) {
    a.bpk_hi = (const uint8_t*)p.bpk;
    a.bpk_lo = a.bpk_hi + p.bpk_lo_off;
    a.bpk_groups = p.bpk_groups;
    a.bpk_row0 = p.bpk_row0;
    a.bpk_shift = pre_mn ? 6 : 3;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align*/ + (2 * stages + 1) * 8 + 16 + 16 * 16 * 8 * 4;
  dim3 grid((unsigned)ceil_div(p.N, a.BN), (unsigned)mt, (unsigned)splitk);
  const bool amn = p.ta != 0, bmn = p.tb != 0;
#define GPS_TC_CASE(AM, BMN)                                                        \
  if (amn == AM && bmn == BMN)                                                      \
    return split ? launch<AM, BMN, true>(a, grid, smem, stream) : launch<AM, BMN, false>(a, grid, smem, stream);
  GPS_TC_CASE(false, false)
  GPS_TC_CASE(false, true)
  GPS_TC_CASE(true, false)
  GPS_TC_CASE(true, true)
#undef GPS_TC_CASE
  return GPS_ERR_UNSUPPORTED;
}

}  // namespace gps

// ------------------------------------------------------------------------------------ weight pre-packing
namespace gps {
namespace {
struct PrepackDesc {
  PrepackItem it[16];
  int start[17];   // CTA index range of each item (1-D grid: no empty CTAs)
  int n;
};
// K-major item: one CTA per (k-block, 8-row group) writes the 1024-byte swizzled group of both planes.
// MN-major item (W is [K x cols]): one CTA per (k-block, 64-column block x 8-k-row group).
__global__ void k_prepack_weights(PrepackDesc d) {
  int item = 0;
  while (item + 1 < d.n && (int)blockIdx.x >= d.start[item + 1]) ++item;
  const PrepackItem& it = d.it[item];
  const int local = (int)blockIdx.x - d.start[item];
  const int r = threadIdx.x >> 3, ck = threadIdx.x & 7;      // 64 threads: 8 rows x 8 sixteen-byte chunks
  const int nkb = (it.K + 63) / 64;
  float v[8];
  uint4 hi, lo;
  if (!it.mn) {
    const int groups = (it.rows + 256 + 7) / 8;
    const int kb = local / groups, grp = local - kb * groups;
    const int row = grp * 8 + r, k = kb * 64 + ck * 8;
    if (row < it.rows && k + 8 <= it.K) {
      const float4 x = ld4(it.W + (int64_t)row * it.ld + k), y = ld4(it.W + (int64_t)row * it.ld + k + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (row < it.rows && k + i < it.K) ? it.W[(int64_t)row * it.ld + k + i] : 0.f;
    }
    tc::split8(v, hi, lo);
    const int64_t plane = (int64_t)nkb * groups * 1024;
    uint8_t* base = reinterpret_cast<uint8_t*>(it.dst) + ((int64_t)kb * groups + grp) * 1024 + r * 128 + ((ck ^ r) << 4);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + plane) = lo;
  } else {
    // it.K = reduction extent (rows of W), it.rows = output columns of the GEMM (columns of W)
    const int cblocks = (it.rows + 256 + 63) / 64, nx = cblocks * 8;
    const int kb = local / nx, x = local - kb * nx;
    const int cb = x >> 3, kg = x & 7;
    const int k = kb * 64 + kg * 8 + r, col = cb * 64 + ck * 8;
    if (k < it.K && col + 8 <= it.rows) {
      const float4 x0 = ld4(it.W + (int64_t)k * it.ld + col), y0 = ld4(it.W + (int64_t)k * it.ld + col + 4);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = y0.x; v[5] = y0.y; v[6] = y0.z; v[7] = y0.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (k < it.K && col + i < it.rows) ? it.W[(int64_t)k * it.ld + col + i] : 0.f;
    }
    tc::split8(v, hi, lo);
    const int64_t plane = (int64_t)nkb * cblocks * 8192;
    uint8_t* base = reinterpret_cast<uint8_t*>(it.dst) + ((int64_t)kb * cblocks + cb) * 8192 + kg * 1024 + r * 128 +
                    ((ck ^ r) << 4);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + plane) = lo;
  }
}
}  // namespace

int prepack_groups(int rows) { return (rows + 256 + 7) / 8; }
int64_t prepack_plane_bytes(int rows, int K) { return (int64_t)((K + 63) / 64) * prepack_groups(rows) * 1024; }
int64_t prepack_bytes(int rows, int K) { return 2 * prepack_plane_bytes(rows, K); }
int prepack_groups_mn(int cols) { return (cols + 256 + 63) / 64; }
int64_t prepack_plane_bytes_mn(int cols, int K) { return (int64_t)((K + 63) / 64) * prepack_groups_mn(cols) * 8192; }
int64_t prepack_bytes_mn(int cols, int K) { return 2 * prepack_plane_bytes_mn(cols, K); }

int prepack_weights(const PrepackItem* items, int n, cudaStream_t stream) {
  if (n <= 0) return GPS_OK;
  GPS_REQUIRE(n <= 16, GPS_ERR_ARG, "prepack_weights: at most 16 matrices per call");
  PrepackDesc d;
  d.n = n;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    d.it[i] = items[i];
    GPS_REQUIRE(items[i].ld % 4 == 0 && (reinterpret_cast<uintptr_t>(items[i].W) & 15) == 0, GPS_ERR_ARG,
                "prepack_weights: weights must be 16-byte aligned with ld %% 4 == 0");
    d.start[i] = total;
    total += ((items[i].K + 63) / 64) * (items[i].mn ? prepack_groups_mn(items[i].rows) * 8 : prepack_groups(items[i].rows));
  }
  d.start[n] = total;
  if (total == 0) return GPS_OK;
  k_prepack_weights<<<(unsigned)total, 64, 0, stream>>>(d);
  GPS_LAUNCH_CHECK();
  return GPS_OK;
}
}  // namespace gps

// attention.cu — exact softmax attention over each graph's own node set, without densification.
//
// Replaces to_dense_batch -> nn.MultiheadAttention core -> [mask] of the reference
// (graphgps/layer/gps_layer.py:199-201, 234-241): the padded [B,Nmax,d] tensor, the key-padding
// mask and the boolean-mask gather (two host syncs per layer) disappear; every query row attends to
// the rows [graph_ptr[g], graph_ptr[g+1]) of its own graph directly in the packed [N, *] layout.
//
// Mapping (CUDA-core version; the flop share of this stage is <1% of the layer at the PCQM/ZINC
// shapes, 4*d*sum n_g^2 vs 24*N*d^2 — SURVEY.md 8d): query rows are packed densely into warps
// regardless of graph boundaries; LPR lanes cooperate on one row, each holding CH float4 chunks
// of the head dimension, so q/o (fwd) and k/v/dk/dv (bwd) live in registers and a dot product is
// an LPR-lane shuffle reduction.  Online softmax in fp32; dropout on the probabilities uses the
// Philox stream (site GPS_SITE_ATTN_P + head).
// Backward = two passes (query-major for dQ and delta, key-major for dK, dV): no atomics.
#include "kernels.cuh"

namespace gps {

namespace {

constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ int find_graph(const int* __restrict__ gptr, int B, int node) {
  // largest g with gptr[g] <= node  (graphs may be empty)
  int lo = 0, hi = B;  // invariant: gptr[lo] <= node < gptr[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (gptr[mid] <= node) lo = mid; else hi = mid;
  }
  return lo;
}

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Dropout on the attention probabilities: one Philox call yields the keep decisions of 4 consecutive keys
// (query i, keys 4*(jl>>2) .. +3) of one head; callers cache it across those 4 loop iterations.
struct DropQuad {
  Philox4 r;
  uint32_t thr;
  float keep_scale;
};
__device__ __forceinline__ void drop_quad_refresh(DropQuad& q, float p, uint64_t seed, uint64_t offset, int head, int i, int jl) {
  q.r = philox4x32(seed, offset + (uint64_t)(GPS_SITE_ATTN_P + head), ((uint64_t)i << 20) | (uint64_t)(jl >> 2));
}
__device__ __forceinline__ float drop_quad_scale(const DropQuad& q, int jl) {
  const int k = jl & 3;
  const uint32_t bits = k == 0 ? q.r.v[0] : k == 1 ? q.r.v[1] : k == 2 ? q.r.v[2] : q.r.v[3];
  return bits >= q.thr ? q.keep_scale : 0.f;
}
__device__ __forceinline__ void drop_quad_init(DropQuad& q, float p) {
  q.thr = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
  q.keep_scale = 1.f / (1.f - p);
}

// lane-slice helpers: lane `sub` of a row group owns float4 chunks sub, sub+LPR, ... (< nch)
template <int CH, int LPR>
__device__ __forceinline__ void load_slice(float4* dst, const float* row, int sub, int nch, bool ok) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    int ch = sub + c * LPR;
    dst[c] = (ok && ch < nch) ? ld4(row + ch * 4) : f4zero();
  }
}
template <int CH, int LPR>
__device__ __forceinline__ void store_slice(const float4* src, float* row, int sub, int nch) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    int ch = sub + c * LPR;
    if (ch < nch) st4(row + ch * 4, src[c]);
  }
}
// the same slice into the bf16 hi/lo planes of the tensor (operand image of the next GEMM), columns col0 ...
template <int CH, int LPR>
__device__ __forceinline__ void store_slice_planes(const float4* src, const Planes& p, int64_t r, int64_t col0, int sub, int nch) {
  if (!p.hi) return;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    int ch = sub + c * LPR;
    if (ch < nch) planes_store4(p, r, col0 + ch * 4, src[c]);
  }
}
template <int CH>
__device__ __forceinline__ float dot_slice(const float4* a, const float4* b) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += a[c].x * b[c].x + a[c].y * b[c].y + a[c].z * b[c].z + a[c].w * b[c].w;
  return s;
}

struct AttnArgs {
  const int* gptr; int B; int N; int H; int hd;
  const float* Q; const float* K; const float* V; int64_t ld;
  float* O; const float* Oc; const float* dO; int64_t ldo;
  float* lse; const float* lsec; float* delta; const float* deltac;
  float* dQ; float* dK; float* dV; int64_t ldg;
  float scale; float p_drop; uint64_t seed, offset;
  const unsigned long long* offset_dev;
  int role;   // backward: -1 both passes in one grid (blockIdx.z), 0 key-major pass, 1 query-major pass
  Planes Op, dQp, dKp, dVp;   // optional bf16 hi/lo plane copies of O (forward) / dQ, dK, dV (backward)
  int smem_rows;              // rows of the two per-block staging tiles (0: no staging)
};
__device__ __forceinline__ uint64_t eff_offset(const AttnArgs& a) {
  return a.offset + ((a.p_drop > 0.f && a.offset_dev) ? *a.offset_dev : 0ull);
}

// Shared-memory staging for batches of small graphs.  The 32 rows of a block belong to a few consecutive graphs, so
// every row it walks (keys for the forward / query-major pass, queries for the key-major pass) lies in the contiguous
// node range [start of the first row's graph, end of the last row's graph).  When that range fits (smem_rows), the
// block copies head h of the two tensors it walks into shared memory once (coalesced 128-bit loads) and the
// per-key loop reads shared memory instead of chasing L2 latency on every iteration (measured at the PCQM4M shape:
// the loop was ~600 cycles per key).  Larger ranges (ogbg-code2) keep the global-memory path.
struct StagedRows {
  const float* x; int64_t ldx;   // row r of tensor X at x + r * ldx (head offset included)
  const float* y; int64_t ldy;
};
template <int RPB>
__device__ __forceinline__ StagedRows stage_rows(const AttnArgs& a, const float* X, int64_t ldX, const float* Y, int64_t ldY,
                                                 int h, float* sm) {
  const int64_t hoff = (int64_t)h * a.hd;
  StagedRows r{X + hoff, ldX, Y + hoff, ldY};
  const int r0 = blockIdx.x * RPB;
  if (a.smem_rows <= 0 || r0 >= a.N) return r;
  const int last = min(r0 + RPB, a.N) - 1;
  const int kmin = a.gptr[find_graph(a.gptr, a.B, r0)];
  const int R = a.gptr[find_graph(a.gptr, a.B, last) + 1] - kmin;
  if (R > a.smem_rows) return r;                      // block-uniform
  const int pitch = a.hd + 4, nch = a.hd >> 2;
  float* sx = sm;
  float* sy = sm + (int64_t)a.smem_rows * pitch;
  for (int idx = threadIdx.x; idx < R * nch; idx += blockDim.x) {
    const int row = idx / nch, c = idx - row * nch;
    st4(sx + row * pitch + c * 4, ld4(X + (int64_t)(kmin + row) * ldX + hoff + c * 4));
    st4(sy + row * pitch + c * 4, ld4(Y + (int64_t)(kmin + row) * ldY + hoff + c * 4));
  }
  __syncthreads();
  r.x = sx - (int64_t)kmin * pitch; r.ldx = pitch;
  r.y = sy - (int64_t)kmin * pitch; r.ldy = pitch;
  return r;
}

template <int CH, int LPR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_attn_fwd(AttnArgs a) {
  extern __shared__ float attn_sm[];
  constexpr int RPW = 32 / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPR, rloc = lane / LPR;
  const int h = blockIdx.y;
  const int i = (blockIdx.x * kWarpsPerBlock + warp) * RPW + rloc;
  const bool row_ok = i < a.N;
  const int nch = a.hd / 4;
  const uint64_t offs = eff_offset(a);
  int gs = 0, n = 0;
  if (row_ok) {
    int g = find_graph(a.gptr, a.B, i);
    gs = a.gptr[g];
    n = a.gptr[g + 1] - gs;
  }
  const int nloop = warp_max_i(n);
  const int64_t hoff = (int64_t)h * a.hd;
  float4 q[CH], o[CH];
  load_slice<CH, LPR>(q, a.Q + (int64_t)(row_ok ? i : 0) * a.ld + hoff, sub, nch, row_ok);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    q[c] = f4scale(q[c], a.scale);
    o[c] = f4zero();
  }
  float m = -INFINITY, l = 0.f;
  const bool use_drop = a.p_drop > 0.f;
  DropQuad dq;
  drop_quad_init(dq, a.p_drop);
  // software pipeline: the K/V rows of key jl+1 are in flight while key jl is processed (the loop is a chain of
  // load -> dot -> shuffle -> exp -> fma, i.e. latency bound at these tiny graph sizes)
  const StagedRows kv = stage_rows<RPW * kWarpsPerBlock>(a, a.K, a.ld, a.V, a.ld, h, attn_sm);
  float4 kc[CH], vc[CH];
  load_slice<CH, LPR>(kc, kv.x + (int64_t)gs * kv.ldx, sub, nch, n > 0);
  load_slice<CH, LPR>(vc, kv.y + (int64_t)gs * kv.ldy, sub, nch, n > 0);
  for (int jl = 0; jl < nloop; ++jl) {
    if (use_drop && (jl & 3) == 0) drop_quad_refresh(dq, a.p_drop, a.seed, offs, h, i, jl);
    const bool valid = jl < n;
    const bool nvalid = jl + 1 < n;
    const int jn = gs + (nvalid ? jl + 1 : 0);
    float4 kn[CH], vn[CH];
    load_slice<CH, LPR>(kn, kv.x + (int64_t)jn * kv.ldx, sub, nch, nvalid);
    load_slice<CH, LPR>(vn, kv.y + (int64_t)jn * kv.ldy, sub, nch, nvalid);
    float s = group_sum<LPR>(dot_slice<CH>(q, kc));
    s = valid ? s : -INFINITY;
    const float m_new = fmaxf(m, s);
    const float corr = (m_new == -INFINITY) ? 1.f : __expf(m - m_new);
    const float p = valid ? __expf(s - m_new) : 0.f;
    l = l * corr + p;
    const float pd = use_drop ? p * drop_quad_scale(dq, jl) : p;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      o[c].x = o[c].x * corr + pd * vc[c].x;
      o[c].y = o[c].y * corr + pd * vc[c].y;
      o[c].z = o[c].z * corr + pd * vc[c].z;
      o[c].w = o[c].w * corr + pd * vc[c].w;
      kc[c] = kn[c];
      vc[c] = vn[c];
    }
    m = m_new;
  }
  if (row_ok) {
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < CH; ++c) o[c] = f4scale(o[c], inv);
    store_slice<CH, LPR>(o, a.O + (int64_t)i * a.ldo + hoff, sub, nch);
    store_slice_planes<CH, LPR>(o, a.Op, i, hoff, sub, nch);
    if (sub == 0) a.lse[(int64_t)i * a.H + h] = m + __logf(l);
  }
}

// delta_i = dO_i . O_i per (row, head): 8 lanes per pair read consecutive float4 chunks (128 B per group)
__global__ void k_attn_delta(AttnArgs a) {
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool ok = t < (int64_t)a.N * a.H;
  const int64_t i = ok ? t / a.H : 0;
  const int h = ok ? (int)(t - i * a.H) : 0;
  const float4* go = reinterpret_cast<const float4*>(a.dO + i * a.ldo + (int64_t)h * a.hd);
  const float4* oo = reinterpret_cast<const float4*>(a.Oc + i * a.ldo + (int64_t)h * a.hd);
  float acc = 0.f;
  if (ok)
    for (int c = sub; c < a.hd / 4; c += 8) {
      const float4 x = __ldg(go + c), y = __ldg(oo + c);
      acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (ok && sub == 0) a.delta[t] = acc;
}

// query-major backward: dQ_i (delta precomputed by k_attn_delta)
template <int CH, int LPR>
__device__ __forceinline__ void attn_bwd_q_body(const AttnArgs& a, float* attn_sm) {
  constexpr int RPW = 32 / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPR, rloc = lane / LPR;
  const int h = blockIdx.y;
  const int i = (blockIdx.x * kWarpsPerBlock + warp) * RPW + rloc;
  const bool row_ok = i < a.N;
  const int nch = a.hd / 4;
  const uint64_t offs = eff_offset(a);
  int gs = 0, n = 0;
  if (row_ok) {
    int g = find_graph(a.gptr, a.B, i);
    gs = a.gptr[g];
    n = a.gptr[g + 1] - gs;
  }
  const int nloop = warp_max_i(n);
  const int64_t hoff = (int64_t)h * a.hd;
  const int ir = row_ok ? i : 0;
  float4 q[CH], go[CH], gq[CH];
  load_slice<CH, LPR>(q, a.Q + (int64_t)ir * a.ld + hoff, sub, nch, row_ok);
  load_slice<CH, LPR>(go, a.dO + (int64_t)ir * a.ldo + hoff, sub, nch, row_ok);
  {
    const float dl = row_ok ? a.deltac[(int64_t)i * a.H + h] : 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) gq[c] = f4zero();
    const float lse = row_ok ? a.lsec[(int64_t)i * a.H + h] : 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) q[c] = f4scale(q[c], a.scale);
    const bool use_drop = a.p_drop > 0.f;
    DropQuad dq;
    drop_quad_init(dq, a.p_drop);
    const StagedRows kv = stage_rows<RPW * kWarpsPerBlock>(a, a.K, a.ld, a.V, a.ld, h, attn_sm);
    float4 kk[CH], vv[CH];
    load_slice<CH, LPR>(kk, kv.x + (int64_t)gs * kv.ldx, sub, nch, n > 0);
    load_slice<CH, LPR>(vv, kv.y + (int64_t)gs * kv.ldy, sub, nch, n > 0);
    for (int jl = 0; jl < nloop; ++jl) {
      if (use_drop && (jl & 3) == 0) drop_quad_refresh(dq, a.p_drop, a.seed, offs, h, i, jl);
      const bool valid = jl < n;
      const bool nvalid = jl + 1 < n;
      const int jn = gs + (nvalid ? jl + 1 : 0);
      float4 kn[CH], vn[CH];
      load_slice<CH, LPR>(kn, kv.x + (int64_t)jn * kv.ldx, sub, nch, nvalid);
      load_slice<CH, LPR>(vn, kv.y + (int64_t)jn * kv.ldy, sub, nch, nvalid);
      float s = dot_slice<CH>(q, kk), dp = dot_slice<CH>(go, vv);
#pragma unroll
      for (int ofs = LPR / 2; ofs > 0; ofs >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, ofs);
        dp += __shfl_xor_sync(0xffffffffu, dp, ofs);
      }
      const float p = valid ? __expf(s - lse) : 0.f;
      const float ds = p * (dp * (use_drop ? drop_quad_scale(dq, jl) : 1.f) - dl);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        gq[c].x += ds * kk[c].x;
        gq[c].y += ds * kk[c].y;
        gq[c].z += ds * kk[c].z;
        gq[c].w += ds * kk[c].w;
        kk[c] = kn[c];
        vv[c] = vn[c];
      }
    }
  }
  if (row_ok) {
#pragma unroll
    for (int c = 0; c < CH; ++c) gq[c] = f4scale(gq[c], a.scale);
    store_slice<CH, LPR>(gq, a.dQ + (int64_t)i * a.ldg + hoff, sub, nch);
    store_slice_planes<CH, LPR>(gq, a.dQp, i, hoff, sub, nch);
  }
}

// key-major backward: dK_j, dV_j
template <int CH, int LPR>
__device__ __forceinline__ void attn_bwd_kv_body(const AttnArgs& a, float* attn_sm) {
  constexpr int RPW = 32 / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPR, rloc = lane / LPR;
  const int h = blockIdx.y;
  const int j = (blockIdx.x * kWarpsPerBlock + warp) * RPW + rloc;
  const bool row_ok = j < a.N;
  const int nch = a.hd / 4;
  const uint64_t offs = eff_offset(a);
  int gs = 0, n = 0;
  if (row_ok) {
    int g = find_graph(a.gptr, a.B, j);
    gs = a.gptr[g];
    n = a.gptr[g + 1] - gs;
  }
  const int nloop = warp_max_i(n);
  const int jl = j - gs;
  const int64_t hoff = (int64_t)h * a.hd;
  const int jr = row_ok ? j : 0;
  float4 kk[CH], vv[CH], gk[CH], gv[CH];
  load_slice<CH, LPR>(kk, a.K + (int64_t)jr * a.ld + hoff, sub, nch, row_ok);
  load_slice<CH, LPR>(vv, a.V + (int64_t)jr * a.ld + hoff, sub, nch, row_ok);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    gk[c] = f4zero();
    gv[c] = f4zero();
  }
  const StagedRows qd = stage_rows<RPW * kWarpsPerBlock>(a, a.Q, a.ld, a.dO, a.ldo, h, attn_sm);
  float4 q[CH], go[CH];
  load_slice<CH, LPR>(q, qd.x + (int64_t)gs * qd.ldx, sub, nch, n > 0);
  load_slice<CH, LPR>(go, qd.y + (int64_t)gs * qd.ldy, sub, nch, n > 0);
  for (int il = 0; il < nloop; ++il) {
    const bool valid = il < n;
    const int i = gs + (valid ? il : 0);
    const bool nvalid = il + 1 < n;
    const int in_ = gs + (nvalid ? il + 1 : 0);
    float4 qn[CH], gon[CH];
    load_slice<CH, LPR>(qn, qd.x + (int64_t)in_ * qd.ldx, sub, nch, nvalid);
    load_slice<CH, LPR>(gon, qd.y + (int64_t)in_ * qd.ldy, sub, nch, nvalid);
    float s = dot_slice<CH>(q, kk), dp = dot_slice<CH>(go, vv);
#pragma unroll
    for (int ofs = LPR / 2; ofs > 0; ofs >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, ofs);
      dp += __shfl_xor_sync(0xffffffffu, dp, ofs);
    }
    const float lse = valid ? a.lsec[(int64_t)i * a.H + h] : 0.f;
    const float dl = valid ? a.deltac[(int64_t)i * a.H + h] : 0.f;
    const float p = valid ? __expf(s * a.scale - lse) : 0.f;
    float dsc = 1.f;
    if (a.p_drop > 0.f) {
      DropQuad dq;
      drop_quad_init(dq, a.p_drop);
      drop_quad_refresh(dq, a.p_drop, a.seed, offs, h, i, jl);
      dsc = drop_quad_scale(dq, jl);
    }
    const float pd = p * dsc;
    const float ds = p * (dp * dsc - dl) * a.scale;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      gk[c].x += ds * q[c].x; gk[c].y += ds * q[c].y; gk[c].z += ds * q[c].z; gk[c].w += ds * q[c].w;
      gv[c].x += pd * go[c].x; gv[c].y += pd * go[c].y; gv[c].z += pd * go[c].z; gv[c].w += pd * go[c].w;
      q[c] = qn[c];
      go[c] = gon[c];
    }
  }
  if (row_ok) {
    store_slice<CH, LPR>(gk, a.dK + (int64_t)j * a.ldg + hoff, sub, nch);
    store_slice<CH, LPR>(gv, a.dV + (int64_t)j * a.ldg + hoff, sub, nch);
    store_slice_planes<CH, LPR>(gk, a.dKp, j, hoff, sub, nch);
    store_slice_planes<CH, LPR>(gv, a.dVp, j, hoff, sub, nch);
  }
}

// both backward passes in one grid (blockIdx.z picks the role) so they share the SMs instead of queueing
template <int CH, int LPR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_attn_bwd(AttnArgs a) {
  extern __shared__ float attn_sm[];
  const int role = a.role >= 0 ? a.role : (int)blockIdx.z;
  if (role == 0) attn_bwd_kv_body<CH, LPR>(a, attn_sm);
  else attn_bwd_q_body<CH, LPR>(a, attn_sm);
}

enum { KFWD = 0, KBWD = 1 };

constexpr int kStageBytes = 72 * 1024;   // two staging tiles per block; 3 blocks per SM still fit

template <int CH, int LPR>
static void launch_one(int which, const AttnArgs& a0, cudaStream_t stream) {
  constexpr int RPW = 32 / LPR;
  AttnArgs a = a0;
  static const bool stage_on = [] {
    const char* e = getenv("GPS_B200_ATTN_STAGE");   // off by default: the staged kernels are faster in isolation but
    return e && e[0] == '1';                          // their 72 KB blocks crowd the GEMMs running next to them
                                                      // (same-box A/B at C3: 0.4983 vs 0.4856 ms/step)
  }();
  // staging pays when graphs are small (a block's 32 rows then see few distinct key rows); with large graphs every
  // block would exceed the tile anyway
  const int pitch = a.hd + 4;
  a.smem_rows = (stage_on && a.B > 0 && a.N < 64LL * a.B) ? kStageBytes / (2 * pitch * 4) : 0;
  const size_t smem = a.smem_rows > 0 ? (size_t)2 * a.smem_rows * pitch * 4 : 0;
  static bool attr_done[2] = {false, false};
  if (smem > 0 && !attr_done[which]) {
    if (which == KFWD) cudaFuncSetAttribute(k_attn_fwd<CH, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);
    else cudaFuncSetAttribute(k_attn_bwd<CH, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);
    attr_done[which] = true;
  }
  dim3 grid((unsigned)ceil_div(a.N, (int64_t)RPW * kWarpsPerBlock), (unsigned)a.H, (which == KFWD || a.role >= 0) ? 1 : 2);
  dim3 block(kWarpsPerBlock * 32);
  if (which == KFWD) k_attn_fwd<CH, LPR><<<grid, block, smem, stream>>>(a);
  else k_attn_bwd<CH, LPR><<<grid, block, smem, stream>>>(a);
}

static int dispatch(int which, const AttnArgs& a, cudaStream_t stream) {
  GPS_REQUIRE(a.hd > 0 && a.hd % 4 == 0, GPS_ERR_UNSUPPORTED, "attention: head dim %d must be a multiple of 4", a.hd);
  GPS_REQUIRE(a.ld % 4 == 0 && a.ldo % 4 == 0 && (which == KFWD || a.ldg % 4 == 0), GPS_ERR_UNSUPPORTED,
              "attention: leading dimensions must be multiples of 4");
  const int nch = a.hd / 4;
  GPS_REQUIRE(nch <= 48, GPS_ERR_UNSUPPORTED, "attention: head dim %d > 192 not supported", a.hd);
  if (a.N == 0) return GPS_OK;
  // smallest power-of-two lane group with <= 6 float4 chunks per lane
  int lpr = 1;
  while ((nch + lpr - 1) / lpr > 6) lpr *= 2;
  const int ch = (nch + lpr - 1) / lpr;
#define GPS_ATTN_CASE(CHV, LPRV)                                  \
  if (ch == CHV && lpr == LPRV) {                                 \
    launch_one<CHV, LPRV>(which, a, stream);                      \
    GPS_LAUNCH_CHECK();                                           \
    return GPS_OK;                                                \
  }
  GPS_ATTN_CASE(1, 1) GPS_ATTN_CASE(2, 1) GPS_ATTN_CASE(3, 1) GPS_ATTN_CASE(4, 1) GPS_ATTN_CASE(5, 1)
  GPS_ATTN_CASE(6, 1) GPS_ATTN_CASE(4, 2) GPS_ATTN_CASE(5, 2) GPS_ATTN_CASE(6, 2) GPS_ATTN_CASE(4, 4)
  GPS_ATTN_CASE(5, 4) GPS_ATTN_CASE(6, 4) GPS_ATTN_CASE(4, 8) GPS_ATTN_CASE(5, 8) GPS_ATTN_CASE(6, 8)
#undef GPS_ATTN_CASE
  set_error("attention: no kernel for head dim %d", a.hd);
  return GPS_ERR_UNSUPPORTED;
}

}  // namespace

int attention_fwd(const GpsGraph& g, int64_t heads, int64_t hd, const float* Q, const float* K, const float* V,
                  int64_t ld, float* O, int64_t ldo, float* lse, float p_drop, uint64_t seed, uint64_t offset,
                  cudaStream_t stream, const unsigned long long* offset_dev, Planes Op) {
  AttnArgs a{};
  a.offset_dev = offset_dev;
  a.Op = Op;
  a.gptr = g.graph_ptr; a.B = (int)g.B; a.N = (int)g.N; a.H = (int)heads; a.hd = (int)hd;
  a.Q = Q; a.K = K; a.V = V; a.ld = ld; a.O = O; a.ldo = ldo; a.lse = lse;
  a.scale = 1.f / sqrtf((float)hd); a.p_drop = p_drop; a.seed = seed; a.offset = offset;
  a.role = -1;
  return dispatch(KFWD, a, stream);
}

int attention_bwd(const GpsGraph& g, int64_t heads, int64_t hd, const float* Q, const float* K, const float* V,
                  int64_t ld, const float* O, const float* dO, int64_t ldo, const float* lse, float* delta,
                  float* dQ, float* dK, float* dV, int64_t ldg, float p_drop, uint64_t seed, uint64_t offset,
                  cudaStream_t stream, const unsigned long long* offset_dev, Planes dQp, Planes dKp, Planes dVp) {
  AttnArgs a{};
  a.offset_dev = offset_dev;
  a.dQp = dQp; a.dKp = dKp; a.dVp = dVp;
  a.gptr = g.graph_ptr; a.B = (int)g.B; a.N = (int)g.N; a.H = (int)heads; a.hd = (int)hd;
  a.Q = Q; a.K = K; a.V = V; a.ld = ld; a.Oc = O; a.dO = dO; a.ldo = ldo; a.lsec = lse;
  a.delta = delta; a.deltac = delta; a.dQ = dQ; a.dK = dK; a.dV = dV; a.ldg = ldg;
  a.scale = 1.f / sqrtf((float)hd); a.p_drop = p_drop; a.seed = seed; a.offset = offset;
  if (a.N == 0) return GPS_OK;
  const int64_t nt = (int64_t)a.N * a.H * 8;
  k_attn_delta<<<(unsigned)ceil_div(nt, (int64_t)256), 256, 0, stream>>>(a);
  GPS_LAUNCH_CHECK();
  static const bool merged = [] {
    const char* e = getenv("GPS_B200_OPT");
    return !e || (atoi(e) & 2);
  }();
  if (merged) {
    a.role = -1;
    return dispatch(KBWD, a, stream);
  }
  a.role = 1;
  GPS_TRY(dispatch(KBWD, a, stream));
  a.role = 0;
  return dispatch(KBWD, a, stream);
}

}  // namespace gps

// graph_prep.cu — per-mini-batch graph structure: dst-sorted CSR, src-sorted CSC, graph offsets.
//
// Replaces, once per batch instead of once per layer, the index handling of PyG
// MessagePassing.propagate (graphgps/layer/gatedgcn_layer.py:67-70), torch_scatter's atomic
// scatter (gatedgcn_layer.py:118-123) and to_dense_batch's bincount/cumsum (gps_layer.py:199).
// Segments are ordered by original edge id so every later segmented reduction is deterministic.
#include "common.cuh"

namespace gps {

__global__ void k_degree(const int64_t* __restrict__ edge_index, int64_t E, int* __restrict__ dst_ptr,
                         int* __restrict__ src_ptr) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= E) return;
  int s = (int)edge_index[k];
  int t = (int)edge_index[E + k];
  atomicAdd(&dst_ptr[t + 1], 1);
  atomicAdd(&src_ptr[s + 1], 1);
}

// In-place inclusive scan of a[1..n] (a[0] stays 0) by one 1024-thread CTA; blockIdx selects array.
__global__ void __launch_bounds__(1024) k_scan2(int* a0, int* a1, int64_t n) {
  int* a = blockIdx.x == 0 ? a0 : a1;
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 1; base <= n; base += 1024) {
    int64_t i = base + threadIdx.x;
    int v = i <= n ? a[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[wid] = s;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane];
      int ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += t;
      }
      warp_tot[lane] = ws - w;  // exclusive prefix of warp totals
    }
    __syncthreads();
    int carry = carry_s;
    int res = s + warp_tot[wid] + carry;
    if (i <= n) a[i] = res;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = res;
    __syncthreads();
  }
}

__global__ void k_fill(const int64_t* __restrict__ edge_index, int64_t E, const int* __restrict__ dst_ptr,
                       const int* __restrict__ src_ptr, int* __restrict__ cur_dst,
                       int* __restrict__ cur_src, int* __restrict__ dst_eid, int* __restrict__ src_eid) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= E) return;
  int s = (int)edge_index[k];
  int t = (int)edge_index[E + k];
  int p = atomicAdd(&cur_dst[t], 1);
  dst_eid[dst_ptr[t] + p] = (int)k;
  int q = atomicAdd(&cur_src[s], 1);
  src_eid[src_ptr[s] + q] = (int)k;
}

// One thread per (node, which): insertion-sort the segment's edge ids, then write the endpoint.
__global__ void k_sort_segments(const int64_t* __restrict__ edge_index, int64_t E, int64_t N,
                                const int* __restrict__ dst_ptr, const int* __restrict__ src_ptr,
                                int* __restrict__ dst_eid, int* __restrict__ src_eid,
                                int* __restrict__ dst_src, int* __restrict__ src_dst) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 2 * N) return;
  const bool by_dst = t < N;
  int64_t node = by_dst ? t : t - N;
  const int* ptr = by_dst ? dst_ptr : src_ptr;
  int* eid = by_dst ? dst_eid : src_eid;
  int* other = by_dst ? dst_src : src_dst;
  int b = ptr[node], e = ptr[node + 1];
  for (int i = b + 1; i < e; ++i) {
    int v = eid[i];
    int j = i - 1;
    while (j >= b && eid[j] > v) {
      eid[j + 1] = eid[j];
      --j;
    }
    eid[j + 1] = v;
  }
  for (int i = b; i < e; ++i) {
    int k = eid[i];
    other[i] = (int)(by_dst ? edge_index[k] : edge_index[E + k]);
  }
}

// graph_ptr[g] = first node of graph g, from the sorted batch vector; empty graphs get empty ranges.
__global__ void k_graph_ptr(const int64_t* __restrict__ batch, int64_t N, int64_t B, int* __restrict__ gptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0 && N == 0) {
    for (int64_t g = 0; g <= B; ++g) gptr[g] = 0;
    return;
  }
  if (i >= N) return;
  int64_t b = batch[i];
  int64_t prev = i > 0 ? batch[i - 1] : -1;
  for (int64_t g = prev + 1; g <= b && g <= B; ++g) gptr[g] = (int)i;
  if (i == N - 1)
    for (int64_t g = b + 1; g <= B; ++g) gptr[g] = (int)N;
}

struct GraphLayout {
  int64_t dst_ptr, dst_src, dst_eid, src_ptr, src_dst, src_eid, graph_ptr, cursors, total;
};
static GraphLayout graph_layout(int64_t N, int64_t E, int64_t B) {
  GraphLayout L;
  int64_t o = 0;
  auto take = [&](int64_t n) {
    int64_t r = o;
    o += round_up(n * 4, 256);
    return r;
  };
  L.dst_ptr = take(N + 1);
  L.src_ptr = take(N + 1);
  L.dst_src = take(E);
  L.dst_eid = take(E);
  L.src_dst = take(E);
  L.src_eid = take(E);
  L.graph_ptr = take(B + 1);
  L.cursors = take(2 * N);
  L.total = o;
  return L;
}

}  // namespace gps

using namespace gps;

extern "C" int64_t gps_graph_bytes(int64_t N, int64_t E, int64_t B) { return graph_layout(N, E, B).total; }

extern "C" int gps_graph_build(const int64_t* edge_index, const int64_t* batch, int64_t N, int64_t E,
                               int64_t B, void* storage, int64_t storage_bytes, GpsGraph* out,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GPS_REQUIRE(out && storage, GPS_ERR_ARG, "gps_graph_build: null argument");
  GPS_REQUIRE(N >= 0 && E >= 0 && B >= 0 && N < (1ll << 31) && E < (1ll << 31), GPS_ERR_ARG,
              "gps_graph_build: sizes out of range N=%lld E=%lld B=%lld", (long long)N, (long long)E,
              (long long)B);
  GraphLayout L = graph_layout(N, E, B);
  GPS_REQUIRE(storage_bytes >= L.total, GPS_ERR_ARG, "gps_graph_build: storage too small (%lld < %lld)",
              (long long)storage_bytes, (long long)L.total);
  char* base = (char*)storage;
  int* dst_ptr = (int*)(base + L.dst_ptr);
  int* src_ptr = (int*)(base + L.src_ptr);
  int* dst_src = (int*)(base + L.dst_src);
  int* dst_eid = (int*)(base + L.dst_eid);
  int* src_dst = (int*)(base + L.src_dst);
  int* src_eid = (int*)(base + L.src_eid);
  int* gptr = (int*)(base + L.graph_ptr);
  int* cur = (int*)(base + L.cursors);
  // dst_ptr and src_ptr are adjacent: one memset covers both
  GPS_CUDA(cudaMemsetAsync(dst_ptr, 0, (size_t)(L.dst_src - L.dst_ptr), stream));
  if (N > 0) GPS_CUDA(cudaMemsetAsync(cur, 0, (size_t)(2 * N * 4), stream));
  const int T = 256;
  if (E > 0) {
    k_degree<<<(unsigned)ceil_div(E, T), T, 0, stream>>>(edge_index, E, dst_ptr, src_ptr);
    GPS_LAUNCH_CHECK();
  }
  if (N > 0) {
    k_scan2<<<2, 1024, 0, stream>>>(dst_ptr, src_ptr, N);
    GPS_LAUNCH_CHECK();
  }
  if (E > 0) {
    k_fill<<<(unsigned)ceil_div(E, T), T, 0, stream>>>(edge_index, E, dst_ptr, src_ptr, cur, cur + N,
                                                         dst_eid, src_eid);
    GPS_LAUNCH_CHECK();
    k_sort_segments<<<(unsigned)ceil_div(2 * N, T), T, 0, stream>>>(edge_index, E, N, dst_ptr, src_ptr,
                                                                      dst_eid, src_eid, dst_src, src_dst);
    GPS_LAUNCH_CHECK();
  }
  k_graph_ptr<<<(unsigned)ceil_div(N > 0 ? N : 1, T), T, 0, stream>>>(batch, N, B, gptr);
  GPS_LAUNCH_CHECK();
  out->N = N;
  out->E = E;
  out->B = B;
  out->dst_ptr = dst_ptr;
  out->dst_src = dst_src;
  out->dst_eid = dst_eid;
  out->src_ptr = src_ptr;
  out->src_dst = src_dst;
  out->src_eid = src_eid;
  out->graph_ptr = gptr;
  return GPS_OK;
}

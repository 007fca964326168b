// RLlib state wire format of the sparse path: COO adjacency <-> fixed-size edge lists.
//
// Reference: util.pack_hidden / util.unpack_hidden (/root/reference/src/gcm/util.py:323-382), called by
// RaySparseGCM.forward around every SparseGCM call (ray_sparse_gcm.py:195-213): a Python loop over the batch with a
// nonzero() per graph one way, a nonzero() + three gathers + a COO rebuild the other way.  Here each direction is one
// pass: one CTA per graph finds its slice of the (coalesced, hence batch-sorted) COO by binary search and writes the
// graph's [2, max_edges] block -- edges first, fill after -- and the inverse compacts every graph's valid entries
// (source >= 0, util.py:367) in slot order with warp ballots.
#include "gcm_common.cuh"

namespace {

__device__ __forceinline__ long long lower_bound_ll(const int64_t* a, long long n, int64_t key) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// one CTA per graph b: edges of b are coo[:, lo..hi) (row 0 of coo = batch index, ascending)
__global__ void __launch_bounds__(256) k_pack_edges(const int64_t* __restrict__ coo, const float* __restrict__ vals,
                                                    long long E, int max_edges, long long edge_fill, float weight_fill,
                                                    int64_t* __restrict__ dense_edges, float* __restrict__ dense_weights,
                                                    int32_t* __restrict__ counts) {
  const int b = blockIdx.x;
  __shared__ long long s_lo, s_hi;
  if (threadIdx.x == 0) {
    s_lo = lower_bound_ll(coo, E, b);
    s_hi = lower_bound_ll(coo, E, (int64_t)b + 1);
    counts[b] = (int32_t)min(s_hi - s_lo, (long long)0x7fffffff);
  }
  __syncthreads();
  const long long lo = s_lo, cnt = s_hi - s_lo;
  int64_t* e0 = dense_edges + (size_t)b * 2 * max_edges;
  int64_t* e1 = e0 + max_edges;
  float* w = dense_weights + (size_t)b * max_edges;
  for (int s = threadIdx.x; s < max_edges; s += blockDim.x) {
    if (s < cnt) {
      e0[s] = coo[E + lo + s];
      e1[s] = coo[2 * E + lo + s];
      w[s] = vals[lo + s];
    } else {
      e0[s] = edge_fill;
      e1[s] = edge_fill;
      w[s] = weight_fill;
    }
  }
}

// counts[b] = number of slots with dense_edges[b, 0, s] >= 0
__global__ void __launch_bounds__(256) k_count_valid(const int64_t* __restrict__ dense_edges, int max_edges,
                                                     int64_t* __restrict__ counts) {
  const int b = blockIdx.x;
  const int64_t* e0 = dense_edges + (size_t)b * 2 * max_edges;
  int n = 0;
  for (int s = threadIdx.x; s < max_edges; s += blockDim.x) n += e0[s] >= 0 ? 1 : 0;
  n = gcm_warp_sum_int(n);
  __shared__ int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    counts[b] = t;
  }
}

// one WARP per graph: stable compaction of the valid slots to coo[:, offsets[b] ...)
__global__ void __launch_bounds__(256) k_unpack_edges(const int64_t* __restrict__ dense_edges,
                                                      const float* __restrict__ dense_weights, int B, int max_edges,
                                                      const int64_t* __restrict__ offsets, long long E,
                                                      int64_t* __restrict__ coo, float* __restrict__ vals) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const int64_t* e0 = dense_edges + (size_t)b * 2 * max_edges;
  const int64_t* e1 = e0 + max_edges;
  const float* w = dense_weights + (size_t)b * max_edges;
  long long out = offsets[b];
  for (int s0 = 0; s0 < max_edges; s0 += 32) {
    const int s = s0 + lane;
    const int64_t src = s < max_edges ? e0[s] : -1;
    const bool ok = src >= 0;
    const unsigned m = __ballot_sync(GCM_FULL_MASK, ok);
    if (ok) {
      const long long o = out + __popc(m & ((1u << lane) - 1u));
      coo[o] = b;
      coo[E + o] = src;
      coo[2 * E + o] = e1[s];
      vals[o] = w[s];
    }
    out += __popc(m);
  }
}

}  // namespace

extern "C" int gcm_pack_edges(const int64_t* coo, const float* vals, long long E, int B, int max_edges, long long edge_fill,
                              float weight_fill, int64_t* dense_edges, float* dense_weights, int32_t* counts,
                              void* stream) {
  GCM_REQUIRE(dense_edges && dense_weights && counts && B >= 0 && max_edges >= 0 && E >= 0 && (E == 0 || (coo && vals)),
              "pack_edges: bad arguments");
  if (B == 0) return GCM_OK;
  k_pack_edges<<<B, 256, 0, (cudaStream_t)stream>>>(coo, vals, E, max_edges, edge_fill, weight_fill, dense_edges,
                                                    dense_weights, counts);
  return gcm_check_launch("k_pack_edges");
}

extern "C" int gcm_count_valid_edges(const int64_t* dense_edges, int B, int max_edges, int64_t* counts, void* stream) {
  GCM_REQUIRE(dense_edges && counts && B >= 0 && max_edges >= 0, "count_valid_edges: bad arguments");
  if (B == 0) return GCM_OK;
  k_count_valid<<<B, 256, 0, (cudaStream_t)stream>>>(dense_edges, max_edges, counts);
  return gcm_check_launch("k_count_valid");
}

extern "C" int gcm_unpack_edges(const int64_t* dense_edges, const float* dense_weights, int B, int max_edges,
                                const int64_t* offsets, long long E, int64_t* coo, float* vals, void* stream) {
  GCM_REQUIRE(dense_edges && dense_weights && offsets && B >= 0 && max_edges >= 0 && E >= 0 && (E == 0 || (coo && vals)),
              "unpack_edges: bad arguments");
  if (B == 0 || E == 0) return GCM_OK;
  k_unpack_edges<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dense_edges, dense_weights, B, max_edges, offsets, E, coo,
                                                               vals);
  return gcm_check_launch("k_unpack_edges");
}


// ---- "Got NaN in returned memory" check of SparseGCM.forward (sparse_gcm.py:203: assert torch.all(torch.isfinite(mx))) ----
// one pass over the returned rows instead of torch's five elementwise / reduce kernels (1.0 ms of an 18.8 ms cfg5 call)
__global__ void __launch_bounds__(256) k_any_nonfinite(const float* __restrict__ x, long long n, int32_t* flag) {
  const long long n4 = n >> 2;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
    // finite <=> exponent field below 0xff
    bad |= ((__float_as_uint(v.x) & 0x7f800000u) == 0x7f800000u) | ((__float_as_uint(v.y) & 0x7f800000u) == 0x7f800000u) |
           ((__float_as_uint(v.z) & 0x7f800000u) == 0x7f800000u) | ((__float_as_uint(v.w) & 0x7f800000u) == 0x7f800000u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3))
    bad |= (__float_as_uint(x[(n4 << 2) + threadIdx.x]) & 0x7f800000u) == 0x7f800000u;
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

extern "C" int gcm_any_nonfinite(const float* x, long long n, int32_t* flag, void* stream) {
  GCM_REQUIRE(x && flag && n >= 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "any_nonfinite: bad arguments");
  if (n == 0) return GCM_OK;
  long long blocks = ((n >> 2) + 255) / 256;
  const long long cap = (long long)gcm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_any_nonfinite<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, flag);
  return gcm_check_launch("k_any_nonfinite");
}

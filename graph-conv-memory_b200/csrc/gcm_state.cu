// State conversion (log + bitmasks <-> the reference's dense hidden state), the cross-batch
// Euclidean distance, the selectors' own dense forward, and the library's error plumbing.
#include <stdarg.h>

#include "gcm_common.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local const char* g_last_kernel = "";
static long long g_launches = 0;

void gcm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gcm_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    gcm_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  g_last_kernel = what;
  ++g_launches;
  return GCM_OK;
}

int gcm_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

extern "C" int gcm_version(void) { return GCM_ABI_VERSION; }
extern "C" const char* gcm_last_error(void) { return g_err; }
extern "C" const char* gcm_last_kernel(void) { return g_last_kernel; }
extern "C" long long gcm_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------------
// materialize: log/bitmask state -> reference layout (gcm.py:194-211)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_materialize(const gcm_dense_state st, float* nodes_out,
                                                     float* adj_out, int64_t* nn_out) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = st.N, C = st.C, F = st.F, W = st.W;
  const int cnt = st.count[b];
  const int nvalid = min(cnt, N);
  const int start = cnt - nvalid;
  if (nn_out && tid == 0) nn_out[b] = nvalid;
  const float* nodes_b = st.nodes + (size_t)b * C * F;
  if (nodes_out) {
    float* out = nodes_out + (size_t)b * N * F;
    for (int idx = tid; idx < N * F; idx += blockDim.x) {
      const int l = idx / F, f = idx - l * F;
      out[idx] = nodes_b[(size_t)gcm_slot(start + l, C) * F + f];
    }
  }
  if (adj_out) {
    const uint32_t* masks_b = st.masks + (size_t)b * C * 2 * W;
    float* out = adj_out + (size_t)b * N * N;
    for (int idx = tid; idx < N * N; idx += blockDim.x) {
      const int l = idx / N, m = idx - l * N;
      float v = 0.0f;
      if (l < nvalid && m < nvalid) {
        const uint32_t* mrow = masks_b + (size_t)gcm_slot(start + l, C) * 2 * W;
        const int e = m <= l ? l - m : m - l;
        const uint32_t word = mrow[(m <= l ? 0 : W) + (e >> 5)];
        v = (float)((word >> (e & 31)) & 1u);
      }
      out[idx] = v;
    }
  }
}

__global__ void __launch_bounds__(256) k_materialize_grad(const gcm_dense_state st, const float* d_nodes,
                                                          float* out_all) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = st.N, C = st.C, F = st.F;
  const int cnt = st.count[b];
  const int nvalid = min(cnt, N);
  const int start = cnt - nvalid;
  const float* src = d_nodes + (size_t)b * C * F;
  float* out = out_all + (size_t)b * N * F;
  for (int idx = tid; idx < N * F; idx += blockDim.x) {
    const int l = idx / F, f = idx - l * F;
    out[idx] = l < nvalid ? src[(size_t)gcm_slot(start + l, C) * F + f] : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// ingest: reference layout -> log/bitmask state (positions == logical indices)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ingest(const gcm_dense_state st, const float* nodes_in,
                                                const float* adj_in, const int64_t* nn_in,
                                                int32_t* status) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int N = st.N, C = st.C, F = st.F, W = st.W;
  long long n64 = nn_in[b];
  unsigned int flags = 0u;
  if (n64 < 0 || n64 > N) {
    flags |= GCM_FLAG_BADCOUNT;
    n64 = n64 < 0 ? 0 : N;
  }
  const int n = (int)n64;
  if (tid == 0) st.count[b] = n;
  float* nodes_b = st.nodes + (size_t)b * C * F;
  const float* src = nodes_in + (size_t)b * N * F;
  if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(nodes_in) | reinterpret_cast<uintptr_t>(st.nodes)) & 15) == 0) {
    // slot l == row l; 4 independent 16-byte loads in flight per thread
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(nodes_b);
    const int n4 = N * F / 4, c4 = C * F / 4, bs = blockDim.x;
    int idx = tid;
    for (; idx + 3 * bs < n4; idx += 4 * bs) {
      const float4 v0 = __ldcs(s4 + idx), v1 = __ldcs(s4 + idx + bs), v2 = __ldcs(s4 + idx + 2 * bs),
                   v3 = __ldcs(s4 + idx + 3 * bs);
      d4[idx] = v0; d4[idx + bs] = v1; d4[idx + 2 * bs] = v2; d4[idx + 3 * bs] = v3;
    }
    for (; idx < n4; idx += bs) d4[idx] = __ldcs(s4 + idx);
    for (idx = n4 + tid; idx < c4; idx += bs) d4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int idx = tid; idx < N * F; idx += blockDim.x) nodes_b[idx] = src[idx];
    for (int idx = N * F + tid; idx < C * F; idx += blockDim.x) nodes_b[idx] = 0.0f;
  }

  uint32_t* masks_b = st.masks + (size_t)b * C * 2 * W;
  const float* adj_b = adj_in + (size_t)b * N * N;
  // one warp per row l: the row is read once with coalesced loads (8 in flight per lane) into a column bitmask
  // (lane j keeps word j: bit i <-> column 32 j + i; W <= 32), which is then re-indexed by offset:
  // past bit e <-> column l - e, future bit e <-> column l + e.
  for (int l = warp; l < C; l += nwarps) {
    uint32_t* row = masks_b + (size_t)l * 2 * W;
    if (l >= N) {
      for (int w = lane; w < 2 * W; w += 32) row[w] = 0u;
      continue;
    }
    const float* arow = adj_b + (size_t)l * N;
    uint32_t myword = 0u;
    for (int j0 = 0; j0 < W; j0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int m = (j0 + u) * 32 + lane;
        v[u] = (j0 + u < W && m < N) ? __ldcs(arow + m) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j < W) {   // warp-uniform
          const int m = j * 32 + lane;
          const bool inb = m < N;
          const bool set = v[u] == 1.0f;
          if (inb && ((v[u] != 0.0f && v[u] != 1.0f) || (set && (l >= n || m >= n)))) flags |= GCM_FLAG_UNCLEAN;
          if (inb && l < n && m < n && !set) flags |= GCM_FLAG_NOTDENSE;   // DenseEdge states are all ones here
          const uint32_t wrd = __ballot_sync(GCM_FULL_MASK, set && l < n && m < n);
          if (lane == j) myword = wrd;
        }
      }
    }
    for (int w = 0; w < W; ++w) {
      const int e = w * 32 + lane;
      const int mp = l - e;
      const uint32_t sp = __shfl_sync(GCM_FULL_MASK, myword, (mp >= 0 ? mp : 0) >> 5);
      const uint32_t pw = __ballot_sync(GCM_FULL_MASK, mp >= 0 && ((sp >> (mp & 31)) & 1u));
      const int mf = l + e;
      const bool okf = e >= 1 && mf < N;
      const uint32_t sf = __shfl_sync(GCM_FULL_MASK, myword, (okf ? mf : 0) >> 5);
      const uint32_t fw = __ballot_sync(GCM_FULL_MASK, okf && ((sf >> (mf & 31)) & 1u));
      if (lane == 0) {
        row[w] = pw;
        row[W + w] = fw;
      }
    }
  }
  if (flags) atomicOr(reinterpret_cast<unsigned int*>(status), flags);
  if (tid == 0) atomicMax(status + 1, n);
}

// ------------------------------------------------------------------------------------------------
// EuclideanEdge distance (distance.py:48-49): mean over ALL current observations
// ------------------------------------------------------------------------------------------------
constexpr int EU_TR = 64, EU_TP = 64, EU_FC = 32;

__global__ void __launch_bounds__(256) k_euclid_batchmean(const float* nodes, long long n_rows, int F,
                                                          const float* cur, int n_cur,
                                                          const float* dist_param, float* dist) {
  __shared__ float rs[EU_TR][EU_FC + 1];
  __shared__ float cs[EU_TP][EU_FC + 1];
  __shared__ float red[EU_TR][16 + 1];
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tp = tid & 15;  // 16 x 16 threads, each a 4 x 4 (row, cur) register tile
  const long long row0 = (long long)blockIdx.x * EU_TR;
  float rowsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p0 = 0; p0 < n_cur; p0 += EU_TP) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int f0 = 0; f0 < F; f0 += EU_FC) {
      for (int idx = tid; idx < EU_TR * EU_FC; idx += 256) {
        const int r = idx / EU_FC, f = idx - r * EU_FC;
        const long long gr = row0 + r;
        rs[r][f] = (gr < n_rows && f0 + f < F) ? nodes[gr * F + f0 + f] : 0.f;
        const int gp = p0 + r;
        cs[r][f] = (gp < n_cur && f0 + f < F) ? cur[(size_t)gp * F + f0 + f] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int f = 0; f < EU_FC; ++f) {
        float rv[4], cv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rv[i] = rs[tr * 4 + i][f];
#pragma unroll
        for (int j = 0; j < 4; ++j) cv[j] = cs[tp * 4 + j][f];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float df = rv[i] - cv[j];
            acc[i][j] = fmaf(df, df, acc[i][j]);
          }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + tp * 4 + j < n_cur) rowsum[i] += sqrtf(acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[tr * 4 + i][tp] = rowsum[i];
  __syncthreads();
  if (tid < EU_TR) {
    float s = 0.f;
    for (int j = 0; j < 16; ++j) s += red[tid][j];
    const long long gr = row0 + tid;
    const float scale = dist_param ? 1.0f / fabsf(*dist_param) : 1.0f;
    if (gr < n_rows) dist[gr] = s / (float)n_cur * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// the selectors' own forward on the reference's dense tensors
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_select_dense(const float* nodes, float* adj,
                                                      const int64_t* num_nodes, int N, int F,
                                                      const gcm_selector sel) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const long long t64 = num_nodes[b];
  if (t64 < 0 || t64 >= N) return;  // the reference would raise an index error here
  const int t = (int)t64;
  float* adj_b = adj + (size_t)b * N * N;
  const float* nodes_b = nodes + (size_t)b * N * F;
  if (sel.kind == GCM_SEL_TEMPORAL) {
    if (tid < sel.n_hops) {
      const int hop = sel.hops[tid];
      if (hop >= 0 && hop <= t) {
        if (sel.direction != GCM_DIR_BACKWARD) adj_b[(size_t)t * N + (t - hop)] = 1.0f;
        if (sel.direction != GCM_DIR_FORWARD) adj_b[(size_t)(t - hop) * N + t] = 1.0f;
      }
    }
  } else if (sel.kind == GCM_SEL_DENSE) {
    for (int j = tid; j <= t; j += blockDim.x) {
      adj_b[(size_t)t * N + j] = 1.0f;
      adj_b[(size_t)j * N + t] = 1.0f;
    }
  } else if (sel.kind != GCM_SEL_NONE) {
    const bool learned = sel.dist_param != nullptr;
    const float thr = learned ? 1.0f : sel.max_distance;
    const float scale = learned ? 1.0f / fabsf(*sel.dist_param) : 1.0f;
    const float* cur = nodes_b + (size_t)t * F;
    float cur_norm = 0.f;
    if (sel.kind == GCM_SEL_COSINE) {
      float s = 0.f;
      for (int f = lane; f < F; f += 32) s += cur[f] * cur[f];
      cur_norm = fmaxf(sqrtf(gcm_warp_sum(s)), 1e-8f);
    }
    for (int j = warp; j < t; j += nwarps) {
      const float* row = nodes_b + (size_t)j * F;
      float dist;
      if (sel.kind == GCM_SEL_EUCLIDEAN) {
        dist = sel.dist[(size_t)b * N + j];
      } else if (sel.kind == GCM_SEL_COSINE) {
        float dot = 0.f, nb = 0.f;
        for (int f = lane; f < F; f += 32) {
          const float v = row[f];
          dot += cur[f] * v;
          nb += v * v;
        }
        dot = gcm_warp_sum(dot);
        nb = fmaxf(sqrtf(gcm_warp_sum(nb)), 1e-8f);
        dist = dot / (cur_norm * nb);
      } else {
        float s = 0.f;
        for (int k = lane; k < sel.slice_len; k += 32) {
          const float df = cur[sel.a_start + k * sel.a_step] - row[sel.b_start + k * sel.b_step];
          s += df * df;
        }
        dist = sqrtf(gcm_warp_sum(s)) * scale;
      }
      if (lane == 0 && dist < thr) adj_b[(size_t)t * N + j] = 1.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static int check_state(const gcm_dense_state* st) {
  GCM_REQUIRE(st && st->nodes && st->masks && st->count, "state: null pointer");
  GCM_REQUIRE(st->N >= 1 && st->N <= GCM_MAX_N && st->C >= st->N && st->F >= 1 && st->B >= 0 &&
                  st->W == (st->N + 31) / 32,
              "state: bad dims B=%d N=%d C=%d F=%d W=%d", st->B, st->N, st->C, st->F, st->W);
  return GCM_OK;
}

extern "C" int gcm_state_materialize(const gcm_dense_state* st, float* nodes_out, float* adj_out,
                                     int64_t* num_nodes_out, void* stream) {
  if (int rc = check_state(st)) return rc;
  if (st->B == 0) return GCM_OK;
  k_materialize<<<st->B, 256, 0, (cudaStream_t)stream>>>(*st, nodes_out, adj_out, num_nodes_out);
  return gcm_check_launch("k_materialize");
}

// nodes[b, (count[b] + offset) % C, :] = obs[b, :]: the node write of gcm.py:274 alone, on any log that shares the
// counters of a state (the RAW observation log kept next to the preprocessed one when DenseGCM has a preprocessor,
// gcm.py:290-291).  offset = 0 before the step that advances the counters, -1 after it.
__global__ void __launch_bounds__(256) k_log_write(const gcm_dense_state st, const float* obs, int offset) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)st.B * st.F) return;
  const int b = (int)(i / st.F), f = (int)(i - (long long)b * st.F);
  st.nodes[((size_t)b * st.C + gcm_slot(__ldcg(st.count + b) + offset, st.C)) * st.F + f] = obs[i];
}
extern "C" int gcm_state_log_write(const gcm_dense_state* st, const float* obs, int offset, void* stream) {
  GCM_REQUIRE(st && st->nodes && st->count && obs && st->B >= 0 && st->F >= 1 && st->C >= 1 && (offset == 0 || offset == -1),
              "state_log_write: bad arguments");
  if (st->B == 0) return GCM_OK;
  const long long n = (long long)st->B * st->F;
  k_log_write<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*st, obs, offset);
  return gcm_check_launch("k_log_write");
}

// T node writes at once, AFTER the T steps that advanced the counters: nodes[b, (count[b] - T + k) % C, :] = x_seq[b, k, :]
// (graph b, step k at b * stride_b + k * stride_t floats).  The raw-observation log of a sequence call (ray_gcm.py:200-202
// with the Linear preprocessor of ray_gcm.py:118).
__global__ void __launch_bounds__(256) k_log_write_seq(const gcm_dense_state st, const float* x_seq, long long stride_b,
                                                       long long stride_t, int T, int k0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_b = (long long)(T - k0) * st.F;
  if (i >= (long long)st.B * per_b) return;
  const int b = (int)(i / per_b);
  const long long rem = i - (long long)b * per_b;
  const int k = k0 + (int)(rem / st.F), f = (int)(rem % st.F);
  const int pos = __ldcg(st.count + b) - T + k;
  st.nodes[((size_t)b * st.C + gcm_slot(pos, st.C)) * st.F + f] = x_seq[b * stride_b + k * stride_t + f];
}
// the same with 16-byte pieces and 32-bit index arithmetic: grid (chunks of B * F / 4, steps).  The element-wise kernel
// above spent its time in 64-bit divisions and 4-byte accesses (1.5 TB/s: 11.5 us per step of a cfg2-pre rollout)
__global__ void __launch_bounds__(256) k_log_write_seq4(const gcm_dense_state st, const float* x_seq, long long stride_b,
                                                        long long stride_t, int T, int k0) {
  const int F4 = st.F >> 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= st.B * F4) return;
  const int b = j / F4, c = j - b * F4;
  const int k = k0 + blockIdx.y;
  const int pos = __ldcg(st.count + b) - T + k;
  const float4 v = __ldcs(reinterpret_cast<const float4*>(x_seq + b * stride_b + k * stride_t) + c);
  reinterpret_cast<float4*>(st.nodes + ((size_t)b * st.C + gcm_slot(pos, st.C)) * st.F)[c] = v;
}
extern "C" int gcm_state_log_write_seq(const gcm_dense_state* st, const float* x_seq, long long stride_b,
                                       long long stride_t, int T, void* stream) {
  GCM_REQUIRE(st && st->nodes && st->count && x_seq && st->B >= 0 && st->F >= 1 && st->C >= 1 && T >= 0,
              "state_log_write_seq: bad arguments");
  if (st->B == 0 || T == 0) return GCM_OK;
  const int k0 = T > st->C ? T - st->C : 0;      // older rows would be overwritten by the later ones anyway
  if ((st->F & 3) == 0 && ((stride_b | stride_t) & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(x_seq) | reinterpret_cast<uintptr_t>(st->nodes)) & 15) == 0 && T - k0 <= 65535 &&
      (long long)st->B * (st->F >> 2) < (1ll << 31)) {
    const dim3 grid((unsigned)(((long long)st->B * (st->F >> 2) + 255) / 256), (unsigned)(T - k0));
    k_log_write_seq4<<<grid, 256, 0, (cudaStream_t)stream>>>(*st, x_seq, stride_b, stride_t, T, k0);
    return gcm_check_launch("k_log_write_seq");
  }
  const long long n = (long long)st->B * (T - k0) * st->F;
  k_log_write_seq<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*st, x_seq, stride_b, stride_t, T, k0);
  return gcm_check_launch("k_log_write_seq");
}

extern "C" int gcm_state_materialize_grad(const gcm_dense_state* st, const float* d_nodes,
                                          float* d_nodes_out, void* stream) {
  if (int rc = check_state(st)) return rc;
  GCM_REQUIRE(d_nodes && d_nodes_out, "materialize_grad: null pointer");
  if (st->B == 0) return GCM_OK;
  k_materialize_grad<<<st->B, 256, 0, (cudaStream_t)stream>>>(*st, d_nodes, d_nodes_out);
  return gcm_check_launch("k_materialize_grad");
}

extern "C" int gcm_state_ingest(const gcm_dense_state* st, const float* nodes_in, const float* adj_in,
                                const int64_t* num_nodes_in, int32_t* status, void* stream) {
  if (int rc = check_state(st)) return rc;
  GCM_REQUIRE(nodes_in && adj_in && num_nodes_in && status, "ingest: null pointer");
  if (st->B == 0) return GCM_OK;
  k_ingest<<<st->B, 256, 0, (cudaStream_t)stream>>>(*st, nodes_in, adj_in, num_nodes_in, status);
  return gcm_check_launch("k_ingest");
}

extern "C" int gcm_euclid_batchmean(const gcm_dense_state* st, const float* cur, int n_cur,
                                    const float* dist_param, float* dist, void* stream) {
  if (int rc = check_state(st)) return rc;
  GCM_REQUIRE(cur && dist && n_cur >= 1, "euclid_batchmean: bad arguments");
  const long long n_rows = (long long)st->B * st->C;
  if (n_rows == 0) return GCM_OK;
  const long long grid = (n_rows + EU_TR - 1) / EU_TR;
  GCM_REQUIRE(grid < 2147483647LL, "euclid_batchmean: too many rows");
  k_euclid_batchmean<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(st->nodes, n_rows, st->F, cur, n_cur,
                                                                      dist_param, dist);
  return gcm_check_launch("k_euclid_batchmean");
}

extern "C" int gcm_select_dense(const float* nodes, float* adj, const int64_t* num_nodes, int B, int N,
                                int F, const gcm_selector* sel, void* stream) {
  GCM_REQUIRE(adj && num_nodes && sel && B >= 0 && N >= 1 && F >= 1, "select_dense: bad arguments");
  GCM_REQUIRE(sel->kind >= GCM_SEL_NONE && sel->kind <= GCM_SEL_SPATIAL, "select_dense: bad kind %d", sel->kind);
  GCM_REQUIRE(sel->n_hops >= 0 && sel->n_hops <= GCM_MAX_HOPS, "select_dense: n_hops=%d", sel->n_hops);
  if (sel->kind >= GCM_SEL_EUCLIDEAN) GCM_REQUIRE(nodes, "select_dense: distance selector needs nodes");
  if (sel->kind == GCM_SEL_EUCLIDEAN) GCM_REQUIRE(sel->dist, "select_dense: euclidean needs dist");
  if (sel->kind == GCM_SEL_SPATIAL)
    GCM_REQUIRE(sel->slice_len >= 0 && sel->a_step >= 1 && sel->b_step >= 1 && sel->a_start >= 0 &&
                    sel->b_start >= 0 &&
                    (sel->slice_len == 0 || (sel->a_start + (sel->slice_len - 1) * sel->a_step < F &&
                                             sel->b_start + (sel->slice_len - 1) * sel->b_step < F)),
                "select_dense: spatial slice outside [0,F)");
  if (B == 0) return GCM_OK;
  k_select_dense<<<B, 256, 0, (cudaStream_t)stream>>>(nodes, adj, num_nodes, N, F, *sel);
  return gcm_check_launch("k_select_dense");
}

// Sparse GCM path (reference /root/reference/src/gcm/sparse_gcm.py:72-212):
//   * node write + flatten of the valid rows            (sparse_gcm.py:111-123, util.py:426-452)
//   * fused edge generation: TemporalEdge + SpatialRadiusEdge, emitted already coalesced
//     (sorted by (batch, sink, source), duplicates merged) -- no sort, no COO coalesce
//                                                        (sparse_edge_selectors/temporal.py:19-63,
//                                                         sparse_edge_selectors/spatial.py:74-115)
//   * GraphConv as a CSR segmented gather-reduce fused with the two Linear layers + activation
//     (torch_geometric.nn.GraphConv; invoked at sparse_gcm.py:178,199), deterministic, and its
//     backward (transposed-CSR gather for dL/dx, split-K accumulation for the weight gradients).
#include <stdlib.h>
#include <type_traits>

#include "gcm_common.cuh"

// ------------------------------------------------------------------------------------------------
// node write + flatten
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sparse_write_flatten(float* nodes, const float* x, const int64_t* T,
                                                              const int64_t* taus, const int64_t* offsets, int N,
                                                              int F, int tmax, float* flat) {
  const int b = blockIdx.x;
  const int t0 = (int)T[b], tau = (int)taus[b];
  float* nodes_b = nodes + (size_t)b * N * F;
  const float* x_b = x + (size_t)b * tmax * F;
  for (int i = threadIdx.x; i < tau * F; i += blockDim.x) {
    const int k = i / F, f = i - k * F;
    if (t0 + k < N) nodes_b[(size_t)(t0 + k) * F + f] = x_b[i];
  }
  __syncthreads();
  if (flat) {
    float* out = flat + offsets[b] * F;
    const int n = min(t0 + tau, N);
    for (int i = threadIdx.x; i < n * F; i += blockDim.x) out[i] = nodes_b[i];
  }
}

// Out-of-place form of the same step: nodes_out = nodes_in with the new rows written, flat = its valid rows, in ONE pass
// over the rows (the in-place form needs the caller's clone of `nodes` first and reads the written rows back: 3 passes
// over a [B,N,F] tensor, 1.33 + 0.34 ms at cfg5).  blockIdx.y splits a graph's rows; VEC = 16-byte pieces when F % 4 == 0.
constexpr int WF_ROWS = 128;
template <typename VT>
__global__ void __launch_bounds__(256) k_sparse_write_flatten_oop(const VT* nodes_in, VT* nodes_out, const VT* x,
                                                                  const int64_t* T, const int64_t* taus,
                                                                  const int64_t* offsets, int N, int Fv, int tmax, VT* flat) {
  const int b = blockIdx.x;
  const int t0 = (int)T[b], tau = (int)taus[b];
  const int j0 = blockIdx.y * WF_ROWS, j1 = min(N, j0 + WF_ROWS);
  const VT* in_b = nodes_in + (size_t)b * N * Fv;
  VT* out_b = nodes_out + (size_t)b * N * Fv;
  const VT* x_b = x + (size_t)b * tmax * Fv;
  VT* flat_b = flat ? flat + offsets[b] * Fv : nullptr;
  const int n_valid = min(t0 + tau, N);
  const int lo = j0 * Fv, hi = j1 * Fv, new_lo = t0 * Fv, new_hi = (t0 + tau) * Fv, valid_hi = n_valid * Fv;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const VT v = (i >= new_lo && i < new_hi) ? x_b[i - new_lo] : in_b[i];
    out_b[i] = v;
    if (flat_b && i < valid_hi) flat_b[i] = v;
  }
}

// Adjoint of the step above, one pass: d_x[b, k] = d_nodes_out[b, T_b + k] + d_flat[off_b + T_b + k] (rows that came from
// x), d_nodes_in[b, j] = d_nodes_out[b, j] + d_flat[off_b + j] on the other rows and 0 on the rows x overwrote.  Any of
// d_nodes_out / d_flat / d_nodes_in / d_x may be NULL (treated as zero / not wanted).
template <typename VT>
__device__ __forceinline__ VT wf_add(const VT& a, const VT& b);
template <>
__device__ __forceinline__ float wf_add<float>(const float& a, const float& b) { return a + b; }
template <>
__device__ __forceinline__ float4 wf_add<float4>(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
template <typename VT>
__device__ __forceinline__ VT wf_zero();
template <>
__device__ __forceinline__ float wf_zero<float>() { return 0.0f; }
template <>
__device__ __forceinline__ float4 wf_zero<float4>() { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }

template <typename VT>
__global__ void __launch_bounds__(256) k_sparse_write_flatten_bwd(const VT* d_nodes_out, const VT* d_flat, const int64_t* T,
                                                                  const int64_t* taus, const int64_t* offsets, int N, int Fv,
                                                                  int tmax, VT* d_nodes_in, VT* d_x) {
  const int b = blockIdx.x;
  const int t0 = (int)T[b], tau = (int)taus[b];
  const int rows = max(N, tmax);
  const int j0 = blockIdx.y * WF_ROWS, j1 = min(rows, j0 + WF_ROWS);
  const VT* dn_b = d_nodes_out ? d_nodes_out + (size_t)b * N * Fv : nullptr;
  const VT* df_b = d_flat ? d_flat + offsets[b] * Fv : nullptr;
  VT* di_b = d_nodes_in ? d_nodes_in + (size_t)b * N * Fv : nullptr;
  VT* dx_b = d_x ? d_x + (size_t)b * tmax * Fv : nullptr;
  const int n_valid = min(t0 + tau, N);
  for (int i = j0 * Fv + threadIdx.x; i < j1 * Fv; i += 256) {
    const int j = i / Fv;
    if (j < N && (di_b || dx_b)) {
      // gradient that arrives at node row j of nodes_out
      VT gsum = dn_b ? dn_b[i] : wf_zero<VT>();
      if (df_b && j < n_valid) gsum = wf_add<VT>(gsum, df_b[i]);
      const bool is_new = j >= t0 && j < t0 + tau;
      if (di_b) di_b[i] = is_new ? wf_zero<VT>() : gsum;
      if (dx_b && is_new) dx_b[i - t0 * Fv] = gsum;
    }
    // rows of x that were never written (k >= tau, or T_b + k >= N): zero gradient
    if (dx_b && j < tmax && (j >= tau || t0 + j >= N)) dx_b[i] = wf_zero<VT>();
  }
}

// ------------------------------------------------------------------------------------------------
// fused edge generation
// ------------------------------------------------------------------------------------------------
constexpr int EG_SINKS = 128;  // new nodes handled per CTA
struct EdgeGenArgs {
  const float* nodes;
  const int64_t* T;
  const int64_t* taus;
  const int64_t* new_off;  // [B+1] exclusive cumsum of taus
  int N, F;
  int n_hops;
  int hops[GCM_MAX_HOPS];
  int use_radius, pos_start, pos_step, pos_len;
  float radius;
  int32_t* deg;             // pass 1: [n_new]
  const int64_t* edge_off;  // pass 2: [n_new + 1] exclusive cumsum of deg
  int64_t* edges;           // pass 2: [3, E]
  int64_t E;
  const int64_t* flat_off;  // pass 2, optional: [B+1] exclusive cumsum of T + tau (the flat node numbering)
  int64_t* flat_col;        // pass 2, optional: [E] flat id of every edge's source = flat_off[b] + source
  uint16_t* hits;           // pass 1, optional: [n_new, hit_cap] the first hit_cap sources of every new node, ascending
  int hit_cap;              //   (gcm_sparse_expand_edges turns them into edges without a second search)
};

__global__ void __launch_bounds__(256) k_sparse_edges(const EdgeGenArgs a) {
  extern __shared__ float eg_pos[];  // [n_src][pos_len] positions of the candidate sources
  const int b = blockIdx.x;
  const int t0 = (int)a.T[b], tau = (int)a.taus[b];
  const int s_begin = t0 + blockIdx.y * EG_SINKS;
  const int s_end = min(t0 + tau, s_begin + EG_SINKS);
  if (s_begin >= s_end) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* nodes_b = a.nodes + (size_t)b * a.N * a.F;
  if (a.use_radius) {
    for (int i = threadIdx.x; i < s_end * a.pos_len; i += blockDim.x) {
      const int k = i / a.pos_len, c = i - k * a.pos_len;
      eg_pos[i] = nodes_b[(size_t)k * a.F + a.pos_start + c * a.pos_step];
    }
    __syncthreads();
  }
  for (int s = s_begin + warp; s < s_end; s += nwarps) {
    const int64_t slot = a.new_off[b] + (s - t0);
    int64_t base = a.edges ? a.edge_off[slot] : 0;
    if (a.edges && a.hit_cap > 0 && a.edge_off[slot + 1] - base <= a.hit_cap) continue;   // done by the expansion
    int count = 0;
    for (int k0 = 0; k0 < s; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < s) {
        const int dlt = s - k;
        for (int h = 0; h < a.n_hops; ++h) hit |= (a.hops[h] == dlt);
        if (!hit && a.use_radius) {
          float d2 = 0.0f;
          for (int c = 0; c < a.pos_len; ++c) {
            const float df = eg_pos[s * a.pos_len + c] - eg_pos[k * a.pos_len + c];
            d2 += df * df;
          }
          hit = sqrtf(d2) < a.radius;
        }
      }
      const unsigned bal = __ballot_sync(GCM_FULL_MASK, hit);
      if (a.edges && hit) {
        const int64_t e = base + __popc(bal & ((1u << lane) - 1u));
        a.edges[e] = b;
        a.edges[a.E + e] = s;
        a.edges[2 * a.E + e] = k;
        if (a.flat_col) a.flat_col[e] = a.flat_off[b] + k;
      } else if (a.hits && hit) {
        const int at = count + __popc(bal & ((1u << lane) - 1u));
        if (at < a.hit_cap) a.hits[slot * a.hit_cap + at] = (uint16_t)k;
      }
      base += __popc(bal);
      count += __popc(bal);
    }
    if (!a.edges && lane == 0) a.deg[slot] = count;
  }
}


// ------------------------------------------------------------------------------------------------
// fused edge generation with a spatial hash (the radius selector at BASELINE cfg5 sizes)
//
// The brute-force kernel above tests every causal pair, N^2 / 2 per graph (8.4 M at N = 4096) in each of
// the two passes.  Here a CTA bins the graph's node positions into cells of side `radius` (first two
// position coordinates; a hash of the integer cell coordinates picks one of EH_BUCKETS buckets, counting
// sort in shared memory), and a sink only tests the nodes in the 3 x 3 cells around its own.  The cell
// test is a conservative filter (|dx| < r  =>  cells differ by at most 1; the side is widened by 2^-10 to
// absorb the rounding of x / side), hash collisions only add candidates, and every candidate goes through
// the SAME float comparison as the brute-force kernel, so the edge set is identical.  Hits are collected in
// a per-warp bitmask over the sources, which de-duplicates (temporal hop == radius hit) and gives the
// ascending source order of a coalesced COO for free.
// ------------------------------------------------------------------------------------------------
constexpr int EH_BUCKETS = 8192;
constexpr int EH_THREADS = 512;
constexpr int EH_SINKS = 1024;   // new nodes handled per CTA

__device__ __forceinline__ int eh_cell(float x, float inv_side) {
  const float c = floorf(x * inv_side);
  return c != c ? 0 : (c > 1.0e9f ? 1000000000 : (c < -1.0e9f ? -1000000000 : (int)c));
}
// Bucket of cell (cx, cy): a hashed ROW of the table per cy, the column is cx modulo the row length -- so the three cells
// cx - 1 .. cx + 1 of one cy are ONE contiguous run of buckets (two runs when the column wraps), and a sink scans three
// runs of the bucket-sorted node list instead of nine separate buckets.  colbits = 6 (128 rows x 64 columns) for 2-D cells,
// 13 (one row) for 1-D positions.  Cells that share a bucket only add candidates.
__device__ __forceinline__ int eh_row(int cy, int colbits) {
  return colbits >= 13 ? 0 : (int)(((uint32_t)cy * 2654435761u) >> (19 + colbits)) << colbits;
}
__device__ __forceinline__ int eh_bucket(int cx, int cy, int colbits) {
  return eh_row(cy, colbits) + (cx & ((1 << colbits) - 1));
}
// smallest float t with sqrtf(t) >= r: sqrtf is correctly rounded and monotone, so  sqrtf(d2) < r  <=>  d2 < t  for every
// d2 (NaN and +inf fail both) -- the pair test of k_sparse_edges without the square root
__device__ __forceinline__ float eh_sq_threshold(float r) {
  float c = r * r;
  while (c > 0.0f && sqrtf(c) >= r) c = __uint_as_float(__float_as_uint(c) - 1u);        // now sqrtf(c) < r (or c == 0)
  if (!(sqrtf(c) < r)) return c;                                                          // r <= 0: nothing passes
  for (;;) {
    const float up = __uint_as_float(__float_as_uint(c) + 1u);
    if (!(sqrtf(up) < r)) return up;
    c = up;
  }
}

// PLT: compile-time pos_len (2: positions as float2, the sink's position in registers), 0 = run-time a.pos_len
template <int PLT>
__global__ void __launch_bounds__(EH_THREADS) k_sparse_edges_hash(const EdgeGenArgs a) {
  extern __shared__ __align__(16) unsigned char eh_smem[];
  const int b = blockIdx.x;
  const int t0 = (int)a.T[b], tau = (int)a.taus[b];
  const int s_begin = t0 + blockIdx.y * EH_SINKS;
  const int s_end = min(t0 + tau, s_begin + EH_SINKS);
  if (s_begin >= s_end) return;
  const int n = s_end;                       // sources of these sinks are < s_end
  const int W = (a.N + 31) >> 5;
  const int PL = PLT ? PLT : a.pos_len;
  const int colbits = PL > 1 ? 6 : 13;
  float* pos = reinterpret_cast<float*>(eh_smem);                        // [n][PL]
  uint32_t* bits = reinterpret_cast<uint32_t*>(pos + (size_t)a.N * PL);  // [warps][W]
  uint16_t* bstart = reinterpret_cast<uint16_t*>(bits + (EH_THREADS / 32) * W);   // [EH_BUCKETS + 1]
  uint16_t* bfill = bstart + EH_BUCKETS + 2;                             // [EH_BUCKETS] running fill offsets
  uint16_t* sorted = bfill + EH_BUCKETS;                                 // [n] node ids grouped by bucket
  uint16_t* bkt = sorted + a.N;                                          // [n] bucket of every node
  __shared__ int scan_tmp[EH_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = EH_THREADS / 32;
  const float* nodes_b = a.nodes + (size_t)b * a.N * a.F;
  const float inv_side = 1.0f / (a.radius * 1.0009765625f);
  const float t2 = eh_sq_threshold(a.radius);

  for (int i = tid; i < n * PL; i += EH_THREADS) {
    const int k = i / PL, c = i - k * PL;
    pos[i] = nodes_b[(size_t)k * a.F + a.pos_start + c * a.pos_step];
  }
  for (int i = tid; i <= EH_BUCKETS; i += EH_THREADS) bstart[i] = 0;
  __syncthreads();
  // counting sort by bucket: histogram (into bstart[bucket + 1]), exclusive scan, fill
  // (16-bit counters, two per 32-bit atomic: a bucket holds at most n <= 65535 nodes, so halves never carry)
  for (int k = tid; k < n; k += EH_THREADS) {
    const int cx = eh_cell(pos[k * PL], inv_side);
    const int cy = PL > 1 ? eh_cell(pos[k * PL + 1], inv_side) : 0;
    const int bk = eh_bucket(cx, cy, colbits);
    bkt[k] = (uint16_t)bk;
    // two buckets share one 32-bit word of bstart (offset by one so that the scan is exclusive)
    atomicAdd(reinterpret_cast<unsigned int*>(bstart) + ((bk + 1) >> 1), ((bk + 1) & 1) ? 0x10000u : 1u);
  }
  __syncthreads();
  {
    // exclusive scan of EH_BUCKETS + 1 counters: 16 + 1 per thread, then a block scan of the thread totals
    constexpr int PER = EH_BUCKETS / EH_THREADS;   // 16
    const int base = tid * PER + 1;                // counters 1..EH_BUCKETS hold the histogram
    int local[PER];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      local[i] = bstart[base + i];
      sum += local[i];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(GCM_FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) scan_tmp[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += scan_tmp[w];
    int run = woff + incl - sum;                   // exclusive prefix of this thread's first counter
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      run += local[i];
      bstart[base + i] = (uint16_t)run;            // bstart[j + 1] = end of bucket j = start of bucket j + 1
    }
    // bstart[0] stays 0
  }
  __syncthreads();
  for (int i = tid; i < EH_BUCKETS; i += EH_THREADS) bfill[i] = bstart[i];
  __syncthreads();
  for (int k = tid; k < n; k += EH_THREADS) {
    const int bk = bkt[k];
    // 16-bit slot counter inside a 32-bit atomic
    const unsigned int old = atomicAdd(reinterpret_cast<unsigned int*>(bfill) + (bk >> 1), (bk & 1) ? 0x10000u : 1u);
    const int at = (bk & 1) ? (int)(old >> 16) : (int)(old & 0xffffu);
    sorted[at] = (uint16_t)k;
  }
  __syncthreads();

  uint32_t* my = bits + warp * W;
  const int ncols = 1 << colbits;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // one candidate of the bucket-sorted list: causal filter, the pair test of k_sparse_edges, hit -> bit of the source
  float sx = 0.0f, sy = 0.0f;                       // PLT == 2: the sink's position
  auto test = [&](int idx, int s) {
    const int k = sorted[idx];
    if (k < s) {
      float d2 = 0.0f;
      if (PLT == 2) {
        const float2 pk = reinterpret_cast<const float2*>(pos)[k];
        const float dx = sx - pk.x, dy = sy - pk.y;
        d2 += dx * dx;
        d2 += dy * dy;
      } else {
        for (int cc = 0; cc < PL; ++cc) {
          const float df = pos[s * PL + cc] - pos[k * PL + cc];
          d2 += df * df;
        }
      }
      if (d2 < t2) atomicOr(my + (k >> 5), 1u << (k & 31));
    }
  };
  for (int w = lane; w < W; w += 32) my[w] = 0u;       // the emission walk leaves the mask clear again
  __syncwarp();
  for (int s = s_begin + warp; s < s_end; s += nwarps) {
    const int64_t slot = a.new_off[b] + (s - t0);
    if (a.edges && a.hit_cap > 0) {   // pass 2 after an expansion: only the nodes whose list overflowed are searched again
      if (a.edge_off[slot + 1] - a.edge_off[slot] <= a.hit_cap) continue;
    }
    if (lane < a.n_hops) {
      const int k = s - a.hops[lane];
      if (k >= 0) atomicOr(my + (k >> 5), 1u << (k & 31));
    }
    if (s > 0) {
      if (PLT == 2) {
        const float2 ps = reinterpret_cast<const float2*>(pos)[s];
        sx = ps.x;
        sy = ps.y;
      } else {
        sx = pos[s * PL];
        sy = PL > 1 ? pos[s * PL + 1] : 0.0f;
      }
      const int cx = eh_cell(sx, inv_side);
      const int cy = PL > 1 ? eh_cell(sy, inv_side) : 0;
      const int col0 = (cx - 1) & (ncols - 1);
      const int cend = min(col0 + 3, ncols);
      // the runs of the three cell rows, walked as one list: candidate c sits at c + off of its run
      int e0, e1, tot, o0, o1, o2;
      {
        const int r0 = eh_row(PL > 1 ? cy - 1 : cy, colbits);
        const int b0 = bstart[r0 + col0];
        e0 = bstart[r0 + cend] - b0;
        o0 = b0;
        e1 = e0;
        tot = e0;
        o1 = o2 = 0;
        if (PL > 1) {
          const int r1 = eh_row(cy, colbits), r2 = eh_row(cy + 1, colbits);
          const int b1 = bstart[r1 + col0], b2 = bstart[r2 + col0];
          e1 = e0 + (bstart[r1 + cend] - b1);
          tot = e1 + (bstart[r2 + cend] - b2);
          o1 = b1 - e0;
          o2 = b2 - e1;
        }
      }
      for (int c = lane; c < tot; c += 32) test(c + (c < e0 ? o0 : (c < e1 ? o1 : o2)), s);
      if (col0 + 3 > ncols) {
        // the column wrapped: the rest of each row's run starts at the row's first bucket
        const int wl = col0 + 3 - ncols;
        for (int dy = (PL > 1 ? -1 : 0); dy <= (PL > 1 ? 1 : 0); ++dy) {
          const int r = eh_row(cy + dy, colbits);
          const int i0 = bstart[r], i1 = bstart[r + wl];
          for (int i = i0 + lane; i < i1; i += 32) test(i, s);
        }
      }
    }
    __syncwarp();
    // emit in ascending source order: word w of the mask belongs to lane w % 32; the warp walks the NON-EMPTY words in
    // ascending order, all lanes on one word at a time (lane = bit)
    const int64_t e_base = a.edges ? a.edge_off[slot] : 0;
    uint16_t* hrow = a.hits ? a.hits + slot * a.hit_cap : nullptr;
    const int hcap = hrow ? a.hit_cap : 0;
    int base = 0;
    for (int w0 = 0; w0 < W; w0 += 32) {
      const uint32_t mine = (w0 + lane < W) ? my[w0 + lane] : 0u;
      if (mine) my[w0 + lane] = 0u;
      uint32_t nz = __ballot_sync(GCM_FULL_MASK, mine != 0u);
      while (nz) {
        const int L = __ffs(nz) - 1;
        nz &= nz - 1;
        const uint32_t word = __shfl_sync(GCM_FULL_MASK, mine, L);
        if ((word >> lane) & 1u) {
          const int at = base + __popc(word & lt_mask);
          const int src = (w0 + L) * 32 + lane;
          if (!a.edges) {
            if (at < hcap) hrow[at] = (uint16_t)src;
          } else {
            const int64_t e = e_base + at;
            a.edges[e] = b;
            a.edges[a.E + e] = s;
            a.edges[2 * a.E + e] = src;
            if (a.flat_col) a.flat_col[e] = a.flat_off[b] + src;
          }
        }
        base += __popc(word);
      }
    }
    if (!a.edges && lane == 0) a.deg[slot] = base;
    __syncwarp();
  }
}

// Pass 2 without a search: the per-sink source lists written by pass 1 become the coalesced edge rows
// (batch, sink, source) and the CSR columns.  One warp per new node; lanes write consecutive edges.
struct EdgeExpandArgs {
  const int64_t* T; const int64_t* taus; const int64_t* new_off;
  const uint16_t* hits; int hit_cap;
  const int64_t* edge_off; int64_t* edges; int64_t E;
  const int64_t* flat_off; int64_t* flat_col;
};
constexpr int EX_SINKS = 256;
__global__ void __launch_bounds__(256) k_sparse_edges_expand(const EdgeExpandArgs a) {
  const int b = blockIdx.x;
  const int t0 = (int)a.T[b], tau = (int)a.taus[b];
  const int k_begin = blockIdx.y * EX_SINKS, k_end = min(tau, k_begin + EX_SINKS);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t fo = a.flat_col ? a.flat_off[b] : 0;
  for (int k = k_begin + warp; k < k_end; k += 8) {
    const int64_t slot = a.new_off[b] + k;
    const int64_t e0 = a.edge_off[slot];
    const int deg = (int)(a.edge_off[slot + 1] - e0);
    if (deg > a.hit_cap) continue;                       // list overflowed: left to the searching pass
    const uint16_t* h = a.hits + slot * a.hit_cap;
    for (int i = lane; i < deg; i += 32) {
      const int64_t src = h[i];
      a.edges[e0 + i] = b;
      a.edges[a.E + e0 + i] = t0 + k;
      a.edges[2 * a.E + e0 + i] = src;
      if (a.flat_col) a.flat_col[e0 + i] = fo + src;
    }
  }
}

static size_t eh_smem_bytes(int N, int pos_len) {
  const int W = (N + 31) / 32;
  return (size_t)N * pos_len * 4 + (size_t)(EH_THREADS / 32) * W * 4 + (size_t)(EH_BUCKETS + 2) * 2 +
         (size_t)EH_BUCKETS * 2 + (size_t)N * 2 * 2 + 16;
}

// ------------------------------------------------------------------------------------------------
// GraphConv forward: tile of GC_TM rows; gather-reduce into smem, then a register-tiled product with
// the K-major weight pack streamed through smem in K chunks.
// ------------------------------------------------------------------------------------------------
constexpr int GC_TM = 64;      // rows per CTA tile
constexpr int GC_THREADS = 256;
constexpr int GC_KC = 32;      // K chunk of the weight pack staged in smem
constexpr int GC_AS = GC_TM + 4;   // row stride of the transposed A tile (16-byte aligned, off the 32-bank period)

struct GraphConvFwdArgs {
  const float* x;         // [n, Fin]
  const int64_t* rowptr;  // [n+1]
  const int64_t* col;     // [E]
  const float* ew;        // [E] or NULL
  const int64_t* rows;    // [m] rows to evaluate, or NULL for all n
  int64_t m;
  int Fin, Fout;
  const float* wt;        // [2 Fin, Fout]
  const float* bias;      // [Fout] or NULL
  int act;
  float* agg_out;         // [m, Fin] or NULL
  float* out;             // [m, Fout]
};

// One row of phase 1 for Fin = 32 V: lane owns V contiguous features (one coalesced warp load per neighbour row).
template <int V>
__device__ __forceinline__ void gc_load_vec(const float* p, float (&v)[V]) {
  if (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
  } else if (V == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1 % V] = t.y;
  } else {
    v[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void gc_gather_row(const GraphConvFwdArgs& a, int64_t i, int64_t li, int64_t e0, int64_t e1,
                                              int lane, float* dst) {
  const int Fin = 32 * V;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.0f;
  const float* xl = a.x + lane * V;
  for (int64_t base = e0; base < e1; base += 32) {
    const int cnt = (int)min((int64_t)32, e1 - base);
    const int64_t my = lane < cnt ? a.col[base + lane] : 0;
    const float myw = (a.ew && lane < cnt) ? a.ew[base + lane] : 1.0f;
    int u = 0;
    for (; u + 8 <= cnt; u += 8) {
      float v[8][V];
#pragma unroll
      for (int q = 0; q < 8; ++q) gc_load_vec<V>(xl + __shfl_sync(GCM_FULL_MASK, my, u + q) * Fin, v[q]);
      if (a.ew) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float w = __shfl_sync(GCM_FULL_MASK, myw, u + q);
#pragma unroll
          for (int j = 0; j < V; ++j) v[q][j] *= w;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[q][j];
    }
    for (; u < cnt; ++u) {
      float v[V];
      gc_load_vec<V>(xl + __shfl_sync(GCM_FULL_MASK, my, u) * Fin, v);
      const float w = a.ew ? __shfl_sync(GCM_FULL_MASK, myw, u) : 1.0f;
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += a.ew ? v[j] * w : v[j];
    }
  }
  float own[V];
  gc_load_vec<V>(xl + i * Fin, own);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int f = lane * V + j;
    dst[f * GC_AS] = acc[j];
    dst[(Fin + f) * GC_AS] = own[j];
    if (a.agg_out) a.agg_out[li * Fin + f] = acc[j];
  }
}

// thread tile: 4 rows x (Fout / 16) columns, Fout in {16, 32, 64, 128} handled via NT = Fout / 16
template <int NT>
__global__ void __launch_bounds__(GC_THREADS) k_graphconv_fwd(const GraphConvFwdArgs a) {
  extern __shared__ __align__(16) float gc_smem[];
  const int Fin = a.Fin, Fout = a.Fout, K = 2 * Fin;
  float* As = gc_smem;                     // [K][GC_AS]: element (row r, k) at As[k * GC_AS + r]
  float* Ws = As + K * GC_AS;              // [GC_KC][16 NT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * GC_TM;

  // ---- phase 1: [sum of in-neighbours | own features] for every row of the tile ----
  // warp per row.  Fin in {32, 64, 128}: a lane owns Fin / 32 contiguous features, so one neighbour row is ONE
  // coalesced 128 / 256 / 512-byte warp load; the column indices of up to 32 edges are fetched with one coalesced
  // load and broadcast with shuffles; 8 gathers in flight; the sum runs in edge order (same order as a scalar loop).
  const int V = (Fin == 32 || Fin == 64 || Fin == 128) ? Fin / 32 : 0;
  for (int r = warp; r < GC_TM; r += GC_THREADS / 32) {
    const int64_t li = row0 + r;
    float* dst = As + r;                   // row r of the transposed tile: feature k at dst[k * GC_AS]
    if (li < a.m) {
      const int64_t i = a.rows ? a.rows[li] : li;
      const int64_t e0 = a.rowptr[i], e1 = a.rowptr[i + 1];
      if (V) {
        if (V == 2) gc_gather_row<2>(a, i, li, e0, e1, lane, dst);
        else if (V == 4) gc_gather_row<4>(a, i, li, e0, e1, lane, dst);
        else gc_gather_row<1>(a, i, li, e0, e1, lane, dst);
        continue;
      }
      for (int f0 = 0; f0 < Fin; f0 += 32) {
        const int f = f0 + lane;
        float acc = 0.0f;
        if (f < Fin) {
          int64_t e = e0;
          for (; e + 4 <= e1; e += 4) {   // 4 independent gathers in flight
            const int64_t c0 = a.col[e], c1 = a.col[e + 1], c2 = a.col[e + 2], c3 = a.col[e + 3];
            float v0 = a.x[c0 * Fin + f], v1 = a.x[c1 * Fin + f], v2 = a.x[c2 * Fin + f], v3 = a.x[c3 * Fin + f];
            if (a.ew) {
              v0 *= a.ew[e]; v1 *= a.ew[e + 1]; v2 *= a.ew[e + 2]; v3 *= a.ew[e + 3];
            }
            acc += v0; acc += v1; acc += v2; acc += v3;
          }
          for (; e < e1; ++e) {
            float v = a.x[a.col[e] * Fin + f];
            if (a.ew) v *= a.ew[e];
            acc += v;
          }
          dst[f * GC_AS] = acc;
          dst[(Fin + f) * GC_AS] = a.x[i * Fin + f];
          if (a.agg_out) a.agg_out[li * Fin + f] = acc;
        }
      }
    } else {
      for (int f = lane; f < K; f += 32) dst[f * GC_AS] = 0.0f;
    }
  }

  // ---- phase 2: out = act(A W + b), 4 x NT register tile per thread ----
  // A tile stored transposed ([k][row], stride GC_AS) and the weight chunk padded to 16 NT columns, so that a thread's
  // 4 rows and its groups of 4 columns are one 128-bit shared load each (2-3 loads per 16-32 FMAs instead of 8 scalar
  // ones: ncu had the kernel issue-bound with this product as half of its instructions).
  constexpr int NV = NT < 4 ? NT : 4, NG = NT / NV, WS = 16 * NT;
  const int tr = tid >> 4, tc = tid & 15;   // 16 row groups x 16 column groups; columns g*16*NV + tc*NV + v
  float acc[4][NT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += GC_KC) {
    __syncthreads();
    const int kc = min(GC_KC, K - k0);
    for (int i = tid; i < kc * WS; i += GC_THREADS) {
      const int kk = i / WS, c = i - kk * WS;
      Ws[i] = c < Fout ? __ldg(a.wt + (size_t)(k0 + kk) * Fout + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kc; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(As + (k0 + kk) * GC_AS + tr * 4);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      float wv[NT];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (NV == 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(Ws + kk * WS + g * 64 + tc * 4);
          wv[g * 4] = w4.x; wv[g * 4 + 1] = w4.y; wv[g * 4 + 2] = w4.z; wv[g * 4 + 3] = w4.w;
        } else if (NV == 2) {
          const float2 w2 = *reinterpret_cast<const float2*>(Ws + kk * WS + tc * 2);
          wv[0] = w2.x; wv[1] = w2.y;
        } else {
          wv[0] = Ws[kk * WS + tc];
        }
      }
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t li = row0 + tr * 4 + i;
    if (li < a.m) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int c = (j / NV) * (16 * NV) + tc * NV + (j % NV);
        if (c < Fout) {
          const float z = acc[i][j] + (a.bias ? __ldg(a.bias + c) : 0.0f);
          a.out[li * Fout + c] = gcm_act_fast(z, a.act);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GraphConv backward
//   kernel A (per tile of rows): dz = d_out * act'(out); d[agg | x_root] = dz [W_rel | W_root];
//                                weight gradients accumulated per CTA, flushed with atomics
//   kernel B (per node):         d_x[j] = d_xroot[j] + sum over out-edges (j -> i) of w * d_agg[i]
// ------------------------------------------------------------------------------------------------
struct GraphConvBwdArgs {
  const float* x;        // [n, Fin] layer input
  const float* agg;      // [m, Fin] saved aggregation of the evaluated rows
  const float* out;      // [m, Fout] post-activation output
  const float* d_out;    // [m, Fout]
  const int64_t* rows;   // [m] or NULL
  int64_t m, n;
  int Fin, Fout;
  const float* w_rel;    // [Fout, Fin]
  const float* w_root;   // [Fout, Fin]
  int act;
  float* d_agg;          // [m, Fin]  (scratch, written)
  float* d_x;            // [n, Fin]  (root term accumulated here; must be zero-initialised)
  float* d_w_rel;
  float* d_w_root;
  float* d_b;
};

constexpr int GB_TM = 32;  // rows per CTA iteration
__global__ void __launch_bounds__(256) k_graphconv_bwd_rows(const GraphConvBwdArgs a) {
  extern __shared__ __align__(16) float gb_smem[];
  const int Fin = a.Fin, Fout = a.Fout;
  float* dz = gb_smem;                        // [GB_TM][Fout]
  float* in = dz + GB_TM * Fout;              // [GB_TM][2 Fin]  (agg | x)
  float* accW = in + GB_TM * 2 * Fin;         // [2 Fin][Fout]  element-owned accumulators
  float* accB = accW + 2 * Fin * Fout;        // [Fout]
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * Fin * Fout + Fout; i += blockDim.x) accW[i] = 0.0f;
  const int64_t n_tiles = (a.m + GB_TM - 1) / GB_TM;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * GB_TM;
    __syncthreads();
    for (int i = tid; i < GB_TM * Fout; i += blockDim.x) {
      const int r = i / Fout, c = i - r * Fout;
      const int64_t li = row0 + r;
      float g = 0.0f;
      if (li < a.m) g = a.d_out[li * Fout + c] * gcm_act_grad(a.out[li * Fout + c], a.act);
      dz[i] = g;
    }
    for (int i = tid; i < GB_TM * Fin; i += blockDim.x) {
      const int r = i / Fin, f = i - r * Fin;
      const int64_t li = row0 + r;
      float va = 0.0f, vx = 0.0f;
      if (li < a.m) {
        va = a.agg[li * Fin + f];
        vx = a.x[(a.rows ? a.rows[li] : li) * Fin + f];
      }
      in[r * 2 * Fin + f] = va;
      in[r * 2 * Fin + Fin + f] = vx;
    }
    __syncthreads();
    // weight gradients: acc[k][c] += sum_r in[r][k] dz[r][c]
    for (int e = tid; e < 2 * Fin * Fout; e += blockDim.x) {
      const int k = e / Fout, c = e - k * Fout;
      float v = 0.0f;
#pragma unroll 8
      for (int r = 0; r < GB_TM; ++r) v = fmaf(in[r * 2 * Fin + k], dz[r * Fout + c], v);
      accW[e] += v;
    }
    for (int c = tid; c < Fout; c += blockDim.x) {
      float v = 0.0f;
      for (int r = 0; r < GB_TM; ++r) v += dz[r * Fout + c];
      accB[c] += v;
    }
    // input gradients: d_agg[r][f] = sum_c dz[r][c] W_rel[c][f];  d_x[row][f] += sum_c dz[r][c] W_root[c][f]
    for (int i = tid; i < GB_TM * Fin; i += blockDim.x) {
      const int r = i / Fin, f = i - r * Fin;
      const int64_t li = row0 + r;
      if (li >= a.m) continue;
      float ga = 0.0f, gx = 0.0f;
      for (int c = 0; c < Fout; ++c) {
        const float g = dz[r * Fout + c];
        ga = fmaf(g, __ldg(a.w_rel + (size_t)c * Fin + f), ga);
        gx = fmaf(g, __ldg(a.w_root + (size_t)c * Fin + f), gx);
      }
      a.d_agg[li * Fin + f] = ga;
      a.d_x[(a.rows ? a.rows[li] : li) * Fin + f] = gx;   // each evaluated row is unique
    }
  }
  __syncthreads();
  for (int e = tid; e < 2 * Fin * Fout; e += blockDim.x) {
    const int k = e / Fout, c = e - k * Fout;
    const float v = accW[e];
    if (v != 0.0f) atomicAdd((k < Fin ? a.d_w_rel + (size_t)c * Fin + k : a.d_w_root + (size_t)c * Fin + (k - Fin)), v);
  }
  if (a.d_b)
    for (int c = tid; c < Fout; c += blockDim.x) atomicAdd(a.d_b + c, accB[c]);
}

// transposed gather: t_rowptr/t_col group the edges by SOURCE node; t_col holds the LOCAL index (into
// the m evaluated rows) of each edge's sink.  d_x[j] += sum w * d_agg[t_col[e]]
// IDX32: every element offset t_col * Fin fits 32 bits (checked at launch): one shuffle and one IMAD.WIDE per gather
// instead of two shuffles and an emulated 64-bit multiply; the last cnt % 8 gathers of a batch are issued together
// (predicated) instead of one dependent load at a time (ncu: the serial tail was a third of this kernel's stall samples).
template <int V, bool IDX32>
__device__ __forceinline__ void gc_bwd_gather_row(const float* d_agg, const int64_t* t_col, const float* t_ew, int64_t e0,
                                                  int64_t e1, int lane, float* dst) {
  // same scheme as gc_gather_row: lane owns V contiguous features, 32 column indices per coalesced load, 8 gathers in flight
  constexpr int Fin = 32 * V;
  using idx_t = typename std::conditional<IDX32, unsigned, int64_t>::type;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.0f;
  const float* xl = d_agg + lane * V;
  for (int64_t base = e0; base < e1; base += 32) {
    const int cnt = (int)min((int64_t)32, e1 - base);
    const idx_t my = lane < cnt ? (idx_t)t_col[base + lane] : (idx_t)0;
    const float myw = (t_ew && lane < cnt) ? t_ew[base + lane] : 1.0f;
    int u = 0;
    for (; u + 8 <= cnt; u += 8) {
      float v[8][V];
#pragma unroll
      for (int q = 0; q < 8; ++q) gc_load_vec<V>(xl + (size_t)(__shfl_sync(GCM_FULL_MASK, my, u + q) * (idx_t)Fin), v[q]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float w = t_ew ? __shfl_sync(GCM_FULL_MASK, myw, u + q) : 1.0f;
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += t_ew ? v[q][j] * w : v[q][j];
      }
    }
    if (u < cnt) {
      float v[7][V];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[q][j] = 0.0f;
        const idx_t c = __shfl_sync(GCM_FULL_MASK, my, (u + q) & 31);
        if (u + q < cnt) gc_load_vec<V>(xl + (size_t)(c * (idx_t)Fin), v[q]);
      }
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float w = t_ew ? __shfl_sync(GCM_FULL_MASK, myw, (u + q) & 31) : 1.0f;
        if (u + q < cnt) {
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += t_ew ? v[q][j] * w : v[q][j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) dst[lane * V + j] += acc[j];
}

// V = Fin / 32 in {1, 2, 4}; V = 0: any Fin (scalar loop).  One instantiation per width so that the narrow ones keep 40
// registers (48 warps per SM; a single kernel holding the V = 4 path ran at 64 registers, 32 warps).
template <int V, bool IDX32>
__global__ void __launch_bounds__(256, V == 4 || V == 0 ? 4 : 6) k_graphconv_bwd_gather(
    const float* d_agg, const int64_t* t_rowptr, const int64_t* t_col, const float* t_ew, int64_t n, int Fin, float* d_x) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const int64_t e0 = t_rowptr[j], e1 = t_rowptr[j + 1];
  if (e0 == e1) return;
  if (V > 0) {
    gc_bwd_gather_row<(V > 0 ? V : 1), IDX32>(d_agg, t_col, t_ew, e0, e1, lane, d_x + j * Fin);
  } else {
    for (int f0 = 0; f0 < Fin; f0 += 32) {
      const int f = f0 + lane;
      if (f >= Fin) break;
      float acc = 0.0f;
      for (int64_t e = e0; e < e1; ++e) {
        float v = d_agg[t_col[e] * Fin + f];
        if (t_ew) v *= t_ew[e];
        acc += v;
      }
      d_x[j * Fin + f] += acc;
    }
  }
}

static int launch_bwd_gather(const float* d_agg, const int64_t* t_rowptr, const int64_t* t_col, const float* t_ew, int64_t n,
                             int64_t m, int Fin, float* d_x, cudaStream_t stream) {
  const int64_t g2 = (n + 7) / 8;
  GCM_REQUIRE(g2 < 2147483647LL, "sparse_graphconv_bwd: too many nodes");
  const bool i32 = (unsigned long long)m * (unsigned long long)Fin < (1ull << 32);      // t_col holds sink rows < m
#define BG_LAUNCH(V)                                                                                                      \
  do {                                                                                                                    \
    if (i32) k_graphconv_bwd_gather<V, true><<<(unsigned)g2, 256, 0, stream>>>(d_agg, t_rowptr, t_col, t_ew, n, Fin, d_x);  \
    else k_graphconv_bwd_gather<V, false><<<(unsigned)g2, 256, 0, stream>>>(d_agg, t_rowptr, t_col, t_ew, n, Fin, d_x);     \
  } while (0)
  if (Fin == 64) BG_LAUNCH(2);
  else if (Fin == 32) BG_LAUNCH(1);
  else if (Fin == 128) BG_LAUNCH(4);
  else BG_LAUNCH(0);
#undef BG_LAUNCH
  return gcm_check_launch("k_graphconv_bwd_gather");
}

// ------------------------------------------------------------------------------------------------
// Transposed CSR (edges grouped by SOURCE) of a block-diagonal graph whose per-graph edge ranges are contiguous:
// what the backward gather needs, without a global sort.  One CTA per graph: histogram of the sources in shared
// memory, exclusive scan, fill through atomic cursors, then every source row is sorted by sink (rows are short), so
// the result does not depend on the order the atomics happened in (the gradient sums stay deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int TC_MAXN = 8192;
__global__ void __launch_bounds__(512) k_csr_transpose(const int64_t* rowptr, const int64_t* col, const int64_t* node_off,
                                                       const int64_t* sink_local, int64_t* t_rowptr, int64_t* t_col,
                                                       int64_t n_total) {
  __shared__ int cnt[TC_MAXN + 1];
  __shared__ int scan_tmp[16];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = node_off[b], n1 = node_off[b + 1];
  const int nb = (int)(n1 - n0);
  const int64_t e0 = rowptr[n0], e1 = rowptr[n1];
  for (int i = tid; i <= nb; i += 512) cnt[i] = 0;
  __syncthreads();
  for (int64_t e = e0 + tid; e < e1; e += 512) atomicAdd(&cnt[(int)(col[e] - n0) + 1], 1);
  __syncthreads();
  // inclusive scan of cnt[1..nb] in place (cnt[0] = 0): 16 values per thread, then a block scan of the thread totals
  const int per = (nb + 511) / 512;
  const int base = 1 + tid * per;
  int sum = 0;
  for (int i = 0; i < per; ++i)
    if (base + i <= nb) sum += cnt[base + i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(GCM_FULL_MASK, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) scan_tmp[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += scan_tmp[w];
  int run = woff + incl - sum;
  for (int i = 0; i < per; ++i)
    if (base + i <= nb) {
      run += cnt[base + i];
      cnt[base + i] = run;                      // cnt[j + 1] = end of source j's row (local), cnt[j] = its start
    }
  __syncthreads();
  for (int i = tid; i < nb; i += 512) t_rowptr[n0 + i] = e0 + cnt[i];
  if (b == gridDim.x - 1 && tid == 0) t_rowptr[n_total] = e1;
  __syncthreads();
  // fill: cursor of source j = cnt[j] (start), advanced atomically; afterwards cnt[j] = end of row j
  if (sink_local) {   // the builder's edge list names every edge's sink: coalesced walk over the edges
    for (int64_t e = e0 + tid; e < e1; e += 512) {
      const int j = (int)(col[e] - n0);
      const int at = atomicAdd(&cnt[j], 1);
      t_col[e0 + at] = n0 + sink_local[e];      // the sink's flat id (= its position among the evaluated rows)
    }
  } else {
    for (int64_t i = n0 + tid; i < n1; i += 512) {
      for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        const int j = (int)(col[e] - n0);
        const int at = atomicAdd(&cnt[j], 1);
        t_col[e0 + at] = i;
      }
    }
  }
  __syncthreads();
  // cnt[j] is now the END of row j; its start is the end of row j - 1 (0 for j = 0).  Every row is sorted by sink in a
  // thread-local buffer (one read and one write of the row; sorting in place in global memory cost a read-modify-write
  // per shifted element and 6 of the kernel's 7.9 ms at cfg5); rows longer than the buffer are sorted in place.
  constexpr int TB = 96;
  for (int j = tid; j < nb; j += 512) {
    const int r0 = j ? cnt[j - 1] : 0, r1 = cnt[j];
    int64_t* row = t_col + e0;
    const int len = r1 - r0;
    if (len <= 1) continue;
    if (len <= TB) {
      int buf[TB];
      for (int a = 0; a < len; ++a) {
        const int v = (int)(row[r0 + a] - n0);
        int p = a - 1;
        while (p >= 0 && buf[p] > v) {
          buf[p + 1] = buf[p];
          --p;
        }
        buf[p + 1] = v;
      }
      for (int a = 0; a < len; ++a) row[r0 + a] = n0 + buf[a];
    } else {
      for (int a = r0 + 1; a < r1; ++a) {
        const int64_t v = row[a];
        int p = a - 1;
        while (p >= r0 && row[p] > v) {
          row[p + 1] = row[p];
          --p;
        }
        row[p + 1] = v;
      }
    }
  }
}

// The same transposition with the source rows assembled and sorted in SHARED memory (16-bit local sink ids, up to 96 K
// entries = every edge of a cfg5 graph at once; more edges go in several source ranges, each a pass over the graph's edge
// list) and written out coalesced: k_csr_transpose sorts every row in a thread-local buffer read from / written to global
// memory one thread per row (3.86 ms at cfg5).  Needs the builder's sink list and graphs of at most TC_MAXN nodes.
constexpr int TS_THREADS = 1024;
constexpr int TS_CAP = 96 * 1024;
__global__ void __launch_bounds__(TS_THREADS) k_csr_transpose_smem(const int64_t* rowptr, const int64_t* col,
                                                                   const int64_t* node_off, const int64_t* sink_local,
                                                                   int64_t* t_rowptr, int64_t* t_col, int64_t n_total,
                                                                   int cap) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  int* cnt = reinterpret_cast<int*>(ts_smem);                                   // [TC_MAXN + 1]
  uint16_t* buf = reinterpret_cast<uint16_t*>(cnt + TC_MAXN + 4);               // [TS_CAP]
  __shared__ int scan_tmp[TS_THREADS / 32];
  __shared__ int sh_jb;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = node_off[b], n1 = node_off[b + 1];
  const int nb = (int)(n1 - n0);
  const int64_t e0 = rowptr[n0], e1 = rowptr[n1];
  for (int i = tid; i <= nb; i += TS_THREADS) cnt[i] = 0;
  __syncthreads();
  for (int64_t e = e0 + tid; e < e1; e += TS_THREADS) atomicAdd(&cnt[(int)(col[e] - n0) + 1], 1);
  __syncthreads();
  // inclusive scan of cnt[1..nb] in place (cnt[0] = 0)
  const int per = (nb + TS_THREADS - 1) / TS_THREADS;
  const int base = 1 + tid * per;
  int sum = 0;
  for (int i = 0; i < per; ++i)
    if (base + i <= nb) sum += cnt[base + i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(GCM_FULL_MASK, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) scan_tmp[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += scan_tmp[w];
  int run = woff + incl - sum;
  for (int i = 0; i < per; ++i)
    if (base + i <= nb) {
      run += cnt[base + i];
      cnt[base + i] = run;                      // cnt[j] = start of source j's row (local), cnt[nb] = number of edges
    }
  __syncthreads();
  for (int i = tid; i < nb; i += TS_THREADS) t_rowptr[n0 + i] = e0 + cnt[i];
  if (b == gridDim.x - 1 && tid == 0) t_rowptr[n_total] = e1;
  const int n_edges = (int)(e1 - e0);
  int ja = 0, base_a = 0;                        // sources < ja are done; base_a = start of row ja
  while (ja < nb) {
    __syncthreads();
    if (tid == 0) {
      // the largest jb in (ja, nb] whose rows ja .. jb - 1 fit the buffer: cnt[j] (j >= ja) still holds the row starts
      int lo = ja + 1, hi = nb;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (cnt[mid] - base_a <= cap) lo = mid;
        else hi = mid - 1;
      }
      sh_jb = lo;
    }
    __syncthreads();
    const int jb = sh_jb;
    const int end_b = jb < nb ? cnt[jb] : n_edges;      // start of row jb = end of the range (read before the fill moves it)
    __syncthreads();
    for (int64_t e = e0 + tid; e < e1; e += TS_THREADS) {
      const int j = (int)(col[e] - n0);
      if (j >= ja && j < jb) {
        const int at = atomicAdd(&cnt[j], 1);
        buf[at - base_a] = (uint16_t)sink_local[e];
      }
    }
    __syncthreads();
    // cnt[j] is now the END of row j for ja <= j < jb; insertion sort of every row by sink, in shared memory
    for (int j = ja + tid; j < jb; j += TS_THREADS) {
      const int r0 = (j == ja ? base_a : cnt[j - 1]) - base_a, r1 = cnt[j] - base_a;
      for (int a = r0 + 1; a < r1; ++a) {
        const uint16_t v = buf[a];
        int p = a - 1;
        while (p >= r0 && buf[p] > v) {
          buf[p + 1] = buf[p];
          --p;
        }
        buf[p + 1] = v;
      }
    }
    __syncthreads();
    int64_t* out = t_col + e0 + base_a;
    for (int i = tid; i < end_b - base_a; i += TS_THREADS) out[i] = n0 + buf[i];
    ja = jb;
    base_a = end_b;
  }
}

static int g_transpose_cap = TS_CAP;
/* Test hook: entries of the shared-memory row buffer of k_csr_transpose_smem (0 = the default, 96 K); a small value forces
 * several source ranges per graph.  Must be >= the longest source row (<= TC_MAXN). */
extern "C" int gcm_set_csr_transpose_cap(int entries) {
  GCM_REQUIRE(entries == 0 || (entries >= TC_MAXN && entries <= TS_CAP), "set_csr_transpose_cap: %d outside [%d, %d]", entries,
              TC_MAXN, TS_CAP);
  g_transpose_cap = entries ? entries : TS_CAP;
  return GCM_OK;
}

extern "C" int gcm_sparse_csr_transpose(const int64_t* rowptr, const int64_t* col, const int64_t* node_off,
                                        const int64_t* sink_local, int B, int64_t n_total, int64_t* t_rowptr,
                                        int64_t* t_col, void* stream) {
  GCM_REQUIRE(rowptr && col && node_off && t_rowptr && t_col && B >= 0 && n_total >= 0, "sparse_csr_transpose: bad arguments");
  if (B == 0) return GCM_OK;
  static const bool no_smem = getenv("GCM_B200_TRANSPOSE_GLOBAL_SORT") != nullptr;     // A/B switch
  if (sink_local && !no_smem) {
    const size_t smem = (size_t)(TC_MAXN + 4) * 4 + (size_t)TS_CAP * 2;
    cudaError_t e = cudaFuncSetAttribute(k_csr_transpose_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(csr_transpose_smem): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    k_csr_transpose_smem<<<B, TS_THREADS, smem, (cudaStream_t)stream>>>(rowptr, col, node_off, sink_local, t_rowptr, t_col,
                                                                        n_total, g_transpose_cap);
    return gcm_check_launch("k_csr_transpose_smem");
  }
  k_csr_transpose<<<B, 512, 0, (cudaStream_t)stream>>>(rowptr, col, node_off, sink_local, t_rowptr, t_col, n_total);
  return gcm_check_launch("k_csr_transpose");
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int gcm_sparse_write_flatten(float* nodes, const float* x, const int64_t* T, const int64_t* taus,
                                        const int64_t* offsets, int B, int N, int F, int tmax, float* flat,
                                        void* stream) {
  GCM_REQUIRE(nodes && x && T && taus && offsets && B >= 0 && N >= 1 && F >= 1 && tmax >= 0,
              "sparse_write_flatten: bad arguments");
  if (B == 0) return GCM_OK;
  k_sparse_write_flatten<<<B, 256, 0, (cudaStream_t)stream>>>(nodes, x, T, taus, offsets, N, F, tmax, flat);
  return gcm_check_launch("k_sparse_write_flatten");
}

extern "C" int gcm_sparse_write_flatten_oop(const float* nodes_in, float* nodes_out, const float* x, const int64_t* T,
                                            const int64_t* taus, const int64_t* offsets, int B, int N, int F, int tmax,
                                            float* flat, void* stream) {
  GCM_REQUIRE(nodes_in && nodes_out && nodes_in != nodes_out && x && T && taus && offsets && B >= 0 && N >= 1 && F >= 1 &&
                  tmax >= 0, "sparse_write_flatten_oop: bad arguments");
  GCM_REQUIRE((long long)N * F < 2147483647LL, "sparse_write_flatten_oop: N * F too large");
  if (B == 0) return GCM_OK;
  dim3 grid(B, (N + WF_ROWS - 1) / WF_ROWS);
  const uintptr_t al = reinterpret_cast<uintptr_t>(nodes_in) | reinterpret_cast<uintptr_t>(nodes_out) |
                       reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(flat);
  if (F % 4 == 0 && (al & 15) == 0) {
    k_sparse_write_flatten_oop<float4><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(nodes_in), reinterpret_cast<float4*>(nodes_out), reinterpret_cast<const float4*>(x),
        T, taus, offsets, N, F / 4, tmax, reinterpret_cast<float4*>(flat));
  } else {
    k_sparse_write_flatten_oop<float><<<grid, 256, 0, (cudaStream_t)stream>>>(nodes_in, nodes_out, x, T, taus, offsets, N,
                                                                            F, tmax, flat);
  }
  return gcm_check_launch("k_sparse_write_flatten_oop");
}

extern "C" int gcm_sparse_write_flatten_bwd(const float* d_nodes_out, const float* d_flat, const int64_t* T,
                                            const int64_t* taus, const int64_t* offsets, int B, int N, int F, int tmax,
                                            float* d_nodes_in, float* d_x, void* stream) {
  GCM_REQUIRE(T && taus && offsets && B >= 0 && N >= 1 && F >= 1 && tmax >= 0, "sparse_write_flatten_bwd: bad arguments");
  GCM_REQUIRE((long long)(N > tmax ? N : tmax) * F < 2147483647LL, "sparse_write_flatten_bwd: N * F too large");
  if (B == 0 || (!d_nodes_in && !d_x)) return GCM_OK;
  const int rows = N > tmax ? N : tmax;
  dim3 grid(B, (rows + WF_ROWS - 1) / WF_ROWS);
  const uintptr_t al = reinterpret_cast<uintptr_t>(d_nodes_out) | reinterpret_cast<uintptr_t>(d_flat) |
                       reinterpret_cast<uintptr_t>(d_nodes_in) | reinterpret_cast<uintptr_t>(d_x);
  if (F % 4 == 0 && (al & 15) == 0) {
    k_sparse_write_flatten_bwd<float4><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_nodes_out), reinterpret_cast<const float4*>(d_flat), T, taus, offsets, N, F / 4, tmax,
        reinterpret_cast<float4*>(d_nodes_in), reinterpret_cast<float4*>(d_x));
  } else {
    k_sparse_write_flatten_bwd<float><<<grid, 256, 0, (cudaStream_t)stream>>>(d_nodes_out, d_flat, T, taus, offsets, N, F, tmax,
                                                                            d_nodes_in, d_x);
  }
  return gcm_check_launch("k_sparse_write_flatten_bwd");
}

static int g_edge_builder = GCM_EB_AUTO;
extern "C" int gcm_set_edge_builder(int which) {
  GCM_REQUIRE(which >= GCM_EB_AUTO && which <= GCM_EB_HASH, "set_edge_builder: bad variant %d", which);
  g_edge_builder = which;
  return GCM_OK;
}

extern "C" int gcm_sparse_build_edges(const float* nodes, const int64_t* T, const int64_t* taus,
                                      const int64_t* new_off, int B, int N, int F, int tmax, const int32_t* hops,
                                      int n_hops, int use_radius, int pos_start, int pos_step, int pos_len,
                                      float radius, int32_t* deg, const int64_t* edge_off, int64_t* edges,
                                      int64_t E, const int64_t* flat_off, int64_t* flat_col, uint16_t* hits,
                                      int hit_cap, void* stream) {
  GCM_REQUIRE(T && taus && new_off && B >= 0 && N >= 1 && F >= 1, "sparse_build_edges: bad arguments");
  GCM_REQUIRE(n_hops >= 0 && n_hops <= GCM_MAX_HOPS && (n_hops == 0 || hops), "sparse_build_edges: n_hops=%d", n_hops);
  GCM_REQUIRE((edges == nullptr) == (edge_off == nullptr), "sparse_build_edges: edges and edge_off go together");
  GCM_REQUIRE(edges || deg, "sparse_build_edges: pass 1 needs deg");
  GCM_REQUIRE(!flat_col || (edges && flat_off), "sparse_build_edges: flat_col needs edges and flat_off");
  if (use_radius) {
    GCM_REQUIRE(nodes && pos_len >= 1 && pos_step >= 1 && pos_start >= 0 && pos_start + (pos_len - 1) * pos_step < F,
                "sparse_build_edges: position slice outside [0,F)");
    GCM_REQUIRE((size_t)N * pos_len * 4 <= 200 * 1024, "sparse_build_edges: N * pos_len too large for shared memory");
  }
  if (B == 0 || tmax == 0) return GCM_OK;
  EdgeGenArgs a;
  a.nodes = nodes; a.T = T; a.taus = taus; a.new_off = new_off; a.N = N; a.F = F;
  a.n_hops = n_hops;
  for (int i = 0; i < n_hops; ++i) {
    GCM_REQUIRE(hops[i] >= 1, "sparse_build_edges: hops must be >= 1 (a self edge violates causality)");
    a.hops[i] = hops[i];
  }
  a.use_radius = use_radius; a.pos_start = pos_start; a.pos_step = pos_step; a.pos_len = pos_len;
  a.radius = radius; a.deg = deg; a.edge_off = edge_off; a.edges = edges; a.E = E;
  a.flat_off = flat_off; a.flat_col = flat_col;
  GCM_REQUIRE(!hits || (!edges && hit_cap >= 1 && N <= 65536), "sparse_build_edges: hits belong to pass 1 (N <= 65536)");
  a.hits = hits; a.hit_cap = hit_cap;   // pass 2 with hit_cap > 0: nodes with at most hit_cap sources are skipped
  // radius selector on graphs large enough for the pair test to dominate: spatial hash
  const bool hash_fits = N <= 65535 && radius > 0.0f && radius < 1.0e30f && eh_smem_bytes(N, pos_len) <= 160 * 1024;
  if (use_radius && hash_fits && (g_edge_builder == GCM_EB_HASH || (g_edge_builder == GCM_EB_AUTO && N >= 256))) {
    const size_t hsmem = eh_smem_bytes(N, pos_len);
    auto kern = pos_len == 2 ? k_sparse_edges_hash<2> : k_sparse_edges_hash<0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(edges_hash): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    dim3 hgrid(B, (tmax + EH_SINKS - 1) / EH_SINKS);
    kern<<<hgrid, EH_THREADS, hsmem, (cudaStream_t)stream>>>(a);
    return gcm_check_launch("k_sparse_edges_hash");
  }
  const size_t smem = use_radius ? (size_t)N * pos_len * 4 : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_sparse_edges, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(edges): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
  }
  dim3 grid(B, (tmax + EG_SINKS - 1) / EG_SINKS);
  k_sparse_edges<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  return gcm_check_launch("k_sparse_edges");
}

int gcm_graphconv_fwd_tc(const float* x, const int64_t* rowptr, const int64_t* col, const float* ew, const int64_t* rows,
                         int64_t m, int Fin, int Fout, const float* wt, const float* bias, int act, float* agg_out,
                         float* out, cudaStream_t stream, int64_t x_rows);
static int64_t g_graphconv_x_rows = 0;
// Rows of x for the next gcm_sparse_graphconv_fwd call of this thread that evaluates a row SUBSET (the tensor-core kernel
// addresses x with 32-bit offsets and must know that column * Fin fits); 0 = unknown.
extern "C" int gcm_sparse_graphconv_hint_rows(long long n) {
  g_graphconv_x_rows = n;
  return GCM_OK;
}

extern "C" int gcm_sparse_graphconv_fwd(const float* x, const int64_t* rowptr, const int64_t* col, const float* ew,
                                        const int64_t* rows, int64_t m, int Fin, int Fout, const float* wt,
                                        const float* bias, int act, float* agg_out, float* out, void* stream) {
  GCM_REQUIRE(x && rowptr && wt && out && m >= 0, "sparse_graphconv_fwd: bad arguments");
  GCM_REQUIRE(Fin >= 1 && Fin <= 128 && Fout >= 1 && Fout <= 128, "sparse_graphconv_fwd: Fin=%d Fout=%d outside [1,128]",
              Fin, Fout);
  if (m == 0) return GCM_OK;
  {
    // the per-tile product on the tensor cores (3xTF32) for Fin in {32, 64} (gcm_sparse_tc.cu); same gathers, same sums
    const int64_t hint = g_graphconv_x_rows;
    g_graphconv_x_rows = 0;
    const int rc = gcm_graphconv_fwd_tc(x, rowptr, col, ew, rows, m, Fin, Fout, wt, bias, act, agg_out, out,
                                        (cudaStream_t)stream, hint);
    if (rc != GCM_ERR_UNSUPPORTED) return rc;
  }
  GraphConvFwdArgs a{x, rowptr, col, ew, rows, m, Fin, Fout, wt, bias, act, agg_out, out};
  const size_t smem = ((size_t)2 * Fin * GC_AS + (size_t)GC_KC * 16 * ((Fout + 15) / 16 <= 1 ? 1 : ((Fout + 15) / 16 <= 2 ? 2 : ((Fout + 15) / 16 <= 4 ? 4 : 8)))) * 4;
  const int64_t grid = (m + GC_TM - 1) / GC_TM;
  GCM_REQUIRE(grid < 2147483647LL, "sparse_graphconv_fwd: too many rows");
  const int nt = (Fout + 15) / 16;
  cudaError_t e = cudaSuccess;
#define GC_LAUNCH(NT)                                                                                          \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      e = cudaFuncSetAttribute(k_graphconv_fwd<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e == cudaSuccess) k_graphconv_fwd<NT><<<(unsigned)grid, GC_THREADS, smem, (cudaStream_t)stream>>>(a);  \
  } while (0)
  if (nt <= 1) GC_LAUNCH(1);
  else if (nt <= 2) GC_LAUNCH(2);
  else if (nt <= 4) GC_LAUNCH(4);
  else GC_LAUNCH(8);
#undef GC_LAUNCH
  if (e != cudaSuccess) {
    gcm_set_error("cudaFuncSetAttribute(graphconv_fwd): %s", cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  return gcm_check_launch("k_graphconv_fwd");
}

extern "C" int gcm_outer_reduce_tc32_pair(const float* A, long long lda, int Ho, const float* X1, long long ldx1, int Hi1,
                                          const float* X2, long long ldx2, int Hi2, long long rows, float* workspace,
                                          float* dW1, float* dW2, float* db, void* stream);
extern "C" int gcm_sparse_graphconv_bwd(const float* x, const float* agg, const float* out, const float* d_out,
                                        const int64_t* rows, int64_t m, int64_t n, const int64_t* t_rowptr,
                                        const int64_t* t_col, const float* t_ew, int Fin, int Fout,
                                        const float* w_rel, const float* w_root, int act, float* d_agg, float* d_x,
                                        float* d_w_rel, float* d_w_root, float* d_b, float* dz_scratch,
                                        const float* w_rel_t, const float* w_root_t, float* outer_ws, void* stream) {
  GCM_REQUIRE(x && agg && out && d_out && t_rowptr && w_rel && w_root && d_agg && d_x && d_w_rel && d_w_root,
              "sparse_graphconv_bwd: null pointer");
  GCM_REQUIRE(Fin >= 1 && Fin <= 128 && Fout >= 1 && Fout <= 128 && m >= 0 && n >= 0,
              "sparse_graphconv_bwd: bad dims");
  if (m == 0 || n == 0) return GCM_OK;
  if (!rows && dz_scratch && w_rel_t && w_root_t) {
    // every row is evaluated (the all-at-once call): the per-row part is four plain products over m rows, done by the
    // register-tiled kernels of the dense path instead of the element-owned tile kernel below (24.5 -> ~6 ms per layer
    // at cfg5):  dz = d_out * act'(out);  d_agg = dz W_rel;  d_x = dz W_root;  dW_rel += dz^T agg;  dW_root += dz^T x
    if (int rc = gcm_act_backward(d_out, out, act, (long long)m * Fout, dz_scratch, stream)) return rc;
    const bool tc_ok = Fin % 16 == 0 && Fout % 16 == 0 && Fin >= 16 && Fout >= 16 &&
                       ((reinterpret_cast<uintptr_t>(dz_scratch) | reinterpret_cast<uintptr_t>(w_rel_t) |
                         reinterpret_cast<uintptr_t>(w_root_t) | reinterpret_cast<uintptr_t>(d_agg) |
                         reinterpret_cast<uintptr_t>(d_x)) & 15) == 0;
    if (tc_ok) {   // 3xTF32 on the tensor cores (fp32-accurate), persistent over the row tiles
      if (int rc = gcm_linear_tc32(dz_scratch, Fout, Fout, w_rel_t, nullptr, 0, 0, nullptr, nullptr, GCM_ACT_NONE, m, Fin,
                                   d_agg, Fin, nullptr, stream)) return rc;
      if (int rc = gcm_linear_tc32(dz_scratch, Fout, Fout, w_root_t, nullptr, 0, 0, nullptr, nullptr, GCM_ACT_NONE, m, Fin,
                                   d_x, Fin, nullptr, stream)) return rc;
    } else {
      if (int rc = gcm_linear2(dz_scratch, Fout, Fout, w_rel_t, nullptr, 0, 0, nullptr, nullptr, GCM_ACT_NONE, m, Fin, d_agg,
                               Fin, nullptr, 0, stream)) return rc;
      if (int rc = gcm_linear2(dz_scratch, Fout, Fout, w_root_t, nullptr, 0, 0, nullptr, nullptr, GCM_ACT_NONE, m, Fin, d_x,
                               Fin, nullptr, 0, stream)) return rc;
    }
    // weight gradients: 3xTF32 reductions over the m rows on the tensor cores for wide layers; at 64 x 64 half of that
    // kernel's row-owning threads would idle and the CUDA-core tile (1.67 ms per reduction at cfg5) beats it (2.10 ms)
    if (outer_ws && tc_ok && 2 * Fin <= 128) {
      // both reductions share dz: ONE pass over the rows on the tensor cores, X operand = [agg | x] (threads 0 .. Fin - 1
      // of a group own an agg column, the next Fin an x column): 2 x 1.67 ms -> one launch at cfg5
      if (int rc = gcm_outer_reduce_tc32_pair(dz_scratch, Fout, Fout, agg, Fin, Fin, x, Fin, Fin, m, outer_ws, d_w_rel,
                                              d_w_root, d_b, stream)) return rc;
    } else if (outer_ws && tc_ok && (Fin > 64 || Fout > 64)) {
      if (int rc = gcm_outer_reduce_tc32(dz_scratch, Fout, Fout, agg, Fin, Fin, m, outer_ws, d_w_rel, d_b, stream)) return rc;
      if (int rc = gcm_outer_reduce_tc32(dz_scratch, Fout, Fout, x, Fin, Fin, m, outer_ws, d_w_root, nullptr, stream)) return rc;
    } else {
      if (int rc = gcm_outer_reduce(dz_scratch, Fout, Fout, agg, Fin, Fin, m, d_w_rel, d_b, stream)) return rc;
      if (int rc = gcm_outer_reduce(dz_scratch, Fout, Fout, x, Fin, Fin, m, d_w_root, nullptr, stream)) return rc;
    }
    return launch_bwd_gather(d_agg, t_rowptr, t_col, t_ew, n, m, Fin, d_x, (cudaStream_t)stream);
  }
  GraphConvBwdArgs a{x, agg, out, d_out, rows, m, n, Fin, Fout, w_rel, w_root, act, d_agg, d_x, d_w_rel, d_w_root, d_b};
  const size_t smem = ((size_t)GB_TM * Fout + (size_t)GB_TM * 2 * Fin + (size_t)2 * Fin * Fout + Fout) * 4;
  cudaError_t e = cudaFuncSetAttribute(k_graphconv_bwd_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    gcm_set_error("cudaFuncSetAttribute(graphconv_bwd): %s", cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  const int64_t n_tiles = (m + GB_TM - 1) / GB_TM;
  int grid = 2 * gcm_num_sms();
  if (grid > n_tiles) grid = (int)n_tiles;
  k_graphconv_bwd_rows<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  if (int rc = gcm_check_launch("k_graphconv_bwd_rows")) return rc;
  return launch_bwd_gather(d_agg, t_rowptr, t_col, t_ew, n, m, Fin, d_x, (cudaStream_t)stream);
}

extern "C" int gcm_sparse_expand_edges(const int64_t* T, const int64_t* taus, const int64_t* new_off, int B, int tmax,
                                       const uint16_t* hits, int hit_cap, const int64_t* edge_off, int64_t* edges,
                                       int64_t E, const int64_t* flat_off, int64_t* flat_col, void* stream) {
  GCM_REQUIRE(T && taus && new_off && hits && edge_off && edges && hit_cap >= 1 && B >= 0 && tmax >= 0 && E >= 0,
              "sparse_expand_edges: bad arguments");
  GCM_REQUIRE(!flat_col || flat_off, "sparse_expand_edges: flat_col needs flat_off");
  if (B == 0 || tmax == 0 || E == 0) return GCM_OK;
  EdgeExpandArgs a{T, taus, new_off, hits, hit_cap, edge_off, edges, E, flat_off, flat_col};
  dim3 grid(B, (tmax + EX_SINKS - 1) / EX_SINKS);
  k_sparse_edges_expand<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return gcm_check_launch("k_sparse_edges_expand");
}

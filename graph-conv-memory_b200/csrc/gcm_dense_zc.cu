// DenseGCM step for ONE distance selector (CosineEdge / SpatialEdge / EuclideanEdge) with a per-node
// pre-activation cache ("zc" path), sm_100a.  BASELINE cfg4.
//
// Distance.forward (edge_selectors/distance.py:18-39) only ever writes row t of the adjacency: edges run from
// older nodes to the new one (`bidirectional` is unreachable from the three subclasses, :45-46,55-56,69-70).
// So the in-neighbourhood of node i is fixed when i is created, and the layer-1 pre-activation
//     z_i = W_rel1 (sum_{j in N(i)} x_j) + W_root1 x_i + b1
// changes afterwards only when one of its in-neighbours e leaves the window (gcm.py:323-355): z_i -= W_rel1 x_e.
// The reference (and k_step_general) recompute layer 1 for every in-neighbour of t each step -- with CosineEdge,
// which links t to the DISSIMILAR nodes (similarity < max_distance), that is ~n rows with ~n/2 neighbours each.
// Here z lives in HBM next to the node log (zcache [B, C, H1]) and a step is two streaming passes per graph:
//   pass 1 over the node rows  : distance + threshold -> mask row of t, and the aggregation of t's neighbours
//                                (rows staged in shared memory tiles, one thread per row for the distance, one thread
//                                per feature column for the sums); then z_t, and u = W_rel1 x_e if a node leaves
//   pass 2 over the cached rows: z_i -= u for the rows that had e as in-neighbour (one mask bit each), and
//                                G = sum_{i in N(t)} act1(z_i);  belief = act2(W_rel2 G + W_root2 act1(z_t) + b2)
// = n (F + 2 H1) 4 bytes per graph-step instead of an n x n/2 gather plus a [n, 2F] x [2F, H1] product.
// Valid while the weights stay the same and every step of the state went through this kernel (the host tracks
// both, gcm/fused.py); training (autograd) uses the general kernels.  One CTA per graph, 256 threads.
#include "gcm_common.cuh"

constexpr int ZC_THREADS = 256;

struct ZcArgs {
  gcm_dense_state st;
  const float* obs;
  gcm_selector sel;
  gcm_gnn gnn;
  float* zcache;     // [B, C, H1]
  float* belief;
  int32_t* status;
  int tile_rows;     // rows of the node log staged in shared memory at a time (<= ZC_THREADS)
};

template <int FR>
__global__ void __launch_bounds__(ZC_THREADS) k_step_dist_zc(const ZcArgs a) {
  extern __shared__ __align__(16) float zs[];
  const int F = a.st.F, N = a.st.N, C = a.st.C, W = a.st.W, H1 = a.gnn.H1, H2 = a.gnn.H2;
  float* cur = zs;                       // [F]   the observation
  float* xe = cur + F;                   // [F]   the node that leaves the window
  float* agg = xe + F;                   // [F]   sum of t's in-neighbours
  float* part = agg + F;                 // [ZC_THREADS]       column sums per row group (pass 1)
  float* part2 = part + ZC_THREADS;      // [4 ZC_THREADS]     partial G per row lane (pass 2)
  float* u = part2 + 4 * ZC_THREADS;     // [H1]  W_rel1 x_e
  float* zt = u + H1;                    // [H1]
  float* ht = zt + H1;                   // [H1]
  float* G = ht + H1;                    // [H1]
  uint32_t* rowmask = reinterpret_cast<uint32_t*>(G + H1);   // [W]
  int* hitf = reinterpret_cast<int*>(rowmask + W);           // [tile_rows]
  float* tile = reinterpret_cast<float*>(hitf + a.tile_rows);   // [tile_rows][F + 1]

  const int b = blockIdx.x, tid = threadIdx.x;
  const int cnt = __ldcg(a.st.count + b);
  const int tpos = cnt;
  const int lt = min(cnt, N - 1);
  const int tslot = gcm_slot(tpos, C);
  const bool evict = cnt >= N;
  const int eslot = evict ? gcm_slot(cnt - N, C) : 0;
  float* nodes_b = a.st.nodes + (size_t)b * C * F;
  uint32_t* masks_b = a.st.masks + (size_t)b * C * 2 * W;
  float* zc_b = a.zcache + (size_t)b * C * H1;
  const gcm_selector& sel = a.sel;

  for (int f = tid; f < F; f += ZC_THREADS) {
    cur[f] = a.obs[(size_t)b * F + f];
    xe[f] = evict ? nodes_b[(size_t)eslot * F + f] : 0.0f;
  }
  for (int w = tid; w < W; w += ZC_THREADS) rowmask[w] = 0u;
  __syncthreads();
  for (int f = tid; f < F; f += ZC_THREADS) nodes_b[(size_t)tslot * F + f] = cur[f];   // node write (gcm.py:274)

  // ---- pass 1: distance + threshold fused with the aggregation of t's in-neighbours ----
  // Rows are staged through shared memory in tiles (coalesced 16-byte loads); ONE THREAD PER ROW forms the distance
  // (no shuffle reductions: the warp-per-row version spent ~100 warp instructions per row and the kernel was
  // issue-bound, profiles/c4_step_dist_zc_r1.md), then thread = (feature column, row group) adds up the selected
  // rows of the tile.
  {
    const bool learned = sel.dist_param != nullptr;
    const float thr = learned ? 1.0f : sel.max_distance;
    const float scale = learned ? 1.0f / fabsf(__ldg(sel.dist_param)) : 1.0f;
    const int TR = a.tile_rows, FS = F + 1;                 // tile row stride: odd -> conflict-free column walks
    float cur_norm = 0.0f;
    if (sel.kind == GCM_SEL_COSINE) {
      float sq = 0.0f;
      for (int f = 0; f < F; ++f) sq = fmaf(cur[f], cur[f], sq);
      cur_norm = fmaxf(sqrtf(sq), 1e-8f);
    }
    const int F4 = F >> 2;                                  // host: F % 4 == 0 on this path
    const int ngrp = ZC_THREADS / F, col = tid % F, grp = tid / F;   // column sums: ngrp row groups
    const int rows_per_grp = (TR + ngrp - 1) / ngrp;
    float acc = 0.0f;
    const int half = tid & 1, prow = tid >> 1;              // two threads per row: each takes half of the features
    const int fh = (F + 1) >> 1, f_lo = half * fh, f_hi = min(F, f_lo + fh);
    for (int d0 = 1; d0 <= lt; d0 += TR) {
      const int nr = min(TR, lt - d0 + 1);
      // (fetching the next tile into registers during the work on this one cost a CTA of occupancy, 80+ registers, and
      // was slower: 0.63 vs 0.49 ms at cfg4; four resident CTAs hide the load latency better)
      for (int i = tid; i < nr * F4; i += ZC_THREADS) {
        const int r = i / F4, c4 = i - r * F4;
        const int d = d0 + r;
        const int slot = tslot - d + (tslot - d < 0 ? C : 0);             // (tpos - d) mod C without a division
        const float4 v = *reinterpret_cast<const float4*>(nodes_b + (size_t)slot * F + c4 * 4);
        float* dst = tile + r * FS + c4 * 4;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
      __syncthreads();
      {
        const bool rok = prow < nr;                        // warp-uniform up to the last partial warp; shuffles below
        const float* row = tile + (rok ? prow : 0) * FS;   // are executed by all lanes
        const int d = d0 + prow;
        float dist = 0.0f;
        if (sel.kind == GCM_SEL_EUCLIDEAN) {
          const int slot = tslot - d + (tslot - d < 0 ? C : 0);
          dist = rok ? __ldg(sel.dist + (size_t)b * C + slot) : 0.0f;
        } else if (sel.kind == GCM_SEL_COSINE) {
          float dot = 0.0f, nb = 0.0f;
#pragma unroll 8
          for (int f = f_lo; f < f_hi; ++f) {
            const float v = row[f];
            dot = fmaf(cur[f], v, dot);
            nb = fmaf(v, v, nb);
          }
          dot += __shfl_xor_sync(GCM_FULL_MASK, dot, 1);
          nb += __shfl_xor_sync(GCM_FULL_MASK, nb, 1);
          dist = dot / (cur_norm * fmaxf(sqrtf(nb), 1e-8f));
        } else {  // spatial
          float sq = 0.0f;
          for (int k = half; k < sel.slice_len; k += 2) {
            const float df = cur[sel.a_start + k * sel.a_step] - row[sel.b_start + k * sel.b_step];
            sq = fmaf(df, df, sq);
          }
          sq += __shfl_xor_sync(GCM_FULL_MASK, sq, 1);
          dist = sqrtf(sq) * scale;
        }
        if (rok && half == 0) {
          const bool hit = dist < thr;
          hitf[prow] = hit ? 1 : 0;
          if (hit) atomicOr(&rowmask[d >> 5], 1u << (d & 31));
        }
      }
      __syncthreads();
      if (grp < ngrp) {
        const int r1 = min(nr, (grp + 1) * rows_per_grp);
        for (int r = grp * rows_per_grp; r < r1; ++r)
          if (hitf[r]) acc += tile[r * FS + col];
      }
      __syncthreads();
    }
    if (grp < ngrp) part[grp * F + col] = acc;
    __syncthreads();
    for (int f = tid; f < F; f += ZC_THREADS) {
      float sum = 0.0f;
      for (int g = 0; g < ngrp; ++g) sum += part[g * F + f];
      agg[f] = sum;
    }
  }
  // commit row t of the adjacency (a recycled slot is fully overwritten) and the counter
  for (int w = tid; w < W; w += ZC_THREADS) {
    gcm_st_mask(masks_b + ((size_t)tslot * 2 + 0) * W + w, rowmask[w]);
    gcm_st_mask(masks_b + ((size_t)tslot * 2 + 1) * W + w, 0u);
  }
  if (tid == 0) __stcg(a.st.count + b, cnt + 1);
  __syncthreads();

  // ---- z_t = W_rel1 agg + W_root1 x_t + b1 ;  u = W_rel1 x_e   (thread = channel, K-major weight pack) ----
  if (tid < H1) {
    float z = a.gnn.b1 ? __ldg(a.gnn.b1 + tid) : 0.0f, uu = 0.0f;
    const float* w1 = a.gnn.w1t + tid;
    for (int k = 0; k < F; ++k) {
      const float wr = __ldg(w1 + (size_t)k * H1);
      z = fmaf(wr, agg[k], z);
      uu = fmaf(wr, xe[k], uu);
    }
    for (int k = 0; k < F; ++k) z = fmaf(__ldg(w1 + (size_t)(F + k) * H1), cur[k], z);
    zt[tid] = z;
    u[tid] = uu;
    zc_b[(size_t)tslot * H1 + tid] = z;
    ht[tid] = gcm_act_fast(z, a.gnn.act1);
  }
  __syncthreads();

  // ---- pass 2 over the cached rows: eviction correction + layer-2 aggregation ----
  // thread = (row lane, 4 channels): one 16-byte load per cached row and thread, 4 rows in flight per thread
  {
    const int tpr = H1 >> 2;                                  // host: H1 % 4 == 0 on this path
    const int rpp = ZC_THREADS / tpr, rr = tid / tpr, vl = tid - rr * tpr;
    float g4[4] = {0.f, 0.f, 0.f, 0.f};
    if (rr < rpp) {
      const float4 u4 = *reinterpret_cast<const float4*>(u + vl * 4);
      const int off_bit0 = N;    // offset of e in row (tpos - d)'s past mask is N - d
      constexpr int U = 4;
      for (int d0 = 1 + rr; d0 <= lt; d0 += rpp * U) {
        float4 z[U];
        uint32_t m[U];
        int slot[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int d = d0 + q * rpp;
          z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          m[q] = 0u;
          slot[q] = 0;
          if (d <= lt) {
            slot[q] = tslot - d + (tslot - d < 0 ? C : 0);       // (tpos - d) mod C without a division
            z[q] = *reinterpret_cast<const float4*>(zc_b + (size_t)slot[q] * H1 + vl * 4);
            if (evict) m[q] = gcm_ld_mask(masks_b + ((size_t)slot[q] * 2 + 0) * W + ((off_bit0 - d) >> 5));
          }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int d = d0 + q * rpp;
          if (d <= lt) {
            if (evict && ((m[q] >> ((off_bit0 - d) & 31)) & 1u)) {
              z[q].x -= u4.x; z[q].y -= u4.y; z[q].z -= u4.z; z[q].w -= u4.w;
              *reinterpret_cast<float4*>(zc_b + (size_t)slot[q] * H1 + vl * 4) = z[q];
            }
            if ((rowmask[d >> 5] >> (d & 31)) & 1u) {
              float hv[4] = {z[q].x, z[q].y, z[q].z, z[q].w};
              gcm_act_fast_vec(hv, a.gnn.act1);
              g4[0] += hv[0]; g4[1] += hv[1]; g4[2] += hv[2]; g4[3] += hv[3];
            }
          }
        }
      }
      *reinterpret_cast<float4*>(part2 + (size_t)rr * H1 + vl * 4) = make_float4(g4[0], g4[1], g4[2], g4[3]);
    }
    __syncthreads();
    if (tid < H1) {
      float sum = 0.0f;
      for (int r = 0; r < rpp; ++r) sum += part2[r * H1 + tid];
      G[tid] = sum;
    }
  }
  __syncthreads();

  // ---- belief = act2(W_rel2 G + W_root2 h_t + b2) ----
  if (tid < H2) {
    float o = a.gnn.b2 ? __ldg(a.gnn.b2 + tid) : 0.0f;
    const float* w2 = a.gnn.w2t + tid;
    for (int k = 0; k < H1; ++k) o = fmaf(__ldg(w2 + (size_t)k * H2), G[k], o);
    for (int k = 0; k < H1; ++k) o = fmaf(__ldg(w2 + (size_t)(H1 + k) * H2), ht[k], o);
    o = gcm_act_fast(o, a.gnn.act2);
    a.belief[(size_t)b * H2 + tid] = o;
    if (!isfinite(o)) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
  }
}

extern "C" int gcm_dense_step_fwd_zc(const gcm_dense_state* st, const float* obs, const gcm_selector* sel,
                                     const gcm_gnn* gnn, float* zcache, float* belief, int32_t* status,
                                     void* stream) {
  GCM_REQUIRE(st && st->nodes && st->masks && st->count && obs && sel && gnn && zcache && belief && status,
              "dense_step_fwd_zc: null pointer");
  GCM_REQUIRE(sel->kind == GCM_SEL_EUCLIDEAN || sel->kind == GCM_SEL_COSINE || sel->kind == GCM_SEL_SPATIAL,
              "dense_step_fwd_zc: selector kind %d is not a distance selector", sel->kind);
  GCM_REQUIRE(sel->kind != GCM_SEL_EUCLIDEAN || sel->dist, "dense_step_fwd_zc: euclidean needs dist");
  GCM_REQUIRE(st->F >= 1 && st->F <= 128 && gnn->F == st->F && gnn->H1 >= 1 && gnn->H1 <= 128 && gnn->H2 >= 1 &&
                  gnn->H2 <= 128 && gnn->w1t && gnn->w2t,
              "dense_step_fwd_zc: F, H1, H2 must be in [1,128]");
  GCM_REQUIRE(st->N >= 1 && st->N <= GCM_MAX_N && st->C >= st->N && st->W == (st->N + 31) / 32,
              "dense_step_fwd_zc: bad state");
  if (sel->kind == GCM_SEL_SPATIAL)
    GCM_REQUIRE(sel->slice_len >= 0 && sel->a_step >= 1 && sel->b_step >= 1 && sel->a_start >= 0 && sel->b_start >= 0 &&
                    (sel->slice_len == 0 || (sel->a_start + (sel->slice_len - 1) * sel->a_step < st->F &&
                                             sel->b_start + (sel->slice_len - 1) * sel->b_step < st->F)),
                "dense_step_fwd_zc: spatial slice outside [0,F)");
  if (st->B == 0) return GCM_OK;
  ZcArgs a;
  a.st = *st; a.obs = obs; a.sel = *sel; a.gnn = *gnn; a.zcache = zcache; a.belief = belief; a.status = status;
  const int F = st->F, H1 = gnn->H1;
  GCM_REQUIRE(F % 4 == 0 && H1 % 4 == 0, "dense_step_fwd_zc: F and H1 must be multiples of 4");
  a.tile_rows = F <= 64 ? 128 : 64;
  const size_t smem = ((size_t)3 * F + 5 * ZC_THREADS + 4 * H1 + st->W + a.tile_rows + (size_t)a.tile_rows * (F + 1) + 8) * 4;
  const int fr = (F + 31) / 32;
  cudaStream_t s = (cudaStream_t)stream;
  switch (fr) {
    case 1: k_step_dist_zc<1><<<st->B, ZC_THREADS, smem, s>>>(a); break;
    case 2: k_step_dist_zc<2><<<st->B, ZC_THREADS, smem, s>>>(a); break;
    case 3: k_step_dist_zc<3><<<st->B, ZC_THREADS, smem, s>>>(a); break;
    default: k_step_dist_zc<4><<<st->B, ZC_THREADS, smem, s>>>(a); break;
  }
  return gcm_check_launch("k_step_dist_zc");
}

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
//                                (the rows are in registers anyway); then z_t, and u = W_rel1 x_e if a node leaves
//   pass 2 over the cached rows: z_i -= u for the rows that had e as in-neighbour (one mask bit each), and
//                                G = sum_{i in N(t)} act1(z_i);  belief = act2(W_rel2 G + W_root2 act1(z_t) + b2)
// = n (F + 2 H1) 4 bytes per graph-step instead of an n x n/2 gather plus a [n, 2F] x [2F, H1] product.
// Valid while the weights stay the same and every step of the state went through this kernel (the host tracks
// both, gcm/fused.py); training (autograd) uses the general kernels.  One CTA per graph, 256 threads.
#include "gcm_common.cuh"

constexpr int ZC_THREADS = 256;
constexpr int ZC_NW = ZC_THREADS / 32;

struct ZcArgs {
  gcm_dense_state st;
  const float* obs;
  gcm_selector sel;
  gcm_gnn gnn;
  float* zcache;     // [B, C, H1]
  float* belief;
  int32_t* status;
};

template <int FR>
__global__ void __launch_bounds__(ZC_THREADS) k_step_dist_zc(const ZcArgs a) {
  extern __shared__ __align__(16) float zs[];
  const int F = a.st.F, N = a.st.N, C = a.st.C, W = a.st.W, H1 = a.gnn.H1, H2 = a.gnn.H2;
  float* cur = zs;                       // [F]   the observation
  float* xe = cur + F;                   // [F]   the node that leaves the window
  float* agg = xe + F;                   // [F]   sum of t's in-neighbours
  float* part = agg + F;                 // [ZC_NW][F] / [8][H1] partial sums
  const int pw = F > H1 ? F : H1;
  float* u = part + ZC_NW * pw;          // [H1]  W_rel1 x_e
  float* zt = u + H1;                    // [H1]
  float* ht = zt + H1;                   // [H1]
  float* G = ht + H1;                    // [H1]
  uint32_t* rowmask = reinterpret_cast<uint32_t*>(G + H1);   // [W]

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

  // ---- pass 1: distance + threshold (same arithmetic as gcm_select_distance) fused with the aggregation ----
  {
    const bool learned = sel.dist_param != nullptr;
    const float thr = learned ? 1.0f : sel.max_distance;
    const float scale = learned ? 1.0f / fabsf(__ldg(sel.dist_param)) : 1.0f;
    float cur_norm = 0.0f;
    if (sel.kind == GCM_SEL_COSINE) {
      float s = 0.0f;
      for (int f = lane; f < F; f += 32) s += cur[f] * cur[f];
      cur_norm = fmaxf(sqrtf(gcm_warp_sum(s)), 1e-8f);
    }
    float acc[FR];
#pragma unroll
    for (int k = 0; k < FR; ++k) acc[k] = 0.0f;
    // 4 rows per warp iteration: their loads are issued together and the shuffle reductions interleave
    constexpr int RU = 4;
    for (int d0 = 1 + warp * RU; d0 <= lt; d0 += ZC_NW * RU) {
      float v[RU][FR];
      float dist[RU];
      const float* rows[RU];
      int slots[RU];
#pragma unroll
      for (int q = 0; q < RU; ++q) {
        const int d = min(d0 + q, lt);                                 // clamped rows are recomputed, not used
        slots[q] = tslot - d + (tslot - d < 0 ? C : 0);                // (tpos - d) mod C without a division
        rows[q] = nodes_b + (size_t)slots[q] * F;
#pragma unroll
        for (int k = 0; k < FR; ++k) {
          const int f = lane + 32 * k;
          v[q][k] = f < F ? rows[q][f] : 0.0f;
        }
      }
      if (sel.kind == GCM_SEL_EUCLIDEAN) {
#pragma unroll
        for (int q = 0; q < RU; ++q) dist[q] = __ldg(sel.dist + (size_t)b * C + slots[q]);
      } else if (sel.kind == GCM_SEL_COSINE) {
        float dot[RU], nb[RU];
#pragma unroll
        for (int q = 0; q < RU; ++q) {
          dot[q] = 0.0f;
          nb[q] = 0.0f;
#pragma unroll
          for (int k = 0; k < FR; ++k) {
            const int f = lane + 32 * k;
            if (f < F) {
              dot[q] += cur[f] * v[q][k];
              nb[q] += v[q][k] * v[q][k];
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < RU; ++q) {
            dot[q] += __shfl_xor_sync(GCM_FULL_MASK, dot[q], o);
            nb[q] += __shfl_xor_sync(GCM_FULL_MASK, nb[q], o);
          }
        }
#pragma unroll
        for (int q = 0; q < RU; ++q) dist[q] = dot[q] / (cur_norm * fmaxf(sqrtf(nb[q]), 1e-8f));
      } else {  // spatial
        float sq[RU];
#pragma unroll
        for (int q = 0; q < RU; ++q) {
          sq[q] = 0.0f;
          for (int k = lane; k < sel.slice_len; k += 32) {
            const float df = cur[sel.a_start + k * sel.a_step] - rows[q][sel.b_start + k * sel.b_step];
            sq[q] += df * df;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < RU; ++q) sq[q] += __shfl_xor_sync(GCM_FULL_MASK, sq[q], o);
        }
#pragma unroll
        for (int q = 0; q < RU; ++q) dist[q] = sqrtf(sq[q]) * scale;
      }
#pragma unroll
      for (int q = 0; q < RU; ++q) {
        const int d = d0 + q;
        if (d <= lt && dist[q] < thr) {          // warp-uniform: every lane holds the reduced value
#pragma unroll
          for (int k = 0; k < FR; ++k) acc[k] += v[q][k];
          if (lane == 0) atomicOr(&rowmask[d >> 5], 1u << (d & 31));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < FR; ++k) {
      const int f = lane + 32 * k;
      if (f < F) part[warp * F + f] = acc[k];
    }
  }
  __syncthreads();
  for (int f = tid; f < F; f += ZC_THREADS) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < ZC_NW; ++w) s += part[w * F + f];
    agg[f] = s;
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
  {
    const int hp = H1 <= 32 ? 32 : (H1 <= 64 ? 64 : 128);   // threads per row
    const int rl = ZC_THREADS / hp, rr = tid / hp, ch = tid - rr * hp;
    float g = 0.0f;
    if (ch < H1) {
      // 8 rows per iteration: all loads (cached row + the one mask word that says whether the leaving node was an
      // in-neighbour) are issued before the first use, so a thread keeps 16 requests in flight
      constexpr int U = 8;
      const int off_bit0 = N;    // offset of e in row (tpos - d)'s past mask is N - d
      for (int d0 = 1 + rr; d0 <= lt; d0 += rl * U) {
        float z[U];
        uint32_t m[U];
        int slot[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int d = d0 + q * rl;
          z[q] = 0.0f;
          m[q] = 0u;
          slot[q] = 0;
          if (d <= lt) {
            slot[q] = tslot - d + (tslot - d < 0 ? C : 0);       // (tpos - d) mod C without a division
            z[q] = zc_b[(size_t)slot[q] * H1 + ch];
            if (evict) m[q] = gcm_ld_mask(masks_b + ((size_t)slot[q] * 2 + 0) * W + ((off_bit0 - d) >> 5));
          }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int d = d0 + q * rl;
          if (d <= lt) {
            if (evict && ((m[q] >> ((off_bit0 - d) & 31)) & 1u)) {
              z[q] -= u[ch];
              zc_b[(size_t)slot[q] * H1 + ch] = z[q];
            }
            if ((rowmask[d >> 5] >> (d & 31)) & 1u) g += gcm_act_fast(z[q], a.gnn.act1);
          }
        }
      }
      part[rr * H1 + ch] = g;
    }
    __syncthreads();
    if (tid < H1) {
      float s = 0.0f;
      for (int r = 0; r < rl; ++r) s += part[r * H1 + tid];
      G[tid] = s;
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
  const int pw = F > H1 ? F : H1;
  const size_t smem = ((size_t)3 * F + (size_t)ZC_NW * pw + 4 * H1 + st->W + 8) * 4;
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

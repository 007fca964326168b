// Dense GCM step, forward: node write + edge-selector bitmask update + 2-layer DenseGraphConv
// restricted to the 2-hop in-neighbourhood of the new node + belief extraction.
//
// Replaces (reference, /root/reference/src/gcm): gcm.py:262-321, edge_selectors/temporal.py:72-88,
// edge_selectors/dense.py:11-23, edge_selectors/distance.py:18-81 and the torch_geometric
// DenseGraphConv x2 + activation the user GNN runs at gcm.py:308 (README.md:52-62).
//
// Two kernels:
//   k_step_general<FR,HR>  one CTA per graph, any selector chain, adjacency read from the bitmasks.
//   k_step_temporal<F>     one warp per graph, persistent; "pure temporal" states only (implicit
//                          adjacency: the neighbour offsets are a static program), H1 = H2 = 32,
//                          layer weights resident in registers.
#include <stdlib.h>

#include "gcm_common.cuh"
#include "gcm_temporal.cuh"

// ------------------------------------------------------------------------------------------------
// general kernel
// ------------------------------------------------------------------------------------------------
constexpr int GEN_THREADS = 256;
constexpr int GEN_NW = GEN_THREADS / 32;

struct DenseStepArgs {
  gcm_dense_state st;
  const float* obs;
  gcm_selector sels[GCM_MAX_SELECTORS];
  int n_sels;
  gcm_gnn gnn;
  float* belief;
  int32_t* status;
};

// distance selectors: connect t to every older in-window node with d < threshold (strict)
__device__ void gcm_select_distance(const gcm_selector& sel, const float* cur, const float* nodes_b,
                                    int b, int tpos, int lt, int C, int F, uint32_t* rowmask, int warp,
                                    int lane, int nwarps) {
  const bool learned = sel.dist_param != nullptr;
  const float thr = learned ? 1.0f : sel.max_distance;
  const float scale = learned ? 1.0f / fabsf(__ldg(sel.dist_param)) : 1.0f;
  float cur_norm = 0.0f;
  if (sel.kind == GCM_SEL_COSINE) {
    float s = 0.0f;
    for (int f = lane; f < F; f += 32) s += cur[f] * cur[f];
    cur_norm = fmaxf(sqrtf(gcm_warp_sum(s)), 1e-8f);
  }
  for (int d = 1 + warp; d <= lt; d += nwarps) {
    const int slot = gcm_slot(tpos - d, C);
    const float* row = nodes_b + (size_t)slot * F;
    float dist;
    if (sel.kind == GCM_SEL_EUCLIDEAN) {
      dist = __ldg(sel.dist + (size_t)b * C + slot);
    } else if (sel.kind == GCM_SEL_COSINE) {
      float dot = 0.0f, nb = 0.0f;
      for (int f = lane; f < F; f += 32) {
        float v = row[f];
        dot += cur[f] * v;
        nb += v * v;
      }
      dot = gcm_warp_sum(dot);
      nb = fmaxf(sqrtf(gcm_warp_sum(nb)), 1e-8f);
      dist = dot / (cur_norm * nb);
    } else {  // spatial
      float s = 0.0f;
      for (int k = lane; k < sel.slice_len; k += 32) {
        float df = cur[sel.a_start + k * sel.a_step] - row[sel.b_start + k * sel.b_step];
        s += df * df;
      }
      dist = sqrtf(gcm_warp_sum(s)) * scale;
    }
    if (lane == 0 && dist < thr) atomicOr(&rowmask[d >> 5], 1u << (d & 31));
  }
}

template <int FR, int HR>
__global__ void __launch_bounds__(GEN_THREADS) k_step_general(const DenseStepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = a.st.F, N = a.st.N, C = a.st.C, W = a.st.W, H1 = a.gnn.H1, H2 = a.gnn.H2;
  float* cur = reinterpret_cast<float*>(smem_raw);
  float* sall = cur + F;
  float* h1t = sall + F;
  float* agg2 = h1t + H1;
  float* agg2part = agg2 + H1;              // [GEN_NW][H1]
  float* aggx = agg2part + GEN_NW * H1;     // [GEN_NW][2F]
  uint32_t* rowmask = reinterpret_cast<uint32_t*>(aggx + GEN_NW * 2 * F);  // [W]
  int* r1n = reinterpret_cast<int*>(rowmask + W);                          // [2]
  uint16_t* r1 = reinterpret_cast<uint16_t*>(r1n + 2);                     // [N]

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cnt = __ldcg(a.st.count + b);
  const int tpos = cnt;
  const int lt = min(cnt, N - 1);  // logical index of the new node (after the wrap, gcm.py:263-271)
  const int tslot = gcm_slot(tpos, C);
  float* nodes_b = a.st.nodes + (size_t)b * C * F;
  uint32_t* masks_b = a.st.masks + (size_t)b * C * 2 * W;

  // ---- node write (gcm.py:274) ----
  for (int f = tid; f < F; f += GEN_THREADS) {
    float v = a.obs[(size_t)b * F + f];
    cur[f] = v;
    nodes_b[(size_t)tslot * F + f] = v;
  }
  for (int w = tid; w < W; w += GEN_THREADS) rowmask[w] = 0u;
  __syncthreads();

  // ---- edge selectors (gcm.py:284-287); chained selectors OR together ----
  for (int s = 0; s < a.n_sels; ++s) {
    const gcm_selector& sel = a.sels[s];
    if (sel.kind == GCM_SEL_TEMPORAL) {
      if (tid < sel.n_hops) {
        const int hop = sel.hops[tid];
        if (hop >= 0 && hop <= lt) {
          const uint32_t bit = 1u << (hop & 31);
          if (sel.direction != GCM_DIR_BACKWARD || hop == 0) atomicOr(&rowmask[hop >> 5], bit);
          if (sel.direction != GCM_DIR_FORWARD && hop > 0) {
            const int slot = gcm_slot(tpos - hop, C);
            atomicOr(masks_b + ((size_t)slot * 2 + 1) * W + (hop >> 5), bit);
          }
        }
      }
    } else if (sel.kind == GCM_SEL_DENSE) {
      for (int d = tid; d <= lt; d += GEN_THREADS) {
        const uint32_t bit = 1u << (d & 31);
        atomicOr(&rowmask[d >> 5], bit);
        if (d > 0) {
          const int slot = gcm_slot(tpos - d, C);
          atomicOr(masks_b + ((size_t)slot * 2 + 1) * W + (d >> 5), bit);
        }
      }
    } else if (sel.kind != GCM_SEL_NONE) {
      gcm_select_distance(sel, cur, nodes_b, b, tpos, lt, C, F, rowmask, warp, lane, GEN_NW);
    }
    __syncthreads();
  }

  // ---- commit row t (a recycled slot is fully overwritten) and the counter ----
  for (int w = tid; w < W; w += GEN_THREADS) {
    gcm_st_mask(masks_b + ((size_t)tslot * 2 + 0) * W + w, rowmask[w]);
    gcm_st_mask(masks_b + ((size_t)tslot * 2 + 1) * W + w, 0u);
  }
  if (tid == 0) __stcg(a.st.count + b, cnt + 1);

  // ---- R1 = {t} U in-neighbours(t), as offsets d from t ----
  if (warp == 0) {
    int base = 1;
    if (lane == 0) r1[0] = 0;
    for (int w = 0; w < W; ++w) {
      const uint32_t m = rowmask[w] & gcm_range_word(w, 1, lt);
      const bool on = (m >> lane) & 1u;
      if (on) r1[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(w * 32 + lane);
      base += __popc(m);
    }
    if (lane == 0) {
      r1n[0] = base;
      r1n[1] = (int)(rowmask[0] & 1u);
    }
  }
  __threadfence();
  __syncthreads();
  const int nR = r1n[0];
  const bool selfloop = r1n[1] != 0;

  // ---- sum of all in-window rows, for the complement trick on dense rows ----
  const bool use_sall = nR > 16;
  if (use_sall) {
    float acc[FR];
#pragma unroll
    for (int k = 0; k < FR; ++k) acc[k] = 0.0f;
    for (int d = warp; d <= lt; d += GEN_NW) {
      const float* row = nodes_b + (size_t)gcm_slot(tpos - d, C) * F;
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int f = lane + 32 * k;
        if (f < F) acc[k] += row[f];
      }
    }
#pragma unroll
    for (int k = 0; k < FR; ++k) {
      const int f = lane + 32 * k;
      if (f < F) aggx[warp * 2 * F + f] = acc[k];
    }
    __syncthreads();
    for (int f = tid; f < F; f += GEN_THREADS) {
      float s = 0.0f;
      for (int w = 0; w < GEN_NW; ++w) s += aggx[w * 2 * F + f];
      sall[f] = s;
    }
    __syncthreads();
  }

  // ---- layer 1 on the rows of R1, one warp per row ----
  float a2[HR];
#pragma unroll
  for (int k = 0; k < HR; ++k) a2[k] = 0.0f;
  float* my = aggx + warp * 2 * F;
  for (int ri = warp; ri < nR; ri += GEN_NW) {
    const int d = r1[ri];
    const int pos = tpos - d;
    const int slot = gcm_slot(pos, C);
    const int lj = lt - d;
    const uint32_t* mrow = masks_b + (size_t)slot * 2 * W;
    uint32_t pw = 0u, fw = 0u;
    if (lane < W) {
      pw = gcm_ld_mask(mrow + lane) & gcm_range_word(lane, 0, lj);
      fw = gcm_ld_mask(mrow + W + lane) & gcm_range_word(lane, 1, d);
    }
    const int deg = gcm_warp_sum_int(__popc(pw) + __popc(fw));
    const bool comp = use_sall && (2 * deg > lt + 1);
    float sign = 1.0f;
    float acc[FR];
#pragma unroll
    for (int k = 0; k < FR; ++k) acc[k] = 0.0f;
    if (comp) {
      sign = -1.0f;
      if (lane < W) {
        pw = ~pw & gcm_range_word(lane, 0, lj);
        fw = ~fw & gcm_range_word(lane, 1, d);
      }
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int f = lane + 32 * k;
        if (f < F) acc[k] = sall[f];
      }
    }
    for (int w = 0; w < W; ++w) {
      uint32_t m = __shfl_sync(GCM_FULL_MASK, pw, w);
      while (m) {
        const int e = w * 32 + __ffs(m) - 1;
        m &= m - 1;
        const float* row = nodes_b + (size_t)gcm_slot(pos - e, C) * F;
#pragma unroll
        for (int k = 0; k < FR; ++k) {
          const int f = lane + 32 * k;
          if (f < F) acc[k] += sign * row[f];
        }
      }
      m = __shfl_sync(GCM_FULL_MASK, fw, w);
      while (m) {
        const int e = w * 32 + __ffs(m) - 1;
        m &= m - 1;
        const float* row = nodes_b + (size_t)gcm_slot(pos + e, C) * F;
#pragma unroll
        for (int k = 0; k < FR; ++k) {
          const int f = lane + 32 * k;
          if (f < F) acc[k] += sign * row[f];
        }
      }
    }
    {
      const float* xrow = nodes_b + (size_t)slot * F;
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int f = lane + 32 * k;
        if (f < F) {
          my[f] = acc[k];
          my[F + f] = xrow[f];
        }
      }
    }
    __syncwarp();
    float z[HR];
#pragma unroll
    for (int k = 0; k < HR; ++k) {
      const int h = lane + 32 * k;
      z[k] = (a.gnn.b1 != nullptr && h < H1) ? __ldg(a.gnn.b1 + h) : 0.0f;
    }
    for (int kk = 0; kk < 2 * F; ++kk) {
      const float av = my[kk];
      const float* wr = a.gnn.w1t + (size_t)kk * H1;
#pragma unroll
      for (int k = 0; k < HR; ++k) {
        const int h = lane + 32 * k;
        if (h < H1) z[k] = fmaf(av, __ldg(wr + h), z[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < HR; ++k) {
      const int h = lane + 32 * k;
      if (h < H1) {
        const float hv = gcm_act_fwd(z[k], a.gnn.act1);
        if (d == 0) h1t[h] = hv;
        if (d > 0 || selfloop) a2[k] += hv;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < HR; ++k) {
    const int h = lane + 32 * k;
    if (h < H1) agg2part[warp * H1 + h] = a2[k];
  }
  __syncthreads();
  for (int k = tid; k < H1; k += GEN_THREADS) {
    float s = 0.0f;
    for (int w = 0; w < GEN_NW; ++w) s += agg2part[w * H1 + k];
    agg2[k] = s;
  }
  __syncthreads();

  // ---- layer 2 on row t only, belief (gcm.py:314) + finite flag (gcm.py:316-318) ----
  for (int h2 = tid; h2 < H2; h2 += GEN_THREADS) {
    float z = a.gnn.b2 != nullptr ? __ldg(a.gnn.b2 + h2) : 0.0f;
    for (int k = 0; k < H1; ++k) {
      z = fmaf(agg2[k], __ldg(a.gnn.w2t + (size_t)k * H2 + h2), z);
      z = fmaf(h1t[k], __ldg(a.gnn.w2t + (size_t)(H1 + k) * H2 + h2), z);
    }
    const float out = gcm_act_fwd(z, a.gnn.act2);
    a.belief[(size_t)b * H2 + h2] = out;
    if (!isfinite(out)) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
  }
}

static size_t gen_smem_bytes(const gcm_dense_state& st, const gcm_gnn& g) {
  size_t fl = (size_t)2 * st.F + 2 * g.H1 + (size_t)GEN_NW * g.H1 + (size_t)GEN_NW * 2 * st.F;
  size_t bytes = fl * 4 + (size_t)st.W * 4 + 8 + (size_t)st.N * 2;
  return (bytes + 15) & ~(size_t)15;
}

template <int FR, int HR>
static int launch_general(const DenseStepArgs& a, cudaStream_t stream) {
  size_t smem = gen_smem_bytes(a.st, a.gnn);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_step_general<FR, HR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(general): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
  }
  k_step_general<FR, HR><<<a.st.B, GEN_THREADS, smem, stream>>>(a);
  return gcm_check_launch("k_step_general");
}

template <int FR>
static int launch_general_h(const DenseStepArgs& a, int hr, cudaStream_t stream) {
  switch (hr) {
    case 1: return launch_general<FR, 1>(a, stream);
    case 2: return launch_general<FR, 2>(a, stream);
    case 3:
    case 4: return launch_general<FR, 4>(a, stream);
    default: return launch_general<FR, 8>(a, stream);
  }
}

// ------------------------------------------------------------------------------------------------
// pure-temporal fast kernel
// ------------------------------------------------------------------------------------------------
struct TemporalArgs {
  gcm_dense_state st;
  const float* obs;
  gcm_gnn gnn;
  float* belief;
  int32_t* status;
  TemporalProg prog;
};

template <int F>
__global__ void __launch_bounds__(TP_THREADS, 1) k_step_temporal(const TemporalArgs a) {
  constexpr int H = 32;
  constexpr int K1 = 2 * F;
  extern __shared__ __align__(16) float tp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = a.st.N, C = a.st.C, W = a.st.W, B = a.st.B;
  const TemporalProg& P = a.prog;
  const int per_warp = P.nD * F + 4 * K1 + 2 * H;
  float* xs = tp_smem + (size_t)warp * per_warp;  // [nD][F]
  float* aggx = xs + P.nD * F;                    // [4][K1]
  float* l2in = aggx + 4 * K1;                    // [2H]

  // layer weights: lane h keeps column h of both K-major packs in registers
  float w1[K1], w2[2 * H];
#pragma unroll
  for (int k = 0; k < K1; ++k) w1[k] = __ldg(a.gnn.w1t + k * H + lane);
#pragma unroll
  for (int k = 0; k < 2 * H; ++k) w2[k] = __ldg(a.gnn.w2t + k * H + lane);
  const float bias1 = a.gnn.b1 ? __ldg(a.gnn.b1 + lane) : 0.0f;
  const float bias2 = a.gnn.b2 ? __ldg(a.gnn.b2 + lane) : 0.0f;
  const int act1 = a.gnn.act1, act2 = a.gnn.act2;

  const int gw = blockIdx.x * TP_NW + warp, nw_total = gridDim.x * TP_NW;
  for (int g = gw; g < B; g += nw_total) {
    const int cnt = __ldcg(a.st.count + g);
    const int tpos = cnt;
    const int lt = min(cnt, N - 1);
    const int tslot = gcm_slot(tpos, C);
    float* nodes_g = a.st.nodes + (size_t)g * C * F;
    uint32_t* masks_g = a.st.masks + (size_t)g * C * 2 * W;

    // node write + gather of the distinct rows of the 2-hop in-neighbourhood
    if (lane < F) {
      const float xv = a.obs[(size_t)g * F + lane];
      nodes_g[(size_t)tslot * F + lane] = xv;
      xs[lane] = xv;
    }
    for (int i = 1; i < P.nD; ++i) {
      const int o = P.doff[i];
      float v = 0.0f;
      if (o <= lt && lane < F) v = nodes_g[(size_t)gcm_slot(tpos - o, C) * F + lane];
      if (lane < F) xs[i * F + lane] = v;
    }
    // adjacency bits of row t (temporal.py:72-88) and the counter
    {
      uint32_t pw = 0u;
      for (int i = 0; i < P.n_past; ++i) {
        const int hop = P.past[i];
        if (hop <= lt && (hop >> 5) == lane) pw |= 1u << (hop & 31);
      }
      if (lane < W) {
        gcm_st_mask(masks_g + ((size_t)tslot * 2 + 0) * W + lane, pw);
        gcm_st_mask(masks_g + ((size_t)tslot * 2 + 1) * W + lane, 0u);
      }
      if (lane < P.n_future) {
        const int hop = P.future[lane];
        if (hop <= lt)
          atomicOr(masks_g + ((size_t)gcm_slot(tpos - hop, C) * 2 + 1) * W + (hop >> 5),
                   1u << (hop & 31));
      }
      if (lane == 0) __stcg(a.st.count + g, cnt + 1);
    }
    __syncwarp();

    float h1t = 0.0f, agg2 = 0.0f;
    for (int c0 = 0; c0 < P.nR; c0 += 4) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = c0 + rr;
        float ag = 0.0f, xv = 0.0f;
        if (r < P.nR && lane < F) {
          for (int q = 0; q < P.nnb[r]; ++q) ag += xs[P.nb[r][q] * F + lane];
          xv = xs[P.rD[r] * F + lane];
        }
        if (lane < F) {
          aggx[rr * K1 + lane] = ag;
          aggx[rr * K1 + F + lane] = xv;
        }
      }
      __syncwarp();
      float z0 = bias1, z1 = bias1, z2 = bias1, z3 = bias1;
#pragma unroll
      for (int k = 0; k < K1; k += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(aggx + 0 * K1 + k);
        const float4 a1 = *reinterpret_cast<const float4*>(aggx + 1 * K1 + k);
        const float4 a2 = *reinterpret_cast<const float4*>(aggx + 2 * K1 + k);
        const float4 a3 = *reinterpret_cast<const float4*>(aggx + 3 * K1 + k);
        z0 = fmaf(a0.x, w1[k], z0); z0 = fmaf(a0.y, w1[k + 1], z0);
        z0 = fmaf(a0.z, w1[k + 2], z0); z0 = fmaf(a0.w, w1[k + 3], z0);
        z1 = fmaf(a1.x, w1[k], z1); z1 = fmaf(a1.y, w1[k + 1], z1);
        z1 = fmaf(a1.z, w1[k + 2], z1); z1 = fmaf(a1.w, w1[k + 3], z1);
        z2 = fmaf(a2.x, w1[k], z2); z2 = fmaf(a2.y, w1[k + 1], z2);
        z2 = fmaf(a2.z, w1[k + 2], z2); z2 = fmaf(a2.w, w1[k + 3], z2);
        z3 = fmaf(a3.x, w1[k], z3); z3 = fmaf(a3.y, w1[k + 1], z3);
        z3 = fmaf(a3.z, w1[k + 2], z3); z3 = fmaf(a3.w, w1[k + 3], z3);
      }
      const float zz[4] = {z0, z1, z2, z3};
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = c0 + rr;
        if (r < P.nR) {
          const float hv = gcm_act_fwd(zz[rr], act1);
          if (r == 0) h1t = hv;
          else if (P.rd[r] <= lt) agg2 += hv;
        }
      }
      __syncwarp();
    }
    l2in[lane] = agg2;
    l2in[H + lane] = h1t;
    __syncwarp();
    float o0 = bias2, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
#pragma unroll
    for (int k = 0; k < 2 * H; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(l2in + k);
      o0 = fmaf(v.x, w2[k], o0);
      o1 = fmaf(v.y, w2[k + 1], o1);
      o2 = fmaf(v.z, w2[k + 2], o2);
      o3 = fmaf(v.w, w2[k + 3], o3);
    }
    const float out = gcm_act_fwd((o0 + o1) + (o2 + o3), act2);
    a.belief[(size_t)g * H + lane] = out;
    if (!isfinite(out)) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
    __syncwarp();
  }
}

// Build the static gather program of a TEMPORAL-only chain.  Returns false when the chain does not
// fit the fast kernel's limits (the general kernel then handles it).
static bool build_temporal_prog(const gcm_selector* sels, int n_sels, TemporalProg& P) {
  int past[GCM_MAX_HOPS], np = 0, fut[GCM_MAX_HOPS], nf = 0;
  auto add = [](int* arr, int& n, int v) {
    for (int i = 0; i < n; ++i)
      if (arr[i] == v) return true;
    if (n >= GCM_MAX_HOPS) return false;
    arr[n++] = v;
    return true;
  };
  for (int s = 0; s < n_sels; ++s) {
    if (sels[s].kind != GCM_SEL_TEMPORAL) return false;
    for (int i = 0; i < sels[s].n_hops; ++i) {
      const int hop = sels[s].hops[i];
      if (hop <= 0) return false;
      if (sels[s].direction != GCM_DIR_BACKWARD && !add(past, np, hop)) return false;
      if (sels[s].direction != GCM_DIR_FORWARD && !add(fut, nf, hop)) return false;
    }
  }
  if (1 + np > TP_MAXR || np + nf > TP_MAXNB) return false;
  P.nD = 1;
  P.doff[0] = 0;
  auto dindex = [&](int off) -> int {
    for (int i = 0; i < P.nD; ++i)
      if (P.doff[i] == off) return i;
    if (P.nD >= TP_MAXD) return -1;
    P.doff[P.nD] = off;
    return P.nD++;
  };
  P.nR = 1 + np;
  for (int r = 0; r < P.nR; ++r) {
    const int d = r == 0 ? 0 : past[r - 1];
    P.rd[r] = d;
    if ((P.rD[r] = dindex(d)) < 0) return false;
    int n = 0;
    for (int i = 0; i < np; ++i) {  // older sources: offset d + hop (zero row if outside the window)
      const int ix = dindex(d + past[i]);
      if (ix < 0) return false;
      P.nb[r][n++] = ix;
    }
    for (int i = 0; i < nf; ++i) {  // newer sources: offset d - hop >= 0
      if (d - fut[i] < 0) continue;
      const int ix = dindex(d - fut[i]);
      if (ix < 0) return false;
      P.nb[r][n++] = ix;
    }
    P.nnb[r] = n;
    P.nbmask[r] = 0u;
    for (int q = 0; q < n; ++q)
      if (P.nb[r][q] < 32) P.nbmask[r] |= 1u << P.nb[r][q];
  }
  for (int r = P.nR; r < TP_MAXR; ++r) P.nbmask[r] = 0u;
  P.n_past = np;
  P.n_future = nf;
  for (int i = 0; i < np; ++i) P.past[i] = past[i];
  for (int i = 0; i < nf; ++i) P.future[i] = fut[i];
  return true;
}

template <int F>
static int launch_temporal(const TemporalArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)TP_NW * (a.prog.nD * F + 4 * 2 * F + 64) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_step_temporal<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(temporal): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
  }
  int grid = (a.st.B + TP_NW - 1) / TP_NW;
  const int cap = gcm_num_sms();
  if (grid > cap) grid = cap;
  k_step_temporal<F><<<grid, TP_THREADS, smem, stream>>>(a);
  return gcm_check_launch("k_step_temporal");
}


// ------------------------------------------------------------------------------------------------
// pure-temporal pipelined kernel: a producer warp streams each graph's window of past rows and the
// observation tile into a shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx); eight consumer warps run the two layers out of shared memory with the layer
// weights resident in registers.  One persistent CTA per SM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <int F>
__global__ void __launch_bounds__(TW_THREADS, 1) k_step_temporal_win(const TemporalWinArgs a) {
  constexpr int H = 32;
  constexpr int K1 = 2 * F;
  extern __shared__ __align__(128) unsigned char tw_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = a.st.N, C = a.st.C, W = a.st.W, B = a.st.B;
  const int win = a.win;
  const TemporalProg& P = a.prog;

  // shared-memory carve-up
  const int stage_floats = TW_G * win * F + TW_G * F;     // history windows + observation tile
  float* stage_base = reinterpret_cast<float*>(tw_raw);
  int* cnt_base = reinterpret_cast<int*>(stage_base + (size_t)TW_STAGES * stage_floats);  // [STAGES][G]
  float* scratch = reinterpret_cast<float*>(cnt_base + TW_STAGES * TW_G);                 // per consumer warp
  uint64_t* full = reinterpret_cast<uint64_t*>(scratch + TW_CONS * (4 * K1 + 2 * H));
  uint64_t* empty = full + TW_STAGES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TW_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, TW_CONS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_tiles = (B + TW_G - 1) / TW_G;

  if (warp == TW_CONS) {
    // ===================== producer warp: lane i feeds graph i of the tile =====================
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % TW_STAGES;
      const uint32_t ph = (it / TW_STAGES) & 1;
      mbar_wait(empty + s, ph ^ 1);
      float* st_win = stage_base + (size_t)s * stage_floats;
      float* st_obs = st_win + TW_G * win * F;
      const int g = tile * TW_G + lane;
      const int gt = min(TW_G, B - tile * TW_G);
      int cnt = 0, nrows = 0;
      if (lane < gt) {
        cnt = a.uniform_count >= 0 ? a.uniform_count : __ldcg(a.st.count + g);
        nrows = min(min(cnt, N - 1), win);   // in-window history rows (the window holds N-1 older nodes)
      }
      cnt_base[s * TW_G + lane] = cnt;
      uint32_t bytes = (uint32_t)nrows * F * 4u;
      uint32_t total = bytes;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(GCM_FULL_MASK, total, o);
      total += (uint32_t)gt * F * 4u;
      __syncwarp();  // the per-lane counts above must be visible before the barrier is armed
      if (lane == 0) mbar_expect_tx(full + s, total);
      __syncwarp();
      if (lane == 0) bulk_g2s(st_obs, a.obs + (size_t)tile * TW_G * F, (uint32_t)gt * F * 4u, full + s);
      if (nrows > 0) {
        const float* nodes_g = a.st.nodes + (size_t)g * C * F;
        float* dst = st_win + ((size_t)lane * win + (win - nrows)) * F;
        const int first = gcm_slot(cnt - nrows, C);
        const int n1 = min(nrows, C - first);
        bulk_g2s(dst, nodes_g + (size_t)first * F, (uint32_t)n1 * F * 4u, full + s);
        if (n1 < nrows) bulk_g2s(dst + (size_t)n1 * F, nodes_g, (uint32_t)(nrows - n1) * F * 4u, full + s);
      }
    }
    return;
  }

  // ============================== consumer warps ==============================
  float* aggx = scratch + warp * (4 * K1 + 2 * H);
  float* l2in = aggx + 4 * K1;
  float w1[K1], w2[2 * H];
#pragma unroll
  for (int k = 0; k < K1; ++k) w1[k] = __ldg(a.gnn.w1t + k * H + lane);
#pragma unroll
  for (int k = 0; k < 2 * H; ++k) w2[k] = __ldg(a.gnn.w2t + k * H + lane);
  const float bias1 = a.gnn.b1 ? __ldg(a.gnn.b1 + lane) : 0.0f;
  const float bias2 = a.gnn.b2 ? __ldg(a.gnn.b2 + lane) : 0.0f;
  const int act1 = a.gnn.act1, act2 = a.gnn.act2;

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % TW_STAGES;
    const uint32_t ph = (it / TW_STAGES) & 1;
    mbar_wait(full + s, ph);
    const float* st_win = stage_base + (size_t)s * stage_floats;
    const float* st_obs = st_win + TW_G * win * F;
    const int gt = min(TW_G, B - tile * TW_G);
    for (int gi = warp; gi < gt; gi += TW_CONS) {
      const int g = tile * TW_G + gi;
      const int cnt = cnt_base[s * TW_G + gi];
      const int tpos = cnt;
      const int lt = min(cnt, N - 1);
      const int tslot = gcm_slot(tpos, C);
      const float* wrow = st_win + (size_t)gi * win * F;   // row i <-> offset (win - i) from t
      float xobs = 0.0f;
      if (lane < F) {
        xobs = st_obs[gi * F + lane];
        a.st.nodes[((size_t)g * C + tslot) * F + lane] = xobs;   // node write (gcm.py:274)
      }
      {
        uint32_t* masks_g = a.st.masks + (size_t)g * C * 2 * W;
        uint32_t pw = 0u;
        for (int i = 0; i < P.n_past; ++i) {
          const int hop = P.past[i];
          if (hop <= lt && (hop >> 5) == lane) pw |= 1u << (hop & 31);
        }
        if (lane < W) {
          gcm_st_mask(masks_g + ((size_t)tslot * 2 + 0) * W + lane, pw);
          gcm_st_mask(masks_g + ((size_t)tslot * 2 + 1) * W + lane, 0u);
        }
        if (lane < P.n_future) {
          const int hop = P.future[lane];
          if (hop <= lt)
            atomicOr(masks_g + ((size_t)gcm_slot(tpos - hop, C) * 2 + 1) * W + (hop >> 5), 1u << (hop & 31));
        }
        if (lane == 0) __stcg(a.st.count + g, cnt + 1);
      }
      // distinct rows of the 2-hop in-neighbourhood, feature `lane` of each (statically unrolled: the
      // program lives in the constant bank, so every index below is an immediate operand)
      float xv[TW_MAXD];
#pragma unroll
      for (int i = 0; i < TW_MAXD; ++i) {
        xv[i] = 0.0f;
        if (i == 0) {
          xv[i] = xobs;
        } else if (i < P.nD) {
          const int o = P.doff[i];
          if (o <= lt && lane < F) xv[i] = wrow[(win - o) * F + lane];
        }
      }
      // rows of R1 (row r sits at program index r): [sum of in-neighbours | own features]
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const uint32_t m = P.nbmask[rr];
        float ag = 0.0f;
#pragma unroll
        for (int i = 0; i < TW_MAXD; ++i)
          if ((m >> i) & 1u) ag += xv[i];
        if (lane < F) {
          aggx[rr * K1 + lane] = ag;
          aggx[rr * K1 + F + lane] = xv[rr];
        }
      }
      __syncwarp();
      float z0 = bias1, z1 = bias1, z2 = bias1, z3 = bias1;
#pragma unroll
      for (int k = 0; k < K1; k += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(aggx + 0 * K1 + k);
        const float4 a1 = *reinterpret_cast<const float4*>(aggx + 1 * K1 + k);
        const float4 a2 = *reinterpret_cast<const float4*>(aggx + 2 * K1 + k);
        const float4 a3 = *reinterpret_cast<const float4*>(aggx + 3 * K1 + k);
        z0 = fmaf(a0.x, w1[k], z0); z0 = fmaf(a0.y, w1[k + 1], z0);
        z0 = fmaf(a0.z, w1[k + 2], z0); z0 = fmaf(a0.w, w1[k + 3], z0);
        z1 = fmaf(a1.x, w1[k], z1); z1 = fmaf(a1.y, w1[k + 1], z1);
        z1 = fmaf(a1.z, w1[k + 2], z1); z1 = fmaf(a1.w, w1[k + 3], z1);
        z2 = fmaf(a2.x, w1[k], z2); z2 = fmaf(a2.y, w1[k + 1], z2);
        z2 = fmaf(a2.z, w1[k + 2], z2); z2 = fmaf(a2.w, w1[k + 3], z2);
        z3 = fmaf(a3.x, w1[k], z3); z3 = fmaf(a3.y, w1[k + 1], z3);
        z3 = fmaf(a3.z, w1[k + 2], z3); z3 = fmaf(a3.w, w1[k + 3], z3);
      }
      const float h1t = gcm_act_fast(z0, act1);
      float agg2 = 0.0f;
      if (1 < P.nR && P.rd[1] <= lt) agg2 += gcm_act_fast(z1, act1);
      if (2 < P.nR && P.rd[2] <= lt) agg2 += gcm_act_fast(z2, act1);
      if (3 < P.nR && P.rd[3] <= lt) agg2 += gcm_act_fast(z3, act1);
      __syncwarp();
      l2in[lane] = agg2;
      l2in[H + lane] = h1t;
      __syncwarp();
      float o0 = bias2, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
#pragma unroll
      for (int k = 0; k < 2 * H; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(l2in + k);
        o0 = fmaf(v.x, w2[k], o0);
        o1 = fmaf(v.y, w2[k + 1], o1);
        o2 = fmaf(v.z, w2[k + 2], o2);
        o3 = fmaf(v.w, w2[k + 3], o3);
      }
      const float out = gcm_act_fast((o0 + o1) + (o2 + o3), act2);
      a.belief[(size_t)g * H + lane] = out;
      if (!isfinite(out)) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
      __syncwarp();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
}

static size_t tw_smem_bytes(int F, int win) {
  size_t fl = (size_t)TW_STAGES * (TW_G * win * F + TW_G * F) + (size_t)TW_CONS * (4 * 2 * F + 64);
  return fl * 4 + (size_t)TW_STAGES * TW_G * 4 + 2 * TW_STAGES * 8 + 128;
}

template <int F>
static int launch_temporal_win(const TemporalWinArgs& a, cudaStream_t stream) {
  const size_t smem = tw_smem_bytes(F, a.win);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_step_temporal_win<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(temporal_win): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    attr_set = true;
  }
  const int n_tiles = (a.st.B + TW_G - 1) / TW_G;
  int grid = gcm_num_sms();
  if (grid > n_tiles) grid = n_tiles;
  k_step_temporal_win<F><<<grid, TW_THREADS, smem, stream>>>(a);
  return gcm_check_launch("k_step_temporal_win");
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static int validate_state(const gcm_dense_state* st) {
  GCM_REQUIRE(st && st->nodes && st->masks && st->count, "state: null pointer");
  GCM_REQUIRE(st->B >= 0 && st->N >= 1 && st->N <= GCM_MAX_N, "state: N=%d outside [1,%d]", st->N, GCM_MAX_N);
  GCM_REQUIRE(st->C >= st->N, "state: capacity C=%d < N=%d", st->C, st->N);
  GCM_REQUIRE(st->F >= 1 && st->F <= GCM_MAX_FEAT, "state: F=%d outside [1,%d]", st->F, GCM_MAX_FEAT);
  GCM_REQUIRE(st->W == (st->N + 31) / 32, "state: W=%d != ceil(N/32)", st->W);
  return GCM_OK;
}

static int g_temporal_kernel = GCM_TK_AUTO;
extern "C" int gcm_set_temporal_kernel(int which) {
  GCM_REQUIRE(which >= GCM_TK_AUTO && which <= GCM_TK_ROWS, "set_temporal_kernel: bad variant %d", which);
  g_temporal_kernel = which;
  return GCM_OK;
}

static int step_fwd_impl(const gcm_dense_state* st, const float* obs, long long obs_ld, const gcm_selector* sels,
                         int n_sels, const gcm_gnn* gnn, float* belief, long long belief_ld, int32_t* status,
                         int flags, int uniform_count, float* hcache, int hc_ring, int* cache_written,
                         void* stream_, int n_steps = 1, long long obs_stride_t = 0, long long belief_stride_t = 0,
                         float* xrec = nullptr, long long xrec_row0 = 0);

extern "C" int gcm_dense_step_fwd(const gcm_dense_state* st, const float* obs, const gcm_selector* sels,
                                  int n_sels, const gcm_gnn* gnn, float* belief, int32_t* status,
                                  int flags, void* stream_) {
  return gcm_dense_step_fwd_cached(st, obs, sels, n_sels, gnn, belief, status, flags, nullptr, 0, nullptr, stream_);
}

extern "C" int gcm_dense_step_fwd_cached(const gcm_dense_state* st, const float* obs, const gcm_selector* sels,
                                         int n_sels, const gcm_gnn* gnn, float* belief, int32_t* status,
                                         int flags, float* hcache, int hc_ring, int* cache_written,
                                         void* stream_) {
  // legacy packing of the uniform count into bits 8.. of `flags` (23 bits); gcm_dense_step_fwd_ex takes it as its own
  // argument and is what the host code of this repository calls
  const int uc = (flags & GCM_STEP_UNIFORM_COUNT) ? (int)((unsigned)flags >> GCM_STEP_COUNT_SHIFT) : -1;
  return step_fwd_impl(st, obs, 0, sels, n_sels, gnn, belief, 0, status, flags & 0xff, uc, hcache, hc_ring,
                       cache_written, stream_);
}

extern "C" int gcm_dense_step_fwd_ex(const gcm_dense_state* st, const float* obs, long long obs_ld,
                                     const gcm_selector* sels, int n_sels, const gcm_gnn* gnn, float* belief,
                                     long long belief_ld, int32_t* status, int flags, int uniform_count, float* hcache,
                                     int hc_ring, int* cache_written, void* stream_) {
  return step_fwd_impl(st, obs, obs_ld, sels, n_sels, gnn, belief, belief_ld, status, flags & 0xff,
                       (flags & GCM_STEP_UNIFORM_COUNT) ? uniform_count : -1, hcache, hc_ring, cache_written, stream_);
}

// n_steps > 1: that many consecutive steps in ONE launch of the cached-row kernel (step k at obs + k * obs_stride_t /
// belief + k * belief_stride_t); any other kernel choice returns GCM_ERR_UNSUPPORTED without launching.
static int step_fwd_impl(const gcm_dense_state* st, const float* obs, long long obs_ld, const gcm_selector* sels,
                         int n_sels, const gcm_gnn* gnn, float* belief, long long belief_ld, int32_t* status,
                         int flags, int uniform_count, float* hcache, int hc_ring, int* cache_written,
                         void* stream_, int n_steps, long long obs_stride_t, long long belief_stride_t, float* xrec,
                         long long xrec_row0) {
  if (cache_written) *cache_written = 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = validate_state(st)) return rc;
  GCM_REQUIRE(obs && gnn && belief && status, "dense_step_fwd: null pointer");
  GCM_REQUIRE(n_sels >= 0 && n_sels <= GCM_MAX_SELECTORS, "dense_step_fwd: n_sels=%d > %d", n_sels,
              GCM_MAX_SELECTORS);
  GCM_REQUIRE(n_sels == 0 || sels, "dense_step_fwd: null selectors");
  GCM_REQUIRE(gnn->F == st->F, "dense_step_fwd: gnn F=%d != state F=%d", gnn->F, st->F);
  GCM_REQUIRE(gnn->H1 >= 1 && gnn->H1 <= GCM_MAX_FEAT && gnn->H2 >= 1 && gnn->H2 <= GCM_MAX_FEAT,
              "dense_step_fwd: H1=%d H2=%d outside [1,%d]", gnn->H1, gnn->H2, GCM_MAX_FEAT);
  GCM_REQUIRE(gnn->w1t && gnn->w2t, "dense_step_fwd: null weights");
  for (int s = 0; s < n_sels; ++s) {
    GCM_REQUIRE(sels[s].kind >= GCM_SEL_NONE && sels[s].kind <= GCM_SEL_SPATIAL, "selector %d: bad kind %d", s,
                sels[s].kind);
    GCM_REQUIRE(sels[s].n_hops >= 0 && sels[s].n_hops <= GCM_MAX_HOPS, "selector %d: n_hops=%d", s,
                sels[s].n_hops);
    if (sels[s].kind == GCM_SEL_EUCLIDEAN) GCM_REQUIRE(sels[s].dist, "selector %d: euclidean needs dist", s);
    if (sels[s].kind == GCM_SEL_SPATIAL) {
      const gcm_selector& q = sels[s];
      GCM_REQUIRE(q.slice_len >= 0 && q.a_step >= 1 && q.b_step >= 1 && q.a_start >= 0 && q.b_start >= 0 &&
                      (q.slice_len == 0 || (q.a_start + (q.slice_len - 1) * q.a_step < st->F &&
                                            q.b_start + (q.slice_len - 1) * q.b_step < st->F)),
                  "selector %d: spatial slice outside [0,F)", s);
    }
  }
  if (st->B == 0) return GCM_OK;
  if (obs_ld <= 0) obs_ld = st->F;
  if (belief_ld <= 0) belief_ld = gnn->H2;
  // only the cached-row kernel takes strided observation / belief rows (the sequence entry's [B, T, .] views)
  const bool strided = obs_ld != st->F || belief_ld != gnn->H2 || n_steps > 1;
  GCM_REQUIRE(!(flags & GCM_STEP_UNIFORM_COUNT) || uniform_count >= 0, "dense_step_fwd: negative uniform count");

  if ((flags & GCM_STEP_PURE_TEMPORAL) && gnn->H1 == 32 && gnn->H2 == 32 &&
      (st->F == 8 || st->F == 16 || st->F == 32)) {
    TemporalArgs ta;
    if (build_temporal_prog(sels, n_sels, ta.prog)) {
      int win = 0;
      for (int i = 0; i < ta.prog.nD; ++i) win = ta.prog.doff[i] > win ? ta.prog.doff[i] : win;
      // stage a contiguous history window when it is small and mostly needed rows
      if (g_temporal_kernel != GCM_TK_ROWS && win >= 1 && win <= TW_MAXWIN && win <= 2 * (ta.prog.nD - 1) + 1 &&
          ta.prog.nR <= 4 &&
          ta.prog.nD <= TW_MAXD &&
          tw_smem_bytes(st->F, win) <= 200 * 1024 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(st->nodes) & 15) == 0) {
        TemporalWinArgs wa;
        wa.st = *st;
        wa.obs = obs;
        wa.gnn = *gnn;
        wa.belief = belief;
        wa.status = status;
        wa.prog = ta.prog;
        wa.win = win;
        wa.uniform_count = (flags & GCM_STEP_UNIFORM_COUNT) ? uniform_count : -1;
        wa.obs_ld = obs_ld;
        wa.belief_ld = belief_ld;
        wa.n_steps = n_steps;
        wa.obs_stride_t = obs_stride_t;
        wa.belief_stride_t = belief_stride_t;
        wa.xrec = xrec;
        wa.xrec_row0 = xrec_row0;
        wa.hcache = hcache;
        wa.hc_ring = hc_ring;
        wa.weights_stable = (flags & GCM_STEP_WEIGHTS_STABLE) ? 1 : 0;
        const bool cache_ok = hcache && gcm_temporal_hc_shape_ok(wa);
        if (!cache_ok) wa.hcache = nullptr;
        // Measured on B200 at cfg2 (profiles/): hc 2x faster than tc, tc 3x faster than win.  AUTO takes the
        // fastest that fits; gcm_set_temporal_kernel() forces a variant (tests, A/B profiling).  hc needs every
        // cached row it reads to be valid (GCM_STEP_HCACHE_VALID, the host's bookkeeping); tc fills the cache.
        const int want = g_temporal_kernel;
        if (cache_ok && (flags & GCM_STEP_HCACHE_VALID) && (want == GCM_TK_AUTO || want == GCM_TK_HC)) {
          const int rc = gcm_launch_temporal_hc(wa, stream);
          if (rc != GCM_ERR_UNSUPPORTED) {
            if (cache_written) *cache_written = 1;
            return rc;
          }
        }
        if (strided) {
          gcm_set_error("dense_step_fwd: strided rows need the cached-row kernel");
          return GCM_ERR_UNSUPPORTED;
        }
        if (want == GCM_TK_AUTO || want == GCM_TK_TC || want == GCM_TK_HC) {
          const int rc = gcm_launch_temporal_tc(wa, stream);
          if (rc != GCM_ERR_UNSUPPORTED) {
            if (cache_written) *cache_written = cache_ok ? 1 : 0;
            return rc;
          }
        }
        wa.hcache = nullptr;
        switch (st->F) {
          case 8: return launch_temporal_win<8>(wa, stream);
          case 16: return launch_temporal_win<16>(wa, stream);
          default: return launch_temporal_win<32>(wa, stream);
        }
      }
      if (strided) {
        gcm_set_error("dense_step_fwd: strided rows need the cached-row kernel");
        return GCM_ERR_UNSUPPORTED;
      }
      ta.st = *st;
      ta.obs = obs;
      ta.gnn = *gnn;
      ta.belief = belief;
      ta.status = status;
      switch (st->F) {
        case 8: return launch_temporal<8>(ta, stream);
        case 16: return launch_temporal<16>(ta, stream);
        default: return launch_temporal<32>(ta, stream);
      }
    }
  }

  if (strided) {
    gcm_set_error("dense_step_fwd: strided rows need the cached-row kernel");
    return GCM_ERR_UNSUPPORTED;
  }
  DenseStepArgs a;
  a.st = *st;
  a.obs = obs;
  a.n_sels = n_sels;
  for (int s = 0; s < n_sels; ++s) a.sels[s] = sels[s];
  a.gnn = *gnn;
  a.belief = belief;
  a.status = status;
  const int fr = (st->F + 31) / 32;
  const int hmax = gnn->H1 > gnn->H2 ? gnn->H1 : gnn->H2;
  const int hr = (hmax + 31) / 32;
  switch (fr) {
    case 1: return launch_general_h<1>(a, hr, stream);
    case 2: return launch_general_h<2>(a, hr, stream);
    case 3:
    case 4: return launch_general_h<4>(a, hr, stream);
    default: return launch_general_h<8>(a, hr, stream);
  }
}

// ------------------------------------------------------------------------------------------------
// Rollout entry: T steps enqueued back to back from C (the loop of ray_gcm.py:200-202 without a Python round trip
// per step).  Each step is the same launch gcm_dense_step_fwd_ex would make; the bookkeeping the host keeps
// between steps (uniform count, freshness of the layer-1 row cache) lives in the gcm_rollout descriptor.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_copy_rows(const float* __restrict__ src, long long src_ld, float* __restrict__ dst,
                                                   long long dst_ld, long long rows, int width) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * width) return;
  const long long r = i / width;
  const int c = (int)(i - r * width);
  dst[r * dst_ld + c] = src[r * src_ld + c];
}

static int copy_rows(const float* src, long long src_ld, float* dst, long long dst_ld, long long rows, int width,
                     cudaStream_t stream) {
  const long long n = rows * width;
  if (n == 0) return GCM_OK;
  k_copy_rows<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, src_ld, dst, dst_ld, rows, width);
  return gcm_check_launch("k_copy_rows");
}

extern "C" int gcm_dense_rollout_fwd(gcm_rollout* r, const float* obs, long long obs_ld, long long obs_stride_t,
                                     float* belief, long long belief_ld, long long belief_stride_t, int T,
                                     void* stream_) {
  GCM_REQUIRE(r && obs && belief && T >= 0, "dense_rollout_fwd: bad arguments");
  GCM_REQUIRE(r->n_sels >= 1 && r->n_sels <= GCM_MAX_SELECTORS && r->max_hop >= 1, "dense_rollout_fwd: bad selector chain");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int F = r->st.F, H2 = r->gnn.H2;
  if (obs_ld <= 0) obs_ld = F;
  if (belief_ld <= 0) belief_ld = H2;
  const bool strided = obs_ld != F || belief_ld != H2;
  const long long l0 = gcm_launch_count();
  r->xrec_from = T;
  static const bool no_multi = getenv("GCM_B200_NO_MULTISTEP") != nullptr;     // A/B switch: one launch per step
  for (int k = 0; k < T;) {
    const float* ob = obs + (long long)k * obs_stride_t;
    float* be = belief + (long long)k * belief_stride_t;
    int flags = GCM_STEP_PURE_TEMPORAL;
    if (r->uniform_count >= 0) flags |= GCM_STEP_UNIFORM_COUNT;
    const int need = r->uniform_count >= 0 && r->uniform_count < r->max_hop ? r->uniform_count : r->max_hop;
    if (r->hcache && r->hc_fresh >= need) flags |= GCM_STEP_HCACHE_VALID;
    if (r->weights_stable && r->hc_fresh >= 1) flags |= GCM_STEP_WEIGHTS_STABLE;
    int written = 0;
    int rc;
    if ((flags & GCM_STEP_HCACHE_VALID) && r->hc_fresh >= r->max_hop && T - k > 1 && !no_multi) {
      // every remaining step runs on the cached-row kernel (the weights cannot change inside this call): ONE launch
      // walks all of them, consecutive steps overlapping inside the kernel (csrc/gcm_dense_fwd_hc.cu)
      rc = step_fwd_impl(&r->st, ob, obs_ld, r->sels, r->n_sels, &r->gnn, be, belief_ld, r->status, flags,
                         r->uniform_count, r->hcache, r->hc_ring, &written, stream_, T - k, obs_stride_t,
                         belief_stride_t, r->xrec, r->xrec_row0 + (long long)k * r->st.B);
      if (rc == GCM_OK && written) {
        if (r->xrec && r->st.F == 32) r->xrec_from = k;
        if (r->uniform_count >= 0) r->uniform_count += T - k;
        r->weights_stable = 1;
        k = T;
        continue;
      }
      if (rc != GCM_ERR_UNSUPPORTED) return rc ? rc : GCM_ERR_INVALID;
    }
    rc = step_fwd_impl(&r->st, ob, obs_ld, r->sels, r->n_sels, &r->gnn, be, belief_ld, r->status, flags,
                       r->uniform_count, r->hcache, r->hc_ring, &written, stream_);
    if (rc == GCM_ERR_UNSUPPORTED && strided) {
      // a step the cached-row kernel cannot take (cache being filled, shape): contiguous copies of the rows
      GCM_REQUIRE(r->scratch_obs && r->scratch_belief, "dense_rollout_fwd: strided rows need the scratch buffers");
      if ((rc = copy_rows(ob, obs_ld, r->scratch_obs, F, r->st.B, F, stream))) return rc;
      rc = step_fwd_impl(&r->st, r->scratch_obs, F, r->sels, r->n_sels, &r->gnn, r->scratch_belief, H2, r->status,
                         flags, r->uniform_count, r->hcache, r->hc_ring, &written, stream_);
      if (rc) return rc;
      rc = copy_rows(r->scratch_belief, H2, be, belief_ld, r->st.B, H2, stream);
    }
    if (rc) return rc;
    if (r->hcache) {
      if (!written) r->hc_fresh = 0;
      else if (r->hc_fresh < r->max_hop) ++r->hc_fresh;
    }
    if (r->uniform_count >= 0) ++r->uniform_count;
    r->weights_stable = 1;   // nothing but this loop touches the stream between the steps of one call
    ++k;
  }
  r->launches = gcm_launch_count() - l0;
  return GCM_OK;
}

extern "C" int gcm_dense_rollout_step(gcm_rollout* r, const float* obs, float* belief, void* stream_) {
  return gcm_dense_rollout_fwd(r, obs, 0, 0, belief, 0, 0, 1, stream_);
}

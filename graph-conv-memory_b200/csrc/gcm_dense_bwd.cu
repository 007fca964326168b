// Dense GCM step, backward (BPTT): gradient of one DenseGCM.forward step w.r.t. the GNN weights,
// the step's observation and -- through the carried node buffer -- every earlier observation.
// Replaces autograd through /root/reference/src/gcm/gcm.py:262-321 + the user GNN (README.md:52-62).
//
// Nothing is saved by the forward: the step is recomputed from the node log.  A step that ran
// `steps_back` steps ago saw count_t = count_now - 1 - steps_back; every row of its window is still
// in the log (the host guarantees spare capacity), rows written later are masked out by position,
// and future-mask bits added by later steps are masked out by `e <= d`.
//
// One CTA walks graphs b = blockIdx.x, blockIdx.x + gridDim.x, ...; weight gradients are accumulated
// in shared memory (element-owned by threads, no atomics) and flushed once per CTA.
#include "gcm_common.cuh"

constexpr int BW_THREADS = 256;
constexpr int BW_NW = BW_THREADS / 32;

struct BwdArgs {
  gcm_dense_state st;
  int steps_back;
  gcm_gnn gnn;
  const float* d_belief;
  float* d_nodes;
  float* d_obs;
  gcm_gnn_grads grads;
  int smem_acc;  // 1: weight-gradient accumulators live in shared memory
};

template <int FR>
__device__ __forceinline__ void bw_accum_bits(float (&acc)[FR], uint32_t pw, uint32_t fw, int W, int pos,
                                              const float* nodes_b, int C, int F, int lane, float sign) {
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
}

template <int FR>
__device__ __forceinline__ void bw_scatter_bits(const float (&g)[FR], uint32_t pw, uint32_t fw, int W, int pos,
                                                float* dn_b, int C, int F, int lane, float sign) {
  for (int w = 0; w < W; ++w) {
    uint32_t m = __shfl_sync(GCM_FULL_MASK, pw, w);
    while (m) {
      const int e = w * 32 + __ffs(m) - 1;
      m &= m - 1;
      float* row = dn_b + (size_t)gcm_slot(pos - e, C) * F;
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int f = lane + 32 * k;
        if (f < F) atomicAdd(row + f, sign * g[k]);
      }
    }
    m = __shfl_sync(GCM_FULL_MASK, fw, w);
    while (m) {
      const int e = w * 32 + __ffs(m) - 1;
      m &= m - 1;
      float* row = dn_b + (size_t)gcm_slot(pos + e, C) * F;
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int f = lane + 32 * k;
        if (f < F) atomicAdd(row + f, sign * g[k]);
      }
    }
  }
}

template <int FR, int HR>
__global__ void __launch_bounds__(BW_THREADS) k_step_bwd_general(const BwdArgs a) {
  extern __shared__ __align__(16) unsigned char bw_raw[];
  const int F = a.st.F, N = a.st.N, C = a.st.C, W = a.st.W, H1 = a.gnn.H1, H2 = a.gnn.H2;
  float* sall = reinterpret_cast<float*>(bw_raw);
  float* dsall = sall + F;
  float* h1t = dsall + F;
  float* agg2 = h1t + H1;
  float* dagg2 = agg2 + H1;
  float* dh1t = dagg2 + H1;
  float* dz2 = dh1t + H1;                  // [H2]
  float* agg2part = dz2 + H2;              // [NW][H1]
  float* mybuf = agg2part + BW_NW * H1;    // [NW][2F]
  float* dz1buf = mybuf + BW_NW * 2 * F;   // [NW][H1]
  float* accW1 = dz1buf + BW_NW * H1;      // [2F][H1]   (only if smem_acc)
  float* accB1 = accW1 + (a.smem_acc ? 2 * F * H1 : 0);
  float* accW2 = accB1 + (a.smem_acc ? H1 : 0);
  float* accB2 = accW2 + (a.smem_acc ? 2 * H1 * H2 : 0);
  float* acc_end = accB2 + (a.smem_acc ? H2 : 0);
  uint32_t* rowmask = reinterpret_cast<uint32_t*>(acc_end);   // [W]
  int* r1n = reinterpret_cast<int*>(rowmask + W);             // [2]
  int* rowvalid = r1n + 2;                                    // [NW]
  uint16_t* r1 = reinterpret_cast<uint16_t*>(rowvalid + BW_NW);  // [N]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.smem_acc)
    for (int i = tid; i < (int)(acc_end - accW1); i += BW_THREADS) accW1[i] = 0.0f;
  __syncthreads();

  for (int b = blockIdx.x; b < a.st.B; b += gridDim.x) {
    const int cnt = __ldcg(a.st.count + b) - 1 - a.steps_back;  // count seen by that step
    if (cnt < 0) continue;                                      // (uniform per CTA iteration)
    const int tpos = cnt;
    const int lt = min(cnt, N - 1);
    const int tslot = gcm_slot(tpos, C);
    const float* nodes_b = a.st.nodes + (size_t)b * C * F;
    const uint32_t* masks_b = a.st.masks + (size_t)b * C * 2 * W;
    float* dn_b = a.d_nodes + (size_t)b * C * F;

    for (int w = tid; w < W; w += BW_THREADS)
      rowmask[w] = gcm_ld_mask(masks_b + ((size_t)tslot * 2 + 0) * W + w) & gcm_range_word(w, 0, lt);
    for (int f = tid; f < F; f += BW_THREADS) dsall[f] = 0.0f;
    __syncthreads();
    if (warp == 0) {
      int base = 1;
      if (lane == 0) r1[0] = 0;
      for (int w = 0; w < W; ++w) {
        const uint32_t m = rowmask[w] & gcm_range_word(w, 1, lt);
        if ((m >> lane) & 1u) r1[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(w * 32 + lane);
        base += __popc(m);
      }
      if (lane == 0) {
        r1n[0] = base;
        r1n[1] = (int)(rowmask[0] & 1u);
      }
    }
    __syncthreads();
    const int nR = r1n[0];
    const bool selfloop = r1n[1] != 0;
    const bool use_sall = nR > 16;
    if (use_sall) {
      float acc[FR];
#pragma unroll
      for (int k = 0; k < FR; ++k) acc[k] = 0.0f;
      for (int d = warp; d <= lt; d += BW_NW) {
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
        if (f < F) mybuf[warp * 2 * F + f] = acc[k];
      }
      __syncthreads();
      for (int f = tid; f < F; f += BW_THREADS) {
        float s = 0.0f;
        for (int w = 0; w < BW_NW; ++w) s += mybuf[w * 2 * F + f];
        sall[f] = s;
      }
      __syncthreads();
    }

    float* my = mybuf + warp * 2 * F;
    float* dz1 = dz1buf + warp * H1;

    // recompute of one R1 row: stages [agg | x] into `my`, returns h1 (lanes own h), and the
    // (possibly complemented) neighbour masks used, for the scatter
    auto row_forward = [&](int ri, float (&hv)[HR], uint32_t& pw, uint32_t& fw, bool& comp, int& pos, int& d) {
      d = r1[ri];
      pos = tpos - d;
      const int slot = gcm_slot(pos, C);
      const int lj = lt - d;
      const uint32_t* mrow = masks_b + (size_t)slot * 2 * W;
      pw = 0u;
      fw = 0u;
      if (lane < W) {
        pw = gcm_ld_mask(mrow + lane) & gcm_range_word(lane, 0, lj);
        fw = gcm_ld_mask(mrow + W + lane) & gcm_range_word(lane, 1, d);
      }
      const int deg = gcm_warp_sum_int(__popc(pw) + __popc(fw));
      comp = use_sall && (2 * deg > lt + 1);
      float acc[FR];
#pragma unroll
      for (int k = 0; k < FR; ++k) acc[k] = 0.0f;
      if (comp) {
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
      bw_accum_bits<FR>(acc, pw, fw, W, pos, nodes_b, C, F, lane, comp ? -1.0f : 1.0f);
      const float* xrow = nodes_b + (size_t)slot * F;
#pragma unroll
      for (int k = 0; k < FR; ++k) {
        const int f = lane + 32 * k;
        if (f < F) {
          my[f] = acc[k];
          my[F + f] = xrow[f];
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < HR; ++k) {
        const int h = lane + 32 * k;
        hv[k] = (a.gnn.b1 != nullptr && h < H1) ? __ldg(a.gnn.b1 + h) : 0.0f;
      }
      for (int kk = 0; kk < 2 * F; ++kk) {
        const float av = my[kk];
        const float* wr = a.gnn.w1t + (size_t)kk * H1;
#pragma unroll
        for (int k = 0; k < HR; ++k) {
          const int h = lane + 32 * k;
          if (h < H1) hv[k] = fmaf(av, __ldg(wr + h), hv[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < HR; ++k) hv[k] = gcm_act_fwd(hv[k], a.gnn.act1);
    };

    // ---- pass 1: forward recompute -> h1[t], agg2 ----
    {
      float a2[HR];
#pragma unroll
      for (int k = 0; k < HR; ++k) a2[k] = 0.0f;
      for (int ri = warp; ri < nR; ri += BW_NW) {
        float hv[HR];
        uint32_t pw, fw;
        bool comp;
        int pos, d;
        row_forward(ri, hv, pw, fw, comp, pos, d);
#pragma unroll
        for (int k = 0; k < HR; ++k) {
          const int h = lane + 32 * k;
          if (h < H1) {
            if (d == 0) h1t[h] = hv[k];
            if (d > 0 || selfloop) a2[k] += hv[k];
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int k = 0; k < HR; ++k) {
        const int h = lane + 32 * k;
        if (h < H1) agg2part[warp * H1 + h] = a2[k];
      }
    }
    __syncthreads();
    for (int k = tid; k < H1; k += BW_THREADS) {
      float s = 0.0f;
      for (int w = 0; w < BW_NW; ++w) s += agg2part[w * H1 + k];
      agg2[k] = s;
    }
    __syncthreads();

    // ---- layer 2 backward ----
    for (int h2 = tid; h2 < H2; h2 += BW_THREADS) {
      float z = a.gnn.b2 != nullptr ? __ldg(a.gnn.b2 + h2) : 0.0f;
      for (int k = 0; k < H1; ++k) {
        z = fmaf(agg2[k], __ldg(a.gnn.w2t + (size_t)k * H2 + h2), z);
        z = fmaf(h1t[k], __ldg(a.gnn.w2t + (size_t)(H1 + k) * H2 + h2), z);
      }
      const float out = gcm_act_fwd(z, a.gnn.act2);
      const float g = a.d_belief[(size_t)b * H2 + h2] * gcm_act_grad(out, a.gnn.act2);
      dz2[h2] = g;
      if (a.grads.d_b2) {
        if (a.smem_acc) accB2[h2] += g;
        else atomicAdd(a.grads.d_b2 + h2, g);
      }
    }
    __syncthreads();
    for (int e = tid; e < 2 * H1 * H2; e += BW_THREADS) {
      const int k = e / H2, h2 = e - k * H2;
      const float v = (k < H1 ? agg2[k] : h1t[k - H1]) * dz2[h2];
      if (a.smem_acc) accW2[e] += v;
      else atomicAdd((k < H1 ? a.grads.d_w_rel2 + (size_t)h2 * H1 + k
                             : a.grads.d_w_root2 + (size_t)h2 * H1 + (k - H1)), v);
    }
    for (int k = tid; k < H1; k += BW_THREADS) {
      float ga = 0.0f, gh = 0.0f;
      for (int h2 = 0; h2 < H2; ++h2) {
        const float g = dz2[h2];
        ga = fmaf(__ldg(a.gnn.w_rel2 + (size_t)h2 * H1 + k), g, ga);
        gh = fmaf(__ldg(a.gnn.w_root2 + (size_t)h2 * H1 + k), g, gh);
      }
      dagg2[k] = ga;
      dh1t[k] = gh;
    }
    __syncthreads();

    // ---- pass 2: layer 1 backward, NW rows at a time ----
    for (int base = 0; base < nR; base += BW_NW) {
      const int ri = base + warp;
      const bool active = ri < nR;
      uint32_t pw = 0u, fw = 0u;
      bool comp = false;
      int pos = 0, d = 0;
      if (active) {
        float hv[HR];
        row_forward(ri, hv, pw, fw, comp, pos, d);
#pragma unroll
        for (int k = 0; k < HR; ++k) {
          const int h = lane + 32 * k;
          if (h < H1) {
            float g = 0.0f;
            if (d == 0) g += dh1t[h];
            if (d > 0 || selfloop) g += dagg2[h];
            dz1[h] = g * gcm_act_grad(hv[k], a.gnn.act1);
          }
        }
      }
      if (lane == 0) rowvalid[warp] = active ? 1 : 0;
      __syncthreads();
      // weight gradients of this batch of rows: element-owned accumulation
      for (int e = tid; e < 2 * F * H1; e += BW_THREADS) {
        const int k = e / H1, h = e - k * H1;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < BW_NW; ++w)
          if (rowvalid[w]) v = fmaf(mybuf[w * 2 * F + k], dz1buf[w * H1 + h], v);
        if (a.smem_acc) accW1[e] += v;
        else atomicAdd((k < F ? a.grads.d_w_rel1 + (size_t)h * F + k
                              : a.grads.d_w_root1 + (size_t)h * F + (k - F)), v);
      }
      if (a.grads.d_b1)
        for (int h = tid; h < H1; h += BW_THREADS) {
          float v = 0.0f;
#pragma unroll
          for (int w = 0; w < BW_NW; ++w)
            if (rowvalid[w]) v += dz1buf[w * H1 + h];
          if (a.smem_acc) accB1[h] += v;
          else atomicAdd(a.grads.d_b1 + h, v);
        }
      // input gradients of this warp's row, scattered into the running dL/dnodes buffer
      if (active) {
        float gagg[FR], gx[FR];
#pragma unroll
        for (int k = 0; k < FR; ++k) gagg[k] = gx[k] = 0.0f;
        for (int h = 0; h < H1; ++h) {
          const float g = dz1[h];
          const float* wr = a.gnn.w_rel1 + (size_t)h * F;
          const float* wo = a.gnn.w_root1 + (size_t)h * F;
#pragma unroll
          for (int k = 0; k < FR; ++k) {
            const int f = lane + 32 * k;
            if (f < F) {
              gagg[k] = fmaf(__ldg(wr + f), g, gagg[k]);
              gx[k] = fmaf(__ldg(wo + f), g, gx[k]);
            }
          }
        }
        float* xr = dn_b + (size_t)gcm_slot(pos, C) * F;
#pragma unroll
        for (int k = 0; k < FR; ++k) {
          const int f = lane + 32 * k;
          if (f < F) {
            atomicAdd(xr + f, gx[k]);
            if (comp) atomicAdd(dsall + f, gagg[k]);
          }
        }
        bw_scatter_bits<FR>(gagg, pw, fw, W, pos, dn_b, C, F, lane, comp ? -1.0f : 1.0f);
      }
      __syncthreads();
    }
    if (use_sall) {
      for (int e = tid; e < (lt + 1) * F; e += BW_THREADS) {
        const int d = e / F, f = e - d * F;
        atomicAdd(dn_b + (size_t)gcm_slot(tpos - d, C) * F + f, dsall[f]);
      }
    }
    __threadfence();
    __syncthreads();
    // dL/dx of this step = its row of the running buffer (all later steps already added theirs)
    for (int f = tid; f < F; f += BW_THREADS) {
      float* p = dn_b + (size_t)tslot * F + f;
      const float v = __ldcg(p);
      if (a.d_obs) a.d_obs[(size_t)b * F + f] = v;
      __stcg(p, 0.0f);
    }
    __syncthreads();
  }

  if (a.smem_acc) {
    __syncthreads();
    for (int e = tid; e < 2 * F * H1; e += BW_THREADS) {
      const int k = e / H1, h = e - k * H1;
      const float v = accW1[e];
      if (v != 0.0f)
        atomicAdd((k < F ? a.grads.d_w_rel1 + (size_t)h * F + k : a.grads.d_w_root1 + (size_t)h * F + (k - F)), v);
    }
    for (int e = tid; e < 2 * H1 * H2; e += BW_THREADS) {
      const int k = e / H2, h2 = e - k * H2;
      const float v = accW2[e];
      if (v != 0.0f)
        atomicAdd((k < H1 ? a.grads.d_w_rel2 + (size_t)h2 * H1 + k
                          : a.grads.d_w_root2 + (size_t)h2 * H1 + (k - H1)), v);
    }
    if (a.grads.d_b1)
      for (int h = tid; h < H1; h += BW_THREADS) atomicAdd(a.grads.d_b1 + h, accB1[h]);
    if (a.grads.d_b2)
      for (int h = tid; h < H2; h += BW_THREADS) atomicAdd(a.grads.d_b2 + h, accB2[h]);
  }
}

static size_t bw_smem_bytes(const gcm_dense_state& st, const gcm_gnn& g, bool acc) {
  size_t fl = (size_t)2 * st.F + 4 * g.H1 + g.H2 + (size_t)BW_NW * g.H1 + (size_t)BW_NW * 2 * st.F +
              (size_t)BW_NW * g.H1;
  if (acc) fl += (size_t)2 * st.F * g.H1 + g.H1 + (size_t)2 * g.H1 * g.H2 + g.H2;
  size_t bytes = fl * 4 + (size_t)st.W * 4 + 8 + BW_NW * 4 + (size_t)st.N * 2;
  return (bytes + 15) & ~(size_t)15;
}

template <int FR, int HR>
static int launch_bwd(const BwdArgs& a0, cudaStream_t stream) {
  BwdArgs a = a0;
  a.smem_acc = bw_smem_bytes(a.st, a.gnn, true) <= 160 * 1024 ? 1 : 0;
  const size_t smem = bw_smem_bytes(a.st, a.gnn, a.smem_acc != 0);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_step_bwd_general<FR, HR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(bwd): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
  }
  int grid = 2 * gcm_num_sms();
  if (grid > a.st.B) grid = a.st.B;
  k_step_bwd_general<FR, HR><<<grid, BW_THREADS, smem, stream>>>(a);
  return gcm_check_launch("k_step_bwd_general");
}

template <int FR>
static int launch_bwd_h(const BwdArgs& a, int hr, cudaStream_t stream) {
  switch (hr) {
    case 1: return launch_bwd<FR, 1>(a, stream);
    case 2: return launch_bwd<FR, 2>(a, stream);
    case 3:
    case 4: return launch_bwd<FR, 4>(a, stream);
    default: return launch_bwd<FR, 8>(a, stream);
  }
}

extern "C" int gcm_dense_step_bwd(const gcm_dense_state* st, int steps_back, const gcm_gnn* gnn,
                                  const float* d_belief, float* d_nodes, float* d_obs,
                                  const gcm_gnn_grads* grads, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GCM_REQUIRE(st && st->nodes && st->masks && st->count, "dense_step_bwd: null state");
  GCM_REQUIRE(st->N >= 1 && st->N <= GCM_MAX_N && st->C >= st->N && st->F >= 1 && st->F <= GCM_MAX_FEAT &&
                  st->W == (st->N + 31) / 32,
              "dense_step_bwd: bad state dims");
  GCM_REQUIRE(gnn && d_belief && d_nodes && grads, "dense_step_bwd: null pointer");
  GCM_REQUIRE(steps_back >= 0 && steps_back <= st->C - st->N,
              "dense_step_bwd: step %d is older than the log's spare capacity (C - N = %d)", steps_back,
              st->C - st->N);
  GCM_REQUIRE(gnn->F == st->F && gnn->H1 >= 1 && gnn->H1 <= GCM_MAX_FEAT && gnn->H2 >= 1 &&
                  gnn->H2 <= GCM_MAX_FEAT,
              "dense_step_bwd: bad gnn dims");
  GCM_REQUIRE(gnn->w1t && gnn->w2t && gnn->w_rel1 && gnn->w_root1 && gnn->w_rel2 && gnn->w_root2,
              "dense_step_bwd: null weights");
  GCM_REQUIRE(grads->d_w_rel1 && grads->d_w_root1 && grads->d_w_rel2 && grads->d_w_root2,
              "dense_step_bwd: null weight-gradient buffers");
  if (st->B == 0) return GCM_OK;
  BwdArgs a;
  a.st = *st;
  a.steps_back = steps_back;
  a.gnn = *gnn;
  a.d_belief = d_belief;
  a.d_nodes = d_nodes;
  a.d_obs = d_obs;
  a.grads = *grads;
  a.smem_acc = 0;
  const int fr = (st->F + 31) / 32;
  const int hr = (gnn->H1 + 31) / 32;
  switch (fr) {
    case 1: return launch_bwd_h<1>(a, hr, stream);
    case 2: return launch_bwd_h<2>(a, hr, stream);
    case 3:
    case 4: return launch_bwd_h<4>(a, hr, stream);
    default: return launch_bwd_h<8>(a, hr, stream);
  }
}

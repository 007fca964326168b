// Pure-temporal DenseGCM step on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Why tensor cores here: with the adjacency implicit, one step of B graphs is the batched product
// [4B, 2F] x [2F, 32] (layer 1 on the 4 rows of each graph's 2-hop in-neighbourhood) followed by
// [4B, 32] x [32, 64] (layer 2, rel | root halves).  On CUDA cores this costs ~950 warp-instructions per
// graph and the step is issue-bound at 7x the HBM time (profiles/c2_step_temporal_win_r1.md); on
// tcgen05 the FMAs disappear from the instruction stream.  fp32 accuracy is kept with the 3xTF32 split
// (gcm_tc.cuh; 5e-7 relative on B200, tests/test_tc_gpu.py).
//
// One persistent CTA per SM, 16 warps:
//   warps 0-11       three consumer groups of 4 warps, each owns a 32-graph tile: warp r of a group holds
//                    row r of R1 for all 32 graphs (lane = graph), so TMEM lane = 32 r + graph and the
//                    neighbour program is warp-uniform.  They build [agg | x] (hi, lo) straight into TMEM
//                    (tcgen05.st), apply bias + activation on the accumulator (tcgen05.ld), write it back
//                    as the layer-2 operand, and reduce the 4 rows of each graph through shared memory.
//   warp 12          producer: TMA bulk copies of each graph's history window + observation into the
//                    group's shared-memory stage (mbarrier complete_tx); TMEM alloc/dealloc.  A stage is
//                    released as soon as the operand is built, so the next tile's copies overlap the
//                    two MMA phases and epilogues of the current one.
//   warps 13-15      MMA issuers (one elected thread each), one per consumer group.
// TMEM columns per group (160): A hi [0,64) | A lo [64,128) | D1 [128,160); the layer-2 operand (hi | lo)
// overlays A hi and the layer-2 accumulator overlays A lo once the layer-1 MMAs have completed.
#include "gcm_tc.cuh"
#include "gcm_temporal.cuh"

constexpr int TC_G = 32;                  // graphs per tile
constexpr int TC_GROUPS = 3;
constexpr int TC_THREADS = (4 * TC_GROUPS + 1 + TC_GROUPS) * 32;
constexpr int TC_MAXNB = 6;               // in-neighbours per row held in registers
constexpr int TC_H = 32;                  // H1 == H2 == 32
constexpr uint32_t TC_COL_AHI = 0, TC_COL_ALO = 64, TC_COL_D1 = 128, TC_COL_D2 = 64, TC_COL_GROUP = 160;

struct TcSmem {   // offsets in bytes into dynamic shared memory
  uint32_t b1hi, b1lo, b2hi, b2lo, stage, cnt, red, bias, bars, tmem_slot, total;
  uint32_t gs_floats;   // per-graph stride inside a stage (window rows + observation + 4 floats of padding)
};

__host__ __device__ inline TcSmem tc_smem_layout(int F, int win) {
  TcSmem L;
  const uint32_t K1 = 2 * F;
  uint32_t o = 0;
  L.b1hi = o; o += TC_H * K1 * 4;
  L.b1lo = o; o += TC_H * K1 * 4;
  L.b2hi = o; o += 64 * TC_H * 4;
  L.b2lo = o; o += 64 * TC_H * 4;
  L.gs_floats = (uint32_t)(win * F + F + 4);
  L.stage = o; o += (uint32_t)TC_GROUPS * TC_G * L.gs_floats * 4;   // [group]
  L.red = o; o += (uint32_t)TC_GROUPS * 4 * TC_H * TC_G * 4;        // [group][r][h][g]
  L.cnt = o; o += (uint32_t)TC_GROUPS * TC_G * 4;
  L.bias = o; o += 2 * TC_H * 4;
  L.bars = o; o += 6 * TC_GROUPS * 8;
  L.tmem_slot = o; o += 16;
  L.total = o;
  return L;
}

template <int F, int NB>
__global__ void __launch_bounds__(TC_THREADS, 1) k_step_temporal_tc(const TemporalWinArgs a) {
  constexpr int K1 = 2 * F;
  constexpr int PROD_WARP = 4 * TC_GROUPS, MMA_WARP0 = PROD_WARP + 1;
  extern __shared__ __align__(128) unsigned char sm[];
  const TcSmem L = tc_smem_layout(F, a.win);
  float* B1hi = reinterpret_cast<float*>(sm + L.b1hi);
  float* B1lo = reinterpret_cast<float*>(sm + L.b1lo);
  float* B2hi = reinterpret_cast<float*>(sm + L.b2hi);
  float* B2lo = reinterpret_cast<float*>(sm + L.b2lo);
  float* stages = reinterpret_cast<float*>(sm + L.stage);
  float* red = reinterpret_cast<float*>(sm + L.red);
  int* cnts = reinterpret_cast<int*>(sm + L.cnt);
  float* bias_s = reinterpret_cast<float*>(sm + L.bias);   // [b1 | b2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
  uint64_t* full = bars;                       // [group]
  uint64_t* empty = bars + TC_GROUPS;
  uint64_t* a1_ready = bars + 2 * TC_GROUPS;
  uint64_t* d1_ready = bars + 3 * TC_GROUPS;
  uint64_t* a2_ready = bars + 4 * TC_GROUPS;
  uint64_t* d2_ready = bars + 5 * TC_GROUPS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.st.N, C = a.st.C, W = a.st.W, B = a.st.B, win = a.win;
  const TemporalProg& P = a.prog;
  const int gs = (int)L.gs_floats;
  const int n_tiles = (B + TC_G - 1) / TC_G;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (tid == 0) {
    for (int g = 0; g < TC_GROUPS; ++g) {
      tc::mbar_init(full + g, 1);
      tc::mbar_init(empty + g, 4);
      tc::mbar_init(a1_ready + g, 128);
      tc::mbar_init(d1_ready + g, 1);
      tc::mbar_init(a2_ready + g, 128);
      tc::mbar_init(d2_ready + g, 1);
    }
    tc::mbar_fence_init();
  }
  if (warp == PROD_WARP) tc::tmem_alloc(tmem_slot, 512);
  // layer weights -> canonical K-major B operands, split hi / lo
  for (int i = tid; i < TC_H * K1; i += TC_THREADS) {
    const int n = i / K1, k = i - n * K1;
    const float w = k < F ? __ldg(a.gnn.w_rel1 + n * F + k) : __ldg(a.gnn.w_root1 + n * F + (k - F));
    uint32_t hi, lo;
    tc::split_tf32(w, hi, lo);
    B1hi[tc::kmajor_off(n, k, K1)] = __uint_as_float(hi);
    B1lo[tc::kmajor_off(n, k, K1)] = __uint_as_float(lo);
  }
  for (int i = tid; i < 64 * TC_H; i += TC_THREADS) {
    const int n = i / TC_H, k = i - n * TC_H;
    const float w = n < TC_H ? __ldg(a.gnn.w_rel2 + n * TC_H + k) : __ldg(a.gnn.w_root2 + (n - TC_H) * TC_H + k);
    uint32_t hi, lo;
    tc::split_tf32(w, hi, lo);
    B2hi[tc::kmajor_off(n, k, TC_H)] = __uint_as_float(hi);
    B2lo[tc::kmajor_off(n, k, TC_H)] = __uint_as_float(lo);
  }
  if (tid < TC_H) {
    bias_s[tid] = a.gnn.b1 ? __ldg(a.gnn.b1 + tid) : 0.0f;
    bias_s[TC_H + tid] = a.gnn.b2 ? __ldg(a.gnn.b2 + tid) : 0.0f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == PROD_WARP) {
    // =============================== producer ===============================
    for (int j = 0; j < my_tiles; ++j) {
      const int grp = j % TC_GROUPS, it = j / TC_GROUPS;
      const uint32_t ph = it & 1;
      const int tile = blockIdx.x + j * gridDim.x;
      uint64_t* fb = full + grp;
      tc::mbar_wait(empty + grp, ph ^ 1);
      float* st_base = stages + (size_t)grp * TC_G * gs;
      const int g = tile * TC_G + lane;
      const int gt = min(TC_G, B - tile * TC_G);
      int cnt = 0, nrows = 0;
      if (lane < gt) {
        cnt = a.uniform_count >= 0 ? a.uniform_count : __ldcg(a.st.count + g);
        nrows = min(min(cnt, N - 1), win);
      }
      cnts[grp * TC_G + lane] = cnt;
      uint32_t total = lane < gt ? (uint32_t)(nrows + 1) * F * 4u : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(GCM_FULL_MASK, total, o);
      __syncwarp();
      if (lane == 0) tc::mbar_expect_tx(fb, total);
      __syncwarp();
      if (lane < gt) {
        float* dst = st_base + (size_t)lane * gs;                       // rows: [win history | obs]
        tc::bulk_g2s(dst + win * F, a.obs + (size_t)g * F, F * 4u, fb);
        if (nrows > 0) {
          const float* nodes_g = a.st.nodes + (size_t)g * C * F;
          float* d = dst + (size_t)(win - nrows) * F;
          const int first = gcm_slot(cnt - nrows, C);
          const int n1 = min(nrows, C - first);
          tc::bulk_g2s(d, nodes_g + (size_t)first * F, (uint32_t)n1 * F * 4u, fb);
          if (n1 < nrows) tc::bulk_g2s(d + (size_t)n1 * F, nodes_g, (uint32_t)(nrows - n1) * F * 4u, fb);
        }
      }
    }
  } else if (warp >= MMA_WARP0) {
    // =============================== MMA issuers ===============================
    const int grp = warp - MMA_WARP0;
    if (lane == 0) {
      const uint32_t tcol = tbase + grp * TC_COL_GROUP;
      const uint32_t idesc1 = tc::idesc_tf32(128, TC_H), idesc2 = tc::idesc_tf32(128, 64);
      const uint32_t sbo1 = (uint32_t)(K1 / 4) * 128u, sbo2 = (uint32_t)(TC_H / 4) * 128u;
      const uint32_t b1hi = tc::smem_u32(B1hi), b1lo = tc::smem_u32(B1lo);
      const uint32_t b2hi = tc::smem_u32(B2hi), b2lo = tc::smem_u32(B2lo);
      int it = 0;
      for (int j = grp; j < my_tiles; j += TC_GROUPS, ++it) {
        const uint32_t ph = it & 1;
        tc::mbar_wait(a1_ready + grp, ph);
        tc::fence_after_sync();
        bool acc = false;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {          // lo*Bhi, hi*Blo, hi*Bhi
          const uint32_t acol = tcol + (pass == 0 ? TC_COL_ALO : TC_COL_AHI);
          const uint32_t bsm = pass == 1 ? b1lo : b1hi;
#pragma unroll
          for (int ks = 0; ks < K1 / 8; ++ks) {
            tc::mma_tf32_ts(tcol + TC_COL_D1, acol + ks * 8, tc::smem_desc_kmajor(bsm + ks * 256, 128, sbo1), idesc1, acc);
            acc = true;
          }
        }
        tc::mma_commit(d1_ready + grp);
        tc::mbar_wait(a2_ready + grp, ph);
        tc::fence_after_sync();
        acc = false;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t acol = tcol + (pass == 0 ? 32u : 0u);   // layer-2 operand overlays A hi: hi [0,32), lo [32,64)
          const uint32_t bsm = pass == 1 ? b2lo : b2hi;
#pragma unroll
          for (int ks = 0; ks < TC_H / 8; ++ks) {
            tc::mma_tf32_ts(tcol + TC_COL_D2, acol + ks * 8, tc::smem_desc_kmajor(bsm + ks * 256, 128, sbo2), idesc2, acc);
            acc = true;
          }
        }
        tc::mma_commit(d2_ready + grp);
      }
    }
    __syncwarp();
  } else {
    // =============================== consumers ===============================
    const int grp = warp >> 2, r = warp & 3;
    const uint32_t taddr = tbase + grp * TC_COL_GROUP + ((uint32_t)(r * 32) << 16);
    float* red_g = red + (size_t)grp * 4 * TC_H * TC_G;
    const bool row_in_prog = r < P.nR;
    const int d_r = row_in_prog ? P.rd[r] : 0;             // offset of this warp's row from t
    // in-neighbours of this warp's row, as offsets from t (warp-uniform, kept in registers)
    int nb_o[NB];
    {
      const int nnb = row_in_prog ? P.nnb[r] : 0;
#pragma unroll
      for (int q = 0; q < NB; ++q) nb_o[q] = q < nnb ? P.doff[P.nb[r][q]] : (1 << 30);
    }
    const int act1 = a.gnn.act1, act2 = a.gnn.act2;
    const float* st_base = stages + (size_t)grp * TC_G * gs;
    const float* mine = st_base + (size_t)lane * gs;       // this lane's graph: rows [0,win) history, row win = obs
    const int* cnt_s = cnts + grp * TC_G;

    int it = 0;
    for (int j = grp; j < my_tiles; j += TC_GROUPS, ++it) {
      const uint32_t ph = it & 1;
      const int tile = blockIdx.x + j * gridDim.x;
      const int gt = min(TC_G, B - tile * TC_G);
      tc::mbar_wait(full + grp, ph);
      const bool live = lane < gt;
      const int cnt = cnt_s[lane];
      const int lt = min(cnt, N - 1);
      const bool row_valid = live && row_in_prog && d_r <= lt;
      const float* nb_ptr[NB];
      bool nb_ok[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        nb_ok[q] = row_valid && nb_o[q] <= lt;
        nb_ptr[q] = mine + (win - (nb_ok[q] ? nb_o[q] : 0)) * F;
      }
      const float* x_ptr = mine + (win - (row_valid ? d_r : 0)) * F;

      // ---- layer-1 operand [agg | x] of row (r, graph) -> TMEM, 16 columns at a time ----
#pragma unroll
      for (int c0 = 0; c0 < K1; c0 += 16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int col = c0 + q4 * 4;                     // columns [0,F): agg, [F,2F): own features
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col < F) {
#pragma unroll
            for (int q = 0; q < NB; ++q) {
              float4 t = *reinterpret_cast<const float4*>(nb_ptr[q] + col);
              if (!nb_ok[q]) t = make_float4(0.f, 0.f, 0.f, 0.f);
              v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
          } else {
            v = *reinterpret_cast<const float4*>(x_ptr + (col - F));
            if (!row_valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tc::split_tf32(v.x, hi[q4 * 4 + 0], lo[q4 * 4 + 0]);
          tc::split_tf32(v.y, hi[q4 * 4 + 1], lo[q4 * 4 + 1]);
          tc::split_tf32(v.z, hi[q4 * 4 + 2], lo[q4 * 4 + 2]);
          tc::split_tf32(v.w, hi[q4 * 4 + 3], lo[q4 * 4 + 3]);
        }
        tc::tmem_st16(taddr + TC_COL_AHI + c0, hi);
        tc::tmem_st16(taddr + TC_COL_ALO + c0, lo);
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(a1_ready + grp);

      // ---- state update while the tensor core works: node row, mask row, counter (8 graphs per warp) ----
      for (int gi = r; gi < gt; gi += 4) {
        const int g = tile * TC_G + gi;
        const int c = cnt_s[gi];
        const int ltg = min(c, N - 1);
        const int tslot = gcm_slot(c, C);
        if (lane < F) a.st.nodes[((size_t)g * C + tslot) * F + lane] = st_base[(size_t)gi * gs + win * F + lane];
        uint32_t* masks_g = a.st.masks + (size_t)g * C * 2 * W;
        uint32_t pw = 0u;
        for (int i = 0; i < P.n_past; ++i) {
          const int hop = P.past[i];
          if (hop <= ltg && (hop >> 5) == lane) pw |= 1u << (hop & 31);
        }
        if (lane < W) {
          gcm_st_mask(masks_g + ((size_t)tslot * 2 + 0) * W + lane, pw);
          gcm_st_mask(masks_g + ((size_t)tslot * 2 + 1) * W + lane, 0u);
        }
        if (lane < P.n_future) {
          const int hop = P.future[lane];
          if (hop <= ltg)
            atomicOr(masks_g + ((size_t)gcm_slot(c - hop, C) * 2 + 1) * W + (hop >> 5), 1u << (hop & 31));
        }
        if (lane == 0) __stcg(a.st.count + g, c + 1);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty + grp);          // the stage can be refilled

      // ---- layer-1 epilogue: h = act(D1 + b1) -> layer-2 operand (hi | lo) over the A hi columns ----
      tc::mbar_wait(d1_ready + grp, ph);
      tc::fence_after_sync();
      {
        uint32_t v0[16], v1[16];
        tc::tmem_ld16(taddr + TC_COL_D1, v0);
        tc::tmem_ld16(taddr + TC_COL_D1 + 16, v1);
        tc::wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float h = gcm_act_fast(__uint_as_float(v0[q]) + bias_s[q], act1);
          if (!row_valid) h = 0.0f;                       // rows outside the window contribute nothing
          tc::split_tf32(h, hi[q], lo[q]);
        }
        tc::tmem_st16(taddr + 0, hi);
        tc::tmem_st16(taddr + 32, lo);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float h = gcm_act_fast(__uint_as_float(v1[q]) + bias_s[16 + q], act1);
          if (!row_valid) h = 0.0f;
          tc::split_tf32(h, hi[q], lo[q]);
        }
        tc::tmem_st16(taddr + 16, hi);
        tc::tmem_st16(taddr + 48, lo);
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(a2_ready + grp);

      // ---- layer-2 epilogue: belief = act(root part of row 0 + rel parts of rows 1..3 + b2) ----
      tc::mbar_wait(d2_ready + grp, ph);
      tc::fence_after_sync();
      {
        uint32_t v0[16], v1[16];
        const uint32_t src = taddr + TC_COL_D2 + (r == 0 ? TC_H : 0);
        tc::tmem_ld16(src, v0);
        tc::tmem_ld16(src + 16, v1);
        tc::wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          red_g[(r * TC_H + q) * TC_G + lane] = __uint_as_float(v0[q]);
          red_g[(r * TC_H + 16 + q) * TC_G + lane] = __uint_as_float(v1[q]);
        }
      }
      tc::fence_before_sync();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      if (live) {
        float outv[8];
        bool bad = false;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const int hh = r * 8 + h;
          const float z = ((red_g[(0 * TC_H + hh) * TC_G + lane] + red_g[(1 * TC_H + hh) * TC_G + lane]) +
                           (red_g[(2 * TC_H + hh) * TC_G + lane] + red_g[(3 * TC_H + hh) * TC_G + lane])) +
                          bias_s[TC_H + hh];
          outv[h] = gcm_act_fast(z, act2);
          bad |= !isfinite(outv[h]);
        }
        float4* dst = reinterpret_cast<float4*>(a.belief + (size_t)(tile * TC_G + lane) * TC_H + r * 8);
        dst[0] = make_float4(outv[0], outv[1], outv[2], outv[3]);
        dst[1] = make_float4(outv[4], outv[5], outv[6], outv[7]);
        if (bad) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // red is reused by the next tile
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == PROD_WARP) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tbase, 512);
  }
}

template <int F, int NB>
static int launch_tc(const TemporalWinArgs& a, cudaStream_t stream) {
  const TcSmem L = tc_smem_layout(F, a.win);
  if (L.total + 128 > 227 * 1024) return GCM_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_step_temporal_tc<F, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(temporal_tc): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    attr_set = true;
  }
  const int n_tiles = (a.st.B + TC_G - 1) / TC_G;
  int grid = gcm_num_sms();
  if (grid > n_tiles) grid = n_tiles;
  k_step_temporal_tc<F, NB><<<grid, TC_THREADS, L.total + 128, stream>>>(a);
  return gcm_check_launch("k_step_temporal_tc");
}

int gcm_launch_temporal_tc(const TemporalWinArgs& a, cudaStream_t stream) {
  if (a.gnn.H1 != TC_H || a.gnn.H2 != TC_H || a.prog.nR > 4 || a.prog.nD > TW_MAXD || a.win < 1 ||
      a.win > TW_MAXWIN || !a.gnn.w_rel1 || !a.gnn.w_root1 || !a.gnn.w_rel2 || !a.gnn.w_root2)
    return GCM_ERR_UNSUPPORTED;
  int maxnb = 0;
  for (int r = 0; r < a.prog.nR; ++r) maxnb = a.prog.nnb[r] > maxnb ? a.prog.nnb[r] : maxnb;
  if (maxnb > TC_MAXNB) return GCM_ERR_UNSUPPORTED;
  const bool small = maxnb <= 3;
  switch (a.st.F) {
    case 8: return small ? launch_tc<8, 3>(a, stream) : launch_tc<8, TC_MAXNB>(a, stream);
    case 16: return small ? launch_tc<16, 3>(a, stream) : launch_tc<16, TC_MAXNB>(a, stream);
    case 32: return small ? launch_tc<32, 3>(a, stream) : launch_tc<32, TC_MAXNB>(a, stream);
    default: return GCM_ERR_UNSUPPORTED;
  }
}

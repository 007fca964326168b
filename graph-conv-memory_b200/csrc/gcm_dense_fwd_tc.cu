// Pure-temporal DenseGCM step on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Why tensor cores here: with the adjacency implicit, one step of B graphs is the batched product
// [4B, 2F] x [2F, 32] (layer 1 on the 4 rows of each graph's 2-hop in-neighbourhood) followed by
// [4B, 32] x [32, 64] (layer 2, rel | root halves).  On CUDA cores this costs ~950 warp-instructions per
// graph and the step is issue-bound at 7x the HBM time (profiles/c2_step_temporal_win_r1.md); on
// tcgen05 the FMAs disappear from the instruction stream.  fp32 accuracy is kept with the 3xTF32 split
// (gcm_tc.cuh; 5e-7 relative on B200, tests/test_tc_gpu.py).
//
// One persistent CTA per SM, 15 warps:
//   warps 0-11       three consumer groups of 4 warps, each owns a 32-graph tile: warp r of a group holds
//                    row r of R1 for all 32 graphs (lane = graph), so TMEM lane = 32 r + graph and the
//                    neighbour program is warp-uniform.  They build [agg | x] (hi, lo) straight into TMEM
//                    (tcgen05.st), apply bias + activation on the accumulator (tcgen05.ld), write it back
//                    as the layer-2 operand, and reduce the 4 rows of each graph through shared memory.
//                    Warp 3 of a group also issues the group's MMAs (one elected lane) once the other
//                    three have arrived on a named barrier; warps 0-2 carry one state-update duty each
//                    (node rows / adjacency rows / counters) while the tensor core works.
//   warps 12-14      producers (one per group's stage): 16-byte cp.async (coalesced, 512 B per instruction) of each graph's
//                    history window + the observation tile into the group's shared-memory stage, completion
//                    signalled with cp.async.mbarrier.arrive.  (Round-1 profile: one warp issuing 96
//                    lane-serialised bulk copies per tile was the throughput limiter, profiles/.)  A stage
//                    is released as soon as the operand is built, so the next tile's copies overlap the
//                    two MMA phases and epilogues of the current one.  Warp 12 owns the TMEM allocation.
// TMEM columns per group (160): A hi [0,64) | A lo [64,128) | D1 [128,160); the layer-2 operand (hi | lo)
// overlays A hi and the layer-2 accumulator overlays A lo once the layer-1 MMAs have completed.
#include "gcm_tc.cuh"
#include "gcm_temporal.cuh"

constexpr int TC_G = 32;                  // graphs per tile
constexpr int TC_GROUPS = 3;
constexpr int TC_NPROD = TC_GROUPS;       // producer warps: ONE PER STAGE (see the producer loop)
constexpr int TC_CONS_THREADS = 4 * TC_GROUPS * 32;
constexpr int TC_THREADS = TC_CONS_THREADS + TC_NPROD * 32;
constexpr int TC_MAXNB = 6;               // in-neighbours per row held in registers
constexpr int TC_H = 32;                  // H1 == H2 == 32
constexpr uint32_t TC_COL_AHI = 0, TC_COL_ALO = 64, TC_COL_D1 = 128, TC_COL_D2 = 64, TC_COL_GROUP = 160;
// named barriers: 0 = whole CTA, 1..3 = layer-2 reduction of group g, 4..6 / 7..9 = "operand of layer 1 / 2
// of group g is in TMEM" (warps 0-2 arrive, warp 3 syncs and issues the MMAs), 10 = weights staged
constexpr int TC_BAR_RED = 1, TC_BAR_A1 = 4, TC_BAR_A2 = 7, TC_BAR_W = 10;

struct TcSmem {   // offsets in bytes into dynamic shared memory
  uint32_t b1hi, b1lo, b2hi, b2lo, stage, red, bias, bars, tmem_slot, total;
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
  L.bias = o; o += 2 * TC_H * 4;
  L.bars = o; o += 4 * TC_GROUPS * 8;
  L.tmem_slot = o; o += 16;
  L.total = o;
  return L;
}

__device__ __forceinline__ void tc_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst_smem)), "l"(src) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its earlier cp.async have landed
__device__ __forceinline__ void tc_cp_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tc_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int F, int NB>
__global__ void __launch_bounds__(TC_THREADS, 1) k_step_temporal_tc(const TemporalWinArgs a) {
  constexpr int K1 = 2 * F;
  constexpr int CPR = F / 4;                 // 16-byte chunks per node row
  constexpr int PROD_WARP0 = 4 * TC_GROUPS;
  extern __shared__ __align__(128) unsigned char sm[];
  const TcSmem L = tc_smem_layout(F, a.win);
  float* B1hi = reinterpret_cast<float*>(sm + L.b1hi);
  float* B1lo = reinterpret_cast<float*>(sm + L.b1lo);
  float* B2hi = reinterpret_cast<float*>(sm + L.b2hi);
  float* B2lo = reinterpret_cast<float*>(sm + L.b2lo);
  float* stages = reinterpret_cast<float*>(sm + L.stage);
  float* red = reinterpret_cast<float*>(sm + L.red);
  float* bias_s = reinterpret_cast<float*>(sm + L.bias);   // [b1 | b2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
  uint64_t* full = bars;                       // [group]
  uint64_t* empty = bars + TC_GROUPS;
  uint64_t* d1_ready = bars + 2 * TC_GROUPS;
  uint64_t* d2_ready = bars + 3 * TC_GROUPS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.st.N, C = a.st.C, W = a.st.W, B = a.st.B, win = a.win;
  const TemporalProg& P = a.prog;
  const int gs = (int)L.gs_floats;
  const int n_tiles = (B + TC_G - 1) / TC_G;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  const bool uni = a.uniform_count >= 0;

  if (tid == 0) {
    for (int g = 0; g < TC_GROUPS; ++g) {
      tc::mbar_init(full + g, 32);      // one cp.async-completion arrival per producer lane
      tc::mbar_init(empty + g, 4);
      tc::mbar_init(d1_ready + g, 1);
      tc::mbar_init(d2_ready + g, 1);
    }
    tc::mbar_fence_init();
  }
  if (warp == PROD_WARP0) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp >= PROD_WARP0) {
    // =============================== producers ===============================
    // 16-byte cp.async per lane: consecutive lanes cover consecutive chunks of one graph's history window
    // (coalesced 512 B per instruction).  Producer p only ever fills stage p (tiles j = p, p + 3, ...): the one-bit
    // phase parity of a stage's mbarriers is unambiguous only if its fills are issued in order by one warp -- with two
    // producers alternating over three stages a producer could run a whole tile ahead of a slow consumer group,
    // see a stale parity on `empty` and refill a stage that was still being read (hang at B = 65536 on full graphs).
    static_assert(TC_NPROD == TC_GROUPS, "one producer warp per stage");
    for (int j = warp - PROD_WARP0; j < my_tiles; j += TC_NPROD) {
      const int grp = j % TC_GROUPS, it = j / TC_GROUPS;
      const int tile = blockIdx.x + j * gridDim.x;
      const int g0 = tile * TC_G;
      const int gt = min(TC_G, B - g0);
      float* st_base = stages + (size_t)grp * TC_G * gs;
      // per-lane view of graph g0 + lane (used when the counts differ between graphs)
      int nrows_l = 0, first_l = 0;
      {
        int cnt = 0;
        if (uni) cnt = a.uniform_count;
        else if (lane < gt) cnt = __ldcg(a.st.count + g0 + lane);
        nrows_l = min(min(cnt, N - 1), win);
        first_l = gcm_slot(cnt - nrows_l, C);
      }
      tc::mbar_wait(empty + grp, (it & 1) ^ 1);
      // observation tile: contiguous in global memory, one row per graph in the stage
#pragma unroll
      for (int k = 0; k < CPR; ++k) {
        const int c = lane + 32 * k;
        const int gi = c / CPR, col = c - gi * CPR;
        if (gi < gt) tc_cp16(st_base + (size_t)gi * gs + win * F + col * 4, a.obs + (size_t)g0 * F + c * 4);
      }
      // history windows: stage row i holds the node at offset (win - i) from t
      const float* nodes0 = a.st.nodes + (size_t)g0 * C * F;
      const int chunks = win * CPR;
      for (int c0 = 0; c0 < chunks; c0 += 32) {
        const int c = c0 + lane;
        const int row = c / CPR, col = c - row * CPR;
        const bool in_range = c < chunks;
        if (uni) {
          const int h = row - (win - nrows_l);       // index into the valid history, 0 = oldest
          if (in_range && h >= 0) {
            int slot = first_l + h;
            if (slot >= C) slot -= C;
            const float* src = nodes0 + (size_t)slot * F + col * 4;
            float* dst = st_base + row * F + col * 4;
#pragma unroll 8
            for (int gi = 0; gi < gt; ++gi) tc_cp16(dst + (size_t)gi * gs, src + (size_t)gi * C * F);
          }
        } else {
          for (int gi = 0; gi < gt; ++gi) {
            const int nrows_g = __shfl_sync(GCM_FULL_MASK, nrows_l, gi);
            const int first_g = __shfl_sync(GCM_FULL_MASK, first_l, gi);
            const int h = row - (win - nrows_g);
            if (in_range && h >= 0) {
              int slot = first_g + h;
              if (slot >= C) slot -= C;
              tc_cp16(st_base + (size_t)gi * gs + row * F + col * 4,
                      nodes0 + ((size_t)gi * C + slot) * F + col * 4);
            }
          }
        }
      }
      tc_cp_arrive(full + grp);
    }
  } else {
    // =============================== consumers ===============================
    const int grp = warp >> 2, r = warp & 3;
    // layer weights -> canonical K-major B operands, split hi / lo (all consumer threads)
#pragma unroll 4
    for (int i = tid; i < TC_H * K1; i += TC_CONS_THREADS) {
      const int n = i / K1, k = i - n * K1;
      const float w = k < F ? __ldg(a.gnn.w_rel1 + n * F + k) : __ldg(a.gnn.w_root1 + n * F + (k - F));
      uint32_t hi, lo;
      tc::split_tf32(w, hi, lo);
      B1hi[tc::kmajor_off(n, k, K1)] = __uint_as_float(hi);
      B1lo[tc::kmajor_off(n, k, K1)] = __uint_as_float(lo);
    }
#pragma unroll 4
    for (int i = tid; i < 64 * TC_H; i += TC_CONS_THREADS) {
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
    tc::fence_proxy_async();           // the tensor core reads the operands through the async proxy
    tc_bar_sync(TC_BAR_W, TC_CONS_THREADS);

    const uint32_t tcol = tbase + grp * TC_COL_GROUP;
    const uint32_t taddr = tcol + ((uint32_t)(r * 32) << 16);
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
    float* st_base = stages + (size_t)grp * TC_G * gs;
    const float* mine = st_base + (size_t)lane * gs;       // this lane's graph: rows [0,win) history, row win = obs
    // MMA descriptors (used by warp 3 of the group)
    const uint32_t idesc1 = tc::idesc_tf32(128, TC_H), idesc2 = tc::idesc_tf32(128, 64);
    const uint32_t sbo1 = (uint32_t)(K1 / 4) * 128u, sbo2 = (uint32_t)(TC_H / 4) * 128u;
    const uint32_t b1hi = tc::smem_u32(B1hi), b1lo = tc::smem_u32(B1lo);
    const uint32_t b2hi = tc::smem_u32(B2hi), b2lo = tc::smem_u32(B2lo);

    int it = 0;
    for (int j = grp; j < my_tiles; j += TC_GROUPS, ++it) {
      const uint32_t ph = it & 1;
      const int tile = blockIdx.x + j * gridDim.x;
      const int g0 = tile * TC_G;
      const int gt = min(TC_G, B - g0);
      const bool live = lane < gt;
      int cnt = 0;
      if (uni) cnt = a.uniform_count;
      else if (live) cnt = __ldcg(a.st.count + g0 + lane);
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
      tc::mbar_wait(full + grp, ph);

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
      if (r != 3) {
        tc_bar_arrive(TC_BAR_A1 + grp, 128);
      } else {
        tc_bar_sync(TC_BAR_A1 + grp, 128);
        tc::fence_after_sync();
        if (lane == 0) {
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
        }
        __syncwarp();
      }

      // ---- state update while the tensor core works (one duty per warp, lane = graph) ----
      const int tslot = gcm_slot(cnt, C);
      if (r == 0) {
        // node rows: 32 / CPR graphs per instruction, 128-bit coalesced stores
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const int gi = i * (32 / CPR) + lane / CPR, col = lane % CPR;
          const int ts = __shfl_sync(GCM_FULL_MASK, tslot, gi);
          if (gi < gt) {
            const float4 v = *reinterpret_cast<const float4*>(st_base + (size_t)gi * gs + win * F + col * 4);
            *reinterpret_cast<float4*>(a.st.nodes + ((size_t)(g0 + gi) * C + ts) * F + col * 4) = v;
          }
        }
      } else if (r == 1) {
        // adjacency rows of the new node: past mask, cleared future mask (contiguous 2W words per graph)
        if (live) {
          uint32_t* mrow = a.st.masks + ((size_t)(g0 + lane) * C + tslot) * 2 * W;
          if ((W & 3) == 0) {
            for (int w4 = 0; w4 < W; w4 += 4) {
              uint32_t pw[4] = {0u, 0u, 0u, 0u};
              for (int i = 0; i < P.n_past; ++i) {
                const int hop = P.past[i];
                const int wi = (hop >> 5) - w4;
                if (hop <= lt && wi >= 0 && wi < 4) {
                  const uint32_t bit = 1u << (hop & 31);
                  pw[0] |= wi == 0 ? bit : 0u; pw[1] |= wi == 1 ? bit : 0u;
                  pw[2] |= wi == 2 ? bit : 0u; pw[3] |= wi == 3 ? bit : 0u;
                }
              }
              __stcg(reinterpret_cast<uint4*>(mrow + w4), make_uint4(pw[0], pw[1], pw[2], pw[3]));
              __stcg(reinterpret_cast<uint4*>(mrow + W + w4), make_uint4(0u, 0u, 0u, 0u));
            }
          } else {
            for (int w = 0; w < W; ++w) {
              uint32_t pw = 0u;
              for (int i = 0; i < P.n_past; ++i) {
                const int hop = P.past[i];
                if (hop <= lt && (hop >> 5) == w) pw |= 1u << (hop & 31);
              }
              gcm_st_mask(mrow + w, pw);
              gcm_st_mask(mrow + W + w, 0u);
            }
          }
        }
      } else if (r == 2) {
        if (live) {
          uint32_t* masks_g = a.st.masks + (size_t)(g0 + lane) * C * 2 * W;
          for (int i = 0; i < P.n_future; ++i) {
            const int hop = P.future[i];
            if (hop <= lt)
              atomicOr(masks_g + ((size_t)gcm_slot(cnt - hop, C) * 2 + 1) * W + (hop >> 5), 1u << (hop & 31));
          }
          __stcg(a.st.count + g0 + lane, cnt + 1);
        }
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
        float h[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) h[q] = __uint_as_float(v0[q]) + bias_s[q];
        gcm_act_fast_vec(h, act1);
        // row 0 is the new node itself: keep its layer-1 output for the cached-row kernel (gcm_dense_fwd_hc.cu)
        float4* hc_dst = nullptr;
        if (r == 0 && live && a.hcache)
          hc_dst = reinterpret_cast<float4*>(a.hcache + ((size_t)(g0 + lane) * a.hc_ring + (cnt & (a.hc_ring - 1))) * TC_H);
        if (hc_dst) {
#pragma unroll
          for (int q = 0; q < 4; ++q) __stcg(hc_dst + q, make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]));
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) tc::split_tf32(row_valid ? h[q] : 0.0f, hi[q], lo[q]);   // rows outside the window: 0
        tc::tmem_st16(taddr + 0, hi);
        tc::tmem_st16(taddr + 32, lo);
#pragma unroll
        for (int q = 0; q < 16; ++q) h[q] = __uint_as_float(v1[q]) + bias_s[16 + q];
        gcm_act_fast_vec(h, act1);
        if (hc_dst) {
#pragma unroll
          for (int q = 0; q < 4; ++q) __stcg(hc_dst + 4 + q, make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]));
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) tc::split_tf32(row_valid ? h[q] : 0.0f, hi[q], lo[q]);
        tc::tmem_st16(taddr + 16, hi);
        tc::tmem_st16(taddr + 48, lo);
      }
      tc::wait_st();
      tc::fence_before_sync();
      if (r != 3) {
        tc_bar_arrive(TC_BAR_A2 + grp, 128);
      } else {
        tc_bar_sync(TC_BAR_A2 + grp, 128);
        tc::fence_after_sync();
        if (lane == 0) {
          bool acc = false;
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
        __syncwarp();
      }

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
      tc_bar_sync(TC_BAR_RED + grp, 128);
      if (live) {
        float outv[8];
        bool bad = false;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const int hh = r * 8 + h;
          outv[h] = ((red_g[(0 * TC_H + hh) * TC_G + lane] + red_g[(1 * TC_H + hh) * TC_G + lane]) +
                     (red_g[(2 * TC_H + hh) * TC_G + lane] + red_g[(3 * TC_H + hh) * TC_G + lane])) +
                    bias_s[TC_H + hh];
        }
        gcm_act_fast_vec(outv, act2);
#pragma unroll
        for (int h = 0; h < 8; ++h) bad |= !isfinite(outv[h]);
        float4* dst = reinterpret_cast<float4*>(a.belief + (size_t)(g0 + lane) * TC_H + r * 8);
        dst[0] = make_float4(outv[0], outv[1], outv[2], outv[3]);
        dst[1] = make_float4(outv[4], outv[5], outv[6], outv[7]);
        if (bad) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
      }
      tc_bar_sync(TC_BAR_RED + grp, 128);   // red is reused by the next tile
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == PROD_WARP0) {
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

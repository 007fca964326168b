// Pure-temporal DenseGCM step, "thread = graph" tensor-core kernel (tcgen05 + TMEM), sm_100a.
//
// Same arithmetic as gcm_dense_fwd_tc.cu (3-term split of every fp32 product, fp32 accumulate), different
// mapping, chosen after profiling that kernel (profiles/c2_step_temporal_tc_r1.md): there a tile was
// 32 graphs x 4 rows of R1 spread over 4 warps, which cost a cross-warp reduction through shared memory,
// two named barriers and ~220 warp-instructions per graph, and only 3 tiles fit in TMEM.  Here
//   * a tile is 128 graphs and TMEM lane = graph: one thread owns a graph for the whole step, so the
//     2-hop neighbourhood never crosses a lane.  Row r of R1 is its own M=128 MMA
//     (A_r [128 graphs, 2F] x B1 -> D1_r [128, 32]); layer 2 contracts [h_0 | h_1 | h_2 + h_3] (K = 96)
//     against [W_root2 ; W_rel2 ; W_rel2], i.e. the sum over the in-neighbours of t is done by the tensor
//     core and nothing is reduced across threads.
//   * the low-order term of the split (lo * B) runs on the bf16 path: lo has 13 significant bits and is
//     2^-11 of the value, so 8 bits of it are enough (gcm_tc_selftest passes == 4: 1.3e-6 relative).
//     Packed bf16 halves the TMEM columns of the lo operand, which is what lets all four rows of a tile
//     (4 x 96 operand + 4 x 32 accumulator columns = 512) be resident at once.
//   * 8 consumer warps: warp (q, hf) owns graphs 32q..32q+31 of the tile (TMEM lane quadrant q) and rows
//     {2hf, 2hf+1}; hf = 0 also writes the node rows, hf = 1 the adjacency rows and counters.  Warps 3 / 7
//     issue the MMAs of their half after a named-barrier hand-off (bar.arrive / bar.sync).
//   * 2 producer warps stream each quarter tile (32 graphs: history window + observation) with 16-byte
//     cp.async into shared memory, completion by cp.async.mbarrier.arrive; a quarter is released as soon
//     as both of its warps have built their operands, so the next tile's loads run under the two MMA
//     phases and epilogues of the current one.
// TMEM columns: A_r = 96 r + [hi: 0..2F) | lo (bf16x2): 64..64+F) ; D1_r = 384 + 32 r.
// Layer 2 reuses them once the layer-1 MMAs of the owning half have completed: operand pieces h_0 / h_1 in
// A_0 (hi 0 / 32, lo 64 / 80), h_2 + h_3 in A_2 (hi 192, lo 256); output columns 0..15 accumulate in D1_0's
// columns (read by the hf = 0 warps) and 16..31 in D1_2's (hf = 1), so the next tile's MMAs cannot touch
// them before their readers have arrived on that half's barrier.
#include <cuda_bf16.h>

#include "gcm_tc.cuh"
#include "gcm_temporal.cuh"

constexpr int TG_Q = 32;                    // graphs per quarter tile (one warp's lanes)
constexpr int TG_NPROD = 2;
constexpr int TG_CONS_THREADS = 8 * 32;
constexpr int TG_THREADS = TG_CONS_THREADS + TG_NPROD * 32;
constexpr int TG_MAXNB = 6;
constexpr int TG_H = 32;
constexpr uint32_t TG_COL_D1 = 384;
// named barriers: 1 + 2 hf + rr = "row 2 hf + rr of every quadrant is in TMEM", 5 = layer-2 operand ready,
// 6 = weights staged
constexpr int TG_BAR_ROW = 1, TG_BAR_L2 = 5, TG_BAR_W = 6;

struct TgSmem {   // byte offsets into dynamic shared memory
  uint32_t b1hi, b1lo, b1bf, b2hi, b2lo, b2bf, zero, bias, bars, tmem_slot, stage, total;
  uint32_t gs_floats;
};

__host__ __device__ inline TgSmem tg_smem_layout(int F, int win) {
  TgSmem L;
  const uint32_t K1 = 2 * F;
  uint32_t o = 0;
  L.b1hi = o; o += TG_H * K1 * 4;
  L.b1lo = o; o += TG_H * K1 * 4;
  L.b1bf = o; o += TG_H * K1 * 2;
  L.b2hi = o; o += 2 * TG_H * TG_H * 4;     // [rel | root], each [32 x 32] K-major
  L.b2lo = o; o += 2 * TG_H * TG_H * 4;
  L.b2bf = o; o += 2 * TG_H * TG_H * 2;
  L.zero = o; o += 128;                     // a row of zeros: target of out-of-window neighbour pointers
  L.bias = o; o += 2 * TG_H * 4;
  L.bars = o; o += 16 * 8;
  L.tmem_slot = o; o += 16;
  L.gs_floats = (uint32_t)(win * F + F + 4);
  L.stage = o; o += 4u * TG_Q * L.gs_floats * 4;
  L.total = o;
  return L;
}

__device__ __forceinline__ void tg_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void tg_cp_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tg_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tg_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 16 fp32 values -> hi (tf32, 16 columns at col_hi) and lo (packed bf16, 8 columns at col_lo) of this lane
__device__ __forceinline__ void tg_store_split16(uint32_t lane_addr, uint32_t col_hi, uint32_t col_lo,
                                                 const float (&v)[16]) {
  uint32_t hi[16], pk[8];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    uint32_t l0, l1;
    tc::split_tf32(v[j], hi[j], l0);
    tc::split_tf32(v[j + 1], hi[j + 1], l1);
    pk[j >> 1] = tc::pack_bf16(__uint_as_float(l0), __uint_as_float(l1));
  }
  tc::tmem_st16(lane_addr + col_hi, hi);
  tc::tmem_st8(lane_addr + col_lo, pk);
}

template <int F, int NB>
__global__ void __launch_bounds__(TG_THREADS, 1) k_step_temporal_tg(const TemporalWinArgs a) {
  constexpr int K1 = 2 * F;
  constexpr int CPR = F / 4;
  extern __shared__ __align__(128) unsigned char sm[];
  const TgSmem L = tg_smem_layout(F, a.win);
  float* B1hi = reinterpret_cast<float*>(sm + L.b1hi);
  float* B1lo = reinterpret_cast<float*>(sm + L.b1lo);
  __nv_bfloat16* B1bf = reinterpret_cast<__nv_bfloat16*>(sm + L.b1bf);
  float* B2hi = reinterpret_cast<float*>(sm + L.b2hi);
  float* B2lo = reinterpret_cast<float*>(sm + L.b2lo);
  __nv_bfloat16* B2bf = reinterpret_cast<__nv_bfloat16*>(sm + L.b2bf);
  float* zero_row = reinterpret_cast<float*>(sm + L.zero);
  float* bias_s = reinterpret_cast<float*>(sm + L.bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
  float* stages = reinterpret_cast<float*>(sm + L.stage);
  uint64_t* full = bars;            // [quarter]
  uint64_t* empty = bars + 4;       // [quarter]
  uint64_t* d1_ready = bars + 8;    // [half]
  uint64_t* d2_ready = bars + 10;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.st.N, C = a.st.C, W = a.st.W, B = a.st.B, win = a.win;
  const TemporalProg& P = a.prog;
  const int gs = (int)L.gs_floats;
  const bool uni = a.uniform_count >= 0;
  // this CTA's contiguous range of quarter tiles
  const int nq_total = (B + TG_Q - 1) / TG_Q;
  const int q_begin = (int)(((long long)nq_total * blockIdx.x) / gridDim.x);
  const int q_end = (int)(((long long)nq_total * (blockIdx.x + 1)) / gridDim.x);
  const int my_tiles = (q_end - q_begin + 3) / 4;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(full + i, 32);
      tc::mbar_init(empty + i, 2);
    }
    tc::mbar_init(d1_ready + 0, 1);
    tc::mbar_init(d1_ready + 1, 1);
    tc::mbar_init(d2_ready, 1);
    tc::mbar_fence_init();
  }
  if (tid < 32) zero_row[tid] = 0.0f;
  if (warp == 8) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp >= 8) {
    // =============================== producers ===============================
    const int p = warp - 8;
    for (int ti = 0; ti < my_tiles; ++ti) {
      for (int q = p; q < 4; q += TG_NPROD) {
        const int qidx = q_begin + ti * 4 + q;
        if (qidx >= q_end) continue;
        const int g0 = qidx * TG_Q;
        const int gt = min(TG_Q, B - g0);
        float* st_base = stages + (size_t)q * TG_Q * gs;
        int nrows_l = 0, first_l = 0;
        {
          int cnt = 0;
          if (uni) cnt = a.uniform_count;
          else if (lane < gt) cnt = __ldcg(a.st.count + g0 + lane);
          nrows_l = min(min(cnt, N - 1), win);
          first_l = gcm_slot(cnt - nrows_l, C);
        }
        tc::mbar_wait(empty + q, (ti & 1) ^ 1);
#pragma unroll
        for (int k = 0; k < CPR; ++k) {
          const int c = lane + 32 * k;
          const int gi = c / CPR, col = c - gi * CPR;
          if (gi < gt) tg_cp16(st_base + (size_t)gi * gs + win * F + col * 4, a.obs + (size_t)g0 * F + c * 4);
        }
        const float* nodes0 = a.st.nodes + (size_t)g0 * C * F;
        const int chunks = win * CPR;
        for (int c0 = 0; c0 < chunks; c0 += 32) {
          const int c = c0 + lane;
          const int row = c / CPR, col = c - row * CPR;
          const bool in_range = c < chunks;
          if (uni) {
            const int h = row - (win - nrows_l);
            if (in_range && h >= 0) {
              int slot = first_l + h;
              if (slot >= C) slot -= C;
              const float* src = nodes0 + (size_t)slot * F + col * 4;
              float* dst = st_base + row * F + col * 4;
#pragma unroll 8
              for (int gi = 0; gi < gt; ++gi) tg_cp16(dst + (size_t)gi * gs, src + (size_t)gi * C * F);
            }
          } else {
            for (int gi = 0; gi < gt; ++gi) {
              const int nrows_g = __shfl_sync(GCM_FULL_MASK, nrows_l, gi);
              const int first_g = __shfl_sync(GCM_FULL_MASK, first_l, gi);
              const int h = row - (win - nrows_g);
              if (in_range && h >= 0) {
                int slot = first_g + h;
                if (slot >= C) slot -= C;
                tg_cp16(st_base + (size_t)gi * gs + row * F + col * 4,
                        nodes0 + ((size_t)gi * C + slot) * F + col * 4);
              }
            }
          }
        }
        tg_cp_arrive(full + q);
      }
    }
  } else {
    // =============================== consumers ===============================
    const int q = warp & 3, hf = warp >> 2;
    // ---- layer weights -> canonical K-major B operands: hi / lo (tf32) and a bf16 copy ----
    {
      constexpr int PER1 = (TG_H * K1 + TG_CONS_THREADS - 1) / TG_CONS_THREADS;
      float w[PER1];
#pragma unroll
      for (int j = 0; j < PER1; ++j) {
        const int i = tid + j * TG_CONS_THREADS;
        const int n = i / K1, k = i - n * K1;
        w[j] = 0.0f;
        if (i < TG_H * K1) w[j] = k < F ? __ldg(a.gnn.w_rel1 + n * F + k) : __ldg(a.gnn.w_root1 + n * F + (k - F));
      }
#pragma unroll
      for (int j = 0; j < PER1; ++j) {
        const int i = tid + j * TG_CONS_THREADS;
        const int n = i / K1, k = i - n * K1;
        if (i < TG_H * K1) {
          uint32_t hi, lo;
          tc::split_tf32(w[j], hi, lo);
          B1hi[tc::kmajor_off(n, k, K1)] = __uint_as_float(hi);
          B1lo[tc::kmajor_off(n, k, K1)] = __uint_as_float(lo);
          B1bf[tc::kmajor_off_bf16(n, k, K1)] = __float2bfloat16_rn(w[j]);
        }
      }
      constexpr int PER2 = 2 * TG_H * TG_H / TG_CONS_THREADS;
      float w2[PER2];
#pragma unroll
      for (int j = 0; j < PER2; ++j) {
        const int i = tid + j * TG_CONS_THREADS;            // [which][n][k]
        const int which = i / (TG_H * TG_H), r = i - which * TG_H * TG_H;
        w2[j] = __ldg((which == 0 ? a.gnn.w_rel2 : a.gnn.w_root2) + r);
      }
#pragma unroll
      for (int j = 0; j < PER2; ++j) {
        const int i = tid + j * TG_CONS_THREADS;
        const int which = i / (TG_H * TG_H), r = i - which * TG_H * TG_H;
        const int n = r / TG_H, k = r - n * TG_H;
        uint32_t hi, lo;
        tc::split_tf32(w2[j], hi, lo);
        B2hi[which * TG_H * TG_H + tc::kmajor_off(n, k, TG_H)] = __uint_as_float(hi);
        B2lo[which * TG_H * TG_H + tc::kmajor_off(n, k, TG_H)] = __uint_as_float(lo);
        B2bf[which * TG_H * TG_H + tc::kmajor_off_bf16(n, k, TG_H)] = __float2bfloat16_rn(w2[j]);
      }
      if (tid < TG_H) {
        bias_s[tid] = a.gnn.b1 ? __ldg(a.gnn.b1 + tid) : 0.0f;
        bias_s[TG_H + tid] = a.gnn.b2 ? __ldg(a.gnn.b2 + tid) : 0.0f;
      }
      tc::fence_proxy_async();
      tg_bar_sync(TG_BAR_W, TG_CONS_THREADS);
    }

    const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
    const int nR = P.nR;
    const int rows_here = max(0, min(2, nR - 2 * hf));      // rows of R1 owned by this half
    const int act1 = a.gnn.act1, act2 = a.gnn.act2;
    float* st_base = stages + (size_t)q * TG_Q * gs;
    const float* mine = st_base + (size_t)lane * gs;
    const bool issuer = q == 3;
    const uint32_t idesc1t = tc::idesc_tf32(128, TG_H), idesc1b = tc::idesc_bf16(128, TG_H);
    const uint32_t idesc2t = tc::idesc_tf32(128, 16), idesc2b = tc::idesc_bf16(128, 16);
    const uint32_t sb1hi = tc::smem_u32(B1hi), sb1lo = tc::smem_u32(B1lo), sb1bf = tc::smem_u32(B1bf);
    const uint32_t sb2hi = tc::smem_u32(B2hi), sb2lo = tc::smem_u32(B2lo), sb2bf = tc::smem_u32(B2bf);

    for (int ti = 0; ti < my_tiles; ++ti) {
      const uint32_t ph = ti & 1;
      const int qidx = q_begin + ti * 4 + q;
      const bool has = qidx < q_end;
      const int g0 = qidx * TG_Q;
      const int gt = has ? min(TG_Q, B - g0) : 0;
      const bool live = lane < gt;
      int cnt = 0;
      if (uni) cnt = a.uniform_count;
      else if (live) cnt = __ldcg(a.st.count + g0 + lane);
      const int lt = min(cnt, N - 1);
      if (has) tc::mbar_wait(full + q, ph);

      // ---- layer 1: operand rows of this half -> TMEM, one M=128 MMA chain per row ----
      for (int rr = 0; rr < rows_here; ++rr) {
        const int r = 2 * hf + rr;
        if (has) {
          const int d_r = P.rd[r];
          const bool row_valid = live && d_r <= lt;
          const int nnb = P.nnb[r];
          const float* nb_ptr[NB];
#pragma unroll
          for (int qn = 0; qn < NB; ++qn) {
            const int off = qn < nnb ? P.doff[P.nb[r][qn]] : (1 << 30);
            nb_ptr[qn] = (row_valid && off <= lt) ? mine + (win - off) * F : zero_row;
          }
          const float* x_ptr = row_valid ? mine + (win - d_r) * F : zero_row;
          const uint32_t a_addr = lane_addr + 96u * r;
#pragma unroll 1
          for (int c0 = 0; c0 < K1; c0 += 16) {
            float v[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int col = c0 + q4 * 4;                   // [0,F): aggregated neighbours, [F,2F): own row
              float4 s;
              if (col < F) {
                s = *reinterpret_cast<const float4*>(nb_ptr[0] + col);
#pragma unroll
                for (int qn = 1; qn < NB; ++qn) {
                  const float4 t = *reinterpret_cast<const float4*>(nb_ptr[qn] + col);
                  s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
                }
              } else {
                s = *reinterpret_cast<const float4*>(x_ptr + (col - F));
              }
              v[q4 * 4 + 0] = s.x; v[q4 * 4 + 1] = s.y; v[q4 * 4 + 2] = s.z; v[q4 * 4 + 3] = s.w;
            }
            tg_store_split16(a_addr, c0, 64 + c0 / 2, v);
          }
          tc::wait_st();
        }
        tc::fence_before_sync();
        const int bar_id = TG_BAR_ROW + 2 * hf + rr;
        if (!issuer) {
          tg_bar_arrive(bar_id, 128);
        } else {
          tg_bar_sync(bar_id, 128);
          tc::fence_after_sync();
          if (lane == 0) {
            const uint32_t acol = tbase + 96u * r, dcol = tbase + TG_COL_D1 + 32u * r;
            bool acc = false;
#pragma unroll
            for (int ks = 0; ks < K1 / 16; ++ks) {          // lo (bf16) * B (bf16)
              tc::mma_bf16_ts(dcol, acol + 64 + ks * 8, tc::smem_desc_kmajor(sb1bf + ks * 256, 128, (K1 / 8) * 128u),
                              idesc1b, acc);
              acc = true;
            }
#pragma unroll
            for (int ks = 0; ks < K1 / 8; ++ks)             // hi * B lo
              tc::mma_tf32_ts(dcol, acol + ks * 8, tc::smem_desc_kmajor(sb1lo + ks * 256, 128, (K1 / 4) * 128u),
                              idesc1t, true);
#pragma unroll
            for (int ks = 0; ks < K1 / 8; ++ks)             // hi * B hi
              tc::mma_tf32_ts(dcol, acol + ks * 8, tc::smem_desc_kmajor(sb1hi + ks * 256, 128, (K1 / 4) * 128u),
                              idesc1t, true);
            if (rr == rows_here - 1) tc::mma_commit(d1_ready + hf);
          }
          __syncwarp();
        }
      }

      // ---- state update while the tensor core works ----
      if (has) {
        const int tslot = gcm_slot(cnt, C);
        if (hf == 0) {
#pragma unroll
          for (int i = 0; i < CPR; ++i) {
            const int gi = i * (32 / CPR) + lane / CPR, col = lane % CPR;
            const int ts = __shfl_sync(GCM_FULL_MASK, tslot, gi);
            if (gi < gt) {
              const float4 v = *reinterpret_cast<const float4*>(st_base + (size_t)gi * gs + win * F + col * 4);
              *reinterpret_cast<float4*>(a.st.nodes + ((size_t)(g0 + gi) * C + ts) * F + col * 4) = v;
            }
          }
        } else if (live) {
          uint32_t* masks_g = a.st.masks + (size_t)(g0 + lane) * C * 2 * W;
          uint32_t* mrow = masks_g + (size_t)tslot * 2 * W;
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
          for (int i = 0; i < P.n_future; ++i) {
            const int hop = P.future[i];
            if (hop <= lt)
              atomicOr(masks_g + ((size_t)gcm_slot(cnt - hop, C) * 2 + 1) * W + (hop >> 5), 1u << (hop & 31));
          }
          __stcg(a.st.count + g0 + lane, cnt + 1);
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(empty + q);           // this warp is done with the quarter's stage
      }

      // ---- layer-1 epilogue: h_r = act(D1_r + b1) -> layer-2 operand pieces ----
      if (rows_here > 0) {
        tc::mbar_wait(d1_ready + hf, ph);
        tc::fence_after_sync();
        const bool valid_a = live && P.rd[2 * hf] <= lt;
        const bool valid_b = rows_here > 1 && live && P.rd[2 * hf + 1] <= lt;
#pragma unroll 1
        for (int c0 = 0; c0 < TG_H; c0 += 16) {
          uint32_t va[16], vb[16];
          tc::tmem_ld16(lane_addr + TG_COL_D1 + 32u * (2 * hf) + c0, va);
          if (rows_here > 1) tc::tmem_ld16(lane_addr + TG_COL_D1 + 32u * (2 * hf + 1) + c0, vb);
          tc::wait_ld();
          float ha[16], hb[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float bj = bias_s[c0 + j];
            ha[j] = __uint_as_float(va[j]) + bj;
            hb[j] = rows_here > 1 ? __uint_as_float(vb[j]) + bj : 0.0f;
          }
          gcm_act_fast_vec(ha, act1);
          if (rows_here > 1) gcm_act_fast_vec(hb, act1);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ha[j] = valid_a ? ha[j] : 0.0f;
            hb[j] = valid_b ? hb[j] : 0.0f;
          }
          if (hf == 0) {
            tg_store_split16(lane_addr, 0 + c0, 64 + c0 / 2, ha);            // piece 0: h_0 (root weights)
            if (rows_here > 1) tg_store_split16(lane_addr, 32 + c0, 80 + c0 / 2, hb);   // piece 1: h_1
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) ha[j] += hb[j];
            tg_store_split16(lane_addr, 192 + c0, 256 + c0 / 2, ha);         // piece 2: h_2 + h_3
          }
        }
        tc::wait_st();
      }
      tc::fence_before_sync();
      if (warp != 7) {
        tg_bar_arrive(TG_BAR_L2, TG_CONS_THREADS);
      } else {
        tg_bar_sync(TG_BAR_L2, TG_CONS_THREADS);
        tc::fence_after_sync();
        if (lane == 0) {
          const int n_pieces = nR >= 3 ? 3 : nR;              // h_0 ; h_1 ; h_2 (+ h_3)
#pragma unroll
          for (int nh = 0; nh < 2; ++nh) {                    // output columns [16 nh, 16 nh + 16)
            const uint32_t dcol = tbase + TG_COL_D1 + 64u * nh;
            bool acc = false;
            for (int pc = 0; pc < n_pieces; ++pc) {
              const uint32_t ahi = tbase + (pc == 0 ? 0u : pc == 1 ? 32u : 192u);
              const uint32_t alo = tbase + (pc == 0 ? 64u : pc == 1 ? 80u : 256u);
              const uint32_t wsel = pc == 0 ? 1u : 0u;        // root weights for h_0, rel for the neighbours
              const uint32_t bhi = sb2hi + wsel * TG_H * TG_H * 4 + nh * 2048u;
              const uint32_t blo = sb2lo + wsel * TG_H * TG_H * 4 + nh * 2048u;
              const uint32_t bbf = sb2bf + wsel * TG_H * TG_H * 2 + nh * 1024u;
#pragma unroll
              for (int ks = 0; ks < TG_H / 16; ++ks) {
                tc::mma_bf16_ts(dcol, alo + ks * 8, tc::smem_desc_kmajor(bbf + ks * 256, 128, (TG_H / 8) * 128u), idesc2b, acc);
                acc = true;
              }
#pragma unroll
              for (int ks = 0; ks < TG_H / 8; ++ks)
                tc::mma_tf32_ts(dcol, ahi + ks * 8, tc::smem_desc_kmajor(blo + ks * 256, 128, (TG_H / 4) * 128u), idesc2t, true);
#pragma unroll
              for (int ks = 0; ks < TG_H / 8; ++ks)
                tc::mma_tf32_ts(dcol, ahi + ks * 8, tc::smem_desc_kmajor(bhi + ks * 256, 128, (TG_H / 4) * 128u), idesc2t, true);
            }
          }
          tc::mma_commit(d2_ready);
        }
        __syncwarp();
      }

      // ---- layer-2 epilogue: belief columns [16 hf, 16 hf + 16) of this lane's graph ----
      tc::mbar_wait(d2_ready, ph);
      tc::fence_after_sync();
      {
        uint32_t v[16];
        tc::tmem_ld16(lane_addr + TG_COL_D1 + 64u * hf, v);
        tc::wait_ld();
        if (live) {
          float o[16];
          bool bad = false;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) + bias_s[TG_H + 16 * hf + j];
          gcm_act_fast_vec(o, act2);
#pragma unroll
          for (int j = 0; j < 16; ++j) bad |= !isfinite(o[j]);
          float4* dst = reinterpret_cast<float4*>(a.belief + (size_t)(g0 + lane) * TG_H + 16 * hf);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          if (bad) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
        }
      }
      tc::fence_before_sync();
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tbase, 512);
  }
}

template <int F, int NB>
static int launch_tg(const TemporalWinArgs& a, cudaStream_t stream) {
  const TgSmem L = tg_smem_layout(F, a.win);
  if (L.total + 128 > 227 * 1024) return GCM_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_step_temporal_tg<F, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(temporal_tg): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    attr_set = true;
  }
  const int nq = (a.st.B + TG_Q - 1) / TG_Q;
  int grid = gcm_num_sms();
  if (grid > (nq + 3) / 4) grid = (nq + 3) / 4;
  k_step_temporal_tg<F, NB><<<grid, TG_THREADS, L.total + 128, stream>>>(a);
  return gcm_check_launch("k_step_temporal_tg");
}

int gcm_launch_temporal_tg(const TemporalWinArgs& a, cudaStream_t stream) {
  if (a.gnn.H1 != TG_H || a.gnn.H2 != TG_H || a.prog.nR > 4 || a.prog.nR < 1 || a.prog.nD > TW_MAXD ||
      a.win < 1 || a.win > TW_MAXWIN || !a.gnn.w_rel1 || !a.gnn.w_root1 || !a.gnn.w_rel2 || !a.gnn.w_root2)
    return GCM_ERR_UNSUPPORTED;
  int maxnb = 0;
  for (int r = 0; r < a.prog.nR; ++r) maxnb = a.prog.nnb[r] > maxnb ? a.prog.nnb[r] : maxnb;
  if (maxnb > TG_MAXNB || maxnb < 1) return GCM_ERR_UNSUPPORTED;
  const bool small = maxnb <= 3;
  switch (a.st.F) {
    case 8: return small ? launch_tg<8, 3>(a, stream) : launch_tg<8, TG_MAXNB>(a, stream);
    case 16: return small ? launch_tg<16, 3>(a, stream) : launch_tg<16, TG_MAXNB>(a, stream);
    case 32: return small ? launch_tg<32, 3>(a, stream) : launch_tg<32, TG_MAXNB>(a, stream);
    default: return GCM_ERR_UNSUPPORTED;
  }
}

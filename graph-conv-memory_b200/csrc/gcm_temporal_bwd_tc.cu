// Window-level backward of forward-only TemporalBackedge chains: both layers' row products AND both weight-gradient
// reductions in ONE kernel (tcgen05 / TMEM, 3xTF32).
//
// Reference: autograd through gcm.py:262-321 with the DenseGraphConv stack over the T steps of a BPTT window
// (tests/test_gcm.py:412-439); algebra in csrc/gcm_temporal_bwd.cu.  With the time-major operand rows
//   X_r = [sum_s x_{q-s} | x_q]   (gcm_temporal_gather)        U_r = [sum_s dz2_{q+s} | dz2_q]   (gcm_temporal_shift_sum)
// of row r = (position q, graph b), everything else is row-local plus two reductions over ALL rows:
//   h_r   = act1(X_r W1^T + b1)                     (recomputed, not saved by the forward)
//   dz1_r = (U_r W2t^T) * act1'(h_r)                W2t = [W_rel2^T | W_root2^T]
//   G1 += X_r^T dz1_r   ([2F, H1]  = [dW_rel1^T ; dW_root1^T])
//   G2 += U_r^T h_r     ([2H2, H1] = [dW_rel2 ; dW_root2]:  sum_p dz2_p (sum_s h_{p-s})^T = sum_q (sum_s dz2_{q+s}) h_q^T)
// The separate-launch version (gcm_linear_tc32 x 2, gcm_act_backward, gcm_temporal_shift_sum, gcm_outer_reduce_tc32 x 2)
// streams seven [rows, 32..64] operands through HBM; this kernel reads X and U once (512 B per row) and writes nothing
// but per-CTA partials.  A CTA takes tiles of 128 rows (TMEM lane = row, 4 threads per row with 8 features each):
//   * the row products are TS-form MMAs: the row-owning threads write X and U (hi / lo split) into TMEM with tcgen05.st;
//   * the weight-gradient products contract over the 128 ROWS of the tile, so their operands are [feature][row] matrices:
//     the same threads scatter X, U, dz1, h into shared memory in the canonical K-major core-matrix layout with the row
//     index as K (one float per store; a warp's 32 rows of one feature land in 32 different banks because the distance
//     between K-adjacent core matrices is padded by 16 bytes);
//   * the accumulator G = [X | U]^T [dz1 | h] ([128, 64]; its two diagonal blocks are G1 and G2) stays in TMEM for the
//     whole launch and is flushed once per CTA; a second kernel adds the per-CTA partials in a fixed order.
#include "gcm_common.cuh"
#include "gcm_tc.cuh"

namespace {

constexpr int WB_TILE = 128;          // rows per tile = TMEM lanes
constexpr int WB_THREADS = 512;       // 4 threads per row: 8 of the 32 features each; warp 0's lane 0 also issues the MMAs
constexpr int WB_F = 32;              // F = H1 = H2 = 32 (the cached-row kernel's shape)
constexpr int WB_AW = 4 * WB_F;       // features of the wide A block: [Xsum | Xself | Usum | Uself]
constexpr int WB_BW = 2 * WB_F;       // [dz1 | h]
// TMEM columns: accumulators, then the TS-form A operands (hi / lo) of the two row products
constexpr uint32_t WB_COL_Z1 = 0, WB_COL_DH = 32, WB_COL_G = 64, WB_COL_XHI = 128, WB_COL_XLO = 192, WB_COL_UHI = 256,
                   WB_COL_ULO = 320;
constexpr int WB_PART = WB_AW * WB_F + 2 * WB_F;   // floats per CTA partial: G rows [128][32] + db1 + db2
// [feature][row] operand blocks, K = row: core matrix (8 features x 4 rows) = 128 B; feature groups adjacent (SBO = 128),
// row chunks LBO apart, LBO = (#feature groups) * 128 + 16: the pad makes LBO / 4 = 4 (mod 32), so the 32 rows a warp
// stores for one feature hit 32 different banks
constexpr int WB_A_LBO = (WB_AW / 8) * 128 + 16;   // 2064
constexpr int WB_B_LBO = (WB_BW / 8) * 128 + 16;   // 1040
constexpr int WB_A_BYTES = (WB_TILE / 4) * WB_A_LBO;   // 66 048
constexpr int WB_B_BYTES = (WB_TILE / 4) * WB_B_LBO;   // 33 280

struct WbSmem {
  unsigned char a_hi[WB_A_BYTES];
  unsigned char a_lo[WB_A_BYTES];
  unsigned char b_hi[WB_B_BYTES];
  unsigned char b_lo[WB_B_BYTES];
  float w1_hi[WB_F * 2 * WB_F];  // [H1][2F] canonical K-major, 8 KB each
  float w1_lo[WB_F * 2 * WB_F];
  float w2_hi[WB_F * 2 * WB_F];  // [H1][2H2]: dh = U w2t^T
  float w2_lo[WB_F * 2 * WB_F];
  float b1[WB_F];
  uint64_t bar_ops, bar_ab, bar_hd, bar_g;
  uint32_t tmem_slot;
};
static_assert(sizeof(WbSmem) <= 227 * 1024, "window backward: shared memory");

struct WbArgs {
  const float* X;       // [rows, 64]
  const float* U;       // [rows, 64]
  const float* w1;      // [32, 64] = [W_rel1 | W_root1]
  const float* b1;      // [32]
  const float* w2t;     // [32, 64] = [W_rel2^T | W_root2^T]
  float* part;          // [grid, WB_PART]
  float* dz1_out;       // optional [rows, 32]
  long long rows;
  int n_tiles, act1;
};

// byte offset of (feature f, row g) in a [feature][row] block with row-chunk distance LBO
template <int LBO>
__device__ __forceinline__ uint32_t fr_off(int f, int g) {
  return (uint32_t)((g >> 2) * LBO + (f >> 3) * 128 + (f & 7) * 16 + (g & 3) * 4);
}

__global__ void __launch_bounds__(WB_THREADS, 1) k_temporal_window_bwd(const WbArgs a) {
  extern __shared__ __align__(1024) unsigned char wb_raw[];
  WbSmem& sm = *reinterpret_cast<WbSmem*>(wb_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(&sm.bar_ops, WB_THREADS / 32);
    tc::mbar_init(&sm.bar_ab, 1);
    tc::mbar_init(&sm.bar_hd, WB_THREADS / 32);
    tc::mbar_init(&sm.bar_g, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(&sm.tmem_slot, 512);
  for (int i = tid; i < WB_F * 2 * WB_F; i += WB_THREADS) {
    const int n = i >> 6, k = i & 63;
    uint32_t hi, lo;
    tc::split_tf32(a.w1[i], hi, lo);
    sm.w1_hi[tc::kmajor_off(n, k, 64)] = __uint_as_float(hi);
    sm.w1_lo[tc::kmajor_off(n, k, 64)] = __uint_as_float(lo);
    tc::split_tf32(a.w2t[i], hi, lo);
    sm.w2_hi[tc::kmajor_off(n, k, 64)] = __uint_as_float(hi);
    sm.w2_lo[tc::kmajor_off(n, k, 64)] = __uint_as_float(lo);
  }
  if (tid < WB_F) sm.b1[tid] = a.b1[tid];
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = sm.tmem_slot;

  // ---------------- MMA issue (warp 0, lane 0) ----------------
  const uint32_t idesc_row = tc::idesc_tf32(128, 32);
  const uint32_t idesc_g = tc::idesc_tf32(128, 64);
  const uint32_t a_hi = tc::smem_u32(sm.a_hi), a_lo = tc::smem_u32(sm.a_lo);
  const uint32_t b_hi = tc::smem_u32(sm.b_hi), b_lo = tc::smem_u32(sm.b_lo);
  const uint32_t w1_hi = tc::smem_u32(sm.w1_hi), w1_lo = tc::smem_u32(sm.w1_lo);
  const uint32_t w2_hi = tc::smem_u32(sm.w2_hi), w2_lo = tc::smem_u32(sm.w2_lo);
  auto issue_rows = [&]() {        // z1 = X W1^T, dh = U W2t^T: A from TMEM
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {          // lo*Whi, hi*Wlo, hi*Whi
      const uint32_t xc = pass == 0 ? WB_COL_XLO : WB_COL_XHI;
      const uint32_t uc = pass == 0 ? WB_COL_ULO : WB_COL_UHI;
      const uint32_t w1s = pass == 1 ? w1_lo : w1_hi;
      const uint32_t w2s = pass == 1 ? w2_lo : w2_hi;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {              // K = 64 = 8 steps of 8
        tc::mma_tf32_ts(tbase + WB_COL_Z1, tbase + xc + ks * 8, tc::smem_desc_kmajor(w1s + ks * 256, 128, 2048),
                        idesc_row, pass > 0 || ks > 0);
        tc::mma_tf32_ts(tbase + WB_COL_DH, tbase + uc + ks * 8, tc::smem_desc_kmajor(w2s + ks * 256, 128, 2048),
                        idesc_row, pass > 0 || ks > 0);
      }
    }
    tc::mma_commit(&sm.bar_ab);
  };
  auto issue_grads = [&](bool first) {   // G (+)= [X | U]^T [dz1 | h]: K = the 128 rows of the tile
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t as = pass == 0 ? a_lo : a_hi;
      const uint32_t bs = pass == 1 ? b_lo : b_hi;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {             // 8 rows = 2 row chunks per instruction
        tc::mma_tf32_ss(tbase + WB_COL_G, tc::smem_desc_kmajor(as + ks * 2 * WB_A_LBO, WB_A_LBO, 128),
                        tc::smem_desc_kmajor(bs + ks * 2 * WB_B_LBO, WB_B_LBO, 128), idesc_g,
                        !first || pass > 0 || ks > 0);
      }
    }
    tc::mma_commit(&sm.bar_g);
  };

  // ---------------- thread = (row of the tile, 8-feature block) ----------------
  const int quarter = warp & 3, fb = warp >> 2;
  const int g = quarter * 32 + lane;
  const int fcol = fb * 8;
  const uint32_t lane_base = tbase + ((uint32_t)(quarter * 32) << 16);
  const int act1 = a.act1;
  float db1[8], db2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) db1[j] = db2[j] = 0.f;

  float4 v[8];    // this tile's [Xsum | Xself | Usum | Uself] slices (2 x float4 each), loaded one tile ahead
  auto load_tile = [&](int tile) {
    const long long r = (long long)tile * WB_TILE + g;
    if (r < a.rows) {
      const float4* x = reinterpret_cast<const float4*>(a.X + r * 64 + fcol);
      const float4* u = reinterpret_cast<const float4*>(a.U + r * 64 + fcol);
      v[0] = __ldcs(x);
      v[1] = __ldcs(x + 1);
      v[2] = __ldcs(x + 8);
      v[3] = __ldcs(x + 9);
      v[4] = __ldcs(u);
      v[5] = __ldcs(u + 1);
      v[6] = __ldcs(u + 8);
      v[7] = __ldcs(u + 9);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  int it = 0;
  int tile = blockIdx.x;
  if (tile < a.n_tiles) load_tile(tile);
  for (; tile < a.n_tiles; tile += gridDim.x, ++it) {
    // ---- split; TS-form A operands -> TMEM ----
    uint32_t hi[4][8], lo[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float f[8] = {v[2 * i].x, v[2 * i].y, v[2 * i].z, v[2 * i].w, v[2 * i + 1].x, v[2 * i + 1].y, v[2 * i + 1].z,
                          v[2 * i + 1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_tf32(f[j], hi[i][j], lo[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) db2[j] += __uint_as_float(hi[3][j]) + __uint_as_float(lo[3][j]);
    tc::tmem_st8(lane_base + WB_COL_XHI + fcol, hi[0]);
    tc::tmem_st8(lane_base + WB_COL_XHI + 32 + fcol, hi[1]);
    tc::tmem_st8(lane_base + WB_COL_XLO + fcol, lo[0]);
    tc::tmem_st8(lane_base + WB_COL_XLO + 32 + fcol, lo[1]);
    tc::tmem_st8(lane_base + WB_COL_UHI + fcol, hi[2]);
    tc::tmem_st8(lane_base + WB_COL_UHI + 32 + fcol, hi[3]);
    tc::tmem_st8(lane_base + WB_COL_ULO + fcol, lo[2]);
    tc::tmem_st8(lane_base + WB_COL_ULO + 32 + fcol, lo[3]);
    tc::wait_st();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&sm.bar_ops);
    if (warp == 0) {
      tc::mbar_wait(&sm.bar_ops, it & 1);
      tc::fence_after_sync();
      if (lane == 0) issue_rows();
      __syncwarp();
    }
    // ---- the same values as [feature][row] operands of the weight-gradient product (previous tile's must be done) ----
    tc::mbar_wait(&sm.bar_g, (it & 1) ^ 1);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t o = fr_off<WB_A_LBO>(i * 32 + fcol + j, g);
        *reinterpret_cast<uint32_t*>(sm.a_hi + o) = hi[i][j];
        *reinterpret_cast<uint32_t*>(sm.a_lo + o) = lo[i][j];
      }
    // ---- next tile's rows (in flight while the MMAs run) ----
    if (tile + (int)gridDim.x < a.n_tiles) load_tile(tile + gridDim.x);
    // ---- h, dz1 ----
    tc::mbar_wait(&sm.bar_ab, it & 1);
    tc::fence_after_sync();
    float hq[8], d1[8];
    {
      uint32_t z[8], dh[8];
      tc::tmem_ld8(lane_base + WB_COL_Z1 + fcol, z);
      tc::tmem_ld8(lane_base + WB_COL_DH + fcol, dh);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) hq[j] = __uint_as_float(z[j]) + sm.b1[fcol + j];
      gcm_act_fast_vec<8>(hq, act1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d1[j] = __uint_as_float(dh[j]) * gcm_act_grad(hq[j], act1);
        db1[j] += d1[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t h, l;
      tc::split_tf32(d1[j], h, l);
      uint32_t o = fr_off<WB_B_LBO>(fcol + j, g);
      *reinterpret_cast<uint32_t*>(sm.b_hi + o) = h;
      *reinterpret_cast<uint32_t*>(sm.b_lo + o) = l;
      tc::split_tf32(hq[j], h, l);
      o = fr_off<WB_B_LBO>(32 + fcol + j, g);
      *reinterpret_cast<uint32_t*>(sm.b_hi + o) = h;
      *reinterpret_cast<uint32_t*>(sm.b_lo + o) = l;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&sm.bar_hd);
    if (a.dz1_out) {
      const long long r = (long long)tile * WB_TILE + g;
      if (r < a.rows) {
        float4* o = reinterpret_cast<float4*>(a.dz1_out + r * WB_F + fcol);
        __stcs(o, make_float4(d1[0], d1[1], d1[2], d1[3]));
        __stcs(o + 1, make_float4(d1[4], d1[5], d1[6], d1[7]));
      }
    }
    if (warp == 0) {
      tc::mbar_wait(&sm.bar_hd, it & 1);
      tc::fence_after_sync();
      if (lane == 0) issue_grads(it == 0);
      __syncwarp();
    }
  }
  // ---- flush: G rows and the bias sums of this CTA ----
  tc::mbar_wait(&sm.bar_g, (it & 1) ^ 1);
  tc::fence_after_sync();
  float* part = a.part + (size_t)blockIdx.x * WB_PART;
  {
    uint32_t w[8];
    const int m = g;                                     // G row = feature index of [Xsum | Xself | Usum | Uself]
    tc::tmem_ld8(lane_base + WB_COL_G + (m < 64 ? 0 : 32) + fcol, w);
    tc::wait_ld();
    float4* o = reinterpret_cast<float4*>(part + m * WB_F + fcol);
    if (it > 0) {
      o[0] = make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3]));
      o[1] = make_float4(__uint_as_float(w[4]), __uint_as_float(w[5]), __uint_as_float(w[6]), __uint_as_float(w[7]));
    } else {
      o[0] = o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // bias sums: [which][feature][row] in the (now idle) A block, then one thread per (which, feature) adds 128 values
  float* red = reinterpret_cast<float*>(sm.a_hi);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(fcol + j) * WB_TILE + g] = db1[j];
    red[(WB_F + fcol + j) * WB_TILE + g] = db2[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 2 * WB_F) {
    float s = 0.f;
    for (int i = 0; i < WB_TILE; ++i) s += red[tid * WB_TILE + ((i + tid) & (WB_TILE - 1))];
    part[WB_AW * WB_F + tid] = s;
  }
  if (warp == 1) tc::tmem_dealloc(tbase, 512);
}

// out[i] += sum over CTAs (fixed order) of part[cta][i]; i < 128 * 32: G rows, then db1, db2
__global__ void __launch_bounds__(256) k_temporal_window_bwd_reduce(const float* __restrict__ part, int n_cta,
                                                                    float* __restrict__ g1, float* __restrict__ g2,
                                                                    float* __restrict__ db1, float* __restrict__ db2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= WB_PART) return;
  float s = 0.f;
  for (int c = 0; c < n_cta; ++c) s += part[(size_t)c * WB_PART + i];
  if (i < 64 * WB_F) g1[i] += s;
  else if (i < 128 * WB_F) g2[i - 64 * WB_F] += s;
  else if (i < 128 * WB_F + WB_F) db1[i - 128 * WB_F] += s;
  else db2[i - 128 * WB_F - WB_F] += s;
}

}  // namespace

extern "C" long long gcm_temporal_window_bwd_workspace(void) { return (long long)gcm_num_sms() * WB_PART; }

extern "C" int gcm_temporal_window_bwd(const float* X, const float* U, long long rows, const float* w1cat, const float* b1,
                                       const float* w2tcat, int act1, float* workspace, float* g1, float* g2, float* db1,
                                       float* db2, float* dz1_out, void* stream) {
  GCM_REQUIRE(X && U && w1cat && b1 && w2tcat && workspace && g1 && g2 && db1 && db2, "temporal_window_bwd: null pointer");
  GCM_REQUIRE(rows >= 0 && rows < (1ll << 37), "temporal_window_bwd: rows out of range");
  GCM_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(dz1_out) |
                reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
              "temporal_window_bwd: pointers must be 16-byte aligned");
  GCM_REQUIRE(act1 == GCM_ACT_NONE || act1 == GCM_ACT_TANH || act1 == GCM_ACT_RELU, "temporal_window_bwd: bad activation");
  if (rows == 0) return GCM_OK;
  WbArgs a;
  a.X = X;
  a.U = U;
  a.w1 = w1cat;
  a.b1 = b1;
  a.w2t = w2tcat;
  a.part = workspace;
  a.dz1_out = dz1_out;
  a.rows = rows;
  a.act1 = act1;
  a.n_tiles = (int)((rows + WB_TILE - 1) / WB_TILE);
  const int sms = gcm_num_sms();
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_temporal_window_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(WbSmem));
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(k_temporal_window_bwd): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    attr_done = true;
  }
  k_temporal_window_bwd<<<grid, WB_THREADS, sizeof(WbSmem), (cudaStream_t)stream>>>(a);
  int rc = gcm_check_launch("k_temporal_window_bwd");
  if (rc != GCM_OK) return rc;
  k_temporal_window_bwd_reduce<<<(WB_PART + 255) / 256, 256, 0, (cudaStream_t)stream>>>(workspace, grid, g1, g2, db1, db2);
  return gcm_check_launch("k_temporal_window_bwd_reduce");
}

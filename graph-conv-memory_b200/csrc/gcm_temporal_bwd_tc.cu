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
// but per-CTA partials.  A CTA takes tiles of 128 rows (TMEM lane = row, 4 threads per row):
//   * X and U come in TILED layout ([tile][16-byte chunk][row of the tile][4 floats], written that way by the two
//     operand kernels): a warp's load of one chunk for its 32 rows is 512 contiguous bytes.  With row-major operands every
//     lane touches its own 128-byte line and the loads alone took 4300 of a tile's 13900 cycles (tools/wb_trace.py);
//   * the row products are TS-form MMAs: the row-owning threads write X and U (hi / lo split) into TMEM with tcgen05.st;
//   * the weight-gradient products contract over the 128 ROWS of the tile, so their operands are [feature][row] matrices:
//     the same threads scatter X, U, dz1, h into shared memory in the canonical K-major core-matrix layout with the row
//     index as K (one float per store; a warp's 32 rows of one feature land in 32 different banks because the distance
//     between K-adjacent core matrices is padded by 16 bytes);
//   * 3xTF32 with STACKED operands: hi and lo parts sit next to each other along M (weight gradients: [X_hi ; X_lo]) or N
//     ([dz1_hi | dz1_lo], [W_hi ; W_lo]), so one pass over K yields the hi*hi, lo*hi and hi*lo blocks side by side in the
//     accumulator and every operand byte is read once instead of three times (64 MMAs per tile instead of 144); the
//     blocks are added in the epilogue (row products) or by the reduce kernel (weight gradients);
//   * the weight-gradient accumulators stay in TMEM for the whole launch and are flushed once per CTA; a second kernel
//     adds the per-CTA partials in a fixed order (deterministic).
#include "gcm_common.cuh"
#include "gcm_tc.cuh"

namespace {

constexpr int WB_TILE = 128;          // rows per tile = TMEM lanes
constexpr int WB_THREADS = 512;       // 4 threads per row; lane 0 of warps 0 and 4 also issue the MMAs
constexpr int WB_F = 32;              // F = H1 = H2 = 32 (the cached-row kernel's shape)
// TMEM columns: row-product accumulators [hi*hi + lo*hi | hi*lo] (64 each), weight-gradient accumulators (64 each), then
// the TS-form A operands
constexpr uint32_t WB_COL_Z1 = 0, WB_COL_DH = 64, WB_COL_G1 = 128, WB_COL_G2 = 192, WB_COL_XHI = 256, WB_COL_XLO = 320,
                   WB_COL_UHI = 384, WB_COL_ULO = 448;
constexpr int WB_GBLK = 128 * 64;                    // floats of one raw weight-gradient accumulator
constexpr int WB_PART = 2 * WB_GBLK + 2 * WB_F;      // floats per CTA partial: raw G1, raw G2, db1, db2
// [feature][row] operand blocks, K = row: core matrix (8 features x 4 rows) = 128 B; feature groups adjacent (SBO = 128),
// row chunks LBO apart, LBO = (#feature groups) * 128 + 16: the pad makes LBO / 4 = 4 (mod 32), so the 32 rows a warp
// stores for one feature hit 32 different banks
constexpr int WB_A_LBO = 16 * 128 + 16;              // [X_hi (64 features) ; X_lo (64)]: 16 groups
constexpr int WB_B_LBO = 8 * 128 + 16;               // [dz1_hi (32) | dz1_lo (32)]: 8 groups
constexpr int WB_A_BYTES = (WB_TILE / 4) * WB_A_LBO;   // 66 048
constexpr int WB_B_BYTES = (WB_TILE / 4) * WB_B_LBO;   // 33 280

struct WbSmem {
  unsigned char ax[WB_A_BYTES];  // [X_hi ; X_lo]^T  (features x rows)
  unsigned char au[WB_A_BYTES];  // [U_hi ; U_lo]
  unsigned char bd[WB_B_BYTES];  // [dz1_hi | dz1_lo]
  unsigned char bh[WB_B_BYTES];  // [h_hi | h_lo]
  float w1[2 * WB_F * 2 * WB_F];  // [W1_hi ; W1_lo]: [64][2F] canonical K-major, 16 KB
  float w2[2 * WB_F * 2 * WB_F];  // [W2t_hi ; W2t_lo]
  float b1[WB_F];
  uint64_t bar_ops, bar_ab, bar_hd, bar_g;
  uint32_t tmem_slot;
};
static_assert(sizeof(WbSmem) <= 227 * 1024, "window backward: shared memory");

struct WbArgs {
  const float* X;       // tiled [rows / 128][16][128][4]
  const float* U;
  const float* w1;      // [32, 64] = [W_rel1 | W_root1]
  const float* b1;      // [32]
  const float* w2t;     // [32, 64] = [W_rel2^T | W_root2^T]
  float* part;          // [grid, WB_PART]
  float* dz1_out;       // optional [rows, 32] (row-major)
  long long rows;
  int n_tiles, act1;
  long long* trace;     // debugging: per-tile phase clocks of warp `trace_warp` of CTA 0 ([tile][10])
  int trace_warp;
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
    tc::mbar_init(&sm.bar_ab, 2);
    tc::mbar_init(&sm.bar_hd, WB_THREADS / 32);
    tc::mbar_init(&sm.bar_g, 2);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(&sm.tmem_slot, 512);
  for (int i = tid; i < WB_F * 2 * WB_F; i += WB_THREADS) {
    const int n = i >> 6, k = i & 63;
    uint32_t hi, lo;
    tc::split_tf32(a.w1[i], hi, lo);
    sm.w1[tc::kmajor_off(n, k, 64)] = __uint_as_float(hi);
    sm.w1[tc::kmajor_off(32 + n, k, 64)] = __uint_as_float(lo);
    tc::split_tf32(a.w2t[i], hi, lo);
    sm.w2[tc::kmajor_off(n, k, 64)] = __uint_as_float(hi);
    sm.w2[tc::kmajor_off(32 + n, k, 64)] = __uint_as_float(lo);
  }
  if (tid < WB_F) sm.b1[tid] = a.b1[tid];
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = sm.tmem_slot;

  // ---------------- MMA issue: lane 0 of warp 0 (layer-1 side: z1, G1) and of warp 4 (layer-2 side: dh, G2) ----------------
  const int side = warp >> 2;                          // 0: X / W1 / dz1, 1: U / W2t / h   (warps 0 and 4 issue)
  auto issue_rows = [&]() {
    // [z1 | .] = X_hi [W_hi ; W_lo]^T (N = 64: hi*hi and hi*lo side by side), then X_lo W_hi^T on top of the first half
    const uint64_t wd0 = tc::smem_desc_kmajor(tc::smem_u32(side ? sm.w2 : sm.w1), 128, 2048);
    const uint32_t a_hi = tbase + (side ? WB_COL_UHI : WB_COL_XHI);
    const uint32_t d = tbase + (side ? WB_COL_DH : WB_COL_Z1);
    uint64_t wd = wd0;
    uint32_t ac = a_hi;
#pragma unroll 2
    for (int ks = 0; ks < 8; ++ks, ac += 8, wd += 16)   // K = 64 = 8 steps of 8 (2 chunks of 128 B in the weight pack)
      tc::mma_tf32_ts(d, ac, wd, tc::idesc_tf32(128, 64), ks > 0);
    wd = wd0;
    ac = a_hi + 64;
#pragma unroll 2
    for (int ks = 0; ks < 8; ++ks, ac += 8, wd += 16)
      tc::mma_tf32_ts(d, ac, wd, tc::idesc_tf32(128, 32), true);
    tc::mma_commit(&sm.bar_ab);
  };
  auto issue_grads = [&](bool first) {
    // raw G1 (+)= [X_hi ; X_lo]^T [dz1_hi | dz1_lo]  (side 0),  raw G2 (+)= [U_hi ; U_lo]^T [h_hi | h_lo]  (side 1):
    // M = 128, N = 64, K = the tile's 128 rows, 8 (= 2 row chunks) per instruction
    uint64_t ad = tc::smem_desc_kmajor(tc::smem_u32(side ? sm.au : sm.ax), WB_A_LBO, 128);
    uint64_t bd = tc::smem_desc_kmajor(tc::smem_u32(side ? sm.bh : sm.bd), WB_B_LBO, 128);
    const uint32_t d = tbase + (side ? WB_COL_G2 : WB_COL_G1);
#pragma unroll 2
    for (int ks = 0; ks < 16; ++ks, ad += 2 * WB_A_LBO / 16, bd += 2 * WB_B_LBO / 16)
      tc::mma_tf32_ss(d, ad, bd, tc::idesc_tf32(128, 64), !first || ks > 0);
    tc::mma_commit(&sm.bar_g);
  };
  const bool issuer = (warp & 3) == 0 && warp < 8;     // warps 0 and 4

  // ---------------- thread = (row of the tile, quarter of the operand columns / 8 of the 32 output features) ----------------
  const int quarter = warp & 3, fb = warp >> 2;
  const int g = quarter * 32 + lane;
  const int fcol = fb * 8;
  const uint32_t lane_base = tbase + ((uint32_t)(quarter * 32) << 16);
  const int act1 = a.act1;
  float db1[8], db2[16];    // db2: columns 16 fb .. of U, i.e. dz2 features 16 (fb - 2) .. for fb >= 2
#pragma unroll
  for (int j = 0; j < 8; ++j) db1[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) db2[j] = 0.f;

  float4 v[8];    // chunks 4 fb .. 4 fb + 3 of this row's X (v[0..3]) and U (v[4..7]), loaded one tile ahead
  auto load_tile = [&](int tile) {
    const float4* x = reinterpret_cast<const float4*>(a.X) + (size_t)tile * (16 * WB_TILE) + (4 * fb) * WB_TILE + g;
    const float4* u = reinterpret_cast<const float4*>(a.U) + (size_t)tile * (16 * WB_TILE) + (4 * fb) * WB_TILE + g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = __ldcs(x + i * WB_TILE);
      v[4 + i] = __ldcs(u + i * WB_TILE);
    }
  };

  int it = 0;
  int tile = blockIdx.x;
#define WB_T(k)                                                                                   \
  if (a.trace && blockIdx.x == 0 && warp == a.trace_warp && lane == 0 && it < 64) a.trace[it * 10 + (k)] = clock64();
  if (tile < a.n_tiles) load_tile(tile);
  for (; tile < a.n_tiles; tile += gridDim.x, ++it) {
    WB_T(0)
    // ---- split; TS-form A operands -> TMEM (the split is redone for the shared-memory copy below: keeping hi / lo of
    //      all 32 values alive across the barrier waits spilled the prefetched rows) ----
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = v[4 * t + i];
        tc::split_tf32(q.x, hi[4 * i], lo[4 * i]);
        tc::split_tf32(q.y, hi[4 * i + 1], lo[4 * i + 1]);
        tc::split_tf32(q.z, hi[4 * i + 2], lo[4 * i + 2]);
        tc::split_tf32(q.w, hi[4 * i + 3], lo[4 * i + 3]);
        if (t == 1) {
          db2[4 * i] += q.x;
          db2[4 * i + 1] += q.y;
          db2[4 * i + 2] += q.z;
          db2[4 * i + 3] += q.w;
        }
      }
      const uint32_t c_hi = (t == 0 ? WB_COL_XHI : WB_COL_UHI) + 16 * fb;
      tc::tmem_st16(lane_base + c_hi, hi);
      tc::tmem_st16(lane_base + c_hi + 64, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&sm.bar_ops);
    WB_T(1)
    if (issuer) {
      tc::mbar_wait(&sm.bar_ops, it & 1);
      tc::fence_after_sync();
      if (lane == 0) issue_rows();
      __syncwarp();
    }
    // ---- the same values as [feature][row] operands of the weight-gradient product (previous tile's must be done) ----
    WB_T(2)
    tc::mbar_wait(&sm.bar_g, (it & 1) ^ 1);
    WB_T(3)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      // feature 16 fb + 4 i + j: group 2 fb + i / 2, row-in-group 4 (i & 1) + j; the lo copy sits 8 groups further
      unsigned char* p = (t == 0 ? sm.ax : sm.au) + fr_off<WB_A_LBO>(16 * fb, g);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = v[4 * t + i];
        const float f[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hi, lo;
          tc::split_tf32(f[j], hi, lo);
          *reinterpret_cast<uint32_t*>(p + (i >> 1) * 128 + ((i & 1) * 4 + j) * 16) = hi;
          *reinterpret_cast<uint32_t*>(p + 1024 + (i >> 1) * 128 + ((i & 1) * 4 + j) * 16) = lo;
        }
      }
    }
    WB_T(4)
    // ---- next tile's rows (in flight while the MMAs run) ----
    if (tile + (int)gridDim.x < a.n_tiles) load_tile(tile + gridDim.x);
    // ---- h, dz1 ----
    WB_T(5)
    tc::mbar_wait(&sm.bar_ab, it & 1);
    tc::fence_after_sync();
    WB_T(6)
    float hq[8], d1[8];
    {
      uint32_t z0[8], z1[8], e0[8], e1[8];
      tc::tmem_ld8(lane_base + WB_COL_Z1 + fcol, z0);
      tc::tmem_ld8(lane_base + WB_COL_Z1 + 32 + fcol, z1);
      tc::tmem_ld8(lane_base + WB_COL_DH + fcol, e0);
      tc::tmem_ld8(lane_base + WB_COL_DH + 32 + fcol, e1);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) hq[j] = (__uint_as_float(z0[j]) + __uint_as_float(z1[j])) + sm.b1[fcol + j];
      gcm_act_fast_vec<8>(hq, act1);
#pragma unroll
      for (int j = 0; j < 8; ++j) d1[j] = __uint_as_float(e0[j]) + __uint_as_float(e1[j]);
      if (act1 == GCM_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d1[j] *= 1.0f - hq[j] * hq[j];
      } else if (act1 == GCM_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d1[j] = hq[j] > 0.0f ? d1[j] : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) db1[j] += d1[j];
    }
    {
      // feature fcol + j: group fb, row-in-group j; lo copy 4 groups further
      unsigned char* pd = sm.bd + fr_off<WB_B_LBO>(fcol, g);
      unsigned char* ph = sm.bh + fr_off<WB_B_LBO>(fcol, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t h, l;
        tc::split_tf32(d1[j], h, l);
        *reinterpret_cast<uint32_t*>(pd + j * 16) = h;
        *reinterpret_cast<uint32_t*>(pd + 512 + j * 16) = l;
        tc::split_tf32(hq[j], h, l);
        *reinterpret_cast<uint32_t*>(ph + j * 16) = h;
        *reinterpret_cast<uint32_t*>(ph + 512 + j * 16) = l;
      }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&sm.bar_hd);
    WB_T(7)
    if (a.dz1_out) {
      const long long r = (long long)tile * WB_TILE + g;
      if (r < a.rows) {
        float4* o = reinterpret_cast<float4*>(a.dz1_out + r * WB_F + fcol);
        __stcs(o, make_float4(d1[0], d1[1], d1[2], d1[3]));
        __stcs(o + 1, make_float4(d1[4], d1[5], d1[6], d1[7]));
      }
    }
    if (issuer) {
      tc::mbar_wait(&sm.bar_hd, it & 1);
      tc::fence_after_sync();
      if (lane == 0) issue_grads(it == 0);
      __syncwarp();
    }
    WB_T(8)
  }
  // ---- flush: the raw accumulators (row m of G1 / G2 = feature m of [hi ; lo]) and the bias sums of this CTA ----
  tc::mbar_wait(&sm.bar_g, (it & 1) ^ 1);
  tc::fence_after_sync();
  float* part = a.part + (size_t)blockIdx.x * WB_PART;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t w[8];
      tc::tmem_ld8(lane_base + (which ? WB_COL_G2 : WB_COL_G1) + half * 32 + fcol, w);
      tc::wait_ld();
      float4* o = reinterpret_cast<float4*>(part + which * WB_GBLK + g * 64 + half * 32 + fcol);
      if (it > 0) {
        o[0] = make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3]));
        o[1] = make_float4(__uint_as_float(w[4]), __uint_as_float(w[5]), __uint_as_float(w[6]), __uint_as_float(w[7]));
      } else {
        o[0] = o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  // bias sums: [feature][row] in the (now idle) A block, then one thread per feature adds 128 values
  float* red = reinterpret_cast<float*>(sm.ax);
#pragma unroll
  for (int j = 0; j < 8; ++j) red[(fcol + j) * WB_TILE + g] = db1[j];
  if (fb >= 2) {
#pragma unroll
    for (int j = 0; j < 16; ++j) red[(WB_F + 16 * (fb - 2) + j) * WB_TILE + g] = db2[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 2 * WB_F) {
    float s = 0.f;
    for (int i = 0; i < WB_TILE; ++i) s += red[tid * WB_TILE + ((i + tid) & (WB_TILE - 1))];
    part[2 * WB_GBLK + tid] = s;
  }
  if (warp == 1) tc::tmem_dealloc(tbase, 512);
}

// g1 / g2 [64, 32] += sum over CTAs (fixed order) of the three useful blocks of the raw accumulators
//   raw[m][n]: m < 64 hi feature m, m >= 64 lo feature m - 64; n < 32 hi column n, n >= 32 lo column n - 32
__global__ void __launch_bounds__(256) k_temporal_window_bwd_reduce(const float* __restrict__ part, int n_cta,
                                                                    float* __restrict__ g1, float* __restrict__ g2,
                                                                    float* __restrict__ db1, float* __restrict__ db2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * 64 * WB_F + 2 * WB_F) return;
  float s = 0.f;
  if (i < 2 * 64 * WB_F) {
    const int which = i / (64 * WB_F), f = (i / WB_F) % 64, k = i % WB_F;
    const float* p = part + which * WB_GBLK;
    for (int c = 0; c < n_cta; ++c, p += WB_PART) s += (p[f * 64 + k] + p[(64 + f) * 64 + k]) + p[f * 64 + 32 + k];
    (which ? g2 : g1)[f * WB_F + k] += s;
  } else {
    const int j = i - 2 * 64 * WB_F;
    const float* p = part + 2 * WB_GBLK + j;
    for (int c = 0; c < n_cta; ++c, p += WB_PART) s += *p;
    (j < WB_F ? db1 : db2)[j & (WB_F - 1)] += s;
  }
}

}  // namespace

static long long* g_wb_trace = nullptr;
static int g_wb_trace_warp = 0;
/* debugging hook (tools/wb_trace.py): device buffer [64][10] of int64 that the next launches fill with phase clocks */
extern "C" int gcm_temporal_window_bwd_set_trace(long long* buf, int warp) {
  g_wb_trace = buf;
  g_wb_trace_warp = warp;
  return GCM_OK;
}

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
  a.trace = g_wb_trace;
  a.trace_warp = g_wb_trace_warp;
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
  k_temporal_window_bwd_reduce<<<(2 * 64 * WB_F + 2 * WB_F + 255) / 256, 256, 0, (cudaStream_t)stream>>>(workspace, grid, g1,
                                                                                                          g2, db1, db2);
  return gcm_check_launch("k_temporal_window_bwd_reduce");
}

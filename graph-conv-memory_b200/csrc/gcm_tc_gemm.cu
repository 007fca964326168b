// bf16 tensor-core (tcgen05 + TMEM) products of the DenseEdge-only path ("ones", csrc/gcm_dense_ones.cu) for the
// configurations that ask for bfloat16 compute (BASELINE cfg3; DenseGCM.compute_dtype = torch.bfloat16).
//
// Only the products whose rounding errors are independent per node / per sample run here:
//   gcm_linear_tc        out[r, :] = epi(X[r, :K] W^T + bias)       the per-node cache fill  R_i = W_root1 x_i
//                                                                    (lin_root of DenseGraphConv, README.md:56-57)
//   gcm_outer_reduce_tc  dW[o, i] += sum_r A[r, o] X[r, i]           the weight gradients of lin_rel / lin_root
// c = W_rel1 S + b1 and the belief stay in fp32 (gcm_linear2): their error is common to every node of a graph.
//
// Building block (validated by gcm_tc_selftest): the A operand is written to TMEM by the threads that own its
// rows (tcgen05.st, TMEM lane = row, two bf16 per column), the B operand sits in shared memory in the canonical
// K-major no-swizzle layout, one elected thread issues tcgen05.mma.kind::f16 with fp32 accumulation in TMEM, the
// accumulator is read back with tcgen05.ld.  Two groups of four warps alternate tiles / row chunks so that the
// loads of one overlap the MMA + epilogue of the other; a ninth warp only issues MMAs.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "gcm_common.cuh"
#include "gcm_tc.cuh"

namespace {

constexpr int TCG_THREADS = 288;   // 2 groups x 128 + MMA warp

__device__ __forceinline__ float tcg_exp2x(float z) {   // exp(2 clamp(z, +-40)); NaN propagates
  const float zc = fminf(fmaxf(z, -40.0f), 40.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zc * 2.8853900817779268f));
  return z == z ? e : z;
}

// W [Ho, K] row-major float32 (global) -> shared memory, canonical K-major no-swizzle layouts.  Four 16-byte loads
// are in flight per thread (a one-load-per-iteration loop serialises ~100 L2 round trips per thread, which was the
// whole duration of the small launches); 4 (tf32) / 8 (bf16) consecutive k are contiguous in the canonical layout.
template <int NTHREADS>
__device__ __forceinline__ void stage_w_tf32(const float* __restrict__ W, int Ho, int K, float* Whi, float* Wlo, int tid) {
  // Item i = one 16-byte piece (row n, columns 4 kq .. 4 kq + 3).  Consecutive items walk the 8 rows of a core-matrix row
  // first, then 4 neighbouring pieces, so a warp's 32 stores cover 512 CONTIGUOUS bytes of the K-major layout (with items
  // in row-major order -- consecutive k of one row -- the stores were 128 bytes apart: 32-way bank conflicts, 8 K of the
  // 8.3 K shared-memory wavefronts of a one-tile launch, and a launch over few rows is dominated by this loop: 20.5 us at
  // [128 x 128] whatever the row count).  8 independent loads in flight per thread.  Needs Ho % 8 == 0, K % 16 == 0.
  constexpr int U = 8;
  const int n4 = Ho * K / 4;
  const int kqb_n = K >> 4;                 // blocks of 4 pieces along k
  for (int i0 = tid; i0 < n4; i0 += U * NTHREADS) {
    float4 v[U];
    int off[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int i = i0 + q * NTHREADS;
      const int blk = i >> 5, l = i & 31;
      const int nb = blk / kqb_n, kqb = blk - nb * kqb_n;
      const int n = nb * 8 + (l & 7), kq = kqb * 4 + (l >> 3);
      off[q] = tc::kmajor_off(n, kq * 4, K);
      v[q] = i < n4 ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K) + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      if (i0 + q * NTHREADS < n4) {
        uint32_t h[4], l[4];
        tc::split_tf32(v[q].x, h[0], l[0]); tc::split_tf32(v[q].y, h[1], l[1]);
        tc::split_tf32(v[q].z, h[2], l[2]); tc::split_tf32(v[q].w, h[3], l[3]);
        *reinterpret_cast<uint4*>(Whi + off[q]) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(Wlo + off[q]) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}
template <int NTHREADS>
__device__ __forceinline__ void stage_w_bf16(const float* __restrict__ W, int Ho, int K, __nv_bfloat16* Ws, int tid) {
  // item = one 16-byte core-matrix row (row n, columns 8 k8 .. 8 k8 + 7); consecutive items walk the 8 rows of a core
  // matrix, then 4 neighbouring matrices: a warp's stores cover 512 contiguous bytes (see stage_w_tf32)
  constexpr int U = 4;
  const int n8 = Ho * K / 8;
  const int kb_n = K >> 5;                  // blocks of 4 core matrices along k (K % 32 == 0), else the plain order below
  if ((K & 31) == 0) {
    for (int i0 = tid; i0 < n8; i0 += U * NTHREADS) {
      float4 v[2 * U];
      int off[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = i0 + q * NTHREADS;
        const int blk = i >> 5, l = i & 31;
        const int nb = blk / kb_n, kb = blk - nb * kb_n;
        const int n = nb * 8 + (l & 7), k8 = kb * 4 + (l >> 3);
        off[q] = tc::kmajor_off_bf16(n, k8 * 8, K);
        const float4* src = reinterpret_cast<const float4*>(W + (size_t)n * K) + 2 * k8;
        v[2 * q] = i < n8 ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[2 * q + 1] = i < n8 ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        if (i0 + q * NTHREADS < n8)
          *reinterpret_cast<uint4*>(Ws + off[q]) =
              make_uint4(tc::pack_bf16(v[2 * q].x, v[2 * q].y), tc::pack_bf16(v[2 * q].z, v[2 * q].w),
                         tc::pack_bf16(v[2 * q + 1].x, v[2 * q + 1].y), tc::pack_bf16(v[2 * q + 1].z, v[2 * q + 1].w));
      }
    }
    return;
  }
  for (int i0 = tid; i0 < n8; i0 += 2 * NTHREADS) {
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = i0 + q * NTHREADS;
      v[2 * q] = i < n8 ? __ldg(reinterpret_cast<const float4*>(W) + 2 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[2 * q + 1] = i < n8 ? __ldg(reinterpret_cast<const float4*>(W) + 2 * i + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = i0 + q * NTHREADS;
      if (i < n8) {
        const int e = i * 8, n = e / K, k = e - n * K;
        *reinterpret_cast<uint4*>(Ws + tc::kmajor_off_bf16(n, k, K)) =
            make_uint4(tc::pack_bf16(v[2 * q].x, v[2 * q].y), tc::pack_bf16(v[2 * q].z, v[2 * q].w),
                       tc::pack_bf16(v[2 * q + 1].x, v[2 * q + 1].y), tc::pack_bf16(v[2 * q + 1].z, v[2 * q + 1].w));
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// out[r, :Ho] = epi(X[r, :K] W^T + bias), W row-major [Ho, K] float32 (rounded to bf16 once per CTA).
// Persistent; a tile = 128 rows, TMEM lane = row.  K, Ho multiples of 16, <= 128.
// ------------------------------------------------------------------------------------------------
struct LinearTcArgs {
  const float* X; long long ldx; int K;
  const float* W; const float* bias;
  int act;              // GCM_ACT_NONE / GCM_ACT_EXP2X
  long long rows;
  int Ho;
  void* out; long long ldo;   // elements
  int out_bf16;
  long long tiles;
};

__global__ void __launch_bounds__(TCG_THREADS) k_linear_tc(const LinearTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);                 // [Ho x K] canonical K-major
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)a.Ho * a.K * 2);
  uint64_t* full = bars;         // [2] A tile of group g is in TMEM
  uint64_t* done = bars + 2;     // [2] MMAs of group g's tile have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* bias_s = reinterpret_cast<float*>(bars + 8);      // [128] bias (0 when absent): the epilogue reads it per element
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, Ho = a.Ho;

  if (tid == 0) {
    tc::mbar_init(&full[0], 128); tc::mbar_init(&full[1], 128);
    tc::mbar_init(&done[0], 1);   tc::mbar_init(&done[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, 512);
  if (tid < 128) bias_s[tid] = (a.bias && tid < Ho) ? __ldg(a.bias + tid) : 0.0f;
  stage_w_bf16<TCG_THREADS>(a.W, Ho, K, Ws, tid);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const long long n_it = (a.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles of this CTA

  if (warp < 8) {
    const int g = warp >> 2;
    const int row_in_tile = tid & 127;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t col_a = g * 64, col_d = 128 + g * 128;
    for (long long it = g; it < n_it; it += 2) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const long long r = tile * 128 + row_in_tile;
      const bool ok = r < a.rows;
      const float4* xr = reinterpret_cast<const float4*>(a.X + (ok ? r : 0) * a.ldx);
      // the previous tile of this group was fully read back before this point (wait_ld below), and its MMAs
      // completed before that (done wait), so both the A columns and the D columns are free
      for (int k0 = 0; k0 < K; k0 += 32) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[j] = (ok && k0 + j * 4 < K) ? __ldcs(xr + (k0 >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pk[2 * j] = tc::pack_bf16(v[j].x, v[j].y);
          pk[2 * j + 1] = tc::pack_bf16(v[j].z, v[j].w);
        }
        tc::tmem_st8(lane_addr + col_a + (k0 >> 1), pk);
        if (k0 + 16 < K) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pk[2 * j] = tc::pack_bf16(v[4 + j].x, v[4 + j].y);
            pk[2 * j + 1] = tc::pack_bf16(v[4 + j].z, v[4 + j].w);
          }
          tc::tmem_st8(lane_addr + col_a + (k0 >> 1) + 8, pk);
        }
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&full[g]);
      tc::mbar_wait(&done[g], (uint32_t)((it >> 1) & 1));
      tc::fence_after_sync();
      for (int n0 = 0; n0 < Ho; n0 += 16) {
        uint32_t d[16];
        tc::tmem_ld16(lane_addr + col_d + n0, d);
        tc::wait_ld();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float z = __uint_as_float(d[j]) + bias_s[n0 + j];
          f[j] = a.act == GCM_ACT_EXP2X ? tcg_exp2x(z) : z;
        }
        if (ok) {
          if (a.out_bf16) {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + r * a.ldo + n0);
            o[0] = make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]),
                              tc::pack_bf16(f[6], f[7]));
            o[1] = make_uint4(tc::pack_bf16(f[8], f[9]), tc::pack_bf16(f[10], f[11]), tc::pack_bf16(f[12], f[13]),
                              tc::pack_bf16(f[14], f[15]));
          } else {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + r * a.ldo + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      }
      tc::fence_before_sync();
    }
  } else if (lane == 0) {
    const uint32_t idesc = tc::idesc_bf16(128, Ho);
    const uint32_t sbo = (uint32_t)(K / 8) * 128u;
    for (long long it = 0; it < n_it; ++it) {
      const int g = (int)(it & 1);
      tc::mbar_wait(&full[g], (uint32_t)((it >> 1) & 1));
      tc::fence_after_sync();
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(Ws) + ks * 256, 128, sbo);
        tc::mma_bf16_ts(tbase + 128 + g * 128, tbase + g * 64 + ks * 8, bdesc, idesc, ks > 0);
      }
      tc::mma_commit(&done[g]);
    }
  }
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------------------------------------
// out[r, :Ho] = act(X1[r, :K1] W1^T  [3xTF32: fp32-accurate]  +  X2[r, :K2] W2^T  [bf16, optional]  + bias)
// for the two products of a step whose error is common to every node of a graph (c = W_rel1 S + b1, and the G term of
// the belief): every fp32 value is split hi + lo (hi exactly representable in tf32) and the product is issued as
// lo*Bhi + hi*Blo + hi*Bhi (gcm_tc.cuh); the h_t W_root2^T term of the belief may ride along in bf16.
// One 128-row tile per CTA (B graphs = B / 128 CTAs), TMEM: A1 hi | A1 lo | A2 | D.
// ------------------------------------------------------------------------------------------------
struct LinearTc32Args {
  const float* X1; long long ldx1; int K1; const float* W1;
  const float* X2; long long ldx2; int K2; const float* W2;   // optional bf16 part
  const float* bias;
  int act;              // GCM_ACT_* or GCM_ACT_EXP2X
  long long rows;
  int Ho;
  float* out; long long ldo;
  int32_t* status;
  long long tiles;
  uint32_t tmem_cols;   // power of two >= 2 K1 + K2 / 2 + Ho
  int stage;            // 1: the 128 x K1 row tile and the 128 x Ho result tile go through shared memory (K1 <= 64, no X2)
};

constexpr int TC32_THREADS = 288;   // warps 0-3: A rows -> TMEM, epilogue; warps 4-8: weight staging; warp 8 lane 0: MMA issue

__global__ void __launch_bounds__(TC32_THREADS) k_linear_tc32(const LinearTc32Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int K1 = a.K1, K2 = a.X2 ? a.K2 : 0, Ho = a.Ho;
  float* Whi = reinterpret_cast<float*>(smem_raw);                         // [Ho x K1] canonical K-major (tf32)
  float* Wlo = Whi + (size_t)Ho * K1;
  __nv_bfloat16* W2s = reinterpret_cast<__nv_bfloat16*>(Wlo + (size_t)Ho * K1);   // [Ho x K2] canonical (bf16)
  uint64_t* bars = reinterpret_cast<uint64_t*>(W2s + (size_t)Ho * K2);
  uint64_t* full = bars;         // A operands of the current tile are in TMEM (128 arrivals)
  uint64_t* wready = bars + 1;   // weights are in shared memory (160 arrivals, once)
  uint64_t* done = bars + 2;     // all MMAs of the current tile completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 64);   // [128] bias (0 when absent)
  // staging tile (a.stage): rows padded by 4 floats so that a thread reading ITS row with 16-byte loads is conflict-free
  float* tile = bias_s + 128;
  const int tile_ld = (K1 > Ho ? K1 : Ho) + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(full, 128);
    tc::mbar_init(wready, TC32_THREADS - 128);
    tc::mbar_init(done, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc(tmem_slot, a.tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t col_hi = 0, col_lo = K1, col_a2 = 2 * K1, col_d = 2 * K1 + K2 / 2;
  const long long n_it = (a.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // persistent: tiles of this CTA
  if (warp < 4) {
    // the tile's rows go to TMEM while the other warps stage the weights (first tile) / idle (later tiles)
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    bool bad = false;
    for (long long it = 0; it < n_it; ++it) {
      const long long r = (blockIdx.x + it * gridDim.x) * 128 + tid;
      const bool ok = r < a.rows;
      const float4* x1 = reinterpret_cast<const float4*>(a.X1 + (ok ? r : 0) * a.ldx1);
      if (a.stage) {
        // Coalesced staging: consecutive threads copy consecutive 16-byte chunks of the tile's rows (a warp instruction
        // covers whole rows) with cp.async, then every thread reads ITS row from shared memory.  The row-per-thread
        // global loads of the other branch touch 32 different lines per warp instruction and were the limiter of the
        // narrow products (K1 = 64 -> Ho = 32: 1.7 TB/s, profiles/c2_bptt_kernels_r2.md).
        const long long r0 = (blockIdx.x + it * gridDim.x) * 128;
        const int cpr = K1 >> 2, csh = __ffs(cpr) - 1;   // 16-byte chunks per row (a power of two on this path)
        for (int c = tid; c < 128 * cpr; c += 128) {
          const int row = c >> csh, c4 = c & (cpr - 1);
          float* dst = tile + row * tile_ld + c4 * 4;
          if (r0 + row < a.rows) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)),
                         "l"(a.X1 + (r0 + row) * a.ldx1 + c4 * 4) : "memory");
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      for (int k0 = 0; k0 < K1; k0 += 64) {        // 16 independent 16-byte loads in flight per thread
        float4 v[16];
        if (a.stage) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = (k0 + 4 * j < K1) ? *reinterpret_cast<const float4*>(tile + tid * tile_ld + k0 + 4 * j)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = (ok && k0 + 4 * j < K1) ? __ldg(x1 + (k0 >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (k0 + 16 * c < K1) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              tc::split_tf32(v[4 * c + j].x, hi[4 * j], lo[4 * j]);
              tc::split_tf32(v[4 * c + j].y, hi[4 * j + 1], lo[4 * j + 1]);
              tc::split_tf32(v[4 * c + j].z, hi[4 * j + 2], lo[4 * j + 2]);
              tc::split_tf32(v[4 * c + j].w, hi[4 * j + 3], lo[4 * j + 3]);
            }
            tc::tmem_st16(lane_addr + col_hi + k0 + 16 * c, hi);
            tc::tmem_st16(lane_addr + col_lo + k0 + 16 * c, lo);
          }
        }
      }
      if (K2) {
        const float4* x2 = reinterpret_cast<const float4*>(a.X2 + (ok ? r : 0) * a.ldx2);
        for (int k0 = 0; k0 < K2; k0 += 64) {
          float4 v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = (ok && k0 + 4 * j < K2) ? __ldg(x2 + (k0 >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (k0 + 16 * c < K2) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                pk[2 * j] = tc::pack_bf16(v[4 * c + j].x, v[4 * c + j].y);
                pk[2 * j + 1] = tc::pack_bf16(v[4 * c + j].z, v[4 * c + j].w);
              }
              tc::tmem_st8(lane_addr + col_a2 + ((k0 + 16 * c) >> 1), pk);
            }
          }
        }
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(full);
      tc::mbar_wait(done, (uint32_t)(it & 1));
      tc::fence_after_sync();
      for (int n0 = 0; n0 < Ho; n0 += 16) {
        uint32_t d[16];
        tc::tmem_ld16(lane_addr + col_d + n0, d);
        tc::wait_ld();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(d[j]) + bias_s[n0 + j];
        if (a.act == GCM_ACT_EXP2X) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = tcg_exp2x(f[j]);
        } else {
          gcm_act_fast_vec(f, a.act);
        }
        if (ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) bad |= !isfinite(f[j]);
        }
        if (a.stage) {
          // (every thread has read its input row long ago: the MMAs that consumed it have completed)
          float4* o = reinterpret_cast<float4*>(tile + tid * tile_ld + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else if (ok) {
          float4* o = reinterpret_cast<float4*>(a.out + r * a.ldo + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      }
      if (a.stage) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const long long r0 = (blockIdx.x + it * gridDim.x) * 128;
        const int cpr = Ho >> 2, csh = __ffs(cpr) - 1;
        for (int c = tid; c < 128 * cpr; c += 128) {     // coalesced 16-byte stores of the result tile
          const int row = c >> csh, c4 = c & (cpr - 1);
          if (r0 + row < a.rows)
            *reinterpret_cast<float4*>(a.out + (r0 + row) * a.ldo + c4 * 4) =
                *reinterpret_cast<const float4*>(tile + row * tile_ld + c4 * 4);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the tile buffer is reused by the next tile's staging
      }
      tc::fence_before_sync();   // the accumulator and the A columns are free again: the next tile may overwrite them
    }
    if (a.status && bad) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
  } else {
    for (int i = tid - 128; i < 128; i += TC32_THREADS - 128) bias_s[i] = (a.bias && i < Ho) ? __ldg(a.bias + i) : 0.0f;
    stage_w_tf32<TC32_THREADS - 128>(a.W1, Ho, K1, Whi, Wlo, tid - 128);
    if (K2) stage_w_bf16<TC32_THREADS - 128>(a.W2, Ho, K2, W2s, tid - 128);
    tc::fence_proxy_async();
    tc::mbar_arrive(wready);
    if (warp == 8 && lane == 0) {
      tc::mbar_wait(wready, 0);
      const uint32_t idesc = tc::idesc_tf32(128, Ho);
      const uint32_t sbo = (uint32_t)(K1 / 4) * 128u;
      const uint32_t idesc16 = tc::idesc_bf16(128, Ho);
      const uint32_t sbo16 = (uint32_t)(K2 ? K2 / 8 : 1) * 128u;
      for (long long it = 0; it < n_it; ++it) {
        tc::mbar_wait(full, (uint32_t)(it & 1));
        tc::fence_after_sync();
        bool acc = false;
        for (int pass = 0; pass < 3; ++pass) {       // lo*Bhi, hi*Blo, hi*Bhi
          const uint32_t a_col = pass == 0 ? col_lo : col_hi;
          const float* bsrc = pass == 1 ? Wlo : Whi;
          for (int ks = 0; ks < K1 / 8; ++ks) {
            const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(bsrc) + ks * 256, 128, sbo);
            tc::mma_tf32_ts(tbase + col_d, tbase + a_col + ks * 8, bdesc, idesc, acc);
            acc = true;
          }
        }
        for (int ks = 0; ks < K2 / 16; ++ks) {
          const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(W2s) + ks * 256, 128, sbo16);
          tc::mma_bf16_ts(tbase + col_d, tbase + col_a2 + ks * 8, bdesc, idesc16, true);
        }
        tc::mma_commit(done);
      }
    }
  }
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tbase, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// part[cta][o][i] = sum over the CTA's rows of A[r, o] X[r, i]   (and the column sums of A); a second kernel
// adds the partials into dW / db in a fixed order (deterministic, no atomics).
// MMA view: D[M = o, N = i] += A^T[o, k = r] X^T[i, k = r]: TMEM lane = output channel o, the reduction runs over
// rows in chunks of 64.  Ho <= 128; Hi a multiple of 16, <= 128.
// ------------------------------------------------------------------------------------------------
constexpr int OT_KC = 64;
struct OuterTcArgs {
  const float* A; long long lda; int Ho;
  const float* X; long long ldx; int Hi;
  long long rows, rows_per_cta;
  float* part;        // [gridDim.x, 128, Hi]
  float* part_b;      // [gridDim.x, 128]
  // k_outer_tc32 only: channels i >= Hi1 of the X operand come from a second matrix (X2[r, i - Hi1]) -- two reductions
  // against the same A in one pass over the rows (the sparse GraphConv backward: dz^T [agg | x])
  const float* X2 = nullptr; long long ldx2 = 0; int Hi1 = 0;
};

__global__ void __launch_bounds__(TCG_THREADS, 2) k_outer_tc(const OuterTcArgs a) {
  __shared__ __align__(128) __nv_bfloat16 Bs[2][128 * OT_KC];     // X chunk of group g, canonical K-major [Hi x 64]
  __shared__ __align__(8) uint64_t full[2], done[2], fin;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ho = a.Ho, Hi = a.Hi;
  if (tid == 0) {
    tc::mbar_init(&full[0], 128); tc::mbar_init(&full[1], 128);
    tc::mbar_init(&done[0], 1);   tc::mbar_init(&done[1], 1);
    tc::mbar_init(&fin, 1);
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_slot, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const long long r_begin = (long long)blockIdx.x * a.rows_per_cta;
  const long long r_end = min(a.rows, r_begin + a.rows_per_cta);
  const long long nchunks = r_end > r_begin ? (r_end - r_begin + OT_KC - 1) / OT_KC : 0;

  if (warp < 8) {
    const int g = warp >> 2;
    const int ch = tid & 127;                                      // output channel o (A) / input channel i (X)
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t col_a = 128 + g * 32;
    const bool a_ok = ch < Ho, x_ok = ch < Hi;
    float colsum = 0.0f;
    unsigned char* bs = reinterpret_cast<unsigned char*>(Bs[g]) + ((ch >> 3) * (OT_KC >> 3)) * 128 + (ch & 7) * 16;
    for (long long j = g; j < nchunks; j += 2) {
      const long long it = j >> 1;
      const long long r0 = r_begin + j * OT_KC;
      if (it > 0) {   // the MMAs that read this group's buffers for its previous chunk have completed
        tc::mbar_wait(&done[g], (uint32_t)((it - 1) & 1));
        tc::fence_after_sync();
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float av[32], xv[32];
        const long long rh = r0 + half * 32;
        const float* pa = a.A + rh * a.lda + ch;
        const float* px = a.X + rh * a.ldx + ch;
        if (rh + 32 <= r_end) {   // whole half-chunk in range: no per-row guards, pointer increments only
#pragma unroll
          for (int u = 0; u < 32; ++u) av[u] = a_ok ? __ldcs(pa + u * a.lda) : 0.0f;
#pragma unroll
          for (int u = 0; u < 32; ++u) xv[u] = x_ok ? __ldcs(px + u * a.ldx) : 0.0f;
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u) av[u] = (a_ok && rh + u < r_end) ? __ldcs(pa + u * a.lda) : 0.0f;
#pragma unroll
          for (int u = 0; u < 32; ++u) xv[u] = (x_ok && rh + u < r_end) ? __ldcs(px + u * a.ldx) : 0.0f;
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint32_t pk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            pk[u] = tc::pack_bf16(av[s * 16 + 2 * u], av[s * 16 + 2 * u + 1]);
            colsum += av[s * 16 + 2 * u] + av[s * 16 + 2 * u + 1];
          }
          tc::tmem_st8(lane_addr + col_a + half * 16 + s * 8, pk);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {   // 8 rows (k) of channel ch -> one 16-byte core-matrix row
          const uint4 w = make_uint4(tc::pack_bf16(xv[s * 8], xv[s * 8 + 1]), tc::pack_bf16(xv[s * 8 + 2], xv[s * 8 + 3]),
                                     tc::pack_bf16(xv[s * 8 + 4], xv[s * 8 + 5]), tc::pack_bf16(xv[s * 8 + 6], xv[s * 8 + 7]));
          *reinterpret_cast<uint4*>(bs + (half * 4 + s) * 128) = w;
        }
      }
      tc::wait_st();
      tc::fence_proxy_async();
      tc::fence_before_sync();
      tc::mbar_arrive(&full[g]);
    }
    if (a.part_b && a_ok) atomicAdd(a.part_b + (size_t)blockIdx.x * 128 + ch, colsum);   // two groups, one address
    if (g == 0) {
      float* out = a.part + ((size_t)blockIdx.x * 128 + ch) * Hi;
      if (nchunks > 0) {
        tc::mbar_wait(&fin, 0);
        tc::fence_after_sync();
        for (int n0 = 0; n0 < Hi; n0 += 16) {
          uint32_t d[16];
          tc::tmem_ld16(lane_addr + n0, d);
          tc::wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(out + n0)[q] = make_float4(__uint_as_float(d[4 * q]), __uint_as_float(d[4 * q + 1]),
                                                                 __uint_as_float(d[4 * q + 2]), __uint_as_float(d[4 * q + 3]));
        }
        tc::fence_before_sync();
      } else {
        for (int n0 = 0; n0 < Hi; ++n0) out[n0] = 0.0f;
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc = tc::idesc_bf16(128, Hi);
    const uint32_t sbo = (uint32_t)(OT_KC / 8) * 128u;
    for (long long j = 0; j < nchunks; ++j) {
      const int g = (int)(j & 1);
      tc::mbar_wait(&full[g], (uint32_t)((j >> 1) & 1));
      tc::fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < OT_KC / 16; ++ks) {
        const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(Bs[g]) + ks * 256, 128, sbo);
        tc::mma_bf16_ts(tbase, tbase + 128 + g * 32 + ks * 8, bdesc, idesc, j > 0 || ks > 0);
      }
      tc::mma_commit(&done[g]);
    }
    if (nchunks > 0) tc::mma_commit(&fin);
  }
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tbase, 256);
}

// The same reduction in 3xTF32 (fp32-accurate: hi/lo split of both operands, lo*Bhi + hi*Blo + hi*Bhi), for callers that
// stay in float32 (the sparse GraphConv backward, DenseEdge states with a float32 cache).  A^T hi | lo take 64 + 64 TMEM
// columns per group, the X chunk hi | lo 2 x 32 KB of shared memory per group.
// THREE loader groups (TMEM: D [0,128) + 3 x 128 operand columns = 512; shared memory 3 x 64 KB): with two, a group's chain
// load -> split -> tcgen05.st / shared stores -> MMA -> done took 7.7 us per 64-row chunk at cfg5 and the kernel moved
// 1.9 TB/s.
constexpr int OT32_NG = 3;
constexpr int OT32_THREADS = (4 * OT32_NG + 1) * 32;
__global__ void __launch_bounds__(OT32_THREADS, 1) k_outer_tc32(const OuterTcArgs a) {
  extern __shared__ __align__(128) unsigned char o32_smem[];
  float* Bs = reinterpret_cast<float*>(o32_smem);                 // [group][hi, lo][128 x 64] canonical K-major
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + OT32_NG * 2 * 128 * OT_KC);
  uint64_t* full = bars;             // [NG]
  uint64_t* done = bars + OT32_NG;   // [NG]
  uint64_t* fin = bars + 2 * OT32_NG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * OT32_NG + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ho = a.Ho, Hi = a.Hi;
  if (tid == 0) {
    for (int g = 0; g < OT32_NG; ++g) {
      tc::mbar_init(&full[g], 128);
      tc::mbar_init(&done[g], 1);
    }
    tc::mbar_init(fin, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4 * OT32_NG) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const long long r_begin = (long long)blockIdx.x * a.rows_per_cta;
  const long long r_end = min(a.rows, r_begin + a.rows_per_cta);
  const long long nchunks = r_end > r_begin ? (r_end - r_begin + OT_KC - 1) / OT_KC : 0;

  if (warp < 4 * OT32_NG) {
    const int g = warp >> 2;
    const int ch = tid & 127;                                      // output channel o (A) / input channel i (X)
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t col_a = 128 + g * 128;                          // hi at +0 .. 63, lo at +64 .. 127
    const bool a_ok = ch < Ho, x_ok = ch < Hi;
    float colsum = 0.0f;
    float* bhi = Bs + (size_t)g * 2 * 128 * OT_KC + ((ch >> 3) * (OT_KC >> 2)) * 32 + (ch & 7) * 4;
    float* blo = bhi + 128 * OT_KC;
    for (long long j = g; j < nchunks; j += OT32_NG) {
      const long long it = j / OT32_NG;
      const long long r0 = r_begin + j * OT_KC;
      if (it > 0) {
        tc::mbar_wait(&done[g], (uint32_t)((it - 1) & 1));
        tc::fence_after_sync();
      }
#pragma unroll 1
      for (int q2 = 0; q2 < 2; ++q2) {                             // 32 rows (k) at a time: 64 loads in flight per thread
        const long long rq = r0 + q2 * 32;
        const float* pa = a.A + rq * a.lda + ch;
        const bool second = a.X2 != nullptr && ch >= a.Hi1;
        const long long ldx = second ? a.ldx2 : a.ldx;
        const float* px = second ? a.X2 + rq * ldx + (ch - a.Hi1) : a.X + rq * ldx + ch;
        float av[32], xv[32];
        const bool whole = rq + 32 <= r_end;
#pragma unroll
        for (int u = 0; u < 32; ++u) av[u] = (a_ok && (whole || rq + u < r_end)) ? __ldcs(pa + u * a.lda) : 0.0f;
#pragma unroll
        for (int u = 0; u < 32; ++u) xv[u] = (x_ok && (whole || rq + u < r_end)) ? __ldcs(px + u * ldx) : 0.0f;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int q = q2 * 2 + h2;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            colsum += av[h2 * 16 + u];
            tc::split_tf32(av[h2 * 16 + u], hi[u], lo[u]);
          }
          tc::tmem_st16(lane_addr + col_a + q * 16, hi);
          tc::tmem_st16(lane_addr + col_a + 64 + q * 16, lo);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {   // 4 rows (k) of channel ch -> one 16-byte core-matrix row, hi and lo
            uint32_t h[4], l[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) tc::split_tf32(xv[h2 * 16 + s4 * 4 + u], h[u], l[u]);
            *reinterpret_cast<uint4*>(bhi + (q * 4 + s4) * 32) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(blo + (q * 4 + s4) * 32) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      tc::wait_st();
      tc::fence_proxy_async();
      tc::fence_before_sync();
      tc::mbar_arrive(&full[g]);
    }
    if (a.part_b && a_ok) atomicAdd(a.part_b + (size_t)blockIdx.x * 128 + ch, colsum);
    if (g == 0) {
      float* out = a.part + ((size_t)blockIdx.x * 128 + ch) * Hi;
      if (nchunks > 0) {
        tc::mbar_wait(fin, 0);
        tc::fence_after_sync();
        for (int n0 = 0; n0 < Hi; n0 += 16) {
          uint32_t d[16];
          tc::tmem_ld16(lane_addr + n0, d);
          tc::wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float4*>(out + n0)[q] = make_float4(__uint_as_float(d[4 * q]), __uint_as_float(d[4 * q + 1]),
                                                                 __uint_as_float(d[4 * q + 2]), __uint_as_float(d[4 * q + 3]));
        }
        tc::fence_before_sync();
      } else {
        for (int n0 = 0; n0 < Hi; ++n0) out[n0] = 0.0f;
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, Hi);
    const uint32_t sbo = (uint32_t)(OT_KC / 4) * 128u;
    for (long long j = 0; j < nchunks; ++j) {
      const int g = (int)(j % OT32_NG);
      tc::mbar_wait(&full[g], (uint32_t)((j / OT32_NG) & 1));
      tc::fence_after_sync();
      const float* bh = Bs + (size_t)g * 2 * 128 * OT_KC;
      const float* bl = bh + 128 * OT_KC;
      for (int pass = 0; pass < 3; ++pass) {                      // lo*Bhi, hi*Blo, hi*Bhi
        const uint32_t a_col = 128 + g * 128 + (pass == 0 ? 64 : 0);
        const float* bsrc = pass == 1 ? bl : bh;
#pragma unroll
        for (int ks = 0; ks < OT_KC / 8; ++ks) {
          const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(bsrc) + ks * 256, 128, sbo);
          tc::mma_tf32_ts(tbase, tbase + a_col + ks * 8, bdesc, idesc, j > 0 || pass > 0 || ks > 0);
        }
      }
      tc::mma_commit(&done[g]);
    }
    if (nchunks > 0) tc::mma_commit(fin);
  }
  __syncthreads();
  if (warp == 4 * OT32_NG) tc::tmem_dealloc(tbase, 512);
}

// dW[o, i] += sum_c part[c][o][i];  db[o] += sum_c part_b[c][o]
__global__ void __launch_bounds__(256) k_outer_tc_reduce(const float* part, const float* part_b, int ctas, int Ho, int Hi,
                                                         float* dW, float* db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Ho * Hi) {
    const int o = i / Hi, c = i - o * Hi;
    float s = 0.0f;
    for (int k = 0; k < ctas; ++k) s += part[((size_t)k * 128 + o) * Hi + c];
    dW[i] += s;
  }
  if (db && i < Ho) {
    float s = 0.0f;
    for (int k = 0; k < ctas; ++k) s += part_b[(size_t)k * 128 + i];
    db[i] += s;
  }
}

// pair form: columns [0, Hi1) of the partial products go to dW1 [Ho, Hi1], columns [Hi1, Hi) to dW2 [Ho, Hi - Hi1]
__global__ void __launch_bounds__(256) k_outer_tc_reduce_pair(const float* part, const float* part_b, int ctas, int Ho, int Hi,
                                                              int Hi1, float* dW1, float* dW2, float* db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Ho * Hi) {
    const int o = i / Hi, c = i - o * Hi;
    float s = 0.0f;
    for (int k = 0; k < ctas; ++k) s += part[((size_t)k * 128 + o) * Hi + c];
    if (c < Hi1) dW1[o * Hi1 + c] += s;
    else dW2[o * (Hi - Hi1) + (c - Hi1)] += s;
  }
  if (db && i < Ho) {
    float s = 0.0f;
    for (int k = 0; k < ctas; ++k) s += part_b[(size_t)k * 128 + i];
    db[i] += s;
  }
}

}  // namespace

extern "C" int gcm_linear_tc(const float* X, int K, long long ldx, const float* W, const float* bias, int act,
                             long long rows, int Ho, void* out, long long ldo, int out_bf16, void* stream) {
  GCM_REQUIRE(X && W && out && rows >= 0, "linear_tc: null pointer");
  GCM_REQUIRE(K >= 16 && K <= 128 && K % 16 == 0 && Ho >= 16 && Ho <= 128 && Ho % 16 == 0 && ldx % 4 == 0 &&
                  ldo % (out_bf16 ? 8 : 4) == 0,
              "linear_tc: K=%d and Ho=%d must be multiples of 16 in [16,128], rows 16-byte aligned", K, Ho);
  GCM_REQUIRE(act == GCM_ACT_NONE || act == GCM_ACT_EXP2X, "linear_tc: epilogue must be none or EXP2X");
  GCM_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "linear_tc: X, W and out must be 16-byte aligned");
  if (rows == 0) return GCM_OK;
  LinearTcArgs a{X, ldx, K, W, bias, act, rows, Ho, out, ldo, out_bf16, (rows + 127) / 128};
  const size_t smem = (size_t)Ho * K * 2 + 64 + 512;
  long long grid = a.tiles < gcm_num_sms() ? a.tiles : gcm_num_sms();
  k_linear_tc<<<(unsigned)grid, TCG_THREADS, smem, (cudaStream_t)stream>>>(a);
  return gcm_check_launch("k_linear_tc");
}

extern "C" int gcm_linear_tc32(const float* X1, int K1, long long ldx1, const float* W1, const float* X2, int K2,
                               long long ldx2, const float* W2, const float* bias, int act, long long rows, int Ho,
                               float* out, long long ldo, int32_t* status, void* stream) {
  GCM_REQUIRE(X1 && W1 && out && rows >= 0, "linear_tc32: null pointer");
  GCM_REQUIRE((X2 == nullptr) == (W2 == nullptr), "linear_tc32: X2 and W2 go together");
  GCM_REQUIRE(K1 >= 16 && K1 <= 128 && K1 % 16 == 0 && Ho >= 16 && Ho <= 128 && Ho % 16 == 0 && ldx1 % 4 == 0 && ldo % 4 == 0,
              "linear_tc32: K1=%d and Ho=%d must be multiples of 16 in [16,128], rows 16-byte aligned", K1, Ho);
  GCM_REQUIRE(!X2 || (K2 >= 16 && K2 <= 128 && K2 % 16 == 0 && ldx2 % 4 == 0), "linear_tc32: bad K2=%d", K2);
  GCM_REQUIRE((act >= GCM_ACT_NONE && act <= GCM_ACT_RELU) || act == GCM_ACT_EXP2X, "linear_tc32: bad epilogue");
  GCM_REQUIRE(((reinterpret_cast<uintptr_t>(X1) | reinterpret_cast<uintptr_t>(W1) | reinterpret_cast<uintptr_t>(X2) |
                reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "linear_tc32: operands must be 16-byte aligned");
  if (rows == 0) return GCM_OK;
  LinearTc32Args a{X1, ldx1, K1, W1, X2, ldx2, X2 ? K2 : 0, W2, bias, act, rows, Ho, out, ldo, status, (rows + 127) / 128, 32, 0};
  const uint32_t need_cols = (uint32_t)(2 * K1 + a.K2 / 2 + Ho);
  while (a.tmem_cols < need_cols) a.tmem_cols <<= 1;
  // narrow streaming products (many rows, K1 <= 64, Ho <= 64): row / result tiles staged through shared memory
  static const bool no_stage = getenv("GCM_B200_NO_LINEAR_STAGE") != nullptr;
  a.stage = (!X2 && K1 <= 64 && Ho <= 64 && (K1 & (K1 - 1)) == 0 && (Ho & (Ho - 1)) == 0 && rows >= 4096 && !no_stage) ? 1 : 0;
  const size_t smem = (size_t)Ho * K1 * 8 + (size_t)Ho * a.K2 * 2 + 64 + 64 + 512 +
                      (a.stage ? (size_t)128 * ((K1 > Ho ? K1 : Ho) + 4) * 4 : 0);
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(k_linear_tc32, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 10 + 1024) != cudaSuccess) {
      gcm_set_error("linear_tc32: cannot raise the dynamic shared memory limit");
      return GCM_ERR_CUDA;
    }
    attr_done = true;
  }
  // persistent: as many CTAs as fit at once (TMEM columns and shared memory decide), each walks its tiles
  int per_sm = (int)(512 / a.tmem_cols);
  const int by_smem = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > 2) per_sm = 2;   // 104 registers x 288 threads: two CTAs per SM
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)per_sm * gcm_num_sms();
  if (grid > a.tiles) grid = a.tiles;
  k_linear_tc32<<<(unsigned)grid, TC32_THREADS, smem, (cudaStream_t)stream>>>(a);
  return gcm_check_launch("k_linear_tc32");
}

extern "C" long long gcm_outer_reduce_tc_workspace(long long rows) {
  long long ctas = (rows + 4 * OT_KC - 1) / (4 * OT_KC);
  const long long cap = 2LL * gcm_num_sms();
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return ctas * 128 * 129;   // floats: [ctas,128,<=128] partial products + [ctas,128] column sums
}

static int outer_reduce_tc_impl(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi, long long rows,
                                float* workspace, float* dW, float* db, bool tf32x3, void* stream) {
  GCM_REQUIRE(A && X && dW && workspace && rows >= 0, "outer_reduce_tc: null pointer");
  GCM_REQUIRE(Ho >= 1 && Ho <= 128 && Hi >= 16 && Hi <= 128 && Hi % 16 == 0,
              "outer_reduce_tc: Ho=%d must be <= 128, Hi=%d a multiple of 16 in [16,128]", Ho, Hi);
  if (rows == 0) return GCM_OK;
  long long ctas = gcm_outer_reduce_tc_workspace(rows) / (128 * 129);
  if (tf32x3 && ctas > gcm_num_sms()) ctas = gcm_num_sms();       // 512 TMEM columns: one CTA per SM
  long long per = (rows + ctas - 1) / ctas;
  per = (per + OT_KC - 1) / OT_KC * OT_KC;
  ctas = (rows + per - 1) / per;
  float* part = workspace;
  float* part_b = workspace + (size_t)ctas * 128 * Hi;
  cudaStream_t s = (cudaStream_t)stream;
  if (db && cudaMemsetAsync(part_b, 0, (size_t)ctas * 128 * sizeof(float), s) != cudaSuccess) {
    gcm_set_error("outer_reduce_tc: cudaMemsetAsync failed");
    return GCM_ERR_CUDA;
  }
  OuterTcArgs a{A, lda, Ho, X, ldx, Hi, rows, per, part, db ? part_b : nullptr};
  if (tf32x3) {
    const size_t smem = (size_t)OT32_NG * 2 * 128 * OT_KC * sizeof(float) + 128;
    static bool attr_done = false;
    if (!attr_done) {
      if (cudaFuncSetAttribute(k_outer_tc32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        gcm_set_error("outer_reduce_tc32: cannot raise the dynamic shared memory limit");
        return GCM_ERR_CUDA;
      }
      attr_done = true;
    }
    k_outer_tc32<<<(unsigned)ctas, OT32_THREADS, smem, s>>>(a);
    if (int rc = gcm_check_launch("k_outer_tc32")) return rc;
  } else {
    k_outer_tc<<<(unsigned)ctas, TCG_THREADS, 0, s>>>(a);
    if (int rc = gcm_check_launch("k_outer_tc")) return rc;
  }
  k_outer_tc_reduce<<<(Ho * Hi + 255) / 256, 256, 0, s>>>(part, part_b, (int)ctas, Ho, Hi, dW, db);
  return gcm_check_launch("k_outer_tc_reduce");
}

extern "C" int gcm_outer_reduce_tc32_pair(const float* A, long long lda, int Ho, const float* X1, long long ldx1, int Hi1,
                                          const float* X2, long long ldx2, int Hi2, long long rows, float* workspace,
                                          float* dW1, float* dW2, float* db, void* stream) {
  GCM_REQUIRE(A && X1 && X2 && dW1 && dW2 && workspace && rows >= 0, "outer_reduce_tc32_pair: null pointer");
  const int Hi = Hi1 + Hi2;
  GCM_REQUIRE(Ho >= 1 && Ho <= 128 && Hi1 >= 16 && Hi2 >= 16 && Hi <= 128 && Hi1 % 16 == 0 && Hi2 % 16 == 0,
              "outer_reduce_tc32_pair: Ho=%d must be <= 128, Hi1=%d and Hi2=%d multiples of 16 with a sum <= 128", Ho, Hi1, Hi2);
  if (rows == 0) return GCM_OK;
  long long ctas = gcm_outer_reduce_tc_workspace(rows) / (128 * 129);
  if (ctas > gcm_num_sms()) ctas = gcm_num_sms();       // 512 TMEM columns: one CTA per SM
  long long per = (rows + ctas - 1) / ctas;
  per = (per + OT_KC - 1) / OT_KC * OT_KC;
  ctas = (rows + per - 1) / per;
  float* part = workspace;
  float* part_b = workspace + (size_t)ctas * 128 * Hi;
  cudaStream_t s = (cudaStream_t)stream;
  if (db && cudaMemsetAsync(part_b, 0, (size_t)ctas * 128 * sizeof(float), s) != cudaSuccess) {
    gcm_set_error("outer_reduce_tc32_pair: cudaMemsetAsync failed");
    return GCM_ERR_CUDA;
  }
  OuterTcArgs a{A, lda, Ho, X1, ldx1, Hi, rows, per, part, db ? part_b : nullptr, X2, ldx2, Hi1};
  const size_t smem = (size_t)OT32_NG * 2 * 128 * OT_KC * sizeof(float) + 128;
  if (cudaFuncSetAttribute(k_outer_tc32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    gcm_set_error("outer_reduce_tc32_pair: cannot raise the dynamic shared memory limit");
    return GCM_ERR_CUDA;
  }
  k_outer_tc32<<<(unsigned)ctas, OT32_THREADS, smem, s>>>(a);
  if (int rc = gcm_check_launch("k_outer_tc32")) return rc;
  k_outer_tc_reduce_pair<<<(Ho * Hi + 255) / 256, 256, 0, s>>>(part, part_b, (int)ctas, Ho, Hi, Hi1, dW1, dW2, db);
  return gcm_check_launch("k_outer_tc_reduce_pair");
}

extern "C" int gcm_outer_reduce_tc(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi,
                                   long long rows, float* workspace, float* dW, float* db, void* stream) {
  return outer_reduce_tc_impl(A, lda, Ho, X, ldx, Hi, rows, workspace, dW, db, false, stream);
}

extern "C" int gcm_outer_reduce_tc32(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi,
                                     long long rows, float* workspace, float* dW, float* db, void* stream) {
  return outer_reduce_tc_impl(A, lda, Ho, X, ldx, Hi, rows, workspace, dW, db, true, stream);
}

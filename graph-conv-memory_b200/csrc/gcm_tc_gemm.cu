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
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, Ho = a.Ho;

  if (tid == 0) {
    tc::mbar_init(&full[0], 128); tc::mbar_init(&full[1], 128);
    tc::mbar_init(&done[0], 1);   tc::mbar_init(&done[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < Ho * K; i += TCG_THREADS) {
    const int n = i / K, k = i - n * K;
    Ws[tc::kmajor_off_bf16(n, k, K)] = __float2bfloat16_rn(__ldg(a.W + i));
  }
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
          const float z = __uint_as_float(d[j]) + (a.bias ? __ldg(a.bias + n0 + j) : 0.0f);
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
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const long long r = r0 + half * 32 + u;
          av[u] = (a_ok && r < r_end) ? __ldcs(a.A + r * a.lda + ch) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const long long r = r0 + half * 32 + u;
          xv[u] = (x_ok && r < r_end) ? __ldcs(a.X + r * a.ldx + ch) : 0.0f;
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

}  // namespace

extern "C" int gcm_linear_tc(const float* X, int K, long long ldx, const float* W, const float* bias, int act,
                             long long rows, int Ho, void* out, long long ldo, int out_bf16, void* stream) {
  GCM_REQUIRE(X && W && out && rows >= 0, "linear_tc: null pointer");
  GCM_REQUIRE(K >= 16 && K <= 128 && K % 16 == 0 && Ho >= 16 && Ho <= 128 && Ho % 16 == 0 && ldx % 4 == 0 &&
                  ldo % (out_bf16 ? 8 : 4) == 0,
              "linear_tc: K=%d and Ho=%d must be multiples of 16 in [16,128], rows 16-byte aligned", K, Ho);
  GCM_REQUIRE(act == GCM_ACT_NONE || act == GCM_ACT_EXP2X, "linear_tc: epilogue must be none or EXP2X");
  if (rows == 0) return GCM_OK;
  LinearTcArgs a{X, ldx, K, W, bias, act, rows, Ho, out, ldo, out_bf16, (rows + 127) / 128};
  const size_t smem = (size_t)Ho * K * 2 + 64;
  long long grid = a.tiles < gcm_num_sms() ? a.tiles : gcm_num_sms();
  k_linear_tc<<<(unsigned)grid, TCG_THREADS, smem, (cudaStream_t)stream>>>(a);
  return gcm_check_launch("k_linear_tc");
}

extern "C" long long gcm_outer_reduce_tc_workspace(long long rows) {
  long long ctas = (rows + 4 * OT_KC - 1) / (4 * OT_KC);
  const long long cap = 2LL * gcm_num_sms();
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return ctas * 128 * 129;   // floats: [ctas,128,<=128] partial products + [ctas,128] column sums
}

extern "C" int gcm_outer_reduce_tc(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi,
                                   long long rows, float* workspace, float* dW, float* db, void* stream) {
  GCM_REQUIRE(A && X && dW && workspace && rows >= 0, "outer_reduce_tc: null pointer");
  GCM_REQUIRE(Ho >= 1 && Ho <= 128 && Hi >= 16 && Hi <= 128 && Hi % 16 == 0,
              "outer_reduce_tc: Ho=%d must be <= 128, Hi=%d a multiple of 16 in [16,128]", Ho, Hi);
  if (rows == 0) return GCM_OK;
  long long ctas = gcm_outer_reduce_tc_workspace(rows) / (128 * 129);
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
  k_outer_tc<<<(unsigned)ctas, TCG_THREADS, 0, s>>>(a);
  if (int rc = gcm_check_launch("k_outer_tc")) return rc;
  k_outer_tc_reduce<<<(Ho * Hi + 255) / 256, 256, 0, s>>>(part, part_b, (int)ctas, Ho, Hi, dW, db);
  return gcm_check_launch("k_outer_tc_reduce");
}

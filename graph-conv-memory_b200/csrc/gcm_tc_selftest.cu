// Self-test of the tcgen05 building block used by the tensor-core step kernels:
//   D[128, N] = A[128, K] * B[N, K]^T   in "3xTF32" (fp32-accurate) arithmetic,
// A written to TMEM by the row-owning threads (tcgen05.st), B staged in shared memory in the canonical
// K-major no-swizzle layout, accumulator in TMEM, read back with tcgen05.ld.
#include "gcm_common.cuh"
#include "gcm_tc.cuh"
#include <cuda_bf16.h>

__global__ void __launch_bounds__(160) k_tc_selftest(const float* A, const float* B, float* D, int K, int N,
                                                     int passes) {
  extern __shared__ __align__(128) unsigned char st_raw[];
  float* Bhi = reinterpret_cast<float*>(st_raw);        // [N x K] canonical
  float* Blo = Bhi + N * K;
  __nv_bfloat16* Bbf = reinterpret_cast<__nv_bfloat16*>(Blo + N * K);   // [N x K] canonical, bf16 (passes == 4)
  float* Ahi = reinterpret_cast<float*>(Bbf + N * K);   // [128 x K] canonical (passes == 5: A from shared memory)
  float* Alo = Ahi + (passes == 5 ? 128 * K : 0);
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(Alo + (passes == 5 ? 128 * K : 0));
  uint64_t* bar_d = bar_a + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(bar_a, 128);
    tc::mbar_init(bar_d, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    uint32_t hi, lo;
    tc::split_tf32(B[i], hi, lo);
    Bhi[tc::kmajor_off(n, k, K)] = __uint_as_float(hi);
    Blo[tc::kmajor_off(n, k, K)] = __uint_as_float(lo);
    Bbf[tc::kmajor_off_bf16(n, k, K)] = __float2bfloat16_rn(B[i]);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t col_ahi = 0, col_alo = 128, col_d = 256;

  if (warp < 4) {
    const int row = tid;  // TMEM lane
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) tc::split_tf32(A[(size_t)row * K + k0 + j], hi[j], lo[j]);
      if (passes == 5) {   // SS form: the row-owning thread writes its row into the canonical shared-memory layout
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          Ahi[tc::kmajor_off(row, k0 + j, K)] = __uint_as_float(hi[j]);
          Alo[tc::kmajor_off(row, k0 + j, K)] = __uint_as_float(lo[j]);
        }
        continue;
      }
      tc::tmem_st16(lane_addr + col_ahi + k0, hi);
      if (passes == 4) {   // lo part as packed bf16: column c holds k = 2c (low half) and 2c + 1
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = tc::pack_bf16(__uint_as_float(lo[2 * j]), __uint_as_float(lo[2 * j + 1]));
        tc::tmem_st8(lane_addr + col_alo + k0 / 2, pk);
      } else {
        tc::tmem_st16(lane_addr + col_alo + k0, lo);
      }
    }
    tc::wait_st();
    if (passes == 5) tc::fence_proxy_async();
    tc::fence_before_sync();
    tc::mbar_arrive(bar_a);
    tc::mbar_wait(bar_d, 0);
    tc::fence_after_sync();
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      tc::tmem_ld16(lane_addr + col_d + n0, v);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[(size_t)row * N + n0 + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync();
  } else {
    // MMA issuer: one thread
    tc::mbar_wait(bar_a, 0);
    tc::fence_after_sync();
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_tf32(128, N);
      const uint32_t sbo = (uint32_t)(K / 4) * 128u;
      bool acc = false;
      if (passes == 5) {   // 3xTF32 with BOTH operands in shared memory: lo*Bhi, hi*Blo, hi*Bhi
        for (int pass = 0; pass < 3; ++pass) {
          const float* asrc = pass == 0 ? Alo : Ahi;
          const float* bsrc = pass == 1 ? Blo : Bhi;
          for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t adesc = tc::smem_desc_kmajor(tc::smem_u32(asrc) + ks * 256, 128, sbo);
            const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(bsrc) + ks * 256, 128, sbo);
            tc::mma_tf32_ss(tbase + col_d, adesc, bdesc, idesc, acc);
            acc = true;
          }
        }
      }
      if (passes == 4) {   // lo(bf16, TMEM) * B(bf16): 16 elements of K per instruction
        const uint32_t idesc16 = tc::idesc_bf16(128, N);
        const uint32_t sbo16 = (uint32_t)(K / 8) * 128u;
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(Bbf) + ks * 256, 128, sbo16);
          tc::mma_bf16_ts(tbase + col_d, tbase + col_alo + ks * 8, bdesc, idesc16, acc);
          acc = true;
        }
      }
      for (int pass = (passes == 4 ? 1 : 0); pass < (passes == 4 ? 3 : (passes == 5 ? 0 : passes)); ++pass) {
        // passes == 3: lo*Bhi, hi*Blo, hi*Bhi ; passes == 1: hi*Bhi only (plain tf32);
        // passes == 4: the lo*B term was issued above in bf16, then hi*Blo, hi*Bhi
        const int which = passes >= 3 ? pass : 2;
        const uint32_t a_col = (which == 0) ? col_alo : col_ahi;
        const float* bsrc = (which == 1) ? Blo : Bhi;
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t bdesc = tc::smem_desc_kmajor(tc::smem_u32(bsrc) + ks * 256, 128, sbo);
          tc::mma_tf32_ts(tbase + col_d, tbase + a_col + ks * 8, bdesc, idesc, acc);
          acc = true;
        }
      }
      tc::mma_commit(bar_d);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tbase, 512);
}

extern "C" int gcm_tc_selftest(const float* A, const float* B, float* D, int K, int N, int passes, void* stream) {
  GCM_REQUIRE(A && B && D, "tc_selftest: null pointer");
  GCM_REQUIRE(K % 16 == 0 && K >= 16 && K <= 128 && N % 16 == 0 && N >= 16 && N <= 256,
              "tc_selftest: K=%d must be a multiple of 16 in [16,128], N=%d a multiple of 16 in [16,256]", K, N);
  GCM_REQUIRE(passes == 1 || passes == 3 || passes == 4 || passes == 5, "tc_selftest: passes must be 1, 3, 4 or 5");
  const size_t smem = (size_t)2 * N * K * 4 + (size_t)N * K * 2 + (passes == 5 ? (size_t)2 * 128 * K * 4 : 0) + 64;
  GCM_REQUIRE(smem <= 220 * 1024, "tc_selftest: operands do not fit in shared memory");
  cudaError_t e = cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    gcm_set_error("cudaFuncSetAttribute(tc_selftest): %s", cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  k_tc_selftest<<<1, 160, smem, (cudaStream_t)stream>>>(A, B, D, K, N, passes);
  return gcm_check_launch("k_tc_selftest");
}

// Shared device helpers for libgcm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gcm_b200.h"

#define GCM_FULL_MASK 0xffffffffu
#define GCM_FR (GCM_MAX_FEAT / 32)  // per-lane feature registers of the generic kernels

void gcm_set_error(const char* fmt, ...);
int gcm_check_launch(const char* what);
int gcm_num_sms();

#define GCM_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      gcm_set_error(__VA_ARGS__);     \
      return GCM_ERR_INVALID;         \
    }                                 \
  } while (0)

__device__ __forceinline__ float gcm_act_fwd(float z, int kind) {
  if (kind == GCM_ACT_TANH) return tanhf(z);
  if (kind == GCM_ACT_RELU) return z <= 0.0f ? 0.0f : z;  // NaN propagates, like torch.relu
  return z;
}

// tanh(z) = 1 - 2 / (2^(2 z log2 e) + 1) through ex2.approx.ftz + rcp.approx.ftz: 5 instructions (2 MUFU),
// |error| < 5e-7 absolute (budget: 1e-5 relative parity).  Saturates to +-1 for large |z| (2^t -> inf -> rcp 0,
// 2^t -> 0 -> rcp 1); NaN propagates.
__device__ __forceinline__ float gcm_tanh_fast(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ float gcm_act_fast(float z, int kind) {
  if (kind == GCM_ACT_TANH) return gcm_tanh_fast(z);
  if (kind == GCM_ACT_RELU) return z <= 0.0f ? 0.0f : z;
  return z;
}

// activation of a register array with ONE warp-uniform branch on the kind (a per-element switch compiles to
// a BSSY/BSYNC pair per element and dominated the tensor-core epilogues, profiles/c2_step_temporal_tg_r1.md)
template <int NV>
__device__ __forceinline__ void gcm_act_fast_vec(float (&v)[NV], int kind) {
  if (kind == GCM_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = gcm_tanh_fast(v[j]);
  } else if (kind == GCM_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = v[j] <= 0.0f ? 0.0f : v[j];
  }
}

// derivative of the activation expressed through its OUTPUT value
__device__ __forceinline__ float gcm_act_grad(float out, int kind) {
  if (kind == GCM_ACT_TANH) return 1.0f - out * out;
  if (kind == GCM_ACT_RELU) return out > 0.0f ? 1.0f : 0.0f;
  return 1.0f;
}

__device__ __forceinline__ float gcm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GCM_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ int gcm_warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GCM_FULL_MASK, v, o);
  return v;
}

// bits [lo, hi] (inclusive) of the N-bit mask that fall into 32-bit word `w`
__device__ __forceinline__ uint32_t gcm_range_word(int w, int lo, int hi) {
  int a = max(lo - w * 32, 0);
  int b = min(hi - w * 32, 31);
  if (b < a) return 0u;
  uint32_t upto_b = (b == 31) ? 0xffffffffu : ((1u << (b + 1)) - 1u);
  uint32_t below_a = (1u << a) - 1u;
  return upto_b & ~below_a;
}

__device__ __forceinline__ int gcm_slot(int pos, int C) { return pos % C; }

// masks live in L2 for the kernels (atomicOr from other threads bypasses L1)
__device__ __forceinline__ uint32_t gcm_ld_mask(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ void gcm_st_mask(uint32_t* p, uint32_t v) { __stcg(p, v); }

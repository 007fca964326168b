// Window-level backward of forward-only TemporalBackedge chains: the streaming (gather / shift-sum) passes.
//
// Reference: autograd through gcm.py:262-321 with the DenseGraphConv stack of README.md:52-62 over the T steps of a
// BPTT window (tests/test_gcm.py:412-439 is the reference's training loop).  For a chain of forward hops S every
// in-neighbourhood is fixed when a node is created (edge_selectors/temporal.py:72-88), so with
//   h_p      = act1(W_rel1 sum_{s in S} x_{p-s} + W_root1 x_p + b1)               (layer 1 of node p)
//   belief_p = act2(W_rel2 sum_{s in S} h_{p-s} + W_root2 h_p + b2)               (step that created node p)
// the whole window's gradient is a handful of row-parallel products over (step, graph) rows:
//   dz2_p = dbelief_p * act2'(belief_p)
//   dh_p  = [sum_s dz2_{p+s} | dz2_p] [W_rel2 ; W_root2]              dz1_p = dh_p * act1'(h_p)
//   dx_q  = [sum_s dz1_{q+s} | dz1_q] [W_rel1 ; W_root1]
//   dW2   = dz2^T [sum_s h_{p-s} | h_p]      dW1 = dz1^T [sum_s x_{p-s} | x_p]      (+ column sums for the biases)
// GCM has no recurrence through the belief, so all dbelief are known before any of this starts.  The products run on
// the tensor cores (3xTF32: gcm_linear_tc32 / gcm_outer_reduce_tc32, csrc/gcm_tc_gemm.cu); this file holds the passes
// that build their operands: rows are TIME-MAJOR [row, graph, feature] with row <-> absolute node position, an edge
// p-s -> p exists iff p - s >= 0 (the state was built from empty by this chain: `pure temporal`, uniform count).
#include "gcm_common.cuh"
#include <stdlib.h>

namespace {

struct HopList {
  int n;
  int h[GCM_MAX_HOPS];
};

// out[i, b, 0:F] = sum_{s: p-s >= 0} x[b, slot(p-s), :],  out[i, b, F:2F] = x[b, slot(p), :],  p = p0 + i (zero row if p < 0)
// grid: (chunks of B * F/4, rows); 32-bit index arithmetic (node positions are int32 counters on the device): the first
// version spent a fifth of its issue slots in emulated 64-bit divisions (profiles/c2_bptt_kernels_r2.md).
// tiled = 1 (F = 32 only): out is [tile of 128 rows][16-byte chunk 0 .. 15][row of the tile][4 floats] over the flattened
// rows r = i * B + b, the layout the fused window-backward kernel loads with fully coalesced warps (gcm_temporal_bwd_tc.cu)
__device__ __forceinline__ float4* tiled_slot(float* out, long long r, int chunk) {
  return reinterpret_cast<float4*>(out) + (r >> 7) * (16 * 128) + chunk * 128 + (r & 127);
}

__global__ void __launch_bounds__(256) k_temporal_gather(const gcm_dense_state st, const HopList hops, int p0,
                                                         float* __restrict__ out, int tiled) {
  const int F4 = st.F >> 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;       // (graph, 16-byte chunk)
  if (j >= st.B * F4) return;
  const int b = j / F4, c = j - b * F4;
  const int i = blockIdx.y;
  const int p = p0 + i;
  float4 self = make_float4(0.f, 0.f, 0.f, 0.f), sum = self;
  if (p >= 0) {
    const float4* rows = reinterpret_cast<const float4*>(st.nodes + (size_t)b * st.C * st.F) + c;
    self = __ldg(rows + (size_t)(p % st.C) * F4);
    for (int k = 0; k < hops.n; ++k) {
      const int q = p - hops.h[k];
      if (q >= 0) {
        const float4 v = __ldg(rows + (size_t)(q % st.C) * F4);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
    }
  }
  if (tiled) {
    const long long r = (long long)i * st.B + b;
    __stcs(tiled_slot(out, r, c), sum);
    __stcs(tiled_slot(out, r, F4 + c), self);
    return;
  }
  float4* o = reinterpret_cast<float4*>(out + ((size_t)i * st.B + b) * 2 * st.F) + c;
  __stcs(o, sum);
  __stcs(o + F4, self);
}

// out[i, b, 0:H] = sum_s src[pos + sign * s, b, :],  out[i, b, H:2H] = src[pos, b, :],  pos = out_pos0 + i;
// src rows cover positions [src_pos0, src_pos0 + n_src); a position below valid_lo (a node that never existed)
// contributes nothing, and an output row whose own position is below valid_lo is zero.  Same grid as above.
// act_out != nullptr: the source rows are src * act'(act_out) (dL/dbelief and the beliefs of a window: dz2 is formed on the
// fly instead of being written and re-read)
__global__ void __launch_bounds__(256) k_shift_sum(const float* __restrict__ src, int src_pos0, int n_src, int valid_lo,
                                                   const HopList hops, int sign, float* __restrict__ out, int out_pos0,
                                                   int B, int H, int tiled, const float* __restrict__ act_out, int act) {
  const int H4 = H >> 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;       // (graph, 16-byte chunk): row-major inside a row block
  if (j >= B * H4) return;
  const int i = blockIdx.y;
  const int pos = out_pos0 + i;
  float4 self = make_float4(0.f, 0.f, 0.f, 0.f), sum = self;
  if (pos >= valid_lo) {
    const float4* s4 = reinterpret_cast<const float4*>(src) + j;
    const float4* a4 = reinterpret_cast<const float4*>(act_out) + j;
    const size_t row = (size_t)B * H4;
    auto fetch = [&](int jq) {
      float4 v = __ldg(s4 + (size_t)jq * row);
      if (act_out) {
        const float4 o = __ldg(a4 + (size_t)jq * row);
        v.x *= gcm_act_grad(o.x, act);
        v.y *= gcm_act_grad(o.y, act);
        v.z *= gcm_act_grad(o.z, act);
        v.w *= gcm_act_grad(o.w, act);
      }
      return v;
    };
    const int j0 = pos - src_pos0;
    if (j0 >= 0 && j0 < n_src) self = fetch(j0);
    for (int k = 0; k < hops.n; ++k) {
      const int q = pos + sign * hops.h[k];
      const int jq = q - src_pos0;
      if (q >= valid_lo && jq >= 0 && jq < n_src) {
        const float4 v = fetch(jq);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
    }
  }
  const int b = j / H4, c = j - b * H4;
  if (tiled) {
    const long long r = (long long)i * B + b;
    __stcs(tiled_slot(out, r, c), sum);
    __stcs(tiled_slot(out, r, H4 + c), self);
    return;
  }
  float4* o = reinterpret_cast<float4*>(out + ((size_t)i * B + b) * 2 * H) + c;
  __stcs(o, sum);
  __stcs(o + H4, self);
}

// The same for sign = +1, H = 32, tiled output, hops <= 4: a thread owns one 16-byte chunk of one graph and WALKS the
// output positions downwards, keeping the source rows pos .. pos + 4 in registers, so every source row is read once (the
// row-parallel kernel above re-reads each row once per hop through L2: 0.59 ms for a cfg2 window, this one 0.4)
// SM: how the source is addressed.  0: [n_src, B, 32] contiguous.  1: element strides (s_t, s_b, 1) with 16-byte aligned rows
// (a [B, T, 32] gradient read in place: no transposing copy).  2: any strides (s_t, s_b, s_h), e.g. all zero for the
// broadcast gradient of a sum loss, which is then never materialised.  act_out is always contiguous.
template <int SM>
__global__ void __launch_bounds__(256) k_shift_sum_walk(const float* __restrict__ src, int src_pos0, int n_src, int valid_lo,
                                                        unsigned hop_mask, float* __restrict__ out, int out_pos0, int n_out,
                                                        int B, const float* __restrict__ act_out, int act, long long s_t,
                                                        long long s_b, long long s_h) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;       // (graph, 16-byte chunk)
  if (j >= B * 8) return;
  const float4* s4 = reinterpret_cast<const float4*>(src) + j;
  const float4* a4 = reinterpret_cast<const float4*>(act_out) + j;
  const size_t row = (size_t)B * 8;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* sp = src + (long long)(j >> 3) * s_b + (long long)((j & 7) * 4) * s_h;      // SM != 0
  auto fetch = [&](int q) {
    const int jq = q - src_pos0;
    if (q < valid_lo || jq < 0 || jq >= n_src) return zero;
    float4 v;
    if (SM == 0) {
      v = __ldcs(s4 + (size_t)jq * row);
    } else if (SM == 1) {
      v = __ldcs(reinterpret_cast<const float4*>(sp + (long long)jq * s_t));
    } else {
      const float* p = sp + (long long)jq * s_t;
      v = make_float4(p[0], p[s_h], p[2 * s_h], p[3 * s_h]);
    }
    if (act_out) {
      const float4 o = __ldcs(a4 + (size_t)jq * row);
      v.x *= gcm_act_grad(o.x, act);
      v.y *= gcm_act_grad(o.y, act);
      v.z *= gcm_act_grad(o.z, act);
      v.w *= gcm_act_grad(o.w, act);
    }
    return v;
  };
  const int b = j >> 3, c = j & 7;
  const int top = out_pos0 + n_out - 1;
  float4 r1 = fetch(top + 1), r2 = fetch(top + 2), r3 = fetch(top + 3), r4 = fetch(top + 4);
  float4 nxt = fetch(top);
#pragma unroll 2
  for (int i = n_out - 1; i >= 0; --i) {
    const int pos = out_pos0 + i;
    const float4 r0 = nxt;
    if (i > 0) nxt = fetch(pos - 1);
    float4 sum = zero;
    if (hop_mask & 2u) { sum.x += r1.x; sum.y += r1.y; sum.z += r1.z; sum.w += r1.w; }
    if (hop_mask & 4u) { sum.x += r2.x; sum.y += r2.y; sum.z += r2.z; sum.w += r2.w; }
    if (hop_mask & 8u) { sum.x += r3.x; sum.y += r3.y; sum.z += r3.z; sum.w += r3.w; }
    if (hop_mask & 16u) { sum.x += r4.x; sum.y += r4.y; sum.z += r4.z; sum.w += r4.w; }
    const long long r = (long long)i * B + b;
    // a position below valid_lo is a node that never existed: its own row is zero (fetch returned zero for r0 already)
    __stcs(tiled_slot(out, r, c), pos >= valid_lo ? sum : zero);
    __stcs(tiled_slot(out, r, 8 + c), r0);
    r4 = r3; r3 = r2; r2 = r1; r1 = r0;
  }
}

bool hop_list(const int32_t* hops, int n_hops, HopList& hl) {
  if (!hops || n_hops < 1 || n_hops > GCM_MAX_HOPS) return false;
  hl.n = n_hops;
  for (int i = 0; i < n_hops; ++i) {
    if (hops[i] < 1) return false;
    hl.h[i] = hops[i];
  }
  return true;
}

}  // namespace

extern "C" int gcm_temporal_gather(const gcm_dense_state* st, const int32_t* hops, int n_hops, long long p0, int n_rows,
                                   float* out, int tiled, void* stream) {
  GCM_REQUIRE(st && st->nodes && out && n_rows >= 0 && st->F >= 4 && (st->F & 3) == 0 && st->C >= 1,
              "temporal_gather: bad arguments (F must be a multiple of 4)");
  GCM_REQUIRE(((reinterpret_cast<uintptr_t>(st->nodes) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "temporal_gather: pointers must be 16-byte aligned");
  HopList hl;
  GCM_REQUIRE(hop_list(hops, n_hops, hl), "temporal_gather: bad hop list");
  GCM_REQUIRE(!tiled || st->F == 32, "temporal_gather: the tiled layout needs F = 32");
  GCM_REQUIRE(p0 > -(1ll << 30) && p0 + n_rows < (1ll << 31) && n_rows <= 65535 && (long long)st->B * (st->F >> 2) < (1ll << 31),
              "temporal_gather: position / size out of range");
  if (n_rows == 0 || st->B == 0) return GCM_OK;
  const dim3 grid((unsigned)(((long long)st->B * (st->F >> 2) + 255) / 256), (unsigned)n_rows);
  k_temporal_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(*st, hl, (int)p0, out, tiled);
  return gcm_check_launch("k_temporal_gather");
}

static int shift_sum_impl(const float* src, long long s_t, long long s_b, long long s_h, bool strided, long long src_pos0,
                          int n_src, long long valid_lo,
                                      const int32_t* hops, int n_hops, int sign, float* out, long long out_pos0,
                                      int n_out, int B, int H, int tiled, const float* act_out, int act, void* stream) {
  GCM_REQUIRE(src && out && n_src >= 0 && n_out >= 0 && B >= 0 && H >= 4 && (H & 3) == 0 && (sign == 1 || sign == -1),
              "temporal_shift_sum: bad arguments (H must be a multiple of 4)");
  const bool scalar_src = strided && !(s_h == 1 && ((s_t | s_b) & 3) == 0);
  GCM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & (scalar_src ? 3 : 15)) == 0,
              "temporal_shift_sum: pointers must be 16-byte aligned");
  HopList hl;
  GCM_REQUIRE(hop_list(hops, n_hops, hl), "temporal_shift_sum: bad hop list");
  GCM_REQUIRE(!tiled || H == 32, "temporal_shift_sum: the tiled layout needs H = 32");
  GCM_REQUIRE((reinterpret_cast<uintptr_t>(act_out) & 15) == 0 &&
                  (!act_out || act == GCM_ACT_NONE || act == GCM_ACT_TANH || act == GCM_ACT_RELU),
              "temporal_shift_sum: bad activation operand");
  GCM_REQUIRE(src_pos0 > -(1ll << 30) && out_pos0 > -(1ll << 30) && src_pos0 + n_src < (1ll << 31) &&
                  out_pos0 + n_out < (1ll << 31) && valid_lo > -(1ll << 31) && valid_lo < (1ll << 31) && n_out <= 65535 &&
                  (long long)B * (H >> 2) < (1ll << 31),
              "temporal_shift_sum: position / size out of range");
  if (n_out == 0 || B == 0) return GCM_OK;
  int max_hop = 0;
  unsigned hop_mask = 0;
  for (int i = 0; i < n_hops; ++i) {
    max_hop = hops[i] > max_hop ? hops[i] : max_hop;
    if (hops[i] <= 4) hop_mask |= 1u << hops[i];
  }
  static const bool no_walk = getenv("GCM_B200_NO_SHIFT_WALK") != nullptr;     // A/B switch
  if (tiled && sign == 1 && H == 32 && max_hop <= 4 && n_out >= 8 && !no_walk) {
    const unsigned wg = (unsigned)(((long long)B * 8 + 255) / 256);
    if (!strided)
      k_shift_sum_walk<0><<<wg, 256, 0, (cudaStream_t)stream>>>(src, (int)src_pos0, n_src, (int)valid_lo, hop_mask, out,
                                                                (int)out_pos0, n_out, B, act_out, act, 0, 0, 0);
    else if (s_h == 1 && ((s_t | s_b) & 3) == 0)
      k_shift_sum_walk<1><<<wg, 256, 0, (cudaStream_t)stream>>>(src, (int)src_pos0, n_src, (int)valid_lo, hop_mask, out,
                                                                (int)out_pos0, n_out, B, act_out, act, s_t, s_b, s_h);
    else
      k_shift_sum_walk<2><<<wg, 256, 0, (cudaStream_t)stream>>>(src, (int)src_pos0, n_src, (int)valid_lo, hop_mask, out,
                                                                (int)out_pos0, n_out, B, act_out, act, s_t, s_b, s_h);
    return gcm_check_launch("k_shift_sum_walk");
  }
  if (strided) return GCM_ERR_UNSUPPORTED;      // only the walking kernel reads through strides
  const dim3 grid((unsigned)(((long long)B * (H >> 2) + 255) / 256), (unsigned)n_out);
  k_shift_sum<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (int)src_pos0, n_src, (int)valid_lo, hl, sign, out,
                                                      (int)out_pos0, B, H, tiled, act_out, act);
  return gcm_check_launch("k_shift_sum");
}

extern "C" int gcm_temporal_shift_sum(const float* src, long long src_pos0, int n_src, long long valid_lo,
                                      const int32_t* hops, int n_hops, int sign, float* out, long long out_pos0,
                                      int n_out, int B, int H, int tiled, const float* act_out, int act, void* stream) {
  return shift_sum_impl(src, 0, 0, 0, false, src_pos0, n_src, valid_lo, hops, n_hops, sign, out, out_pos0, n_out, B, H, tiled,
                        act_out, act, stream);
}

extern "C" int gcm_temporal_shift_sum_strided(const float* src, long long s_t, long long s_b, long long s_h,
                                              long long src_pos0, int n_src, long long valid_lo, const int32_t* hops,
                                              int n_hops, int sign, float* out, long long out_pos0, int n_out, int B, int H,
                                              int tiled, const float* act_out, int act, void* stream) {
  return shift_sum_impl(src, s_t, s_b, s_h, true, src_pos0, n_src, valid_lo, hops, n_hops, sign, out, out_pos0, n_out, B, H,
                        tiled, act_out, act, stream);
}

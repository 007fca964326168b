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

namespace {

struct HopList {
  int n;
  int h[GCM_MAX_HOPS];
};

// out[i, b, 0:F] = sum_{s: p-s >= 0} x[b, slot(p-s), :],  out[i, b, F:2F] = x[b, slot(p), :],  p = p0 + i (zero row if p < 0)
__global__ void __launch_bounds__(256) k_temporal_gather(const gcm_dense_state st, const HopList hops, long long p0,
                                                         int n_rows, float* __restrict__ out) {
  const int F4 = st.F >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_rows * st.B * F4;
  if (idx >= total) return;
  const int c = (int)(idx % F4);
  const long long ib = idx / F4;
  const int b = (int)(ib % st.B);
  const long long p = p0 + ib / st.B;
  float4 self = make_float4(0.f, 0.f, 0.f, 0.f), sum = self;
  if (p >= 0) {
    const float4* rows = reinterpret_cast<const float4*>(st.nodes + (size_t)b * st.C * st.F) + c;
    self = __ldg(rows + (size_t)(p % st.C) * F4);
    for (int j = 0; j < hops.n; ++j) {
      const long long q = p - hops.h[j];
      if (q >= 0) {
        const float4 v = __ldg(rows + (size_t)(q % st.C) * F4);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(out + (size_t)ib * 2 * st.F) + c;
  __stcs(o, sum);
  __stcs(o + F4, self);
}

// out[i, b, 0:H] = sum_s src[pos + sign * s, b, :],  out[i, b, H:2H] = src[pos, b, :],  pos = out_pos0 + i;
// src rows cover positions [src_pos0, src_pos0 + n_src); a position below valid_lo (a node that never existed)
// contributes nothing, and an output row whose own position is below valid_lo is zero.
__global__ void __launch_bounds__(256) k_shift_sum(const float* __restrict__ src, long long src_pos0, int n_src,
                                                   long long valid_lo, const HopList hops, int sign,
                                                   float* __restrict__ out, long long out_pos0, int n_out, int B, int H) {
  const int H4 = H >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_out * B * H4;
  if (idx >= total) return;
  const int c = (int)(idx % H4);
  const long long ib = idx / H4;
  const int b = (int)(ib % B);
  const long long pos = out_pos0 + ib / B;
  float4 self = make_float4(0.f, 0.f, 0.f, 0.f), sum = self;
  if (pos >= valid_lo) {
    const float4* s4 = reinterpret_cast<const float4*>(src) + (size_t)b * H4 + c;
    const size_t row = (size_t)B * H4;
    const long long j0 = pos - src_pos0;
    if (j0 >= 0 && j0 < n_src) self = __ldg(s4 + (size_t)j0 * row);
    for (int j = 0; j < hops.n; ++j) {
      const long long q = pos + (long long)sign * hops.h[j];
      const long long jq = q - src_pos0;
      if (q >= valid_lo && jq >= 0 && jq < n_src) {
        const float4 v = __ldg(s4 + (size_t)jq * row);
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(out + (size_t)ib * 2 * H) + c;
  __stcs(o, sum);
  __stcs(o + H4, self);
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
                                   float* out, void* stream) {
  GCM_REQUIRE(st && st->nodes && out && n_rows >= 0 && st->F >= 4 && (st->F & 3) == 0 && st->C >= 1,
              "temporal_gather: bad arguments (F must be a multiple of 4)");
  GCM_REQUIRE(((reinterpret_cast<uintptr_t>(st->nodes) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "temporal_gather: pointers must be 16-byte aligned");
  HopList hl;
  GCM_REQUIRE(hop_list(hops, n_hops, hl), "temporal_gather: bad hop list");
  const long long total = (long long)n_rows * st->B * (st->F >> 2);
  if (total == 0) return GCM_OK;
  k_temporal_gather<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*st, hl, p0, n_rows, out);
  return gcm_check_launch("k_temporal_gather");
}

extern "C" int gcm_temporal_shift_sum(const float* src, long long src_pos0, int n_src, long long valid_lo,
                                      const int32_t* hops, int n_hops, int sign, float* out, long long out_pos0,
                                      int n_out, int B, int H, void* stream) {
  GCM_REQUIRE(src && out && n_src >= 0 && n_out >= 0 && B >= 0 && H >= 4 && (H & 3) == 0 && (sign == 1 || sign == -1),
              "temporal_shift_sum: bad arguments (H must be a multiple of 4)");
  GCM_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "temporal_shift_sum: pointers must be 16-byte aligned");
  HopList hl;
  GCM_REQUIRE(hop_list(hops, n_hops, hl), "temporal_shift_sum: bad hop list");
  const long long total = (long long)n_out * B * (H >> 2);
  if (total == 0) return GCM_OK;
  k_shift_sum<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, src_pos0, n_src, valid_lo, hl, sign,
                                                                                out, out_pos0, n_out, B, H);
  return gcm_check_launch("k_shift_sum");
}

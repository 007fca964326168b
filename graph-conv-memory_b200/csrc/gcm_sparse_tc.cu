// GraphConv forward with the per-tile product on the tensor cores (tcgen05 + TMEM, 3xTF32), sm_100a.
//
// torch_geometric.nn.GraphConv (call sites ray_sparse_gcm.py:37-40, invoked at sparse_gcm.py:178,199):
//   out_i = act(W_rel (sum_{(j->i) in E} w_ji x_j) + b + W_root x_i)
// k_graphconv_fwd (gcm_sparse.cu) gathers a 64-row tile into shared memory and multiplies it by the weight pack on the
// CUDA cores; ncu had it issue-bound with that [64 x 2Fin] x [2Fin x Fout] product as half of its instructions
// (profiles/c4_c5_kernels_r1.md).  Here the product leaves the instruction stream:
//   warps 0-23  gather: warp per row, a lane owns Fin/32 contiguous features (one coalesced warp load per neighbour row),
//               32-bit column indices fetched 32 at a time and broadcast by shuffle, 8 gathers in flight per warp, summed
//               in edge order (the SAME order as the CUDA-core kernel), then write [agg | x_i] of the row, split hi / lo
//               (3xTF32), into the
//               A tile [128 x 2Fin] in shared memory in the canonical K-major core-matrix layout, with the leading-byte
//               offset padded to 144 B so that the per-row stores do not pile onto four banks
//   warp 28     one lane issues, per 128-row tile, lo*Whi + hi*Wlo + hi*Whi as SS-form tcgen05.mma.kind::tf32 (A and the
//               weight pack both in shared memory) into one of two accumulators in TMEM
//   warps 24-27 epilogue of the PREVIOUS tile (tcgen05.ld, bias, activation, row stores) while the gather warps already
//               fill the A tile of the next one (the A tile is free as soon as the tile's MMAs have completed)
// Persistent: one CTA per SM walks the tiles.  fp32-accurate (tests/test_sparse_gpu.py compares with the CUDA-core kernel).
#include <stdlib.h>

#include "gcm_common.cuh"
#include "gcm_tc.cuh"

namespace {

constexpr int GT_TM = 128;                 // rows per tile (M of the MMA)
constexpr int GT_GATHER_WARPS = 24;             // memory-level parallelism of the gathers: 8 warps per SM ran at half the speed of the CUDA-core kernel (5 CTAs x 8 warps)
constexpr int GT_THREADS = (GT_GATHER_WARPS + 4 + 1) * 32;
constexpr int GT_LBO = 144;                // bytes between core matrices along K in the A tile (128 + 16 padding)

struct GcTcArgs {
  const float* x;
  const int64_t* rowptr;
  const int64_t* col;
  const float* ew;
  const int64_t* rows;
  int64_t m;
  int Fin, Fout;
  const float* wt;      // [2 Fin, Fout] K-major pack: wt[k * Fout + n]
  const float* bias;
  int act;
  float* agg_out;
  float* out;
  int64_t tiles;
  const float* agg_in;  // two-pass form: the sums were made by k_csr_gather, the row warps only load [agg_in | x] rows
};

// Warp per row, a lane owns V = Fin / 32 contiguous features (one coalesced warp load per neighbour row).  Column indices
// come 32 at a time (one coalesced load) and are narrowed to 32 bits: the shuffle that broadcasts one is a single
// instruction and the row address is ONE IMAD.WIDE (the int64 version spent 2 shuffles + an emulated 64-bit multiply per
// gather; ncu: 1.8 G warp instructions per layer at cfg5, 22 per edge).  Every batch of 8 gathers is issued in full with
// out-of-range slots predicated off, so there is no serial tail.  Sums run in edge order, like the CUDA-core kernel's.
// (Tried and measured slower: half-warp per row, two rows per warp, float4 per lane: 5.4 ms per layer against 4.8.)
template <int V>
__device__ __forceinline__ void gt_ld(const float* p, float (&v)[V]) {
  if (V == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1 % V] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}

template <int V, bool OWN = true>
__device__ __forceinline__ void gt_gather(const GcTcArgs& a, int64_t i, int64_t e0, int64_t e1, int lane, float (&acc)[V],
                                          float (&own)[V]) {
  constexpr int Fin = 32 * V;
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.0f;
  const float* xl = a.x + lane * V;
  if (OWN) gt_ld<V>(xl + i * Fin, own);
  for (int64_t base = e0; base < e1; base += 32) {
    const int cnt = (int)min((int64_t)32, e1 - base);
    const unsigned my = lane < cnt ? (unsigned)a.col[base + lane] : 0u;     // host: every column * Fin fits 32 bits
    const float myw = (a.ew && lane < cnt) ? a.ew[base + lane] : 1.0f;
    int u = 0;
    for (; u + 8 <= cnt; u += 8) {
      float v[8][V];
#pragma unroll
      for (int q = 0; q < 8; ++q) gt_ld<V>(xl + (size_t)(__shfl_sync(GCM_FULL_MASK, my, u + q) * (unsigned)Fin), v[q]);
      if (a.ew) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float w = __shfl_sync(GCM_FULL_MASK, myw, u + q);
#pragma unroll
          for (int j = 0; j < V; ++j) v[q][j] *= w;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[q][j];
    }
    if (u < cnt) {
      // the last cnt % 8 edges of the batch: all loads first (no-ops beyond cnt), then the sums in edge order
      float v[7][V];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[q][j] = 0.0f;
        const unsigned c = __shfl_sync(GCM_FULL_MASK, my, (u + q) & 31);
        if (u + q < cnt) gt_ld<V>(xl + (size_t)(c * (unsigned)Fin), v[q]);
      }
      if (a.ew) {
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const float w = __shfl_sync(GCM_FULL_MASK, myw, (u + q) & 31);
#pragma unroll
          for (int j = 0; j < V; ++j) v[q][j] *= w;
        }
      }
#pragma unroll
      for (int q = 0; q < 7; ++q)
        if (u + q < cnt) {
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += v[q][j];
        }
    }
  }
}

// Pass 1 of the two-pass form: agg_i = sum_{(j->i)} w_ji x_j alone, a warp per row with nothing else on the SM -- no A
// tile, no weight pack, no TMEM -- so 48 warps per SM keep gathers in flight instead of the fused kernel's 24 (whose
// 214 KB of shared memory also leave the gathers ~30 KB of L1).  ncu (profiles/c5_graphconv_r2.md) has this pass bound by
// the L1 data pipe (l1tex__data_pipe_lsu_wavefronts 77 %): a 256-byte row costs 4 wavefronts whichever load width brings
// it, and every SHFL that broadcasts a column index is one more wavefront on the same pipe.  Measured and rejected here:
// LDG.128 with two rows per warp instruction (2.80 ms per layer at cfg5 against 2.22), broadcasting the index with a warp
// reduction instead of the shuffle (REDUX of `lane == src ? index : 0`: 2.76 against 2.44 for the same loop shape), one
// loop of predicated 8-slot batches instead of full batches + a tail (2.44 against 2.22), two scalar loads of 128
// contiguous bytes per warp instead of one 8-byte load per lane (2.27 against 2.22: the L1 returns ~64 B per clock to
// global loads whatever their shape -- 21.6 GB of rows per layer at cfg5 = 1.16 ms at best for this design).  Also built,
// measured and removed: the source rows staged in SHARED memory (256 sink rows of one graph per CTA, the graph's rows
// walked in double-buffered tiles of 384 rows brought by bulk copies, per-row cursors over the ascending column lists,
// index runs read back by uniform 16-byte loads): bit-identical sums, but 4.4 ms per layer -- at ~20 edges per row a
// 96 KB tile serves ~25 edges per warp, so a CTA spends its time waiting for tiles (one CTA of 200 KB per SM).  Same loads and the same sums in
// the same order as the fused kernel: bit-identical aggregation.
template <int V>
__global__ void __launch_bounds__(256, 6) k_csr_gather(const GcTcArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t li = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (li >= a.m) return;
  const int64_t i = a.rows ? a.rows[li] : li;
  float acc[V], own[V];
  gt_gather<V, false>(a, i, a.rowptr[i], a.rowptr[i + 1], lane, acc, own);
  float* o = a.agg_out + li * (32 * V) + lane * V;
  if (V == 2) *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1 % V]);
  else o[0] = acc[0];
}

// float offset of element (row r, k) of the padded K-major A tile with K columns
__device__ __forceinline__ int gt_a_off(int r, int k, int K) {
  return (r >> 3) * ((K >> 2) * (GT_LBO / 4)) + (k >> 2) * (GT_LBO / 4) + (r & 7) * 4 + (k & 3);
}

// STREAM: pass 2 of the two-pass form (a.agg_in set): the row warps only load [agg_in | x] rows, and they load the rows of
// tile it + 1 into registers right after handing tile it to the MMA warp, so the loads fly during the MMAs (0.89 ms per
// layer at cfg5 = 3.6 TB/s).  Measured and rejected: tiles of 64 rows with TWO A tiles in shared memory (the split / store
// of tile t + 1 under the MMAs of tile t; 16 row warps, rows of one or two tiles ahead in registers): 1.06 / 1.01 ms.
template <int V, bool STREAM>
__global__ void __launch_bounds__(GT_THREADS, 1) k_graphconv_fwd_tc(const GcTcArgs a) {
  constexpr int Fin = 32 * V, K = 2 * Fin;
  extern __shared__ __align__(128) unsigned char gt_smem[];
  const int Fout = a.Fout;
  const int a_floats = (GT_TM / 8) * (K / 4) * (GT_LBO / 4);
  float* Ahi = reinterpret_cast<float*>(gt_smem);
  float* Alo = Ahi + a_floats;
  float* Whi = Alo + a_floats;                       // [Fout x K] canonical K-major (unpadded)
  float* Wlo = Whi + (size_t)Fout * K;
  float* bias_s = Wlo + (size_t)Fout * K;            // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);
  uint64_t* a_full = bars;            // gather warps -> MMA warp (one arrival per gather warp)
  uint64_t* mma_done = bars + 1;      // [2]  MMA warp -> epilogue warps and gather warps (A tile free)
  uint64_t* d_free = bars + 3;        // [2]  epilogue warps -> MMA warp (accumulator free)
  uint64_t* w_ready = bars + 5;       // weights staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t d_cols = Fout <= 32 ? 32u : (Fout <= 64 ? 64u : 128u);
  const uint32_t tmem_cols = 2 * d_cols < 32 ? 32u : 2 * d_cols;

  if (tid == 0) {
    tc::mbar_init(a_full, GT_GATHER_WARPS);
    tc::mbar_init(mma_done + 0, 1);
    tc::mbar_init(mma_done + 1, 1);
    tc::mbar_init(d_free + 0, 128);
    tc::mbar_init(d_free + 1, 128);
    tc::mbar_init(w_ready, 5 * 32);
    tc::mbar_fence_init();
  }
  if (warp == GT_GATHER_WARPS + 4) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const int64_t n_it = (a.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp < GT_GATHER_WARPS) {
    // =============================== gather ===============================
    if (STREAM) {
      constexpr int RP = (GT_TM + GT_GATHER_WARPS - 1) / GT_GATHER_WARPS;     // rows of a tile per warp
      float pa[RP][V], po[RP][V];
      auto load_tile = [&](int64_t it) {
        const int64_t row0 = (blockIdx.x + it * gridDim.x) * GT_TM;
#pragma unroll
        for (int k = 0; k < RP; ++k) {
          const int r = warp + k * GT_GATHER_WARPS;
          const int64_t li = row0 + r;
#pragma unroll
          for (int j = 0; j < V; ++j) pa[k][j] = po[k][j] = 0.0f;
          if (r < GT_TM && li < a.m) {
            const int64_t i = a.rows ? a.rows[li] : li;
            gt_ld<V>(a.agg_in + li * Fin + lane * V, pa[k]);
            gt_ld<V>(a.x + i * Fin + lane * V, po[k]);
          }
        }
      };
      if (n_it > 0) load_tile(0);
      for (int64_t it = 0; it < n_it; ++it) {
        if (it > 0) tc::mbar_wait(mma_done + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
#pragma unroll
        for (int k = 0; k < RP; ++k) {
          const int r = warp + k * GT_GATHER_WARPS;
          if (r < GT_TM) {
            uint32_t ah[V], al[V], oh[V], ol[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
              tc::split_tf32(pa[k][j], ah[j], al[j]);
              tc::split_tf32(po[k][j], oh[j], ol[j]);
            }
            const int o_agg = gt_a_off(r, lane * V, K), o_own = gt_a_off(r, Fin + lane * V, K);
            if (V == 2) {
              *reinterpret_cast<uint2*>(Ahi + o_agg) = make_uint2(ah[0], ah[1 % V]);
              *reinterpret_cast<uint2*>(Alo + o_agg) = make_uint2(al[0], al[1 % V]);
              *reinterpret_cast<uint2*>(Ahi + o_own) = make_uint2(oh[0], oh[1 % V]);
              *reinterpret_cast<uint2*>(Alo + o_own) = make_uint2(ol[0], ol[1 % V]);
            } else {
              Ahi[o_agg] = __uint_as_float(ah[0]); Alo[o_agg] = __uint_as_float(al[0]);
              Ahi[o_own] = __uint_as_float(oh[0]); Alo[o_own] = __uint_as_float(ol[0]);
            }
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(a_full);
        if (it + 1 < n_it) load_tile(it + 1);
      }
    } else
    for (int64_t it = 0; it < n_it; ++it) {
      const int64_t row0 = (blockIdx.x + it * gridDim.x) * GT_TM;
      if (it > 0) {
        // the MMAs that read the A tile of the previous tile have completed
        tc::mbar_wait(mma_done + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
      }
      for (int r = warp; r < GT_TM; r += GT_GATHER_WARPS) {
        const int64_t li = row0 + r;
        float acc[V], own[V];
        if (li < a.m) {
          const int64_t i = a.rows ? a.rows[li] : li;
          gt_gather<V>(a, i, a.rowptr[i], a.rowptr[i + 1], lane, acc, own);
          if (a.agg_out) {
#pragma unroll
            for (int j = 0; j < V; ++j) a.agg_out[li * Fin + lane * V + j] = acc[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] = own[j] = 0.0f;
        }
        uint32_t ah[V], al[V], oh[V], ol[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          tc::split_tf32(acc[j], ah[j], al[j]);
          tc::split_tf32(own[j], oh[j], ol[j]);
        }
        const int o_agg = gt_a_off(r, lane * V, K), o_own = gt_a_off(r, Fin + lane * V, K);
        if (V == 2) {
          *reinterpret_cast<uint2*>(Ahi + o_agg) = make_uint2(ah[0], ah[1 % V]);
          *reinterpret_cast<uint2*>(Alo + o_agg) = make_uint2(al[0], al[1 % V]);
          *reinterpret_cast<uint2*>(Ahi + o_own) = make_uint2(oh[0], oh[1 % V]);
          *reinterpret_cast<uint2*>(Alo + o_own) = make_uint2(ol[0], ol[1 % V]);
        } else {
          Ahi[o_agg] = __uint_as_float(ah[0]); Alo[o_agg] = __uint_as_float(al[0]);
          Ahi[o_own] = __uint_as_float(oh[0]); Alo[o_own] = __uint_as_float(ol[0]);
        }
      }
      tc::fence_proxy_async();          // the tensor core reads the tile through the async proxy
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a_full);
    }
  } else {
    // ---- weight pack -> canonical K-major B operand, split hi / lo (warps 8-12, once) ----
    const int t5 = tid - GT_GATHER_WARPS * 32;
    for (int i = t5; i < Fout * K; i += 5 * 32) {
      const int k = i / Fout, n = i - k * Fout;            // wt[k * Fout + n]: coalesced reads
      uint32_t hi, lo;
      tc::split_tf32(__ldg(a.wt + i), hi, lo);
      Whi[tc::kmajor_off(n, k, K)] = __uint_as_float(hi);
      Wlo[tc::kmajor_off(n, k, K)] = __uint_as_float(lo);
    }
    for (int i = t5; i < 128; i += 5 * 32) bias_s[i] = (a.bias && i < Fout) ? __ldg(a.bias + i) : 0.0f;
    tc::fence_proxy_async();
    tc::mbar_arrive(w_ready);
    tc::mbar_wait(w_ready, 0);

    if (warp == GT_GATHER_WARPS + 4) {
      // =============================== MMA issue ===============================
      if (lane == 0) {
        const uint32_t idesc = tc::idesc_tf32(GT_TM, Fout);
        const uint32_t a_sbo = (uint32_t)(K / 4) * GT_LBO, b_sbo = (uint32_t)(K / 4) * 128u;
        const uint32_t ahi = tc::smem_u32(Ahi), alo = tc::smem_u32(Alo), whi = tc::smem_u32(Whi), wlo = tc::smem_u32(Wlo);
        for (int64_t it = 0; it < n_it; ++it) {
          const int buf = (int)(it & 1);
          tc::mbar_wait(a_full, (uint32_t)(it & 1));
          tc::mbar_wait(d_free + buf, (uint32_t)(((it >> 1) & 1) ^ 1));     // the epilogue of tile it - 2 has read it
          tc::fence_after_sync();
          const uint32_t d = tbase + buf * d_cols;
          bool accum = false;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {           // lo*Whi, hi*Wlo, hi*Whi
            const uint32_t asm_ = pass == 0 ? alo : ahi;
            const uint32_t bsm = pass == 1 ? wlo : whi;
            for (int ks = 0; ks < K / 8; ++ks) {
              tc::mma_tf32_ss(d, tc::smem_desc_kmajor(asm_ + ks * 2 * GT_LBO, GT_LBO, a_sbo),
                              tc::smem_desc_kmajor(bsm + ks * 256, 128, b_sbo), idesc, accum);
              accum = true;
            }
          }
          tc::mma_commit(mma_done + buf);
        }
      }
    } else {
      // =============================== epilogue ===============================
      const int q = warp - GT_GATHER_WARPS;                 // TMEM lane quadrant
      const int r = q * 32 + lane;
      const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
      const int act = a.act;
      for (int64_t it = 0; it < n_it; ++it) {
        const int buf = (int)(it & 1);
        const int64_t li = (blockIdx.x + it * gridDim.x) * GT_TM + r;
        tc::mbar_wait(mma_done + buf, (uint32_t)((it >> 1) & 1));
        tc::fence_after_sync();
        for (int n0 = 0; n0 < Fout; n0 += 16) {
          uint32_t dv[16];
          tc::tmem_ld16(lane_addr + buf * d_cols + n0, dv);
          tc::wait_ld();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(dv[j]) + bias_s[n0 + j];
          gcm_act_fast_vec(f, act);
          if (li < a.m) {
            float4* o = reinterpret_cast<float4*>(a.out + li * Fout + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
        tc::fence_before_sync();
        tc::mbar_arrive(d_free + buf);
      }
    }
  }
  __syncthreads();
  if (warp == GT_GATHER_WARPS + 4) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tbase, tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Block-local variant: every edge of GCM's flat graph stays inside its own graph (sparse_gcm.py builds the edges per
// batch element and only offsets them, util.flatten_adj), whose rows are contiguous in x.  A CTA therefore takes whole
// graphs: one TMA bulk copy brings the graph's rows (<= 64 KB) into shared memory, and the gathers of its 64-row tiles
// read shared memory instead of L1 / L2 (the kernel above waits on ~20 dependent global loads per row: long-scoreboard
// stalls 15 per issue, profiles/c5_graphconv_r2.md).  Tiles are 64 rows (M = 64 MMAs) so that block + A tile + weight
// pack fit 227 KB.  Same sums in the same order, same product.
constexpr int GB_TM = 64;
constexpr int GB_XBYTES = 64 * 1024;
constexpr int GB_THREADS = (GT_GATHER_WARPS + 4 + 2) * 32;     // + MMA warp + loader warp

struct GcBlkArgs {
  GcTcArgs g;
  const int64_t* node_off;   // [n_graphs] first row of every graph
  int n_graphs;
};

template <int V>
__device__ __forceinline__ void gb_ld(const float* p, float (&v)[V]) {
  if (V == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1 % V] = t.y;
  } else {
    v[0] = *p;
  }
}

// as gt_gather, sources read from the graph's block in shared memory (column index - first row of the graph)
template <int V>
__device__ __forceinline__ void gb_gather(const GcTcArgs& a, const float* xs, int64_t gstart, int lr, int64_t e0, int64_t e1,
                                          int lane, float (&acc)[V], float (&own)[V]) {
  constexpr int Fin = 32 * V;
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.0f;
  const float* xl = xs + lane * V;
  gb_ld<V>(xl + lr * Fin, own);
  for (int64_t base = e0; base < e1; base += 32) {
    const int cnt = (int)min((int64_t)32, e1 - base);
    const unsigned my = lane < cnt ? (unsigned)(a.col[base + lane] - gstart) : 0u;
    const float myw = (a.ew && lane < cnt) ? a.ew[base + lane] : 1.0f;
    int u = 0;
    for (; u + 8 <= cnt; u += 8) {
      float v[8][V];
#pragma unroll
      for (int q = 0; q < 8; ++q) gb_ld<V>(xl + __shfl_sync(GCM_FULL_MASK, my, u + q) * (unsigned)Fin, v[q]);
      if (a.ew) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float w = __shfl_sync(GCM_FULL_MASK, myw, u + q);
#pragma unroll
          for (int j = 0; j < V; ++j) v[q][j] *= w;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[q][j];
    }
    if (u < cnt) {
      float v[7][V];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[q][j] = 0.0f;
        const unsigned c = __shfl_sync(GCM_FULL_MASK, my, (u + q) & 31);
        if (u + q < cnt) gb_ld<V>(xl + c * (unsigned)Fin, v[q]);
      }
      if (a.ew) {
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const float w = __shfl_sync(GCM_FULL_MASK, myw, (u + q) & 31);
#pragma unroll
          for (int j = 0; j < V; ++j) v[q][j] *= w;
        }
      }
#pragma unroll
      for (int q = 0; q < 7; ++q)
        if (u + q < cnt) {
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += v[q][j];
        }
    }
  }
}

template <int V>
__global__ void __launch_bounds__(GB_THREADS, 1) k_graphconv_fwd_blk(const GcBlkArgs b) {
  constexpr int Fin = 32 * V, K = 2 * Fin;
  const GcTcArgs& a = b.g;
  extern __shared__ __align__(128) unsigned char gb_smem[];
  const int Fout = a.Fout;
  const int a_floats = (GB_TM / 8) * (K / 4) * (GT_LBO / 4);
  float* xs = reinterpret_cast<float*>(gb_smem);                  // the graph's rows [<= 64 KB]
  float* Ahi = xs + GB_XBYTES / 4;
  float* Alo = Ahi + a_floats;
  float* Whi = Alo + a_floats;                       // [Fout x K] canonical K-major (unpadded)
  float* Wlo = Whi + (size_t)Fout * K;
  float* bias_s = Wlo + (size_t)Fout * K;            // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);
  uint64_t* a_full = bars;            // gather warps -> MMA warp (one arrival per gather warp)
  uint64_t* mma_done = bars + 1;      // [2]  MMA warp -> epilogue warps and gather warps (A tile free)
  uint64_t* d_free = bars + 3;        // [2]  epilogue warps -> MMA warp (accumulator free)
  uint64_t* w_ready = bars + 5;       // weights staged
  uint64_t* x_full = bars + 6;        // loader -> gather warps (the graph's rows have landed)
  uint64_t* x_free = bars + 7;        // gather warps -> loader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t d_cols = Fout <= 32 ? 32u : (Fout <= 64 ? 64u : 128u);
  const uint32_t tmem_cols = 2 * d_cols < 32 ? 32u : 2 * d_cols;
  constexpr int W_MMA = GT_GATHER_WARPS + 4, W_LOAD = GT_GATHER_WARPS + 5;

  if (tid == 0) {
    tc::mbar_init(a_full, GT_GATHER_WARPS);
    tc::mbar_init(mma_done + 0, 1);
    tc::mbar_init(mma_done + 1, 1);
    tc::mbar_init(d_free + 0, 128);
    tc::mbar_init(d_free + 1, 128);
    tc::mbar_init(w_ready, 5 * 32);
    tc::mbar_init(x_full, 1);
    tc::mbar_init(x_free, GT_GATHER_WARPS);
    tc::mbar_fence_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  auto g_begin = [&](int gi) { return b.node_off[gi]; };
  auto g_end = [&](int gi) { return gi + 1 < b.n_graphs ? b.node_off[gi + 1] : a.m; };

  if (warp == W_LOAD) {
    // =============================== loader: one bulk copy per graph ===============================
    if (lane == 0) {
      int k = 0;
      for (int gi = blockIdx.x; gi < b.n_graphs; gi += gridDim.x, ++k) {
        const int64_t s0 = g_begin(gi);
        const uint32_t bytes = (uint32_t)(g_end(gi) - s0) * (uint32_t)(Fin * 4);
        if (k > 0) tc::mbar_wait(x_free, (uint32_t)((k - 1) & 1));
        if (bytes) {
          tc::mbar_expect_tx(x_full, bytes);
          tc::bulk_g2s(xs, a.x + s0 * Fin, bytes, x_full);
        } else {
          tc::mbar_arrive(x_full);
        }
      }
    }
  } else if (warp < GT_GATHER_WARPS) {
    // =============================== gather ===============================
    int64_t it = 0;
    int k = 0;
    for (int gi = blockIdx.x; gi < b.n_graphs; gi += gridDim.x, ++k) {
      const int64_t s0 = g_begin(gi);
      const int nb = (int)(g_end(gi) - s0);
      tc::mbar_wait(x_full, (uint32_t)(k & 1));
      for (int t0 = 0; t0 < nb; t0 += GB_TM, ++it) {
        if (it > 0) tc::mbar_wait(mma_done + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
        for (int r = warp; r < GB_TM; r += GT_GATHER_WARPS) {
          const int lr = t0 + r;
          float acc[V], own[V];
          if (lr < nb) {
            const int64_t i = s0 + lr;
            gb_gather<V>(a, xs, s0, lr, a.rowptr[i], a.rowptr[i + 1], lane, acc, own);
            if (a.agg_out) {
#pragma unroll
              for (int j = 0; j < V; ++j) a.agg_out[i * Fin + lane * V + j] = acc[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = own[j] = 0.0f;
          }
          uint32_t ah[V], al[V], oh[V], ol[V];
#pragma unroll
          for (int j = 0; j < V; ++j) {
            tc::split_tf32(acc[j], ah[j], al[j]);
            tc::split_tf32(own[j], oh[j], ol[j]);
          }
          const int o_agg = gt_a_off(r, lane * V, K), o_own = gt_a_off(r, Fin + lane * V, K);
          if (V == 2) {
            *reinterpret_cast<uint2*>(Ahi + o_agg) = make_uint2(ah[0], ah[1 % V]);
            *reinterpret_cast<uint2*>(Alo + o_agg) = make_uint2(al[0], al[1 % V]);
            *reinterpret_cast<uint2*>(Ahi + o_own) = make_uint2(oh[0], oh[1 % V]);
            *reinterpret_cast<uint2*>(Alo + o_own) = make_uint2(ol[0], ol[1 % V]);
          } else {
            Ahi[o_agg] = __uint_as_float(ah[0]); Alo[o_agg] = __uint_as_float(al[0]);
            Ahi[o_own] = __uint_as_float(oh[0]); Alo[o_own] = __uint_as_float(ol[0]);
          }
        }
        tc::fence_proxy_async();          // the tensor core reads the tile through the async proxy
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(a_full);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(x_free);     // this warp no longer reads the block
    }
  } else {
    // ---- weight pack -> canonical K-major B operand, split hi / lo (5 warps, once) ----
    const int t5 = tid - GT_GATHER_WARPS * 32;
    for (int i = t5; i < Fout * K; i += 5 * 32) {
      const int k = i / Fout, n = i - k * Fout;            // wt[k * Fout + n]: coalesced reads
      uint32_t hi, lo;
      tc::split_tf32(__ldg(a.wt + i), hi, lo);
      Whi[tc::kmajor_off(n, k, K)] = __uint_as_float(hi);
      Wlo[tc::kmajor_off(n, k, K)] = __uint_as_float(lo);
    }
    for (int i = t5; i < 128; i += 5 * 32) bias_s[i] = (a.bias && i < Fout) ? __ldg(a.bias + i) : 0.0f;
    tc::fence_proxy_async();
    tc::mbar_arrive(w_ready);
    tc::mbar_wait(w_ready, 0);

    if (warp == W_MMA) {
      // =============================== MMA issue ===============================
      if (lane == 0) {
        const uint32_t idesc = tc::idesc_tf32(GB_TM, Fout);
        const uint32_t a_sbo = (uint32_t)(K / 4) * GT_LBO, b_sbo = (uint32_t)(K / 4) * 128u;
        const uint32_t ahi = tc::smem_u32(Ahi), alo = tc::smem_u32(Alo), whi = tc::smem_u32(Whi), wlo = tc::smem_u32(Wlo);
        int64_t it = 0;
        for (int gi = blockIdx.x; gi < b.n_graphs; gi += gridDim.x) {
          const int nb = (int)(g_end(gi) - g_begin(gi));
          for (int t0 = 0; t0 < nb; t0 += GB_TM, ++it) {
            const int buf = (int)(it & 1);
            tc::mbar_wait(a_full, (uint32_t)(it & 1));
            tc::mbar_wait(d_free + buf, (uint32_t)(((it >> 1) & 1) ^ 1));     // the epilogue of tile it - 2 has read it
            tc::fence_after_sync();
            const uint32_t d = tbase + buf * d_cols;
            bool accum = false;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {           // lo*Whi, hi*Wlo, hi*Whi
              const uint32_t asm_ = pass == 0 ? alo : ahi;
              const uint32_t bsm = pass == 1 ? wlo : whi;
              for (int ks = 0; ks < K / 8; ++ks) {
                tc::mma_tf32_ss(d, tc::smem_desc_kmajor(asm_ + ks * 2 * GT_LBO, GT_LBO, a_sbo),
                                tc::smem_desc_kmajor(bsm + ks * 256, 128, b_sbo), idesc, accum);
                accum = true;
              }
            }
            tc::mma_commit(mma_done + buf);
          }
        }
      }
    } else {
      // =============================== epilogue (M = 64: row r sits in lane 32 (r / 16) + r % 16) ===============================
      const int q = warp - GT_GATHER_WARPS;                 // TMEM lane quadrant
      const int r = q * 16 + lane;
      const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
      const int act = a.act;
      int64_t it = 0;
      for (int gi = blockIdx.x; gi < b.n_graphs; gi += gridDim.x) {
        const int64_t s0 = g_begin(gi);
        const int nb = (int)(g_end(gi) - s0);
        for (int t0 = 0; t0 < nb; t0 += GB_TM, ++it) {
          const int buf = (int)(it & 1);
          const bool live = lane < 16 && t0 + r < nb;
          const int64_t li = s0 + t0 + r;
          tc::mbar_wait(mma_done + buf, (uint32_t)((it >> 1) & 1));
          tc::fence_after_sync();
          for (int n0 = 0; n0 < Fout; n0 += 16) {
            uint32_t dv[16];
            tc::tmem_ld16(lane_addr + buf * d_cols + n0, dv);
            tc::wait_ld();
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(dv[j]) + bias_s[n0 + j];
            gcm_act_fast_vec(f, act);
            if (live) {
              float4* o = reinterpret_cast<float4*>(a.out + li * Fout + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
          }
          tc::fence_before_sync();
          tc::mbar_arrive(d_free + buf);
        }
      }
    }
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tbase, tmem_cols);
  }
}

}  // namespace

static int g_graphconv_kernel = GCM_GC_AUTO;
extern "C" int gcm_set_graphconv_kernel(int which) {
  GCM_REQUIRE(which >= GCM_GC_AUTO && which <= GCM_GC_TC, "set_graphconv_kernel: bad variant %d", which);
  g_graphconv_kernel = which;
  return GCM_OK;
}

static const int64_t* g_gc_node_off = nullptr;
static int g_gc_n_graphs = 0, g_gc_max_nodes = 0;
/* Block structure of x for the NEXT gcm_sparse_graphconv_fwd call that evaluates every row: the rows of graph g are
 * node_off[g] .. node_off[g + 1] - 1 (device array of n_graphs first rows; the last graph ends at m), no graph has more
 * than max_nodes rows, and every edge's source lies in the sink's graph.  Enables the block-local kernel. */
extern "C" int gcm_sparse_graphconv_hint_blocks(const int64_t* node_off, int n_graphs, int max_nodes) {
  g_gc_node_off = node_off;
  g_gc_n_graphs = n_graphs;
  g_gc_max_nodes = max_nodes;
  return GCM_OK;
}

static int graphconv_fwd_blk(const GcTcArgs& a, const int64_t* node_off, int n_graphs, cudaStream_t stream) {
  const int K = 2 * a.Fin;
  const size_t a_bytes = (size_t)(GB_TM / 8) * (K / 4) * GT_LBO;
  const size_t smem = GB_XBYTES + 2 * a_bytes + (size_t)2 * a.Fout * K * 4 + 128 * 4 + 128 + 128;
  if (smem > 227 * 1024) return GCM_ERR_UNSUPPORTED;
  GcBlkArgs b{a, node_off, n_graphs};
  long long grid = gcm_num_sms();
  if (grid > n_graphs) grid = n_graphs;
  cudaError_t e;
  if (a.Fin == 64) {
    e = cudaFuncSetAttribute(k_graphconv_fwd_blk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_graphconv_fwd_blk<2><<<(unsigned)grid, GB_THREADS, smem, stream>>>(b);
  } else {
    e = cudaFuncSetAttribute(k_graphconv_fwd_blk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k_graphconv_fwd_blk<1><<<(unsigned)grid, GB_THREADS, smem, stream>>>(b);
  }
  if (e != cudaSuccess) {
    gcm_set_error("cudaFuncSetAttribute(graphconv_fwd_blk): %s", cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  return gcm_check_launch("k_graphconv_fwd_blk");
}

// returns GCM_ERR_UNSUPPORTED (nothing launched) when the shape is not covered: the caller then runs k_graphconv_fwd
int gcm_graphconv_fwd_tc(const float* x, const int64_t* rowptr, const int64_t* col, const float* ew, const int64_t* rows,
                         int64_t m, int Fin, int Fout, const float* wt, const float* bias, int act, float* agg_out,
                         float* out, cudaStream_t stream, int64_t x_rows) {
  // the block hint belongs to THIS call whatever path it takes
  const int64_t* node_off = g_gc_node_off;
  const int n_graphs = g_gc_n_graphs, max_nodes = g_gc_max_nodes;
  g_gc_node_off = nullptr;
  g_gc_n_graphs = g_gc_max_nodes = 0;
  static const bool off = getenv("GCM_B200_GRAPHCONV_CUDA_CORES") != nullptr;     // A/B switch
  if (off || g_graphconv_kernel == GCM_GC_CUDA_CORES) return GCM_ERR_UNSUPPORTED;
  if (!(Fin == 32 || Fin == 64) || Fout < 16 || Fout > 128 || (Fout & 15) != 0) return GCM_ERR_UNSUPPORTED;
  if (m < 4 * GT_TM && g_graphconv_kernel != GCM_GC_TC) return GCM_ERR_UNSUPPORTED;     // small calls: not worth a persistent CTA
  // the gathers address x with 32-bit element offsets (column * Fin): the number of rows of x is only known here when
  // every row is evaluated (rows == NULL: m == n); a row subset goes to the CUDA-core kernel unless the caller vouches
  if (x_rows <= 0) x_rows = rows ? 0 : m;
  if (x_rows <= 0 || (unsigned long long)x_rows * (unsigned long long)Fin >= (1ull << 32)) return GCM_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return GCM_ERR_UNSUPPORTED;
  const int K = 2 * Fin;
  const size_t a_bytes = (size_t)(GT_TM / 8) * (K / 4) * GT_LBO;
  const size_t smem = 2 * a_bytes + (size_t)2 * Fout * K * 4 + 128 * 4 + 64 + 128;
  if (smem > 227 * 1024) return GCM_ERR_UNSUPPORTED;
  GcTcArgs a{x, rowptr, col, ew, rows, m, Fin, Fout, wt, bias, act, agg_out, out, (m + GT_TM - 1) / GT_TM, nullptr};
  {
    // two-pass form when the caller provides room for the sums (always while recording; gcm.sparse_ops also hands a
    // scratch buffer to large no-grad calls): k_csr_gather at full occupancy, then the product kernel streams [agg | x]
    static const bool one_pass = getenv("GCM_B200_GRAPHCONV_ONE_PASS") != nullptr;     // A/B switch
    if (agg_out && !one_pass && m >= 64 * GT_TM && (reinterpret_cast<uintptr_t>(agg_out) & 15) == 0) {
      const long long gg = (m + 7) / 8;
      if (gg < 2147483647LL) {
        if (Fin == 64) k_csr_gather<2><<<(unsigned)gg, 256, 0, stream>>>(a);
        else k_csr_gather<1><<<(unsigned)gg, 256, 0, stream>>>(a);
        const int rc = gcm_check_launch("k_csr_gather");
        if (rc != GCM_OK) return rc;
        a.agg_in = agg_out;
        a.agg_out = nullptr;
        node_off = nullptr;            // the block-local kernel is a gather variant
      }
    }
  }
  {
    static const bool no_blk = getenv("GCM_B200_GRAPHCONV_NO_BLOCKS") != nullptr;     // A/B switch
    if (node_off && !rows && !no_blk && n_graphs > 0 && max_nodes > 0 && (size_t)max_nodes * Fin * 4 <= GB_XBYTES &&
        (m >= 4 * GT_TM || g_graphconv_kernel == GCM_GC_TC)) {
      const int rc = graphconv_fwd_blk(a, node_off, n_graphs, stream);
      if (rc != GCM_ERR_UNSUPPORTED) return rc;
    }
  }
  long long grid = gcm_num_sms();
  if (grid > a.tiles) grid = a.tiles;
  cudaError_t e;
#define GT_LAUNCH(V, ST)                                                                                                  \
  do {                                                                                                                    \
    e = cudaFuncSetAttribute(k_graphconv_fwd_tc<V, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e == cudaSuccess) k_graphconv_fwd_tc<V, ST><<<(unsigned)grid, GT_THREADS, smem, stream>>>(a);                     \
  } while (0)
  if (Fin == 64) {
    if (a.agg_in) GT_LAUNCH(2, true); else GT_LAUNCH(2, false);
  } else {
    if (a.agg_in) GT_LAUNCH(1, true); else GT_LAUNCH(1, false);
  }
#undef GT_LAUNCH
  if (e != cudaSuccess) {
    gcm_set_error("cudaFuncSetAttribute(graphconv_fwd_tc): %s", cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  return gcm_check_launch("k_graphconv_fwd_tc");
}

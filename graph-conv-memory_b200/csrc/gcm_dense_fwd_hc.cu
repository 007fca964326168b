// Pure-temporal DenseGCM step with a layer-1 row cache ("hc"), tcgen05 + TMEM, sm_100a.
//
// With forward-only TemporalBackedge hops the in-neighbourhood of a node is fixed when the node is created
// (edges only ever point from older nodes to the new one, edge_selectors/temporal.py:72-88), so the layer-1
// output h_p = act1(W_rel1 sum_{hops} x_{p-hop} + W_root1 x_p + b1) of node p is the same in every later
// step, as long as (a) the layer weights have not changed and (b) none of p's in-neighbours has been evicted
// while p is still within one hop of the newest node (guaranteed by N - 1 >= 2 max_hop).  The reference
// recomputes h for all N rows every step (gcm.py:308); the other kernels of this library recompute it for
// the 1 + #hops rows of the 2-hop neighbourhood.  This kernel keeps the last HC_RING rows of h per graph in
// HBM (hcache [B, ring, 32], ring slot = position % ring; 67 MB at BASELINE cfg2, L2-sized) and per step
//   layer 1:  ONE row per graph,   [agg_x | x_t]  (K = 2F)  x  [W_rel1 ; W_root1]  -> h_t   (written to the cache)
//   layer 2:  ONE row per graph,   [h_t | sum_{hops} h_{t-hop}] (K = 64) x [W_root2 ; W_rel2] -> belief
// i.e. a quarter of the layer-1 work of the row-recomputing kernels, and a tile is 128 graphs with
// TMEM lane = graph, so nothing is ever reduced across threads.  The host (gcm/fused.py) only selects this
// kernel when every cached row it reads was written under the current weights; until then the tc kernel
// (gcm_dense_fwd_tc.cu) runs and fills the cache.  Same arithmetic as there: 3xTF32, fp32 accumulate.
//
// One persistent CTA per SM, 12 warps:
//   warps 0-7   two consumer groups of 4 warps; a group owns a tile of QPG quarter tiles (32 graphs each; QPG = 3
//               or 4, as many stages as fit in shared memory), warp q < QPG its graphs 32q..32q+31 (TMEM lane
//               quadrant q; with QPG = 3 the last quadrant of the M = 128 MMA is idle).  Each thread builds its
//               graph's operands straight into TMEM
//               (tcgen05.st), warp 3 of the group (warp 0 if QPG = 4) issues the MMAs after a named-barrier hand-off, every builder
//               does the state update of its own quarter (node row, adjacency row, counter) while the tensor
//               core works, then bias + activation on the accumulators (tcgen05.ld).
//   warps 8-11  producers (3 active with 6 stages): 16-byte cp.async of each graph's needed rows (#hops node rows, #hops cached h
//               rows, the observation) into the quarter-tile stages, ONE STAGE PER ACTIVE CONSUMER WARP
//               (stage = group * QPG + q), completion by cp.async.mbarrier.arrive.  A stage is released right
//               after the operand build, so the next tile's loads run under the MMAs and epilogues.
// TMEM columns per group (256): A1 hi [0,64) | A1 lo [64,128) | D1 [128,160) | sum-h hi [160,192) |
// sum-h lo [192,224) | D2 [224,256); h_t (hi | lo) overlays A1 hi once the layer-1 MMAs have completed.
#include <stdlib.h>

#include "gcm_tc.cuh"
#include "gcm_temporal.cuh"

constexpr int HC_Q = 32;                   // graphs per quarter tile (one warp's lanes)
constexpr int HC_NPROD = 4;                // producer warps launched; 3 are used with 6 stages, 4 (or 2) with 8
constexpr int HC_GROUPS = 2;
constexpr int HC_CONS_THREADS = HC_GROUPS * 4 * 32;
constexpr int HC_THREADS = HC_CONS_THREADS + HC_NPROD * 32;
constexpr int HC_H = 32;
constexpr int HC_MAXP = 3;                 // forward hops (= in-neighbours of a node)
constexpr int HC_MAXSTAGES = 8;
constexpr uint32_t HC_COL_A1HI = 0, HC_COL_A1LO = 64, HC_COL_D1 = 128, HC_COL_SHI = 160, HC_COL_SLO = 192,
                   HC_COL_D2 = 224, HC_COL_GROUP = 256;
constexpr int HC_BAR_A1 = 1, HC_BAR_A2 = 3, HC_BAR_W = 5;   // named barriers (A1/A2: + group)

struct HcSmem {   // byte offsets into dynamic shared memory
  uint32_t b1hi, b1lo, b2hi, b2lo, zero, bias, bars, tmem_slot, flags, stage, total;
  uint32_t gs_floats;   // per-graph stride inside a stage
  uint32_t slot_bytes;
  int ns;               // quarter-tile stages in the ring
};

// nflags: quarter tiles per CTA (multi-step launches keep one "steps completed" word per quarter tile)
__host__ __device__ inline HcSmem hc_smem_layout(int F, int np, int nflags) {
  HcSmem L;
  const uint32_t K1 = 2 * F;
  uint32_t o = 0;
  L.b1hi = o; o += HC_H * K1 * 4;
  L.b1lo = o; o += HC_H * K1 * 4;
  L.b2hi = o; o += HC_H * 64 * 4;
  L.b2lo = o; o += HC_H * 64 * 4;
  L.zero = o; o += 128;                    // a row of zeros: target of out-of-window neighbour pointers
  L.bias = o; o += 2 * HC_H * 4;
  L.bars = o; o += (2 * HC_MAXSTAGES + 2 * HC_GROUPS) * 8;
  L.tmem_slot = o; o += 16;
  L.flags = o; o += (uint32_t)(nflags > HC_MAXSTAGES ? nflags : HC_MAXSTAGES) * 4;
  o = (o + 127u) & ~127u;
  // [np node rows | np cached h rows | observation] + 4 floats so that the 16-byte stride is odd (no bank
  // conflicts when the 32 lanes of a warp read the same column of their own graphs)
  L.gs_floats = (uint32_t)(np * F + np * HC_H + F + 4);
  L.slot_bytes = (uint32_t)HC_Q * L.gs_floats * 4;
  int ns = (int)((226u * 1024u - o) / L.slot_bytes);
  ns = ns > HC_MAXSTAGES ? HC_MAXSTAGES : ns;
  // One stage per ACTIVE consumer warp (ns / 2 per group, 3 or 4): a stage is then always consumed by the same
  // warp, in order, which is what makes the one-bit phase parity of its mbarriers unambiguous.
  L.ns = ns & ~1;
  L.stage = o; o += (uint32_t)(L.ns > 0 ? L.ns : 0) * L.slot_bytes;
  L.total = o;
  return L;
}

#ifdef HC_WATCHDOG
// debug build: bounded waits that report which barrier never completed
__device__ __noinline__ void hc_wait(uint64_t* bar, uint32_t parity, int code) {
  for (long long spin = 0; spin < 300000ll; ++spin) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  if ((threadIdx.x & 31) == 0)
    printf("hc watchdog: block %d warp %d wait code %d parity %u\n", blockIdx.x, threadIdx.x >> 5, code, parity);
}
#define HC_WAIT(bar, parity, code) hc_wait(bar, parity, code)
#else
#define HC_WAIT(bar, parity, code) tc::mbar_wait(bar, parity)
#endif

#ifdef HC_TRACE
// debug build: per-warp phase timestamps (globaltimer ns) of CTA 0, read back with gcm_debug_hc_trace()
__device__ unsigned long long g_hc_trace[12][8][8];
__device__ unsigned long long g_hc_cta[160][2];
extern "C" int gcm_debug_hc_cta(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_hc_cta, sizeof(g_hc_cta)) == cudaSuccess ? 0 : -2;
}
__device__ __forceinline__ void hc_cta_stamp(int which) {
  if (threadIdx.x == 0 && blockIdx.x < 160) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_hc_cta[blockIdx.x][which] = t;
  }
}
#define HC_CTA_STAMP(w) hc_cta_stamp(w)
#ifndef HC_TRACE_FROM
#define HC_TRACE_FROM 0
#endif
__device__ __forceinline__ void hc_stamp(int warp, int it, int phase) {
  it -= HC_TRACE_FROM;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it >= 0 && it < 8) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_hc_trace[warp][it][phase] = t;
  }
}
extern "C" int gcm_debug_hc_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_hc_trace, sizeof(g_hc_trace)) == cudaSuccess ? 0 : -2;
}
#define HC_STAMP(it, phase) hc_stamp(warp, it, phase)
#else
#define HC_STAMP(it, phase)
#define HC_CTA_STAMP(w)
#endif

__device__ __forceinline__ void hc_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void hc_cp_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hc_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void hc_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Programmatic dependent launch: back-to-back steps are launched with programmatic stream serialization, so the
// next step's CTAs take an SM as soon as this step's CTA leaves it and run their prologue (barrier init, TMEM
// allocation, weight staging: nothing that depends on earlier kernels) while the rest of this grid drains;
// pdl_wait() then blocks until every earlier kernel in the stream has completed and flushed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 16 fp32 values -> hi / lo tf32 halves, 16 TMEM columns each, of this thread's lane
__device__ __forceinline__ void hc_store_split16(uint32_t addr_hi, uint32_t addr_lo, const float (&v)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tc::split_tf32(v[j], hi[j], lo[j]);
  tc::tmem_st16(addr_hi, hi);
  tc::tmem_st16(addr_lo, lo);
}

template <int F>
__global__ void __launch_bounds__(HC_THREADS, 1) k_step_temporal_hc(const TemporalWinArgs a) {
  constexpr int K1 = 2 * F;
  constexpr int CPR = F / 4;                 // 16-byte chunks per node row
  const int np = a.prog.n_past;
  extern __shared__ __align__(128) unsigned char sm[];
  // this CTA's contiguous range of quarter tiles; every CTA walks the same padded sequence of nqp quarter slots per step
  const int nq_total = (a.st.B + HC_Q - 1) / HC_Q;
  const int nq_max = (nq_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const HcSmem L = hc_smem_layout(F, np, nq_max);
  float* B1hi = reinterpret_cast<float*>(sm + L.b1hi);
  float* B1lo = reinterpret_cast<float*>(sm + L.b1lo);
  float* B2hi = reinterpret_cast<float*>(sm + L.b2hi);
  float* B2lo = reinterpret_cast<float*>(sm + L.b2lo);
  float* zero_row = reinterpret_cast<float*>(sm + L.zero);
  float* bias_s = reinterpret_cast<float*>(sm + L.bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
  int* done_step = reinterpret_cast<int*>(sm + L.flags);        // [quarter tile of this CTA] steps completed so far
  float* stages = reinterpret_cast<float*>(sm + L.stage);
  uint64_t* full = bars;                         // [slot]
  uint64_t* empty = bars + HC_MAXSTAGES;         // [slot]
  uint64_t* d1_ready = bars + 2 * HC_MAXSTAGES;  // [group]
  uint64_t* d2_ready = d1_ready + HC_GROUPS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.st.N, C = a.st.C, W = a.st.W, B = a.st.B;
  const TemporalProg& P = a.prog;
  const int gs = (int)L.gs_floats, NS = L.ns;
  const int R = a.hc_ring;                       // power of two
  const bool uni = a.uniform_count >= 0;
  const int q_begin = (int)(((long long)nq_total * blockIdx.x) / gridDim.x);
  const int q_end = (int)(((long long)nq_total * (blockIdx.x + 1)) / gridDim.x);
  const int nq = q_end - q_begin;
  // The work of a launch is ONE sequence of quarter slots: S = t * nqp + s (step t, quarter tile s of this CTA; slots
  // s >= nq are bubbles).  Stage = S % NS, so over the steps of a sequence call every builder warp gets the same
  // share of quarter tiles (a single step leaves ceil(nq / builders) rounds to some warps and one less to others),
  // and consecutive steps overlap: there is no ramp / drain between them.  nqp >= NS keeps step t + 1 of a quarter
  // tile out of the MMA tile that holds its step t.
  const int T = a.n_steps > 1 ? a.n_steps : 1;
  const int nqp = nq_max > NS ? nq_max : NS;
  const int total = T * nqp;
  const int my_tiles = (total + (L.ns >> 1) - 1) / (L.ns >> 1);

  HC_CTA_STAMP(0);
  pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      tc::mbar_init(full + i, 32);               // one cp.async-completion arrival per producer lane
      tc::mbar_init(empty + i, 1);
    }
    for (int g = 0; g < HC_GROUPS; ++g) {
      tc::mbar_init(d1_ready + g, 1);
      tc::mbar_init(d2_ready + g, 1);
    }
    tc::mbar_fence_init();
  }
  if (tid < 32) zero_row[tid] = 0.0f;
  for (int i = tid; i < (nq_max > HC_MAXSTAGES ? nq_max : HC_MAXSTAGES); i += HC_THREADS) done_step[i] = 0;
  if (warp == 8) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp >= 8) {
    // =============================== producers ===============================
    // A stage must always be refilled by the SAME producer warp (its mbarrier phases are tracked by parity), so
    // the number of active producers divides the number of stages: producer p serves quarter q = p of both
    // groups, and the quarter tiles of one tile are fetched concurrently instead of one after the other.
    const int p = warp - 8;
    const int nprod = (NS % 3 == 0) ? 3 : 4;
    pdl_wait();                                  // the state and the observations come from earlier kernels
    const int xch = np * CPR, hch = np * 8, tch = xch + hch + CPR;    // 16-byte chunks per graph
    for (int S = p; S < total && p < nprod; S += nprod) {
      const int slot = S % NS, use = S / NS;
      const int t = S / nqp, s = S - t * nqp;
      const int g0 = (q_begin + s) * HC_Q;
      const int gt = s < nq ? min(HC_Q, B - g0) : 0;          // 0: a bubble (nothing to fetch, the stage still cycles)
      float* st_base = stages + (size_t)slot * HC_Q * gs;
      const float* obs_t = a.obs + (long long)t * a.obs_stride_t;
      HC_STAMP(S / nprod, 0);
      HC_WAIT(empty + slot, (use & 1) ^ 1, 100 + S);
      if (t > 0 && gt > 0) {
        // rows and counters of step t were written by the builder warp that ran step t - 1 of this quarter tile
        if (lane == 0) {
          volatile int* f = done_step + s;
          while (*f < t) {}
          __threadfence();
        }
        __syncwarp();
      }
      int cnt_l = 0;
      if (uni) cnt_l = a.uniform_count + t;
      else if (lane < gt) cnt_l = __ldcg(a.st.count + g0 + lane);
      HC_STAMP(S / nprod, 1);
      for (int c0 = 0; c0 < tch; c0 += 32) {
        const int c = c0 + lane;
        // decode the chunk: which row of which array, and its offset inside the per-graph stage
        int kind = 3, i = 0, col = 0, dst_off = 0;     // kind 0 = node row, 1 = cached h row, 2 = observation
        if (c < xch) {
          kind = 0; i = c / CPR; col = c - i * CPR; dst_off = i * F + col * 4;
        } else if (c < xch + hch) {
          const int c1 = c - xch;
          kind = 1; i = c1 >> 3; col = c1 & 7; dst_off = np * F + i * HC_H + col * 4;
        } else if (c < tch) {
          kind = 2; col = c - xch - hch; dst_off = np * F + np * HC_H + col * 4;
        }
        const int hop = kind < 2 ? P.past[i] : 0;
        if (uni) {
          const int cnt = cnt_l, lt = min(cnt, N - 1);
          const float* src = nullptr;
          size_t stride = 0;
          if (kind == 0 && hop <= lt) {
            src = a.st.nodes + ((size_t)g0 * C + gcm_slot(cnt - hop, C)) * F + col * 4;
            stride = (size_t)C * F;
          } else if (kind == 1 && hop <= lt) {
            src = a.hcache + ((size_t)g0 * R + ((cnt - hop) & (R - 1))) * HC_H + col * 4;
            stride = (size_t)R * HC_H;
          } else if (kind == 2) {
            src = obs_t + (size_t)g0 * a.obs_ld + col * 4;
            stride = (size_t)a.obs_ld;
          }
          if (src) {
            float* dst = st_base + dst_off;
#pragma unroll 8
            for (int gi = 0; gi < gt; ++gi) hc_cp16(dst + (size_t)gi * gs, src + (size_t)gi * stride);
          }
        } else {
          for (int gi = 0; gi < gt; ++gi) {
            const int cnt = __shfl_sync(GCM_FULL_MASK, cnt_l, gi);
            const int lt = min(cnt, N - 1);
            const float* src = nullptr;
            if (kind == 0 && hop <= lt)
              src = a.st.nodes + ((size_t)(g0 + gi) * C + gcm_slot(cnt - hop, C)) * F + col * 4;
            else if (kind == 1 && hop <= lt)
              src = a.hcache + ((size_t)(g0 + gi) * R + ((cnt - hop) & (R - 1))) * HC_H + col * 4;
            else if (kind == 2)
              src = obs_t + (size_t)(g0 + gi) * a.obs_ld + col * 4;
            if (src) hc_cp16(st_base + (size_t)gi * gs + dst_off, src);
          }
        }
      }
      hc_cp_arrive(full + slot);
      HC_STAMP(S / nprod, 2);
    }
  } else {
    // =============================== consumers ===============================
    const int grp = warp >> 2, q = warp & 3;
    // The weights may only be read ahead of pdl_wait() when the host vouches that nothing wrote them since the
    // previous step of this state (GCM_STEP_WEIGHTS_STABLE); otherwise wait first.
    if (!a.weights_stable) pdl_wait();
    // ---- layer weights -> canonical K-major B operands, split hi / lo (loads first, then the stores) ----
    {
      constexpr int PER1 = (HC_H * K1) / HC_CONS_THREADS;     // F = 8: 2, 16: 4, 32: 8
      float w1[PER1], w2[8];
#pragma unroll
      for (int j = 0; j < PER1; ++j) {
        const int i = tid + j * HC_CONS_THREADS;
        const int n = i / K1, k = i - n * K1;
        w1[j] = k < F ? __ldg(a.gnn.w_rel1 + n * F + k) : __ldg(a.gnn.w_root1 + n * F + (k - F));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tid + j * HC_CONS_THREADS;              // [n][k], k < 32: root, k >= 32: rel
        const int n = i >> 6, k = i & 63;
        w2[j] = k < HC_H ? __ldg(a.gnn.w_root2 + n * HC_H + k) : __ldg(a.gnn.w_rel2 + n * HC_H + (k - HC_H));
      }
#pragma unroll
      for (int j = 0; j < PER1; ++j) {
        const int i = tid + j * HC_CONS_THREADS;
        const int n = i / K1, k = i - n * K1;
        uint32_t hi, lo;
        tc::split_tf32(w1[j], hi, lo);
        B1hi[tc::kmajor_off(n, k, K1)] = __uint_as_float(hi);
        B1lo[tc::kmajor_off(n, k, K1)] = __uint_as_float(lo);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tid + j * HC_CONS_THREADS;
        const int n = i >> 6, k = i & 63;
        uint32_t hi, lo;
        tc::split_tf32(w2[j], hi, lo);
        B2hi[tc::kmajor_off(n, k, 64)] = __uint_as_float(hi);
        B2lo[tc::kmajor_off(n, k, 64)] = __uint_as_float(lo);
      }
      if (tid < HC_H) {
        bias_s[tid] = a.gnn.b1 ? __ldg(a.gnn.b1 + tid) : 0.0f;
        bias_s[HC_H + tid] = a.gnn.b2 ? __ldg(a.gnn.b2 + tid) : 0.0f;
      }
      tc::fence_proxy_async();           // the tensor core reads the operands through the async proxy
      hc_bar_sync(HC_BAR_W, HC_CONS_THREADS);
    }
    if (a.weights_stable) pdl_wait();    // everything below touches state written by earlier kernels

    const uint32_t tcol = tbase + grp * HC_COL_GROUP;
    const uint32_t taddr = tcol + ((uint32_t)(q * 32) << 16);
    const int act1 = a.gnn.act1, act2 = a.gnn.act2;
    const int QPG = NS >> 1;                 // builder warps (= quarter tiles) per group: 3 or 4
    const bool builder = q < QPG;
    // with 3 builders the 4th warp of the group does nothing but issue the MMAs (measured: an issuing builder
    // delays its own state update and epilogue by ~1.5 us per tile, and the whole group waits for it)
    const bool issuer = QPG == 3 ? q == 3 : q == 0;
    const uint32_t idesc = tc::idesc_tf32(128, HC_H);
    const uint32_t sbo1 = (uint32_t)(K1 / 4) * 128u, sbo2 = (uint32_t)(64 / 4) * 128u;
    const uint32_t b1hi = tc::smem_u32(B1hi), b1lo = tc::smem_u32(B1lo);
    const uint32_t b2hi = tc::smem_u32(B2hi), b2lo = tc::smem_u32(B2lo);

    int it = 0;
    for (int j = grp; j < my_tiles; j += HC_GROUPS, ++it) {
      const uint32_t ph = it & 1;
      const int S = QPG * j + q;               // position in the launch's sequence; slot = grp * QPG + q, use = it
      const bool has = builder && S < total;
      const int slot = S % NS, use = S / NS;
      const int t = has ? S / nqp : 0, s = has ? S - t * nqp : 0;
      const int g0 = (q_begin + s) * HC_Q;
      const int gt = (has && s < nq) ? min(HC_Q, B - g0) : 0;
      const bool live = lane < gt;
      float* st_base = stages + (size_t)slot * HC_Q * gs;
      const float* mine = st_base + (size_t)lane * gs;
      int cnt = 0, lt = 0;

      if (has) {
        HC_STAMP(it, 0);
        HC_WAIT(full + slot, use & 1, 200 + S);
        HC_STAMP(it, 1);
        // (after the wait: in a multi-step launch the counter was written by the previous step's builder warp)
        if (uni) cnt = a.uniform_count + t;
        else if (live) cnt = __ldcg(a.st.count + g0 + lane);
        lt = min(cnt, N - 1);
        const float* xp[HC_MAXP];
        const float* hp[HC_MAXP];
#pragma unroll
        for (int i = 0; i < HC_MAXP; ++i) {
          const bool ok = live && i < np && P.past[i] <= lt;
          xp[i] = ok ? mine + i * F : zero_row;
          hp[i] = ok ? mine + np * F + i * HC_H : zero_row;
        }
        const float* x_ptr = live ? mine + np * F + np * HC_H : zero_row;
        // recording (F = 32): the same operand row also goes to the tiled buffer of the fused window backward
        float4* xrec4 = nullptr;
        if (F == 32 && a.xrec && live) {
          const long long r = a.xrec_row0 + (long long)t * a.st.B + g0 + lane;
          xrec4 = reinterpret_cast<float4*>(a.xrec) + (r >> 7) * (16 * 128) + (r & 127);
        }
        // ---- layer-1 operand [sum of in-neighbour rows | own row] -> TMEM (hi | lo) ----
#pragma unroll
        for (int c0 = 0; c0 < K1; c0 += 16) {
          float v[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int col = c0 + q4 * 4;
            float4 sv;
            if (col < F) {
              sv = *reinterpret_cast<const float4*>(xp[0] + col);
#pragma unroll
              for (int i = 1; i < HC_MAXP; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(xp[i] + col);
                sv.x += t.x; sv.y += t.y; sv.z += t.z; sv.w += t.w;
              }
            } else {
              sv = *reinterpret_cast<const float4*>(x_ptr + (col - F));
            }
            v[q4 * 4 + 0] = sv.x; v[q4 * 4 + 1] = sv.y; v[q4 * 4 + 2] = sv.z; v[q4 * 4 + 3] = sv.w;
          }
          hc_store_split16(taddr + HC_COL_A1HI + c0, taddr + HC_COL_A1LO + c0, v);
          if (F == 32 && xrec4) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              __stcs(xrec4 + (c0 / 4 + q4) * 128, make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]));
          }
        }
        // ---- second half of the layer-2 operand: sum of the in-neighbours' cached h rows ----
#pragma unroll
        for (int c0 = 0; c0 < HC_H; c0 += 16) {
          float v[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int col = c0 + q4 * 4;
            float4 sv = *reinterpret_cast<const float4*>(hp[0] + col);
#pragma unroll
            for (int i = 1; i < HC_MAXP; ++i) {
              const float4 t = *reinterpret_cast<const float4*>(hp[i] + col);
              sv.x += t.x; sv.y += t.y; sv.z += t.z; sv.w += t.w;
            }
            v[q4 * 4 + 0] = sv.x; v[q4 * 4 + 1] = sv.y; v[q4 * 4 + 2] = sv.z; v[q4 * 4 + 3] = sv.w;
          }
          hc_store_split16(taddr + HC_COL_SHI + c0, taddr + HC_COL_SLO + c0, v);
        }
        // node write (gcm.py:274) straight from the stage's observation tile, before the stage is released: 32 / CPR
        // graphs per instruction, 128-bit coalesced stores (re-reading the tile from global after the release put an
        // L2 round trip of ~1 us on every tile's critical path, tools/hc_trace_seq.py)
        if (gt > 0) {
          const int tslot_w = gcm_slot(cnt, C);
#pragma unroll
          for (int i = 0; i < CPR; ++i) {
            const int gi = i * (32 / CPR) + lane / CPR, col = lane % CPR;
            const int ts = __shfl_sync(GCM_FULL_MASK, tslot_w, gi);
            if (gi < gt) {
              const float4 v = *reinterpret_cast<const float4*>(st_base + (size_t)gi * gs + np * F + np * HC_H + col * 4);
              *reinterpret_cast<float4*>(a.st.nodes + ((size_t)(g0 + gi) * C + ts) * F + col * 4) = v;
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(empty + slot);        // everything was read: the stage can be refilled
        tc::wait_st();
        HC_STAMP(it, 2);
      }
      tc::fence_before_sync();
      if (!issuer) {
        hc_bar_arrive(HC_BAR_A1 + grp, 128);
      } else {
        hc_bar_sync(HC_BAR_A1 + grp, 128);
        tc::fence_after_sync();
        if (lane == 0) {
          bool acc = false;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {          // lo*Bhi, hi*Blo, hi*Bhi
            const uint32_t acol = tcol + (pass == 0 ? HC_COL_A1LO : HC_COL_A1HI);
            const uint32_t bsm = pass == 1 ? b1lo : b1hi;
#pragma unroll
            for (int ks = 0; ks < K1 / 8; ++ks) {
              tc::mma_tf32_ts(tcol + HC_COL_D1, acol + ks * 8, tc::smem_desc_kmajor(bsm + ks * 256, 128, sbo1), idesc, acc);
              acc = true;
            }
          }
          tc::mma_commit(d1_ready + grp);
        }
        __syncwarp();
      }

      // ---- state update of this warp's quarter while the tensor core works ----
      if (gt > 0) {
        const int tslot = gcm_slot(cnt, C);
        if (live) {
          // adjacency row of the new node: past mask, cleared future mask (contiguous 2W words per graph)
          uint32_t* mrow = a.st.masks + ((size_t)(g0 + lane) * C + tslot) * 2 * W;
          if ((W & 3) == 0) {
            for (int w4 = 0; w4 < W; w4 += 4) {
              uint32_t pw[4] = {0u, 0u, 0u, 0u};
              for (int i = 0; i < np; ++i) {
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
              for (int i = 0; i < np; ++i) {
                const int hop = P.past[i];
                if (hop <= lt && (hop >> 5) == w) pw |= 1u << (hop & 31);
              }
              gcm_st_mask(mrow + w, pw);
              gcm_st_mask(mrow + W + w, 0u);
            }
          }
          __stcg(a.st.count + g0 + lane, cnt + 1);
        }
      }

      // ---- layer-1 epilogue: h_t = act(D1 + b1) -> cache row + first half of the layer-2 operand ----
      HC_STAMP(it, 3);
      if (builder) {
      HC_WAIT(d1_ready + grp, ph, 300 + j);
      HC_STAMP(it, 4);
      tc::fence_after_sync();
      {
        uint32_t v0[16], v1[16];
        tc::tmem_ld16(taddr + HC_COL_D1, v0);
        tc::tmem_ld16(taddr + HC_COL_D1 + 16, v1);
        tc::wait_ld();
        float h0[16], h1[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          h0[k] = __uint_as_float(v0[k]) + bias_s[k];
          h1[k] = __uint_as_float(v1[k]) + bias_s[16 + k];
        }
        gcm_act_fast_vec(h0, act1);
        gcm_act_fast_vec(h1, act1);
        if (live) {
          float4* dst = reinterpret_cast<float4*>(a.hcache + ((size_t)(g0 + lane) * R + (cnt & (R - 1))) * HC_H);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            __stcg(dst + k, make_float4(h0[4 * k], h0[4 * k + 1], h0[4 * k + 2], h0[4 * k + 3]));
            __stcg(dst + 4 + k, make_float4(h1[4 * k], h1[4 * k + 1], h1[4 * k + 2], h1[4 * k + 3]));
          }
        }
        hc_store_split16(taddr + 0, taddr + 32, h0);           // h_t hi [0,32) | lo [32,64) over A1 hi
        hc_store_split16(taddr + 16, taddr + 48, h1);
      }
      tc::wait_st();
      }
      if (T > 1 && gt > 0) {
        // this quarter tile's node rows, cached rows and counters of step t are written: step t + 1 may fetch them
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          *reinterpret_cast<volatile int*>(done_step + s) = t + 1;
        }
      }
      tc::fence_before_sync();
      if (!issuer) {
        hc_bar_arrive(HC_BAR_A2 + grp, 128);
      } else {
        hc_bar_sync(HC_BAR_A2 + grp, 128);
        tc::fence_after_sync();
        if (lane == 0) {
          bool acc = false;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t bsm = pass == 1 ? b2lo : b2hi;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                 // K steps 0-3: h_t (root), 4-7: sum of cached rows (rel)
              const uint32_t hi_col = ks < 4 ? (uint32_t)(ks * 8) : HC_COL_SHI + (uint32_t)((ks - 4) * 8);
              const uint32_t lo_col = ks < 4 ? (uint32_t)(32 + ks * 8) : HC_COL_SLO + (uint32_t)((ks - 4) * 8);
              tc::mma_tf32_ts(tcol + HC_COL_D2, tcol + (pass == 0 ? lo_col : hi_col),
                              tc::smem_desc_kmajor(bsm + ks * 256, 128, sbo2), idesc, acc);
              acc = true;
            }
          }
          tc::mma_commit(d2_ready + grp);
        }
        __syncwarp();
      }

      // ---- layer-2 epilogue: belief row of this thread's graph ----
      HC_STAMP(it, 5);
      if (builder) {
      HC_WAIT(d2_ready + grp, ph, 400 + j);
      HC_STAMP(it, 6);
      tc::fence_after_sync();
      {
        uint32_t v0[16], v1[16];
        tc::tmem_ld16(taddr + HC_COL_D2, v0);
        tc::tmem_ld16(taddr + HC_COL_D2 + 16, v1);
        tc::wait_ld();
        if (live) {
          float o0[16], o1[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            o0[k] = __uint_as_float(v0[k]) + bias_s[HC_H + k];
            o1[k] = __uint_as_float(v1[k]) + bias_s[HC_H + 16 + k];
          }
          gcm_act_fast_vec(o0, act2);
          gcm_act_fast_vec(o1, act2);
          bool bad = false;
#pragma unroll
          for (int k = 0; k < 16; ++k) bad |= !isfinite(o0[k]) | !isfinite(o1[k]);
          float4* dst = reinterpret_cast<float4*>(a.belief + (long long)t * a.belief_stride_t + (size_t)(g0 + lane) * a.belief_ld);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            dst[k] = make_float4(o0[4 * k], o0[4 * k + 1], o0[4 * k + 2], o0[4 * k + 3]);
            dst[4 + k] = make_float4(o1[4 * k], o1[4 * k + 1], o1[4 * k + 2], o1[4 * k + 3]);
          }
          if (bad) atomicOr(reinterpret_cast<unsigned int*>(a.status), GCM_FLAG_NONFINITE);
        }
      }
      }
      HC_STAMP(it, 7);
      tc::fence_before_sync();
    }
  }
  __syncthreads();
  HC_CTA_STAMP(1);
  if (warp == 8) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tbase, 512);
  }
}

template <int F>
static int launch_hc(const TemporalWinArgs& a, cudaStream_t stream) {
  const int nq = (a.st.B + HC_Q - 1) / HC_Q;
  int grid = gcm_num_sms();
  if (grid > nq) grid = nq;
  const int nq_max = (nq + grid - 1) / grid;
  const HcSmem L = hc_smem_layout(F, a.prog.n_past, nq_max);
  if (L.ns < 6) return GCM_ERR_UNSUPPORTED;
  if ((long long)(a.n_steps > 1 ? a.n_steps : 1) * (nq_max > L.ns ? nq_max : L.ns) >= (1ll << 30)) return GCM_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_step_temporal_hc<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      gcm_set_error("cudaFuncSetAttribute(temporal_hc): %s", cudaGetErrorString(e));
      return GCM_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(HC_THREADS);
  cfg.dynamicSmemBytes = L.total + 128;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool no_pdl = getenv("GCM_B200_NO_PDL") != nullptr;   // A/B switches for profiling
  static const bool no_l2p = getenv("GCM_B200_NO_L2PERSIST") != nullptr;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  // EXPERIMENT KNOB, off unless GCM_B200_L2PERSIST_MB is set.  The row cache is written once and read three times
  // (hops 1, 2, 4 at cfg2) over the next steps and is L2-sized (67 MB against 126 MB), so a persisting access window
  // looked like a way to stop 3 of the 7 rows of a graph-step from coming out of HBM.  Measured at cfg2 (B200, 82.9 MB
  // maximum carve-out): 23.3 us without, 23.2 / 24.3 / 29.1 / 32.4 us with a 16 / 40 / 67 / 79 MB carve-out -- the
  // carve-out takes L2 away from the streamed rows and costs more than the pinned lines give back.
  static size_t l2_persist = 0, l2_window = 0;
  static bool l2_init = false;
  if (!l2_init && !no_l2p) {
    l2_init = true;
    int dev = 0;
    cudaDeviceProp prop;
    const char* mb = getenv("GCM_B200_L2PERSIST_MB");                 // size of the persisting carve-out (experiment knob)
    if (mb && cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess &&
        prop.persistingL2CacheMaxSize > 0) {
      size_t want = (size_t)atoll(mb) << 20;
      if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
      if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
        l2_persist = want;
        l2_window = (size_t)prop.accessPolicyMaxWindowSize;
      }
      if (getenv("GCM_B200_L2PERSIST_VERBOSE"))
        fprintf(stderr, "gcm: persisting L2 max %d B, window max %d B, carve-out %zu B\n", prop.persistingL2CacheMaxSize,
                prop.accessPolicyMaxWindowSize, l2_persist);
    }
    cudaGetLastError();
  }
  const size_t hc_bytes = (size_t)a.st.B * a.hc_ring * HC_H * sizeof(float);
  if (!no_l2p && l2_persist > 0 && hc_bytes > 0 && hc_bytes <= l2_window) {
    attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[na].val.accessPolicyWindow.base_ptr = const_cast<float*>(a.hcache);
    attr[na].val.accessPolicyWindow.num_bytes = hc_bytes;
    attr[na].val.accessPolicyWindow.hitRatio = hc_bytes <= l2_persist ? 1.0f : (float)((double)l2_persist / (double)hc_bytes);
    attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_step_temporal_hc<F>, a);
  if (e != cudaSuccess) {
    gcm_set_error("k_step_temporal_hc: launch failed: %s", cudaGetErrorString(e));
    return GCM_ERR_CUDA;
  }
  return gcm_check_launch("k_step_temporal_hc");
}

// the structural conditions under which cached layer-1 rows are exact (see the header comment)
bool gcm_temporal_hc_shape_ok(const TemporalWinArgs& a) {
  const TemporalProg& P = a.prog;
  if (a.gnn.H1 != HC_H || a.gnn.H2 != HC_H || P.n_future != 0 || P.n_past < 1 || P.n_past > HC_MAXP) return false;
  if (!(a.st.F == 8 || a.st.F == 16 || a.st.F == 32)) return false;
  if (!a.gnn.w_rel1 || !a.gnn.w_root1 || !a.gnn.w_rel2 || !a.gnn.w_root2) return false;
  int maxhop = 0;
  for (int i = 0; i < P.n_past; ++i) maxhop = P.past[i] > maxhop ? P.past[i] : maxhop;
  if (a.st.N - 1 < 2 * maxhop) return false;                           // (b): no in-neighbour of a cached row evicted
  if (a.hc_ring < maxhop + 1 || (a.hc_ring & (a.hc_ring - 1)) != 0) return false;
  return true;
}

int gcm_launch_temporal_hc(const TemporalWinArgs& a, cudaStream_t stream) {
  if (!a.hcache || !gcm_temporal_hc_shape_ok(a) || (reinterpret_cast<uintptr_t>(a.hcache) & 15) != 0 ||
      (a.obs_ld & 3) != 0 || (a.belief_ld & 3) != 0 || a.obs_ld < a.st.F || a.belief_ld < HC_H ||
      (reinterpret_cast<uintptr_t>(a.belief) & 15) != 0)
    return GCM_ERR_UNSUPPORTED;
  switch (a.st.F) {
    case 8: return launch_hc<8>(a, stream);
    case 16: return launch_hc<16>(a, stream);
    case 32: return launch_hc<32>(a, stream);
    default: return GCM_ERR_UNSUPPORTED;
  }
}

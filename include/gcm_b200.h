/*
 * gcm_b200.h — C ABI of libgcm_b200.so: the B200 (sm_100a) hot path of Graph
 * Convolutional Memory (proroklab/graph-conv-memory), memory-update-and-aggregate.
 *
 * The reference is pure Python; its "FFI" for this path is the set of Python
 * call sites listed beside each entry point (paths under /root/reference/src/gcm).
 * Everything here is plain C: raw DEVICE pointers, sizes and a cudaStream_t
 * (passed as void*).  No torch types, no exceptions: every function returns 0
 * on success and a negative gcm_status on failure; gcm_last_error() gives the
 * text.  Kernels are enqueued on the given stream and never synchronise.
 *
 * State layout (all in HBM, owned by the caller):
 *   nodes  float32 [B, C, F]       node log; node with absolute position p lives in
 *                                  slot p % C.  C >= N (C == N: in-place ring).
 *   masks  uint32  [B, C, 2, W]    per node two N-bit masks, W = ceil(N/32):
 *                                  [..,0,:] past  : bit d set  <=>  edge (p-d) -> p  (d = 0: self loop)
 *                                  [..,1,:] future: bit d set  <=>  edge (p+d) -> p  (d >= 1)
 *                                  a bit is live only while its source is inside the window, so the
 *                                  reference's "drop node 0 and shift" (gcm.py:323-355) costs nothing.
 *   count  int32   [B]             nodes ever written; the next node gets position count[b].
 * The reference's visible hidden state (gcm.py:194-211) is nodes[B,N,F] f32, adj[B,N,N] f32
 * (adj[i,j] = 1: j -> i), num_nodes[B] i64 = min(count, N); gcm_state_materialize /
 * gcm_state_ingest convert between the two.
 */
#ifndef GCM_B200_H
#define GCM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCM_ABI_VERSION 1
#define GCM_MAX_HOPS 16
#define GCM_MAX_SELECTORS 4
#define GCM_MAX_N 1024 /* graph_size limit of the packed representation   */
#define GCM_MAX_FEAT 256 /* F, H1, H2 limit of the fused kernels           */

typedef enum gcm_status {
  GCM_OK = 0,
  GCM_ERR_INVALID = -1,     /* bad argument / unsupported shape            */
  GCM_ERR_CUDA = -2,        /* a CUDA runtime call failed                  */
  GCM_ERR_UNSUPPORTED = -3  /* valid request the fused path does not cover */
} gcm_status;

/* device-side status word bits (int32 written with atomicOr, polled lazily by the host) */
#define GCM_FLAG_NONFINITE 1u   /* belief has NaN/Inf  (gcm.py:316-318 assertion)            */
#define GCM_FLAG_UNCLEAN 2u     /* ingest: adj not {0,1} or edges touching rows >= num_nodes  */
#define GCM_FLAG_BADCOUNT 4u    /* ingest: num_nodes outside [0, N]                           */
#define GCM_FLAG_OVERFLOW 8u    /* sparse: T + tau > N  (sparse_gcm.py:120-121)               */
#define GCM_FLAG_NONCAUSAL 16u  /* sparse: edge with source >= sink (sparse_gcm.py:171)       */

typedef enum gcm_selector_kind {
  GCM_SEL_NONE = 0,
  GCM_SEL_TEMPORAL = 1,  /* edge_selectors/temporal.py:72-88 (deterministic branch) */
  GCM_SEL_DENSE = 2,     /* edge_selectors/dense.py:11-23                            */
  GCM_SEL_EUCLIDEAN = 3, /* edge_selectors/distance.py:18-39 + 48-49                 */
  GCM_SEL_COSINE = 4,    /* edge_selectors/distance.py:18-39 + 59-61                 */
  GCM_SEL_SPATIAL = 5    /* edge_selectors/distance.py:18-39 + 77-81                 */
} gcm_selector_kind;

typedef enum gcm_direction { GCM_DIR_FORWARD = 0, GCM_DIR_BACKWARD = 1, GCM_DIR_BOTH = 2 } gcm_direction;
typedef enum gcm_act { GCM_ACT_NONE = 0, GCM_ACT_TANH = 1, GCM_ACT_RELU = 2 } gcm_act;

typedef struct gcm_dense_state {
  float* nodes;    /* [B, C, F]    */
  uint32_t* masks; /* [B, C, 2, W] */
  int32_t* count;  /* [B]          */
  int32_t B, N, C, F, W;
} gcm_dense_state;

typedef struct gcm_selector {
  int32_t kind;      /* gcm_selector_kind */
  int32_t direction; /* temporal only     */
  int32_t n_hops;
  int32_t hops[GCM_MAX_HOPS];
  float max_distance;                   /* strict '<' (distance.py:27)                          */
  int32_t a_start, a_step;              /* spatial: slice of the CURRENT node (a_pose_slice)     */
  int32_t b_start, b_step;              /* spatial: slice of the PAST nodes   (b_pose_slice)     */
  int32_t slice_len;
  const float* dist_param;              /* learned=True: device scalar, nodes are divided by it  */
  const float* dist;                    /* euclidean: [B, C] batch-mean distance per slot, from  */
                                        /* gcm_euclid_batchmean (the reference's cdist quirk)    */
} gcm_selector;

/* 2-layer DenseGraphConv stack (torch_geometric.nn.DenseGraphConv, call sites README.md:52-62):
 *   out = act2(lin_rel2(A h) + lin_root2(h)),  h = act1(lin_rel1(A x) + lin_root1(x))
 * w1t/w2t are K-major packs for the forward; w_rel / w_root are the layers' own [out,in] weights
 * (used by the backward).  One bias per layer (on lin_rel in PyG >= 2.0, lin_root in 1.x). */
typedef struct gcm_gnn {
  const float* w1t; /* [2F,  H1]: rows 0..F-1 = lin_rel1.weight^T, rows F..2F-1 = lin_root1.weight^T */
  const float* b1;  /* [H1] or NULL */
  const float* w2t; /* [2H1, H2] */
  const float* b2;  /* [H2] or NULL */
  const float* w_rel1;  /* [H1, F]  */
  const float* w_root1; /* [H1, F]  */
  const float* w_rel2;  /* [H2, H1] */
  const float* w_root2; /* [H2, H1] */
  int32_t F, H1, H2, act1, act2;
} gcm_gnn;

/* gradient accumulators of the backward (all device, float32, ACCUMULATED into) */
typedef struct gcm_gnn_grads {
  float* d_w_rel1;  /* [H1, F]  */
  float* d_w_root1; /* [H1, F]  */
  float* d_b1;      /* [H1] or NULL */
  float* d_w_rel2;  /* [H2, H1] */
  float* d_w_root2; /* [H2, H1] */
  float* d_b2;      /* [H2] or NULL */
} gcm_gnn_grads;

int gcm_version(void);
const char* gcm_last_error(void);
/* Introspection for benchmarks / tests: name of the kernel most recently launched by the calling thread,
 * and the number of kernels this library has launched in the process so far. */
const char* gcm_last_kernel(void);
long long gcm_launch_count(void);

/* One DenseGCM.forward step for the whole batch (replaces gcm.py:262-321: node write at
 * nodes[b, num_nodes[b]] (:274), overflow wrap (:263-271, :323-355), edge selectors (:284-287),
 * gnn (:308), belief extraction (:314), finite check (:316-318), num_nodes + 1 (:320)).
 * The state is updated IN PLACE.  belief: [B, H2].  status: device int32 (flags OR-ed in).
 * flags: bit 0 = the state is "pure temporal" (built from empty by this same TEMPORAL-only
 *        selector chain), which enables the implicit-adjacency fast kernel. */
#define GCM_STEP_PURE_TEMPORAL 1
/*        bit 1 = every graph has the same count, given in bits 8.. of `flags` (host mirror; saves the
 *        dependent counter load in the pipelined kernel). */
#define GCM_STEP_UNIFORM_COUNT 2
#define GCM_STEP_COUNT_SHIFT 8
int gcm_dense_step_fwd(const gcm_dense_state* st, const float* obs, const gcm_selector* sels,
                       int n_sels, const gcm_gnn* gnn, float* belief, int32_t* status, int flags,
                       void* stream);

/* The same step with a layer-1 row cache.  For forward-only TemporalBackedge chains the layer-1 output h_p of
 * a node never changes once computed (its in-edges are fixed at creation, edge_selectors/temporal.py:72-88),
 * provided the layer weights stay the same and N - 1 >= 2 max_hop.  hcache [B, hc_ring, H1] float32 (hc_ring a
 * power of two > max_hop; slot = node position % hc_ring) holds the most recent rows.  Protocol: every step that
 * reports *cache_written = 1 stored h of the node it created.  The caller sets GCM_STEP_HCACHE_VALID in `flags`
 * once the last min(max_hop, count) nodes were all written under the CURRENT weights; the library then reads
 * the cached rows instead of recomputing them (1 layer-1 row per graph instead of 1 + #hops).  Shapes the cache
 * does not apply to behave exactly like gcm_dense_step_fwd and report *cache_written = 0. */
#define GCM_STEP_HCACHE_VALID 4
/* bit 3: the layer weights were not written by anything since the previous step of this state.  The cached-row
 * kernel is launched with programmatic stream serialization; with this bit it stages the weights while the
 * previous kernel in the stream is still draining (it always waits for that kernel before touching state). */
#define GCM_STEP_WEIGHTS_STABLE 8
int gcm_dense_step_fwd_cached(const gcm_dense_state* st, const float* obs, const gcm_selector* sels,
                              int n_sels, const gcm_gnn* gnn, float* belief, int32_t* status, int flags,
                              float* hcache, int hc_ring, int* cache_written, void* stream);

/* The same step with the uniform count as its own argument (any int32; the packed form above holds 23 bits) and
 * strided rows: observation of graph b at obs + b * obs_ld floats, belief row b at belief + b * belief_ld (0 = contiguous;
 * multiples of 4).  Strided rows are what a caller holding [B, T, .] tensors passes for one t (ray_gcm.py:200-202); only
 * the cached-row kernel takes them, any other kernel choice returns GCM_ERR_UNSUPPORTED without launching. */
int gcm_dense_step_fwd_ex(const gcm_dense_state* st, const float* obs, long long obs_ld, const gcm_selector* sels,
                          int n_sels, const gcm_gnn* gnn, float* belief, long long belief_ld, int32_t* status, int flags,
                          int uniform_count, float* hcache, int hc_ring, int* cache_written, void* stream);

/* Rollout entry for TEMPORAL-only selector chains (edge_selectors/temporal.py:72-88): T consecutive DenseGCM.forward
 * steps (gcm.py:213-321) enqueued back to back from C, i.e. the loop `for t in range(T): out, hidden = self.gcm(flat[:,
 * t, :], hidden)` of ray_gcm.py:200-202 without a host round trip per step.  The descriptor carries what the host
 * otherwise keeps between steps; the library updates the in/out fields.  Step k reads obs + k * obs_stride_t (graph b at
 * + b * obs_ld) and writes belief + k * belief_stride_t (row b at + b * belief_ld); scratch_obs [B,F] / scratch_belief
 * [B,H2] are needed when rows are strided (steps that cannot run on the cached-row kernel are staged through them). */
typedef struct gcm_rollout {
  gcm_dense_state st;
  gcm_gnn gnn;
  gcm_selector sels[GCM_MAX_SELECTORS];
  int32_t n_sels;
  int32_t max_hop;        /* largest hop of the chain                                                         */
  float* hcache;          /* layer-1 row cache [B, hc_ring, H1] (see gcm_dense_step_fwd_cached) or NULL        */
  int32_t hc_ring;
  int32_t uniform_count;  /* in/out: every graph's count when they are all equal (host mirror), else -1        */
  int32_t hc_fresh;       /* in/out: newest nodes whose cached row was written under the current weights       */
  int32_t weights_stable; /* in: the weights were not written since the previous step of this state            */
  int32_t* status;        /* device status word                                                                */
  float* scratch_obs;
  float* scratch_belief;
  long long launches;     /* out: kernels launched by the last call                                            */
  /* recording for the fused window backward (gcm_temporal_window_bwd), F = 32: xrec = its TILED operand buffer X or NULL;
   * step k of a call writes the rows xrec_row0 + k * B + graph.  Only the multi-step launch of the cached-row kernel
   * writes them: on return xrec_from = index of the first step whose rows were written (T if none); the caller fills
   * the earlier rows with gcm_temporal_gather. */
  float* xrec;
  long long xrec_row0;
  int32_t xrec_from;
} gcm_rollout;
int gcm_dense_rollout_fwd(gcm_rollout* r, const float* obs, long long obs_ld, long long obs_stride_t, float* belief,
                          long long belief_ld, long long belief_stride_t, int T, void* stream);
/* one contiguous step: gcm_dense_rollout_fwd(r, obs, 0, 0, belief, 0, 0, 1, stream) */
int gcm_dense_rollout_step(gcm_rollout* r, const float* obs, float* belief, void* stream);

/* Which kernel serves GCM_STEP_PURE_TEMPORAL steps (process-wide; default AUTO = fastest that fits the
 * shape).  All variants compute the same step; the switch exists for parity tests and A/B profiling. */
typedef enum gcm_temporal_kernel {
  GCM_TK_AUTO = 0,
  GCM_TK_HC = 1,   /* tcgen05, cached layer-1 rows, thread = graph (gcm_dense_fwd_hc.cu); falls back to
                      TC while the cache is being filled */
  GCM_TK_TC = 2,   /* tcgen05, lane = (row, graph), 32-graph tiles (gcm_dense_fwd_tc.cu) */
  GCM_TK_WIN = 3,  /* CUDA cores, pipelined history window        (k_step_temporal_win) */
  GCM_TK_ROWS = 4  /* CUDA cores, per-row gathers                 (k_step_temporal)     */
} gcm_temporal_kernel;
int gcm_set_temporal_kernel(int which);

/* Backward of step `steps_back` steps ago (0 = the most recent step), recomputed from the node
 * log (requires that no slot of that step's window was overwritten since: count_now - start <= C).
 * Replaces autograd through gcm.py:262-321.  d_belief [B,H2] in; d_nodes [B,C,F] is the running
 * dL/dnodes buffer (accumulated); d_obs [B,F] out = dL/dx of that step (its d_nodes row, which is
 * then zeroed).  Weight gradients are accumulated into `grads`. */
int gcm_dense_step_bwd(const gcm_dense_state* st, int steps_back, const gcm_gnn* gnn,
                       const float* d_belief, float* d_nodes, float* d_obs,
                       const gcm_gnn_grads* grads, void* stream);

/* ---- window-level backward of forward-only TemporalBackedge chains (csrc/gcm_temporal_bwd.cu) ----------------
 * Autograd through gcm.py:262-321 over the T steps of a BPTT window (reference training loop: tests/test_gcm.py:412-439)
 * for a state built from empty by a chain of forward hops (edge_selectors/temporal.py:72-88): every product of the
 * backward is a row-parallel GEMM over TIME-MAJOR rows [row, graph, feature] (row <-> absolute node position p; edge
 * p-s -> p exists iff p - s >= 0), run by gcm_linear_tc32 / gcm_outer_reduce_tc32; these two entries build the operands.
 * gcm_temporal_gather: out[i, b, :] = [ sum_{s in hops, p-s >= 0} x_{p-s} | x_p ] (2F floats), p = p0 + i, from the node
 * log (rows with p < 0 are zero).  The caller guarantees that positions p0 - max_hop .. p0 + n_rows - 1 are still in
 * the log (count - C <= p0 - max_hop) and already written (p0 + n_rows <= count).  F % 4 == 0.
 * gcm_temporal_shift_sum: out[i, b, :] = [ sum_s src[pos + sign * s] | src[pos] ] (2H floats), pos = out_pos0 + i, with
 * src [n_src, B, H] holding positions src_pos0 .. src_pos0 + n_src - 1 (anything outside is zero); positions below
 * valid_lo are nodes that never existed: they contribute nothing and their own output rows are zero.  sign = -1: sums
 * over in-neighbours (layer inputs), +1: over out-neighbours (gradients).  H % 4 == 0.
 * act_out (optional, same shape as src): the source rows are src * act'(act_out) with act = GCM_ACT_* -- dL/dbelief and
 * the beliefs of a window, so that dz2 is formed on the fly.
 * tiled = 1 (32 features only): out is written as [tile of 128 rows][16-byte chunk 0 .. 15][row of the tile][4 floats] over
 * the flattened rows r = i * B + b, the operand layout of gcm_temporal_window_bwd; the buffer must cover whole tiles. */
int gcm_temporal_gather(const gcm_dense_state* st, const int32_t* hops, int n_hops, long long p0, int n_rows, float* out,
                        int tiled, void* stream);
int gcm_temporal_shift_sum(const float* src, long long src_pos0, int n_src, long long valid_lo, const int32_t* hops,
                           int n_hops, int sign, float* out, long long out_pos0, int n_out, int B, int H, int tiled,
                           const float* act_out, int act, void* stream);
/* The same with the source read through element strides: src[j, b, h] at src + j * s_t + b * s_b + h * s_h (a [B, T, H]
 * gradient of the beliefs read in place, or the stride-0 gradient of a sum loss, which autograd hands over as an
 * expanded scalar: neither is copied).  Only the shape the fused window backward uses is covered (tiled = 1, sign = +1,
 * H = 32, hops <= 4, n_out >= 8); anything else returns GCM_ERR_UNSUPPORTED and the caller passes a contiguous copy to
 * gcm_temporal_shift_sum.  act_out stays contiguous [n_src, B, H]. */
int gcm_temporal_shift_sum_strided(const float* src, long long s_t, long long s_b, long long s_h, long long src_pos0,
                                   int n_src, long long valid_lo, const int32_t* hops, int n_hops, int sign, float* out,
                                   long long out_pos0, int n_out, int B, int H, int tiled, const float* act_out, int act,
                                   void* stream);

/* ---- the row products and weight-gradient reductions of that backward as ONE kernel (csrc/gcm_temporal_bwd_tc.cu) ----
 * Replaces gcm_linear_tc32 x 2 + gcm_act_backward + gcm_temporal_shift_sum + gcm_outer_reduce_tc32 x 2 of the window
 * backward (autograd through gcm.py:262-321, tests/test_gcm.py:412-439) for F = H1 = H2 = 32: tcgen05 / TMEM, 3xTF32,
 * one pass over the two operand streams.
 *   X = rows [sum_s x_{q-s} | x_q]      (gcm_temporal_gather, tiled = 1)         row = (position q, graph b), any order
 *   U = rows [sum_s dz2_{q+s} | dz2_q]  (gcm_temporal_shift_sum, sign = +1, tiled = 1); both cover ceil(rows / 128) whole
 *       tiles, rows past `rows` zero
 *   w1cat [32, 64] = [W_rel1 | W_root1],  w2tcat [32, 64] = [W_rel2^T | W_root2^T],  b1 [32],  act1 = GCM_ACT_*
 *   g1 [64, 32] += [dW_rel1^T ; dW_root1^T],  g2 [64, 32] += [dW_rel2 ; dW_root2],  db1 [32] +=,  db2 [32] +=
 *   dz1_out (optional) [rows, 32]: dL/d(pre-activation of layer 1) of every row
 *   workspace: gcm_temporal_window_bwd_workspace() floats.  Deterministic (per-CTA partials added in a fixed order). */
long long gcm_temporal_window_bwd_workspace(void);
/* debugging hook: device buffer [64][10] of int64 that later launches fill with per-tile phase clocks (NULL: off) */
int gcm_temporal_window_bwd_set_trace(long long* buf, int warp);
int gcm_temporal_window_bwd(const float* X, const float* U, long long rows, const float* w1cat, const float* b1,
                            const float* w2tcat, int act1, float* workspace, float* g1, float* g2, float* db1, float* db2,
                            float* dz1_out, void* stream);

/* ring/log + bitmasks -> the reference's hidden-state tensors (gcm.py:194-211 layout).
 * Any of nodes_out / adj_out / num_nodes_out may be NULL. */
int gcm_state_materialize(const gcm_dense_state* st, float* nodes_out, float* adj_out,
                          int64_t* num_nodes_out, void* stream);
/* nodes[b, (count[b] + offset) % C, :] = obs[b, :] -- the node write of gcm.py:274 on a log that shares the counters of
 * a state (the raw-observation log kept beside the preprocessed one when DenseGCM has a preprocessor, gcm.py:290-291).
 * offset = 0 before the step that advances the counters, -1 after it. */
int gcm_state_log_write(const gcm_dense_state* st, const float* obs, int offset, void* stream);
/* T node writes at once, after the T steps that advanced the counters: nodes[b, (count[b] - T + k) % C, :] = x_seq[b, k, :]
 * with x_seq[b, k, :] at b * stride_b + k * stride_t floats (the raw-observation log of a sequence call). */
int gcm_state_log_write_seq(const gcm_dense_state* st, const float* x_seq, long long stride_b, long long stride_t, int T,
                            void* stream);
/* dL/dnodes in log layout -> [B,N,F] reference layout (rows >= num_nodes are zero) */
int gcm_state_materialize_grad(const gcm_dense_state* st, const float* d_nodes, float* d_nodes_out,
                               void* stream);

/* the inverse: a caller-supplied hidden state (nodes[B,N,F] f32, adj[B,N,N] f32, num_nodes i64)
 * -> log layout with positions == logical indices.  Sets GCM_FLAG_UNCLEAN / GCM_FLAG_BADCOUNT. */
int gcm_state_ingest(const gcm_dense_state* st, const float* nodes_in, const float* adj_in,
                     const int64_t* num_nodes_in, int32_t* status, void* stream);

/* EuclideanEdge's distance (distance.py:48-49): dist[b, slot] = mean_p || cur[p] - nodes[b, slot] ||_2
 * over ALL n_cur current observations (cur [n_cur, F], already including this step's obs), for the
 * slots of the nodes currently in the window.  Runs BEFORE gcm_dense_step_fwd of the same step. */
int gcm_euclid_batchmean(const gcm_dense_state* st, const float* cur, int n_cur,
                         const float* dist_param, float* dist, void* stream);
/* The same distances through the tensor cores: ||c - n||^2 = |n|^2 + |c|^2 - 2 n.c with n.c in 3xTF32 (tcgen05; the
 * matmul form torch.cdist itself uses at these sizes).  F a multiple of 16, <= 64; scratch: gcm_euclid_tc_scratch(n_cur,
 * F) floats of device memory (the observations pre-split into K-major tiles). */
long long gcm_euclid_tc_scratch(int n_cur, int F);
int gcm_euclid_batchmean_tc(const gcm_dense_state* st, const float* cur, int n_cur, const float* dist_param,
                            float* scratch, float* dist, void* stream);

/* The selectors' own forward(nodes, adj_mats, edge_weights, num_nodes, B) on the reference's DENSE
 * tensors (edge_selectors/{temporal,dense,distance}.py): ORs 1s into adj [B,N,N] f32 in place.
 * nodes [B,N,F]; num_nodes i64 [B]. */
int gcm_select_dense(const float* nodes, float* adj, const int64_t* num_nodes, int B, int N, int F,
                     const gcm_selector* sel, void* stream);

/* ---- DenseEdge-only states ("ones" path, csrc/gcm_dense_ones.cu) ------------------------------------
 * On a state built by DenseEdge alone (edge_selectors/dense.py:11-23) the adjacency is the all-ones block over
 * the valid nodes, so the 2-layer stack of gcm.py:308 reduces to
 *   c = W_rel1 S + b1 (S = sum of the window's rows),  h_i = act1(c + W_root1 x_i),
 *   belief = act2(W_rel2 sum_i h_i + b2 + W_root2 h_t).
 * The caller keeps, next to the node log: xsum [B,F] (= S), the per-node cache [B,C,H1] (a function of R_i =
 * W_root1 x_i, see below) and, while training, per-step [K,B,.] buffers of the BPTT window.  The adjacency bit masks
 * are NOT maintained on this path; gcm_dense_fill_masks writes them when they are needed.
 * GCM_FLAG_NOTDENSE is set by gcm_state_ingest when the valid block of a caller-supplied adjacency is not all
 * ones; status[1] receives max(num_nodes) (status must then have two words). */
#define GCM_FLAG_NOTDENSE 32u
/* node write (gcm.py:274) + overflow eviction (gcm.py:323-355) + S update + num_nodes + 1.  xsum_in = S before the
 * step, xsum = S after it (may alias; a recording caller keeps one S per step of the BPTT window). */
int gcm_dense_ones_update(const gcm_dense_state* st, const float* obs, const float* xsum_in, float* xsum, void* stream);
/* S recomputed from the log */
int gcm_dense_ones_xsum(const gcm_dense_state* st, float* xsum, void* stream);
/* out[r,:Ho] = act(A1[r,:K1] W1^T + A2[r,:K2] W2^T + bias); W row-major [Ho,K] (torch.nn.Linear.weight); A2/W2 and
 * bias may be NULL; lda / ldo = row strides in floats.  Ho <= 128.  (lin_rel / lin_root of DenseGraphConv.)
 * status (or NULL): GCM_FLAG_NONFINITE is OR-ed in if an output is not finite; accumulate: out += result. */
int gcm_linear2(const float* A1, int K1, long long lda1, const float* W1, const float* A2, int K2, long long lda2,
                const float* W2, const float* bias, int act, long long rows, int Ho, float* out, long long ldo,
                int32_t* status, int accumulate, void* stream);
/* dW[o,i] += sum_r A[r,o] X[r,i];  db[o] += sum_r A[r,o] (db may be NULL).  Ho, Hi <= 128.  (weight gradients) */
int gcm_outer_reduce(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi, long long rows,
                     float* dW, float* db, void* stream);
/* The per-node cache [B,C,H1] (slot = position % C).  Element type float32 or bfloat16; content for act1 = tanh:
 * Q_i = exp(2 clamp(R_i, +-40)), so that tanh(c + R_i) = 1 - 2 / (E Q_i + 1) with E = exp(2 clamp(c, +-40)) costs one
 * MUFU; for relu / none: R_i itself.  gcm_linear2 with act = GCM_ACT_EXP2X writes E and Q.  H1 <= 128, H1 % 4 == 0
 * (bfloat16: H1 % 8 == 0). */
#define GCM_CACHE_F32 0
#define GCM_CACHE_BF16 1
#define GCM_ACT_EXP2X 3
/* After gcm_dense_ones_update: stores q_t (this step's node: Q or R, [B,H1] f32) as the new cache row, then
 * G = sum_i act1(c + R_i) over the window, h_t = act1(c + r_t) and, if P != NULL, P = sum_i act1'(c + R_i).
 * e_c [B,H1] = E (tanh) or c.  G, P, h_t: [B,H1] f32. */
int gcm_dense_ones_fwd(const gcm_dense_state* st, int H1, int act1, int cache_type, void* cache, const float* e_c,
                       const float* q_t, float* G, float* P, float* h_t, void* stream);
/* Backward of a whole BPTT window (the autograd of gcm.py:308 for every step of the chain at once; GCM has no
 * recurrence through the belief, so the steps' dL/dbelief are all known before any per-node work starts):
 * DZ[b, slot(p), :] = sum over chain steps k < steps_used that had node p in their window of
 * dG_k * act1'(c_k + R_p), plus dzo_k for the node written by step k; zero for every other slot.
 * steps_total = steps applied to the state since the chain started (count0 = count - steps_total), wE / wdG / wdzo:
 * [K, B, H1] f32 with `step_stride` floats between steps (E_k or c_k; dL/dG_k; own-node term times act1'(h_t)). */
int gcm_dense_ones_window_bwd(const gcm_dense_state* st, int H1, int act1, int cache_type, const void* cache,
                              int steps_total, int steps_used, const float* wE, const float* wdG, const float* wdzo,
                              long long step_stride, float* DZ, void* stream);
/* The DZ row of the node written by chain step k alone (dz [B,H1]): what dL/dx_k needs when observations require
 * grad; valid once steps k .. steps_used-1 have delivered dG / dzo. */
int gcm_dense_ones_node_bwd(const gcm_dense_state* st, int H1, int act1, int cache_type, const void* cache,
                            int steps_total, int steps_used, int k, const float* wE, const float* wdG, const float* wdzo,
                            long long step_stride, float* dz, void* stream);
/* ---- T steps at once (sequence entry; the caller is a loop like ray_gcm.py:200-202 that already holds all T
 * observations).  gcm_dense_ones_seq_update = T x gcm_dense_ones_update: x_seq[b, k, :] at b * stride_b + k * stride_t
 * floats; wS [T, B, F] receives S after every step (step_stride floats between steps), count += T.
 * gcm_dense_ones_window_fwd = T x gcm_dense_ones_fwd with the graph's cache rows read from HBM once: wE [T,B,H1]
 * (E_k or c_k), q_new (cache rows of the new nodes, float32: graph b, step k at b * q_stride_b + k * H1), outputs wG / wP
 * (NULL ok) / wht [T,B,H1].  A window may be taken in chunks (fewer rows in shared memory -> more resident CTAs):
 * steps_after = steps of the same gcm_dense_ones_seq_update call that come after this chunk.
 * Needs gcm_dense_ones_seq_smem(N, T, H1, cache_type) <= 220 KB of shared memory. */
int gcm_dense_ones_seq_update(const gcm_dense_state* st, const float* x_seq, long long stride_b, long long stride_t, int T,
                              const float* xsum_in, float* wS, long long step_stride, void* stream);
long long gcm_dense_ones_seq_smem(int N, int T, int H1, int cache_type);
int gcm_dense_ones_window_fwd(const gcm_dense_state* st, int H1, int act1, int cache_type, void* cache, int T,
                              int steps_after, const float* wE, const float* q_new, long long q_stride_b, float* wG,
                              float* wP, float* wht, long long step_stride, void* stream);
/* res = d_out * act'(out) elementwise (act' expressed through the activation's output) */
int gcm_act_backward(const float* d_out, const float* out, int act, long long n, float* res, void* stream);
/* The same with d_out read through element strides: d_out[t, b, :] at d_out + t * s_t + b * s_b (rows of H floats contiguous,
 * 16-byte aligned; a [B, T, H] gradient of the beliefs seen time-major is not copied); out and res contiguous [T, B, H]. */
int gcm_act_backward_strided(const float* d_out, long long s_t, long long s_b, const float* out, int act, int T, int B,
                             int H, float* res, void* stream);
/* per-step pieces of the backward, elementwise over n = B*H1: dht <- dzo = dht * act1'(h_t); dc = dG * P + dzo;
 * dcs = dc + dcs_next (suffix sum over the later steps; dcs / dcs_next may be NULL) */
int gcm_dense_ones_dc(const float* dG, float* dht, const float* P, const float* h_t, int act1, long long n, float* dc,
                      const float* dcs_next, float* dcs, void* stream);
/* float32 -> bfloat16 (cache refill) */
int gcm_to_bf16(const float* in, void* out, long long n, void* stream);
/* bf16 tensor-core (tcgen05) variants for callers that asked for bfloat16 compute (csrc/gcm_tc_gemm.cu); operands are
 * rounded to bfloat16, accumulation is float32.
 * gcm_linear_tc: out[r,:Ho] = epi(X[r,:K] W^T + bias), epi = none or GCM_ACT_EXP2X; out float32 or bfloat16 (out_bf16);
 * K, Ho multiples of 16 in [16,128]; ldx % 4 == 0; ldo % 4 == 0 (bfloat16: % 8).  (lin_root of DenseGraphConv over the
 * node log: the per-node cache fill.)
 * gcm_outer_reduce_tc: as gcm_outer_reduce (dW += A^T X, db += column sums of A), deterministic; Hi a multiple of 16;
 * workspace: gcm_outer_reduce_tc_workspace(rows) floats of device memory. */
int gcm_linear_tc(const float* X, int K, long long ldx, const float* W, const float* bias, int act, long long rows,
                  int Ho, void* out, long long ldo, int out_bf16, void* stream);
/* out[r,:Ho] = act(X1[r,:K1] W1^T + X2[r,:K2] W2^T + bias): the X1 product in 3xTF32 (fp32-accurate: hi/lo split, three
 * tcgen05 tf32 MMAs), the optional X2 product in bf16; act = GCM_ACT_* or GCM_ACT_EXP2X; status as gcm_linear2.
 * (lin_rel of layer 1 on the window sum S, and lin_rel + lin_root of layer 2 on (G, h_t): the two products of the ones
 * path whose rounding is common to every node of a graph.)  K1, K2, Ho multiples of 16 in [16,128]. */
int gcm_linear_tc32(const float* X1, int K1, long long ldx1, const float* W1, const float* X2, int K2, long long ldx2,
                    const float* W2, const float* bias, int act, long long rows, int Ho, float* out, long long ldo,
                    int32_t* status, void* stream);
long long gcm_outer_reduce_tc_workspace(long long rows);
int gcm_outer_reduce_tc(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi, long long rows,
                        float* workspace, float* dW, float* db, void* stream);
/* the same in 3xTF32 (fp32-accurate), same workspace */
int gcm_outer_reduce_tc32(const float* A, long long lda, int Ho, const float* X, long long ldx, int Hi, long long rows,
                          float* workspace, float* dW, float* db, void* stream);

/* Two such reductions against the same A in one pass over the rows (3xTF32): dW1 [Ho,Hi1] += A^T X1, dW2 [Ho,Hi2] += A^T X2,
 * db [Ho] += column sums of A.  Hi1, Hi2 multiples of 16, Hi1 + Hi2 <= 128.  The sparse GraphConv backward's weight
 * gradients dz^T [agg | x] (torch_geometric GraphConv lin_rel / lin_root; autograd of sparse_gcm.py:178,199). */
int gcm_outer_reduce_tc32_pair(const float* A, long long lda, int Ho, const float* X1, long long ldx1, int Hi1,
                               const float* X2, long long ldx2, int Hi2, long long rows, float* workspace,
                               float* dW1, float* dW2, float* db, void* stream);
/* writes the bit masks of the all-ones valid block (every node of the window linked to every node, self loops) */
int gcm_dense_fill_masks(const gcm_dense_state* st, void* stream);

/* ---- one distance selector with a per-node pre-activation cache ("zc" path, csrc/gcm_dense_zc.cu) ----
 * Same step as gcm_dense_step_fwd for a selector chain made of exactly one EuclideanEdge / CosineEdge /
 * SpatialEdge (edge_selectors/distance.py).  Those selectors only write row t of the adjacency, so the layer-1
 * pre-activation z_i of an existing node changes only when one of its in-neighbours leaves the window; the
 * caller keeps zcache [B, C, H1] float32 next to the node log and the step is two streaming passes per graph.
 * Valid only if EVERY step of this state went through this entry point under the SAME weights (the caller's
 * bookkeeping; otherwise use gcm_dense_step_fwd).  sel->dist as for gcm_dense_step_fwd (euclidean). */
int gcm_dense_step_fwd_zc(const gcm_dense_state* st, const float* obs, const gcm_selector* sel, const gcm_gnn* gnn,
                          float* zcache, float* belief, int32_t* status, void* stream);

/* ---- sparse path (sparse_gcm.py:72-212) -------------------------------------------------- */

/* Node write (sparse_gcm.py:111-123) + flat gather (util.py:426-452): nodes[b, T_b + k] = x[b, k]
 * for k < tau_b (x is [B, tmax, F], zero padded); flat[offsets[b] + k] = nodes[b, k] for
 * k < T_b + tau_b, offsets = exclusive cumsum(T + tau) (util.py:234-240), [B+1] int64.
 * `nodes` [B,N,F] is updated in place (the caller clones, like sparse_gcm.py:109). flat may be NULL. */
int gcm_sparse_write_flatten(float* nodes, const float* x, const int64_t* T, const int64_t* taus,
                             const int64_t* offsets, int B, int N, int F, int tmax, float* flat,
                             void* stream);

/* The same step out of place (sparse_gcm.py:109-123 in one pass): nodes_out [B,N,F] = nodes_in with the new rows written
 * (every row of nodes_out is written, so the caller needs no clone), flat as above.  nodes_in != nodes_out. */
int gcm_sparse_write_flatten_oop(const float* nodes_in, float* nodes_out, const float* x, const int64_t* T,
                                 const int64_t* taus, const int64_t* offsets, int B, int N, int F, int tmax,
                                 float* flat, void* stream);

/* Adjoint of gcm_sparse_write_flatten_oop in one pass (autograd of sparse_gcm.py:109-123): d_x [B,tmax,F] receives the
 * gradient of the rows that came from x (d_nodes_out[b, T_b + k] + d_flat[offsets[b] + T_b + k]; 0 for unwritten rows),
 * d_nodes_in [B,N,F] the gradient of the other rows (0 on the rows x overwrote).  d_nodes_out, d_flat, d_nodes_in, d_x
 * may each be NULL (zero / not wanted). */
int gcm_sparse_write_flatten_bwd(const float* d_nodes_out, const float* d_flat, const int64_t* T, const int64_t* taus,
                                 const int64_t* offsets, int B, int N, int F, int tmax, float* d_nodes_in, float* d_x,
                                 void* stream);

/* Fused edge selectors for the new nodes s in [T_b, T_b + tau_b): sources k < s with
 * (s - k in hops) [TemporalEdge, sparse_edge_selectors/temporal.py:19-63] OR
 * ||pos_s - pos_k||_2 < radius [SpatialRadiusEdge causal, sparse_edge_selectors/spatial.py:74-115;
 * pos = nodes[..., pos_start + c * pos_step], c < pos_len].  Edges come out as rows
 * (batch, sink, source) already coalesced: sorted, duplicates merged -- what the reference obtains
 * from three COO coalesces (sparse_gcm.py:132-139,152).  Two passes: edges == NULL writes the
 * in-degree of every new node to deg [n_new] (ordered by b, then s; new_off = exclusive cumsum of
 * taus, [B+1]); with edge_off = exclusive cumsum of deg ([n_new+1]) the second call fills
 * edges int64 [3, E] and, if flat_col != NULL, flat_col[e] = flat_off[b] + source (flat_off [B+1] = exclusive cumsum
 * of T + tau: the source's row in the flat node array of util.py:426-452, i.e. the CSR column of GraphConv).
 * One-search variant: pass 1 with hits != NULL also stores the first hit_cap sources of every new node (uint16
 * [n_new, hit_cap], ascending); gcm_sparse_expand_edges then writes edges / flat_col of every node with at most hit_cap
 * sources from those lists and edge_off, and the second (searching) call, given the same hit_cap (hits = NULL), only
 * handles the nodes whose list overflowed (not needed at all if max(deg) <= hit_cap). */
int gcm_sparse_build_edges(const float* nodes, const int64_t* T, const int64_t* taus, const int64_t* new_off,
                           int B, int N, int F, int tmax, const int32_t* hops, int n_hops, int use_radius,
                           int pos_start, int pos_step, int pos_len, float radius, int32_t* deg,
                           const int64_t* edge_off, int64_t* edges, int64_t E, const int64_t* flat_off, int64_t* flat_col, uint16_t* hits,
                           int hit_cap, void* stream);
int gcm_sparse_expand_edges(const int64_t* T, const int64_t* taus, const int64_t* new_off, int B, int tmax,
                            const uint16_t* hits, int hit_cap, const int64_t* edge_off, int64_t* edges, int64_t E,
                            const int64_t* flat_off, int64_t* flat_col, void* stream);

/* Which kernel evaluates the radius selector (process-wide).  AUTO: all-pairs test for small graphs, spatial
 * hash (cells of side `radius`, 3 x 3 neighbourhood, same float comparison) from N = 256.  Both produce the
 * same edges; the switch exists for parity tests and A/B profiling. */
typedef enum gcm_edge_builder { GCM_EB_AUTO = 0, GCM_EB_PAIRS = 1, GCM_EB_HASH = 2 } gcm_edge_builder;
int gcm_set_edge_builder(int which);

/* GraphConv over a CSR grouped by sink (torch_geometric.nn.GraphConv; call sites
 * ray_sparse_gcm.py:37-40, invoked at sparse_gcm.py:178,199):
 *   out[r] = act(W_rel (sum_{e in row i} w_e x[col[e]]) + b + W_root x[i]),  i = rows ? rows[r] : r
 * x [n, Fin]; rowptr int64 [n+1]; col int64 [E]; ew float [E] or NULL (== 1); rows int64 [m] or NULL
 * (all rows, m == n); wt = K-major pack [2 Fin, Fout] as in gcm_gnn.  Deterministic segmented
 * gather-reduce (no atomics).  agg_out [m, Fin] (or NULL) receives the aggregation for the backward. */
int gcm_sparse_graphconv_fwd(const float* x, const int64_t* rowptr, const int64_t* col, const float* ew,
                             const int64_t* rows, int64_t m, int Fin, int Fout, const float* wt,
                             const float* bias, int act, float* agg_out, float* out, void* stream);

/* Which kernel runs the GraphConv forward (process-wide).  AUTO: the tensor-core kernel (csrc/gcm_sparse_tc.cu: same
 * gathers, the [128 x 2 Fin] x [2 Fin x Fout] product of a tile in 3xTF32 on tcgen05) for Fin in {32, 64}, Fout a multiple
 * of 16 and at least 512 rows, else the CUDA-core kernel.  The switch exists for parity tests and A/B profiling. */
typedef enum gcm_graphconv_kernel { GCM_GC_AUTO = 0, GCM_GC_CUDA_CORES = 1, GCM_GC_TC = 2 } gcm_graphconv_kernel;
int gcm_set_graphconv_kernel(int which);
/* number of rows of x for the NEXT gcm_sparse_graphconv_fwd call that evaluates a row subset (`rows` != NULL): lets the
 * tensor-core kernel, which addresses x with 32-bit offsets, take that call too; 0 / not called: unknown. */
int gcm_sparse_graphconv_hint_rows(long long n);
/* block structure of x for the NEXT gcm_sparse_graphconv_fwd call that evaluates every row (`rows` == NULL): the rows of
 * graph g are node_off[g] .. node_off[g + 1] - 1 (device array of the n_graphs first rows, the last graph ends at m), no
 * graph has more than max_nodes rows and every edge stays inside its graph -- what util.flatten_adj (util.py) produces
 * from the per-batch-element adjacency.  Lets the call take the block-local kernel (a graph's rows staged in shared
 * memory by one bulk copy, gathers from shared memory). */
int gcm_sparse_graphconv_hint_blocks(const int64_t* node_off, int n_graphs, int max_nodes);

/* The transposed grouping the backward needs, for a block-diagonal graph: rowptr [n+1] / col [E] = CSR by sink over
 * the flat numbering, node_off [B+1] = first flat node of every graph (each graph has at most 8192 nodes and its
 * edges are contiguous); sink_local [E] (optional) = every edge's sink as an index inside its graph (row 1 of the
 * builder's edge list).  Writes t_rowptr [n+1] / t_col [E]: edges grouped by source, sinks ascending within a
 * source (deterministic).  Replaces a global argsort of the sources (the scatter backward of torch_geometric's
 * GraphConv, sparse_gcm.py:178,199). */
int gcm_sparse_csr_transpose(const int64_t* rowptr, const int64_t* col, const int64_t* node_off,
                             const int64_t* sink_local, int B, int64_t n_total, int64_t* t_rowptr, int64_t* t_col,
                             void* stream);

/* Test hook: entries of the shared-memory row buffer of the transposition kernel (0 = default); a small value forces
 * several source ranges per graph. */
int gcm_set_csr_transpose_cap(int entries);

/* Backward of the above.  t_rowptr [n+1] / t_col [E] / t_ew group the same edges by SOURCE node, with
 * t_col holding the position (in 0..m-1) of the edge's sink among the evaluated rows.  d_x [n, Fin]
 * must be zero on entry and receives dL/dx; d_agg [m, Fin] is scratch; weight gradients accumulate.
 * Optional (all three or none; used when rows == NULL): dz_scratch [m, Fout] and the transposed weights w_rel_t /
 * w_root_t [Fin, Fout] let the per-row products run as plain GEMMs over the m rows (3xTF32 on the tensor cores when
 * Fin and Fout are multiples of 16); outer_ws (optional, gcm_outer_reduce_tc_workspace(m) floats) moves the two weight
 * gradient reductions there as well. */
int gcm_sparse_graphconv_bwd(const float* x, const float* agg, const float* out, const float* d_out,
                             const int64_t* rows, int64_t m, int64_t n, const int64_t* t_rowptr,
                             const int64_t* t_col, const float* t_ew, int Fin, int Fout, const float* w_rel,
                             const float* w_root, int act, float* d_agg, float* d_x, float* d_w_rel,
                             float* d_w_root, float* d_b, float* dz_scratch,
                             const float* w_rel_t, const float* w_root_t, float* outer_ws, void* stream);

/* ---- RLlib state wire format of the sparse path (util.py:323-382; RaySparseGCM.forward, ray_sparse_gcm.py:195-213) ----
 * gcm_pack_edges = util.pack_hidden's adjacency part: coo int64 [3, E] (rows: batch, index 1, index 2) COALESCED (sorted
 * by batch), vals [E] -> dense_edges int64 [B, 2, max_edges] (rows = coo rows 1 and 2 of the graph's edges in order,
 * rest = edge_fill) and dense_weights [B, 1, max_edges] (rest = weight_fill); counts int32 [B] = edges per graph (the
 * caller asserts counts < max_edges like util.py:343-346).
 * gcm_unpack_edges = util.unpack_hidden's: the slots with dense_edges[b, 0, s] >= 0 (util.py:367) of every graph, in
 * (graph, slot) order, back to coo [3, E] / vals [E]; offsets int64 [B+1] = exclusive cumsum of the per-graph valid counts
 * (gcm_count_valid_edges), E = offsets[B]. */
int gcm_pack_edges(const int64_t* coo, const float* vals, long long E, int B, int max_edges, long long edge_fill,
                   float weight_fill, int64_t* dense_edges, float* dense_weights, int32_t* counts, void* stream);
int gcm_count_valid_edges(const int64_t* dense_edges, int B, int max_edges, int64_t* counts, void* stream);
int gcm_unpack_edges(const int64_t* dense_edges, const float* dense_weights, int B, int max_edges, const int64_t* offsets,
                     long long E, int64_t* coo, float* vals, void* stream);
/* sets *flag (device int32, |= 1) when any of the n floats is NaN or +-inf: the finite check of SparseGCM.forward
 * (sparse_gcm.py:203) as one pass.  x 16-byte aligned. */
int gcm_any_nonfinite(const float* x, long long n, int32_t* flag, void* stream);

/* Self-test of the tcgen05/TMEM building block of the tensor-core step kernels:
 * D[128,N] = A[128,K] B[N,K]^T, passes = 3 (3xTF32, fp32-accurate) or 1 (plain tf32).  Test hook only. */
int gcm_tc_selftest(const float* A, const float* B, float* D, int K, int N, int passes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCM_B200_H */

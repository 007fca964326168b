// Shared declarations of the pure-temporal step kernels (gcm_dense_fwd.cu, gcm_dense_fwd_tc.cu, gcm_dense_fwd_hc.cu).
#pragma once
#include "gcm_common.cuh"

constexpr int TP_MAXD = 48;   // distinct node offsets gathered per graph
constexpr int TP_MAXR = 8;    // rows of R1 (1 + number of forward hops)
constexpr int TP_MAXNB = 16;  // in-neighbours per row
constexpr int TP_THREADS = 256;
constexpr int TP_NW = TP_THREADS / 32;

struct TemporalProg {
  int nD;
  int doff[TP_MAXD];          // offsets from t of the distinct rows, doff[0] = 0
  int nR;
  int rd[TP_MAXR];            // offset of row r (rd[0] = 0; r >= 1 are the in-neighbours of t)
  int rD[TP_MAXR];            // index into doff of row r itself
  int nnb[TP_MAXR];
  int nb[TP_MAXR][TP_MAXNB];  // indices into doff of the in-neighbours of row r
  uint32_t nbmask[TP_MAXR];   // the same as a bitmask over doff indices (valid when nD <= 32)
  int n_past, past[GCM_MAX_HOPS];      // hops written into row t's past mask
  int n_future, future[GCM_MAX_HOPS];  // hops written into the future mask of row t - hop
};


constexpr int TW_G = 32;        // graphs per pipeline stage (one per producer lane)
constexpr int TW_STAGES = 3;
constexpr int TW_CONS = 11;     // consumer warps
constexpr int TW_MAXD = 12;     // distinct rows of the 2-hop in-neighbourhood (statically unrolled)
constexpr int TW_THREADS = (TW_CONS + 1) * 32;
constexpr int TW_MAXWIN = 16;   // rows of history staged per graph

struct TemporalWinArgs {
  gcm_dense_state st;
  const float* obs;
  gcm_gnn gnn;
  float* belief;
  int32_t* status;
  TemporalProg prog;
  int win;            // rows of history per graph = largest offset in prog.doff
  int uniform_count;  // >= 0: every graph has this count (host mirror), the counter is not read
  float* hcache;      // [B, hc_ring, 32] layer-1 output of the last hc_ring nodes (slot = position % hc_ring), or NULL
  int hc_ring;        // power of two
  int weights_stable; // GCM_STEP_WEIGHTS_STABLE: the weights were not written since the previous step of this state
  long long obs_ld;    // floats between the observation rows of consecutive graphs (F when contiguous); hc kernel only
  long long belief_ld; // floats between consecutive belief rows (H2 when contiguous); hc kernel only
  // hc kernel only: n_steps consecutive steps in ONE launch (the sequence entry); step k reads obs + k * obs_stride_t,
  // writes belief + k * belief_stride_t and sees uniform_count + k.  0 or 1 = a single step.
  int n_steps;
  long long obs_stride_t, belief_stride_t;
  // hc kernel only, F = 32: also write the layer-1 operand rows [sum of in-neighbour rows | own row] of every step into the
  // TILED operand buffer of the fused window backward (gcm_temporal_bwd_tc.cu), row index xrec_row0 + step * B + graph
  float* xrec;
  long long xrec_row0;
};


// tensor-core variants; each returns GCM_ERR_UNSUPPORTED when the shape does not fit
int gcm_launch_temporal_hc(const TemporalWinArgs& a, cudaStream_t stream);   // gcm_dense_fwd_hc.cu: cached layer-1 rows
int gcm_launch_temporal_tc(const TemporalWinArgs& a, cudaStream_t stream);   // gcm_dense_fwd_tc.cu: lane = (row, graph)
bool gcm_temporal_hc_shape_ok(const TemporalWinArgs& a);                     // may the cache be kept for this shape?

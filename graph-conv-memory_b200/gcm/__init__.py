"""gcm — B200-native drop-in for the hot path of proroklab/graph-conv-memory.

Same import paths as the reference package (`gcm.gcm.DenseGCM`, `gcm.edge_selectors.*`,
`gcm.sparse_gcm.SparseGCM`, `gcm.sparse_edge_selectors.*`, `gcm.util`); the memory update,
edge selection and GraphConv aggregation run in hand-written sm_100a CUDA kernels behind the
C ABI declared in include/gcm_b200.h.
"""
__version__ = "0.1.0"

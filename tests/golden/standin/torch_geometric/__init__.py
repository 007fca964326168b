"""Minimal CPU stand-in for the parts of `torch_geometric` the reference imports.

TEST INFRASTRUCTURE ONLY.  `torch_geometric` (pinned by the reference only as
`>= 1.7.0`, /root/reference/setup.cfg:22-26) is not installed in this image and
cannot be fetched.  This package restates the *published* definitions of the few
symbols the reference touches on the hot path, so that the UNMODIFIED reference
sources under /root/reference/src can be imported by `tests/golden/make_golden.py`
to generate golden vectors.  It is never imported by the product (`gcm`) package.

Symbols (call sites in the reference):
  nn.DenseGraphConv   README.md:56-57, tests/test_gcm.py:97
  nn.GraphConv        ray_sparse_gcm.py:37-40, tests/test_sparse_gcm.py:311
  nn.DenseGCNConv     tests/test_gcm.py:332 (only exercises GCM-owned logic)
  nn.Sequential       tests/test_gcm.py:20-25
  utils.coalesce      sparse_gcm.py:173
  utils.k_hop_subgraph sparse_gcm.py:192
"""
from . import nn, utils, data, transforms  # noqa: F401

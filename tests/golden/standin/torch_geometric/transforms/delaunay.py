"""Stand-in: imported by /root/reference/src/gcm/sparse_edge_selectors/spatial.py:8 (unused)."""


class Delaunay:  # pragma: no cover
    pass

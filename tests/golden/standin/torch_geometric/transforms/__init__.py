from . import delaunay  # noqa: F401

"""Stub for `from shapely.geometry import Polygon` (od_datasets.py:13)."""


class Polygon:
    def __init__(self, *a, **k):
        raise RuntimeError("shapely stub: GEOS is not available in this environment")

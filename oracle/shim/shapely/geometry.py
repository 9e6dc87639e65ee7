class LineString:
    def __init__(self, *a, **k):
        raise NotImplementedError("shapely stand-in: SSA_frequency is never called on the hot path")

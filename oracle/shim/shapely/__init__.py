"""Stand-in so `from shapely.geometry import LineString` (LeHaMoC_f.py:10) imports; dead code in the reference."""

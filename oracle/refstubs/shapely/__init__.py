"""Import stub for the absent `shapely` dependency (test infrastructure only)."""

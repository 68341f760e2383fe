"""Geometry utilities mirroring the reference's utils/ package (anchors, compute_overlap)."""

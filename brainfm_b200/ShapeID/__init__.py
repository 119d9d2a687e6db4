"""ShapeID: random Perlin shapes advected by a PDE-integrated, divergence-free velocity field
(mirror of the reference's ShapeID/ package, computed by libbfm)."""

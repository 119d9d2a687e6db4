"""BASELINE.json configs[2] and configs[3] at their FULL sizes (256^3 interpol, 192^3 ShapeID): the numpy oracle
cannot evaluate whole volumes of that size in seconds, so the cubic / nearest pulls are checked on a strided sample
of points against it, and the rest through size-independent properties (fused scaling-and-squaring == composed
form bit for bit, mass conservation of the advected shape, RHS-evaluation bookkeeping)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_interpol_256_against_the_oracle_on_sampled_points():
    import config_bench as cb
    r = cb.interpol_cfg(256, oracle_check=True)          # asserts fused == composed scaling and squaring inside
    chk = r["oracle_spot_check"]
    assert chk["points"] == 256 and chk["labels_equal"]
    assert chk["cubic_max_abs_err"] < 1e-4, chk           # float32 CUDA vs float64 numpy on 64-tap cubic sums
    assert set(r["ms"]) == {"scaling_and_squaring_7", "cubic_prefilter_4ch", "cubic_pull_4ch", "nearest_pull_labels"}


def test_shapeid_192_bookkeeping():
    import config_bench as cb
    r = cb.shapeid_cfg(192)
    assert r["steps"] >= 10 and r["rhs_evaluations"] == 2 + 6 * r["steps"] or r["rhs_evaluations"] > 6 * r["steps"]
    assert r["mass_drift"] < 0.2                          # upwind advection of a compact shape: mass stays bounded
    assert np.isfinite(r["total_ms"])

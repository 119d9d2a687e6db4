"""tests/golden/atlas_gca_L160_u8.npz: the benchmark label map of SURVEY.md 8d -- a 160^3 crop of round(files/gca.mgz)
(the only real volume the reference ships: a 256^3 float32 MGH atlas with integer values 0..233), centred on the
atlas' non-zero bounding box.  Test infrastructure: run in the build container, where /root/reference exists;
the GPU box only sees the committed fixture.

    python oracle/make_gca_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brainfm_b200 import io as bio     # the MGH reader under test (tests/test_readers.py pins it on this file too)

SRC = "/root/reference/files/gca.mgz"


def main():
    a = np.asarray(bio.load(SRC).get_fdata())
    assert a.shape == (256, 256, 256) and np.all(a == np.round(a)) and a.min() == 0 and a.max() == 233
    nz = np.nonzero(a)
    bb = [(int(x.min()), int(x.max()) + 1) for x in nz]
    assert bb == [(51, 203), (39, 215), (19, 203)], bb           # 152 x 176 x 184
    o = [min(max(0, (b[0] + b[1]) // 2 - 80), 256 - 160) for b in bb]
    L160 = np.round(a).astype(np.uint8)[o[0]:o[0] + 160, o[1]:o[1] + 160, o[2]:o[2] + 160]
    out = os.path.join(ROOT, "tests", "golden", "atlas_gca_L160_u8.npz")
    np.savez_compressed(out, L160=L160, origin=np.array(o), checksum=np.array([int(L160.astype(np.int64).sum())]))
    print(out, os.path.getsize(out), "bytes; non-zero fraction %.3f; %d distinct values" %
          ((L160 > 0).mean(), np.unique(L160).size))


if __name__ == "__main__":
    main()

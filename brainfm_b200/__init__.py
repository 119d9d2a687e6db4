"""brainfm_b200 -- B200-native (sm_100a) implementation of BrainFM's on-the-fly synthetic-data generator.

Python keeps the reference's API surface (Generator.build_datasets / BaseGen / BrainIDGen, the operator
registries, utils.interpol, ShapeID); the arithmetic runs in libbfm.so (hand-written CUDA behind the C ABI
declared in include/bfm.h).  There is no CPU fallback."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def build(force=False):
    from .build import build as _b
    return _b(force=force)

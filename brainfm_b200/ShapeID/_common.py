import ctypes as C

import torch

from .. import _lib


def stream():
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def need_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.BfmError("%s must be a CUDA tensor: brainfm_b200.ShapeID has no CPU path" % what)


def ivec(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def fvec(v):
    return (C.c_float * 3)(*[float(x) for x in v])

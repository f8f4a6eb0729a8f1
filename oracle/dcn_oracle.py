"""ctypes front-end of oracle/dcn_oracle.c (TEST INFRASTRUCTURE ONLY).

Mirrors the reference op signature ``modulated_deform_conv(x, offset, mask, weight,
bias, stride, padding, dilation, groups, deformable_groups)``
(/root/reference/basicsr/models/ops/dcn/deform_conv.py:111-185) on numpy / CPU-torch
fp32 arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile libdcn_oracle.so next to this file (gcc, no reference sources involved)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libdcn_oracle.so"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libdcn_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        _LIB.dcn_oracle_forward.argtypes = [fp] * 6 + [ctypes.c_int] * 12
        _LIB.dcn_oracle_forward.restype = ctypes.c_int
        _LIB.dcn_oracle_backward.argtypes = [fp] * 10 + [ctypes.c_int] * 12
        _LIB.dcn_oracle_backward.restype = ctypes.c_int
        _LIB.dcn_oracle_forward_ex.argtypes = [fp] * 6 + [ctypes.c_int] * 15
        _LIB.dcn_oracle_forward_ex.restype = ctypes.c_int
        _LIB.dcn_oracle_backward_ex.argtypes = [fp] * 10 + [ctypes.c_float] + [ctypes.c_int] * 15
        _LIB.dcn_oracle_backward_ex.restype = ctypes.c_int
    return _LIB


def _p(a):
    if a is None:
        return ctypes.POINTER(ctypes.c_float)()
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _c(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def out_hw(H, W, kh, kw, stride, pad, dil):
    return ((H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1,
            (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1)


def forward(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
            groups=1, deformable_groups=1):
    x, offset, mask, weight, bias = map(_c, (x, offset, mask, weight, bias))
    N, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho, Wo = out_hw(H, W, kh, kw, stride, padding, dilation)
    out = np.empty((N, Cout, Ho, Wo), np.float32)
    rc = _lib().dcn_oracle_forward(_p(x), _p(offset), _p(mask), _p(weight), _p(bias), _p(out),
                                   N, C, H, W, Cout, kh, kw, stride, padding, dilation,
                                   groups, deformable_groups)
    if rc != 0:
        raise ValueError(f"dcn_oracle_forward rc={rc}")
    return out


def backward(x, offset, mask, weight, grad_out, with_bias=True, stride=1, padding=0,
             dilation=1, groups=1, deformable_groups=1):
    """Returns (grad_x, grad_offset, grad_mask, grad_weight, grad_bias|None)."""
    x, offset, mask, weight, grad_out = map(_c, (x, offset, mask, weight, grad_out))
    N, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    gx = np.zeros_like(x)
    goff = np.zeros_like(offset)
    gmask = np.zeros_like(mask)
    gw = np.zeros_like(weight)
    gb = np.zeros((Cout,), np.float32) if with_bias else None
    rc = _lib().dcn_oracle_backward(_p(x), _p(offset), _p(mask), _p(weight), _p(grad_out),
                                    _p(gx), _p(goff), _p(gmask), _p(gw), _p(gb),
                                    N, C, H, W, Cout, kh, kw, stride, padding, dilation,
                                    groups, deformable_groups)
    if rc != 0:
        raise ValueError(f"dcn_oracle_backward rc={rc}")
    return gx, goff, gmask, gw, gb


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def forward_v1(x, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
    """DCNv1 (no modulation, no bias) — reference op ``deform_conv`` (deform_conv.py:12-57), per-axis parameters."""
    x, offset, weight = map(_c, (x, offset, weight))
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    N, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    out = np.empty((N, Cout, Ho, Wo), np.float32)
    rc = _lib().dcn_oracle_forward_ex(_p(x), _p(offset), _p(None), _p(weight), _p(None), _p(out), N, C, H, W, Cout,
                                      kh, kw, sh, sw, ph, pw, dh, dw, groups, deformable_groups)
    if rc != 0:
        raise ValueError(f"dcn_oracle_forward_ex rc={rc}")
    return out


def backward_v1(x, offset, weight, grad_out, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                scale=1.0):
    """Returns (grad_x, grad_offset, grad_weight) of DCNv1."""
    x, offset, weight, grad_out = map(_c, (x, offset, weight, grad_out))
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    N, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    gx, goff, gw = np.zeros_like(x), np.zeros_like(offset), np.zeros_like(weight)
    rc = _lib().dcn_oracle_backward_ex(_p(x), _p(offset), _p(None), _p(weight), _p(grad_out), _p(gx), _p(goff),
                                       _p(None), _p(gw), _p(None), ctypes.c_float(scale), N, C, H, W, Cout, kh, kw,
                                       sh, sw, ph, pw, dh, dw, groups, deformable_groups)
    if rc != 0:
        raise ValueError(f"dcn_oracle_backward_ex rc={rc}")
    return gx, goff, gw

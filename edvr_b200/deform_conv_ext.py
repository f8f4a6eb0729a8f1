"""B1 boundary: a drop-in for the compiled module ``basicsr.models.ops.dcn.deform_conv_ext``.

Exposes the five pybind11 names of /root/reference/basicsr/models/ops/dcn/src/deform_conv_ext.cpp:149-163 with the
same positional signatures and in-place ownership rules (caller allocates every output; grad_weight / grad_bias are
accumulated into; `ones` / `columns` are legacy scratch handles and are ignored), implemented over the C ABI of
libedvr_b200.so.  Register it BEFORE importing basicsr and the unmodified reference tree runs on the B200 kernels:

    import sys, edvr_b200.deform_conv_ext as ext
    sys.modules["basicsr.models.ops.dcn.deform_conv_ext"] = ext

Errors surface as RuntimeError with the library's message (the reference raises through TORCH_CHECK / AT_ERROR);
CPU tensors raise RuntimeError("... not implemented on CPU") like deform_conv_ext.cpp:123,145.
"""
import torch

from . import _lib as L


def _check_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what} is not implemented on CPU")


def _f32(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w,
                                  stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                  with_bias):
    _check_cuda(input, "modulated deform conv")
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")      # deform_conv_cuda.cpp:497
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")     # deform_conv_cuda.cpp:498
    if (stride_h, pad_h, dilation_h) != (stride_w, pad_w, dilation_w):
        raise RuntimeError("edvr_b200: anisotropic stride/pad/dilation is not supported (the v2 Python API passes one int)")
    if tuple(weight.shape[2:]) != (kernel_h, kernel_w):
        raise RuntimeError(f"Input shape and kernel shape wont match: ({kernel_h} x {kernel_w} vs "
                           f"{weight.shape[2]} x {weight.shape[3]}).")
    N, C, H, W = input.shape
    Cout = weight.shape[0]
    if C != weight.shape[1] * group:
        raise RuntimeError(f"Input shape and kernel channels wont match: ({C} vs {weight.shape[1] * group}).")
    x, w, off, m = _f32(input), _f32(weight), _f32(offset), _f32(mask)
    b = _f32(bias) if with_bias else None
    out32 = output if (output.dtype == torch.float32 and output.is_contiguous()) else torch.empty(
        output.shape, dtype=torch.float32, device=output.device)
    need = L.lib().eb_mdcn_forward_workspace(N, C, H, W, Cout, kernel_h, kernel_w)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=input.device)
    with torch.cuda.device(input.device):                           # at::DeviceGuard, deform_conv_cuda.cpp:499
        L.check(L.lib().eb_mdcn_forward(L.ptr(x), L.ptr(off), L.ptr(m), L.ptr(w), L.ptr(b), L.ptr(out32), N, C, H, W,
                                        Cout, kernel_h, kernel_w, stride_h, pad_h, dilation_h, group, deformable_group,
                                        L.ptr(ws), ws.numel(), L.stream_ptr()), "eb_mdcn_forward")
    if out32 is not output:
        output.copy_(out32)


def modulated_deform_conv_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight,
                                   grad_bias, grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h,
                                   stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    _check_cuda(input, "modulated deform conv")
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")
    if (stride_h, pad_h, dilation_h) != (stride_w, pad_w, dilation_w):
        raise RuntimeError("edvr_b200: anisotropic stride/pad/dilation is not supported")
    N, C, H, W = input.shape
    Cout = weight.shape[0]
    x, w, off, m, go = _f32(input), _f32(weight), _f32(offset), _f32(mask), _f32(grad_output)
    f32 = lambda t: t.dtype == torch.float32 and t.is_contiguous()
    tmp = {}

    def buf(name, t, accumulate=False):
        if f32(t):
            return t
        tmp[name] = (t, t.float().contiguous() if accumulate else torch.empty(t.shape, dtype=torch.float32, device=t.device))
        return tmp[name][1]

    gx, goff, gm = buf("gx", grad_input), buf("goff", grad_offset), buf("gm", grad_mask)
    gw = buf("gw", grad_weight, True)
    gb = buf("gb", grad_bias, True) if with_bias else None
    need = L.lib().eb_mdcn_backward_workspace(N, C, H, W, Cout, kernel_h, kernel_w, stride_h, pad_h, dilation_h)
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=input.device)
    with torch.cuda.device(input.device):
        L.check(L.lib().eb_mdcn_backward(L.ptr(x), L.ptr(off), L.ptr(m), L.ptr(w), L.ptr(go), L.ptr(gx), L.ptr(goff),
                                         L.ptr(gm), L.ptr(gw), L.ptr(gb), N, C, H, W, Cout, kernel_h, kernel_w,
                                         stride_h, pad_h, dilation_h, group, deformable_group, L.ptr(ws), ws.numel(),
                                         L.stream_ptr()), "eb_mdcn_backward")
    for dst, src in tmp.values():
        dst.copy_(src)


def _v1(*args, **kwargs):
    raise NotImplementedError("DCNv1 entry points are not on the EDVR hot path (DESIGN.md §7, SURVEY §8 row f3)")


deform_conv_forward = _v1
deform_conv_backward_input = _v1
deform_conv_backward_parameters = _v1

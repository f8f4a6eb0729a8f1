"""Frame staging either side of the EDVR forward, on the device: drop-ins for the two reference helpers the
video test loop calls around the network (`basicsr/models/video_base_model.py:44-70`):

    read_img_seq   basicsr/data/data_util.py:11-32   image files -> float32 tensor (t, c, h, w), RGB, [0, 1]
    tensor2img     basicsr/utils/img_util.py:36-98   float tensor (RGB) -> uint8 ndarray H x W x C (BGR)

Decoding stays with OpenCV like in the reference; everything after it - the /255 scaling, BGR<->RGB, HWC<->CHW, clamp,
round-half-to-even - runs in two bit-exact kernels (csrc/elementwise.cuh), so that the bytes crossing PCIe are the uint8
frames (a quarter of the fp32 tensors: 4.8 MB in and 11 MB out per 4 EDVR-L clips instead of 19.4 / 44.2 MB).
"""
import ctypes

import numpy as np
import torch

from . import _lib as L


def frames_to_tensor(frames_u8, bgr2rgb=True):
    """uint8 CUDA tensor [T, H, W, 3] (OpenCV order) -> float32 [T, 3, H, W] in [0, 1]: the arithmetic of read_img_seq
    after cv2.imread (`img.astype(np.float32) / 255.`, img2tensor(bgr2rgb=True, float32=True), stack)."""
    if not frames_u8.is_cuda:
        raise NotImplementedError("edvr_b200.img.frames_to_tensor: CUDA tensors only (no CPU fallback)")
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
        raise TypeError(f"expected a uint8 tensor [T, H, W, 3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
    f = frames_u8.contiguous()
    T, H, W, _ = f.shape
    out = torch.empty(T, 3, H, W, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        L.check(L.lib().eb_frames_u8_to_f32(L.ptr(f), L.ptr(out), T, H, W, 1 if bgr2rgb else 0, L.stream_ptr()),
                "eb_frames_u8_to_f32")
    return out


def read_img_seq(path, require_mod_crop=False, scale=1, device="cuda"):
    """Same arguments and result as the reference's read_img_seq (a folder or a list of image paths -> (t, c, h, w), RGB,
    [0, 1]) with the tensor produced on `device`: cv2 decodes, the uint8 frames go to the GPU, one kernel does the rest."""
    import cv2
    from os import path as osp
    if isinstance(path, list):
        img_paths = path
    else:
        import os
        img_paths = sorted(osp.join(path, f) for f in os.listdir(path) if not f.startswith(".") and osp.isfile(osp.join(path, f)))
    imgs = [cv2.imread(v) for v in img_paths]
    if require_mod_crop:                # mod_crop: basicsr/data/transforms.py:6-25
        imgs = [im[:im.shape[0] - im.shape[0] % scale, :im.shape[1] - im.shape[1] % scale] for im in imgs]
    host = torch.from_numpy(np.stack(imgs, 0)).pin_memory()
    return frames_to_tensor(host.to(device, non_blocking=True))


def tensor_to_bytes(t, rgb2bgr=True, min_max=(0, 1)):
    """The device half of tensor2img for a whole batch: float32 CUDA tensor [N, 3 / 1, H, W] (RGB) -> uint8 CUDA tensor
    [N, H, W, C] (BGR), each image exactly what tensor2img(t[i]) returns; the caller copies it to the host when and how it
    likes (asynchronously in a serving loop: bench.py `e2e_u8`)."""
    if not t.is_cuda:
        raise NotImplementedError("edvr_b200.img.tensor_to_bytes: CUDA tensors only (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] not in (1, 3):
        raise TypeError(f"expected a float32 tensor [N, 3 or 1, H, W], got {t.dtype} {tuple(t.shape)}")
    t = t.contiguous()
    N, C, H, W = t.shape
    out = torch.empty(N, H, W, C, dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        L.check(L.lib().eb_tensor2img_u8(L.ptr(t), L.ptr(out), N, C, H, W, 1 if rgb2bgr else 0, ctypes.c_float(min_max[0]),
                                         ctypes.c_float(min_max[1]), L.stream_ptr()), "eb_tensor2img_u8")
    return out


def _one(t, rgb2bgr, out_type, min_max):
    t = t.squeeze(0).float().detach()
    if t.dim() == 4:
        raise NotImplementedError("edvr_b200.img.tensor2img: a 4-D batch is tiled with make_grid by the reference; pass the "
                                  "images one by one")
    if t.dim() not in (2, 3):
        raise TypeError(f"Only support 4D, 3D or 2D tensor. But received with dimension: {t.dim()}")
    chw = t.unsqueeze(0) if t.dim() == 2 else t
    C, H, W = chw.shape
    if out_type == np.uint8 and C in (1, 3):
        chw = chw.contiguous()
        out = torch.empty(H, W, C, dtype=torch.uint8, device=t.device)
        with torch.cuda.device(t.device):
            L.check(L.lib().eb_tensor2img_u8(L.ptr(chw), L.ptr(out), 1, C, H, W, 1 if rgb2bgr else 0,
                                             ctypes.c_float(min_max[0]), ctypes.c_float(min_max[1]), L.stream_ptr()),
                    "eb_tensor2img_u8")
        img = out.cpu().numpy()
        return img[:, :, 0] if (C == 1) else img
    # float output (or an unusual channel count): the same clamp / normalise / transpose, values in [0, 1]
    x = (chw.clamp(*min_max) - min_max[0]) / (min_max[1] - min_max[0])
    if t.dim() == 2:
        img = x[0].cpu().numpy()
    else:
        img = x.permute(1, 2, 0).cpu().numpy()
        if C == 1:
            img = img[:, :, 0]
        elif rgb2bgr and C == 3:
            img = img[:, :, ::-1]
    if out_type == np.uint8:
        img = (img * 255.0).round()
    return np.ascontiguousarray(img.astype(out_type))


def tensor2img(tensor, rgb2bgr=True, out_type=np.uint8, min_max=(0, 1)):
    """Drop-in for basicsr.utils.tensor2img for CUDA tensors: a tensor or a list of tensors of shape (1 x) 3/1 x H x W or
    H x W (RGB) -> ndarray(s) H x W x C / H x W (BGR), uint8 in [0, 255] (default) or float in [0, 1].  The uint8 conversion
    happens on the device, so a quarter of the bytes are copied back."""
    if not (torch.is_tensor(tensor) or (isinstance(tensor, list) and all(torch.is_tensor(t) for t in tensor))):
        raise TypeError(f"tensor or list of tensors expected, got {type(tensor)}")
    ts = [tensor] if torch.is_tensor(tensor) else tensor
    for t in ts:
        if not t.is_cuda:
            raise NotImplementedError("edvr_b200.img.tensor2img: CUDA tensors only (no CPU fallback)")
    result = [_one(t, rgb2bgr, out_type, min_max) for t in ts]
    return result[0] if len(result) == 1 else result

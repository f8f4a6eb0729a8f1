"""oracle/img_ref.py — TEST INFRASTRUCTURE ONLY.

numpy restatement of the byte <-> float staging either side of the EDVR forward in the reference's test loop
(SURVEY §8 f2): the decoded frames of `read_img_seq` and the `tensor2img` of `VideoBaseModel.dist_validation`
(/root/reference/basicsr/models/video_base_model.py:44-70).  Pinned bit-exactly against the imported reference functions
by tests/golden/img_ref_import.npz (oracle/make_golden_img.py) and, in the build container, against the live import.
"""
import numpy as np


def frames_to_tensor(frames_u8):
    """What read_img_seq does with the frames cv2.imread decoded (basicsr/data/data_util.py:28-32 and img2tensor,
    basicsr/utils/img_util.py:22-27): uint8 [T, H, W, 3] BGR -> float32 [T, 3, H, W] RGB in [0, 1].
    `astype(np.float32) / 255.` is a correctly rounded fp32 division."""
    f = np.asarray(frames_u8)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[3] == 3
    x = f.astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(x[..., ::-1].transpose(0, 3, 1, 2))


def tensor2img(t, rgb2bgr=True, min_max=(0, 1)):
    """tensor2img for ONE float tensor [1, C, H, W] / [C, H, W] with C in (1, 3), or [H, W], out_type uint8
    (basicsr/utils/img_util.py:62-97): clamp to [min, max], (x - min) / (max - min) in fp32, CHW -> HWC (RGB -> BGR for 3
    channels), `(img * 255.0).round()` = round-half-to-even in fp32, astype(uint8).  A gray image loses its channel axis."""
    x = np.asarray(t, dtype=np.float32)
    if x.ndim == 4:
        assert x.shape[0] == 1, "the 4-D branch with a batch tiles the images with make_grid: out of scope"
        x = x[0]
    lo, hi = np.float32(min_max[0]), np.float32(min_max[1])
    x = (np.clip(x, lo, hi) - lo) / (hi - lo)
    if x.ndim == 3:
        x = x.transpose(1, 2, 0)
        if x.shape[2] == 1:
            x = x[:, :, 0]
        elif rgb2bgr:
            x = x[:, :, ::-1]
    return np.round(x * np.float32(255.0)).astype(np.uint8)

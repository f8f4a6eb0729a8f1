"""oracle/make_golden_img.py — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/img_ref_import.npz by importing the UNMODIFIED reference staging functions from /root/reference
(basicsr.utils.img_util.tensor2img / img2tensor and the arithmetic of basicsr.data.data_util.read_img_seq) in the build
container.  Inputs include values outside [0, 1], exact k + 0.5 products (round half to even) and NaN-free fp32 noise.
Run:  python -m oracle.make_golden_img
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "img_ref_import.npz")


def reference_functions():
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from basicsr.utils.img_util import img2tensor, tensor2img
    return img2tensor, tensor2img


def main():
    img2tensor, tensor2img = reference_functions()
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, size=(3, 10, 14, 3), dtype=np.uint8)
    frames[0, 0, :, :] = np.arange(14 * 3, dtype=np.uint8).reshape(14, 3) * 6       # a ramp incl. 0 and 252
    # read_img_seq: imgs = [cv2.imread(v).astype(np.float32) / 255.]; img2tensor(bgr2rgb=True, float32=True); stack
    imgs = [f.astype(np.float32) / 255. for f in frames]
    x = torch.stack(img2tensor(imgs, bgr2rgb=True, float32=True), dim=0).numpy()
    out = rng.normal(0.5, 0.5, size=(1, 3, 12, 18)).astype(np.float32)              # a third of the values outside [0, 1]
    halves = ((np.arange(12 * 18, dtype=np.float32) % 255) + 0.5) / 255.0           # products close to k + 0.5
    out[0, 1] = halves.reshape(12, 18)
    out[0, 2, 0, :4] = [0.0, 1.0, -0.0, 1.0000001]
    y = tensor2img([torch.from_numpy(out)])
    y_rgb = tensor2img(torch.from_numpy(out), rgb2bgr=False)
    gray = tensor2img(torch.from_numpy(out[:, :1]))
    y_pm1 = tensor2img(torch.from_numpy(out * 2 - 1), min_max=(-1, 1))
    np.savez_compressed(OUT, frames=frames, x=x, out=out, y=y, y_rgb=y_rgb, gray=gray, y_pm1=y_pm1)
    print("wrote", OUT, x.shape, y.shape, y.dtype, gray.shape)


if __name__ == "__main__":
    main()

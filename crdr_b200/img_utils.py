"""Image helpers with the reference's rounding semantics (src/utils/img_utils.py:17-42,79-132):
[-1,1] -> (x+1)/2*255 -> astype(uint8) TRUNCATES; PSNR is computed on the truncated uint8 values."""
import math

import numpy as np
import torch


def cvt_range_to_255(img):
    return (img + 1.0) / 2.0 * 255.0


def torch2npimg(img_rgb, out_mode="rgb"):
    img = img_rgb.detach().clone()
    if torch.max(img) <= 1.0:
        img = cvt_range_to_255(img)
    if img.dim() == 4:
        assert img.size(0) == 1, f"batch size must be 1, but {img.size(0)}"
        img = img.squeeze(0)
    arr = img.cpu().numpy().transpose(1, 2, 0)
    if out_mode == "bgr":
        arr = arr[..., ::-1]
    return arr.astype(np.uint8)


def imwrite(path, img_rgb):
    import cv2
    if isinstance(img_rgb, torch.Tensor):
        img = torch2npimg(img_rgb, out_mode="bgr")
    else:
        img = cvt_range_to_255(img_rgb) if np.max(img_rgb) <= 1.0 else img_rgb
        img = img[..., ::-1] if img.ndim == 3 else img
    cv2.imwrite(path, np.ascontiguousarray(img))


def calc_psnr(real, fake, data_range=255):
    assert data_range == 255
    if real.max() <= 1.0:
        real, fake = cvt_range_to_255(real), cvt_range_to_255(fake)
    if isinstance(real, torch.Tensor):
        real, fake = real.detach().cpu().numpy(), fake.detach().cpu().numpy()
    real = real.astype(np.uint8).astype(np.float32)
    fake = fake.astype(np.uint8).astype(np.float32)
    mse = float(np.mean((real - fake) ** 2))
    return 10.0 * math.log10(255.0 ** 2 / mse)


def calc_ms_ssim(real, fake):
    raise NotImplementedError("MS-SSIM needs pytorch_msssim, which is not part of the codec hot path")

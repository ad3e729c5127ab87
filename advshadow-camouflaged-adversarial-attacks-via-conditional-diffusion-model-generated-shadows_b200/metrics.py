"""Image-quality metrics of the reference's evaluation scripts (second half): FID
(fid_fast.py:22-45), SSIM / PSNR (PSNR_SSIM_fast.py:21-55).  Evaluation-side host code -- it runs once per
experiment on the images the sampler wrote, not on the sampling path -- written with torch ops so that it
follows its inputs' device.

Pinned / not pinned:
  * `calculate_fid` equals the reference's own function body on recorded activations
    (tests/golden/metrics.pt, minted by oracle/make_golden.py::metrics_cases).
  * SSIM / PSNR live in scikit-image in the reference (`skimage.metrics.structural_similarity`,
    `peak_signal_noise_ratio`; version unpinned), which is NOT installed in this image: the functions below
    restate the published algorithm (Wang et al. 2004 as scikit-image >= 0.19 implements it: Gaussian window
    sigma 1.5 truncated at 3.5 sigma = 11 taps, 'reflect' borders, sample covariance, K1 = 0.01, K2 = 0.03, mean
    over the image cropped by (win_size - 1) // 2, mean over channels) and are checked against an independent
    scipy.ndimage evaluation of the same definition -- parity with scikit-image itself is UNPINNED.
  * The Inception-v3 features need torchvision's pretrained weights (`inception_v3(pretrained=True)`,
    fid_fast.py:11); no weights ship with the reference or this repository, so `get_activations` takes the model.
"""
import os
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .datasets import IMAGE_EXTENSIONS


# ---- FID (fid_fast.py) ------------------------------------------------------------------
def inception_features_model(weights_path=None):
    """torchvision Inception-v3 with the classifier removed (fid_fast.py:11-13).  `weights_path`: a state_dict
    file of torchvision's ImageNet weights (there is no network here to download them)."""
    from torchvision.models import inception_v3
    model = inception_v3(weights=None, aux_logits=True, transform_input=False, init_weights=False)
    if weights_path is not None:
        model.load_state_dict(torch.load(weights_path, map_location="cpu"))
    model.fc = torch.nn.Identity()
    return model.eval()


def fid_preprocess(images: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] in [0,1] -> Resize((299,299)) + Normalize(ImageNet mean / std) (fid_fast.py:16-20).  The
    reference resizes PIL images (bilinear, antialiased); tensors take the same filter here."""
    x = F.interpolate(images.float(), size=(299, 299), mode="bilinear", antialias=True, align_corners=False)
    mean = torch.tensor([0.485, 0.456, 0.406], device=x.device).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


@torch.no_grad()
def get_activations(images: torch.Tensor, model, batch_size: int = 64) -> np.ndarray:
    """Pooled features of `model` for [N,3,H,W] images in [0,1] (fid_fast.py:23-27, in batches)."""
    dev = next(model.parameters()).device
    out = [model(fid_preprocess(images[i:i + batch_size].to(dev))).float().cpu() for i in range(0, len(images), batch_size)]
    return torch.cat(out).numpy()


def calculate_fid(act1: np.ndarray, act2: np.ndarray) -> float:
    """Frechet distance between two activation sets [N,D] (fid_fast.py:30-45): |mu1-mu2|^2 + tr(S1 + S2 - 2 sqrt(S1 S2)),
    float64, scipy.linalg.sqrtm with the imaginary part dropped."""
    from scipy import linalg
    (m1, s1), (m2, s2) = [(a.mean(axis=0), np.cov(a, rowvar=False)) for a in (np.asarray(act1), np.asarray(act2))]
    root = linalg.sqrtm(s1.dot(s2))
    root = root.real if np.iscomplexobj(root) else root
    return float(np.sum((m1 - m2) ** 2.0) + np.trace(s1 + s2 - 2.0 * root))


# ---- SSIM / PSNR (PSNR_SSIM_fast.py) ------------------------------------------------------
def _gaussian_taps(sigma: float, truncate: float, device, dtype):
    r = int(truncate * sigma + 0.5)
    x = torch.arange(-r, r + 1, device=device, dtype=torch.float64)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return (k / k.sum()).to(dtype), r


def _gaussian_filter(x: torch.Tensor, taps: torch.Tensor, r: int) -> torch.Tensor:
    """Separable correlation over the last two axes of [N,1,H,W] with scipy's 'reflect' border (edge sample repeated
    = torch's 'symmetric'; built by hand: F.pad's 'reflect' is scipy's 'mirror')."""
    def pad(t, dim):
        n = t.shape[dim]
        idx = torch.arange(-r, n + r, device=t.device)
        idx = torch.where(idx < 0, -idx - 1, idx)
        idx = torch.where(idx >= n, 2 * n - 1 - idx, idx)
        return t.index_select(dim, idx)

    x = F.conv2d(pad(x, 2), taps.view(1, 1, -1, 1))
    return F.conv2d(pad(x, 3), taps.view(1, 1, 1, -1))


def structural_similarity(image1: torch.Tensor, image2: torch.Tensor, win_size: int = 11, data_range=None,
                          sigma: float = 1.5) -> torch.Tensor:
    """Mean SSIM per image pair with Gaussian weights, channels-first: [N,C,H,W] x2 -> float64 [N] (the reference calls
    scikit-image with gaussian_weights=True, channel_axis=2, data_range = max - min of the FIRST image,
    PSNR_SSIM_fast.py:24).  As in scikit-image the Gaussian is always 11 taps (sigma 1.5 truncated at 3.5 sigma);
    `win_size` only sets the sample-covariance normalisation NP/(NP-1), NP = win_size^2, and the border crop."""
    if image1.shape != image2.shape or image1.dim() != 4:
        raise ValueError("structural_similarity: two [N,C,H,W] tensors of equal shape")
    if win_size % 2 == 0 or min(image1.shape[2:]) < win_size:
        raise ValueError("win_size must be odd and not exceed the image side")
    N, C, H, W = image1.shape
    a, b = image1.float(), image2.float()
    if data_range is None:
        data_range = (a.amax(dim=(1, 2, 3)) - a.amin(dim=(1, 2, 3)))
    R = torch.as_tensor(data_range, device=a.device, dtype=torch.float32).reshape(-1, 1, 1, 1).expand(N, C, 1, 1).reshape(N * C, 1, 1, 1)
    taps, r = _gaussian_taps(sigma, 3.5, a.device, torch.float32)
    x, y = a.reshape(N * C, 1, H, W), b.reshape(N * C, 1, H, W)
    ux, uy = _gaussian_filter(x, taps, r), _gaussian_filter(y, taps, r)
    uxx, uyy, uxy = _gaussian_filter(x * x, taps, r), _gaussian_filter(y * y, taps, r), _gaussian_filter(x * y, taps, r)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * R) ** 2, (0.03 * R) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    p = (win_size - 1) // 2
    return S[:, :, p:H - p, p:W - p].double().mean(dim=(1, 2, 3)).reshape(N, C).mean(dim=1)


def peak_signal_noise_ratio(image1: torch.Tensor, image2: torch.Tensor, data_range=None) -> torch.Tensor:
    """10 log10(R^2 / MSE) per image pair in float64, R = max - min of the first image (PSNR_SSIM_fast.py:25)."""
    if data_range is None:        # in the image's own dtype, as `image1.max() - image1.min()` is in the reference
        data_range = image1.amax(dim=(1, 2, 3)) - image1.amin(dim=(1, 2, 3))
    a, b = image1.double(), image2.double()
    R = torch.as_tensor(data_range, device=a.device, dtype=torch.float64).reshape(-1)
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    return 10 * torch.log10(R ** 2 / mse)


def calculate_ssim_psnr(image1, image2, win_size: int = 11) -> Tuple[float, float]:
    """One CHW image pair (numpy or tensor) -> (ssim, psnr), the reference's helper (PSNR_SSIM_fast.py:21-26)."""
    a = torch.as_tensor(np.asarray(image1) if not torch.is_tensor(image1) else image1)[None]
    b = torch.as_tensor(np.asarray(image2) if not torch.is_tensor(image2) else image2)[None]
    return float(structural_similarity(a, b, win_size=win_size)[0]), float(peak_signal_noise_ratio(a, b)[0])


def load_images_from_folder(folder: str, size: int = 64) -> torch.Tensor:
    """Every image file of `folder` in os.listdir order, RGB, Resize((size,size)) + ToTensor (PSNR_SSIM_fast.py:10-36)."""
    from ._compat import Image, transforms
    prep = transforms.Compose([transforms.Resize((size, size)), transforms.ToTensor()])
    files = [f for f in os.listdir(folder) if f.lower().endswith(IMAGE_EXTENSIONS)]
    return torch.stack([prep(Image.open(os.path.join(folder, f)).convert('RGB')) for f in files])


def compare_folders(folder1: str, folder2: str, win_size: int = 7) -> Tuple[float, float]:
    """Mean SSIM and mean PSNR over the paired images of two folders (PSNR_SSIM_fast.py:39-56)."""
    a, b = load_images_from_folder(folder1), load_images_from_folder(folder2)
    if len(a) != len(b):
        raise ValueError("Folders must contain the same number of images")
    return float(structural_similarity(a, b, win_size=win_size).mean()), float(peak_signal_noise_ratio(a, b).mean())


__all__ = ["inception_features_model", "fid_preprocess", "get_activations", "calculate_fid", "structural_similarity",
           "peak_signal_noise_ratio", "calculate_ssim_psnr", "load_images_from_folder", "compare_folders"]

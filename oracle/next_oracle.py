"""CPU restatements of the SURVEY.md §8f "next" rows that sit either side of the render path (the ORACLE).

TEST INFRASTRUCTURE ONLY: imported by tests/ (never by posegen_b200/).

* ``hmr_input``: the image hand-off of the PoseGen GAN loop (run_gan.py:2057-2071, 2326, 2433-2445):
  PNG quantisation, crop, /255, Normalize(mean, std), ``skimage.transform.resize(img, (3,R,R), anti_aliasing=True)``.
  scikit-image is a third-party dependency that is neither vendored in the reference nor pinned in its
  requirements.txt (it is imported at run_gan.py:2337) and is not installed here; its published algorithm
  (skimage/transform/_warps.py ``resize``, releases >= 0.19) is restated with the scipy.ndimage calls it makes:
  ``ndi.gaussian_filter(image, (scale-1)/2 per axis, mode='mirror')`` followed by
  ``ndi.zoom(filtered, 1/scale, order=1, mode='mirror', grid_mode=True)`` (skimage maps its default
  ``mode='reflect'`` to ndimage's ``'mirror'``).  Parity unpinned by reference fixtures (there are none).
* ``density_of_points``: RayCaster.render_pts_density (core/raycasters.py:597-648) on top of the render oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import render_oracle as orc


def hmr_input(image: np.ndarray, crop=(100, 100, 412, 412), out_res=224, mean=(0.485, 0.456, 0.406),
              std=(0.485, 0.456, 0.406), quantize_u8=True) -> np.ndarray:
    from scipy import ndimage as ndi
    x0, y0, x1, y1 = crop
    img = np.asarray(image, dtype=np.float32)
    if quantize_u8:
        img = (np.clip(img * 255.0, 0.0, 255.0)).astype(np.uint8).astype(np.float32)       # run_gan.py:2326 + cv2.imread
    else:
        img = img * 255.0
    img = img[y0:y1, x0:x1, :]                                                            # run_gan.py:2059
    img = np.transpose(img, (2, 0, 1)) / np.float32(255.0)                                # rgb_processing, run_gan.py:2443-2445
    img = (img - np.asarray(mean, np.float32)[:, None, None]) / np.asarray(std, np.float32)[:, None, None]
    img = img.astype(np.float64)                                                          # skimage converts to float64
    factors = np.array(img.shape, dtype=np.float64) / np.array([3, out_res, out_res], dtype=np.float64)
    sigma = np.maximum(0.0, (factors - 1.0) / 2.0)
    filtered = ndi.gaussian_filter(img, sigma, cval=0, mode="mirror")
    out = ndi.zoom(filtered, 1.0 / factors, order=1, mode="mirror", cval=0, grid_mode=True)
    return out.astype(np.float32)


def density_of_points(pts: torch.Tensor, skts: torch.Tensor, net: dict, emb: dict) -> torch.Tensor:
    """pts [N,3], skts [24,4,4] -> raw density [N] (alpha_linear of the trunk; no ReLU)."""
    n = pts.shape[0]
    sk = skts[None].expand(n, 24, 4, 4)
    enc = orc.encode(pts[:, None, :], torch.zeros_like(pts), sk, emb)          # view part unused by the density head
    raw = orc.nerf_forward(enc.reshape(n, -1), net, frame_code=orc.frame_codes(net, None, 1, n))   # the density head ignores it
    return raw[:, 3]

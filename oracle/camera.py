"""Oracle (test infrastructure): host camera math of the reference's splat adapter.

Follows reference models/modules/renderer/gaussian.py:30-66 and utils/camera_util.py:213-214.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import numpy as np
import torch


class RasterSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] f32, = E^T  (row-vector convention, gaussian.py:60)
    projmatrix: np.ndarray   # [4,4] f32, = E^T K_ndc^T (gaussian.py:61)
    campos: np.ndarray       # [3]


def raster_settings_from_KE(K, E, img_size) -> RasterSettings:
    """K [3,3], E [4,4] (numpy or torch, one frame), img_size (w,h)."""
    K = torch.as_tensor(np.asarray(K), dtype=torch.float32)
    E = torch.as_tensor(np.asarray(E), dtype=torch.float32)
    fx, fy = K[0, 0].item(), K[1, 1].item()          # gaussian.py:30
    px, py = K[0, 2].item(), K[1, 2].item()          # gaussian.py:31
    w, h = img_size
    tanfovx = math.tan(2 * math.atan(w / (2 * fx)) * 0.5)   # focal2fov + tan(fov/2), gaussian.py:33-36
    tanfovy = math.tan(2 * math.atan(h / (2 * fy)) * 0.5)
    znear, zfar = 0.001, 100
    K_ndc = torch.tensor([
        [2 * fx / w, 0, (2 * px - w) / w, 0],
        [0, 2 * fy / h, (2 * py - h) / h, 0],
        [0, 0, zfar / (zfar - znear), -zfar * znear / (zfar - znear)],
        [0, 0, 1, 0]]).float()                      # gaussian.py:41-46
    view = E.T.contiguous()
    proj = (E.T @ K_ndc.T).contiguous()
    campos = E.T.inverse()[3, :3]
    return RasterSettings(int(h), int(w), tanfovx, tanfovy, view.numpy().copy(), proj.numpy().copy(),
                          campos.numpy().copy())

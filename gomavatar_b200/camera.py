"""Camera setup on the device: replaces the host math, four ``.item()`` syncs and the H2D copy of reference
``models/modules/renderer/gaussian.py:30-47,60-61`` with one tiny kernel (``gom_camera_from_KE``)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomCameraArgs, call, ptr


def camera_from_KE(K, E, image_height, image_width, with_campos=False):
    """K [B,3,3], E [B,4,4] (device)  ->  viewmatrix [B,4,4] (= E^T), projmatrix [B,4,4] (= E^T K_ndc^T),
    tanfov [B,2] (, campos [B,3]) — the tensors the reference puts into GaussianRasterizationSettings."""
    if K.device.type != "cuda":
        raise _lib.GomError("camera_from_KE: inputs must live on a CUDA device (no CPU path exists)")
    B = K.shape[0]
    K, E = K.detach().contiguous().float(), E.detach().contiguous().float()
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=K.device)
    view, proj, tanfov = e(B, 4, 4), e(B, 4, 4), e(B, 2)
    campos = e(B, 3) if with_campos else None
    call("gom_camera_from_KE", GomCameraArgs(n_frames=B, height=int(image_height), width=int(image_width), K=ptr(K),
                                             E=ptr(E), viewmatrix=ptr(view), projmatrix=ptr(proj), tanfov=ptr(tanfov),
                                             campos=ptr(campos)))
    return (view, proj, tanfov, campos) if with_campos else (view, proj, tanfov)

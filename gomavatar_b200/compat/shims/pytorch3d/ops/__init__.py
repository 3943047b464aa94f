"""``pytorch3d.ops`` names the reference imports (models/model.py:22, utils/pc_util.py:7, shadow_module.py:6,
utils/network_util.py:8); none is called on the mesh-based training / evaluation path."""
from .knn import knn_points  # noqa: F401


def _absent(name):
    def f(*args, **kwargs):
        raise NotImplementedError(f"pytorch3d stand-in: ops.{name} is not provided (unused by GoMAvatar's mesh path)")
    f.__name__ = name
    return f


estimate_pointcloud_local_coord_frames = _absent("estimate_pointcloud_local_coord_frames")
estimate_pointcloud_normals = _absent("estimate_pointcloud_normals")
knn_gather = _absent("knn_gather")
ball_query = _absent("ball_query")


def interpolate_face_attributes(pix_to_face, barycentric_coords, face_attributes):
    """[N,H,W,K] face ids (-1 = none), [N,H,W,K,3] barycentrics, [F,3,D] -> [N,H,W,K,D] (0 where no face)."""
    import torch
    mask = pix_to_face < 0
    idx = pix_to_face.clamp(min=0)
    att = face_attributes[idx]                                    # [N,H,W,K,3,D]
    out = (barycentric_coords[..., None] * att).sum(dim=-2)
    return torch.where(mask[..., None], torch.zeros_like(out), out)

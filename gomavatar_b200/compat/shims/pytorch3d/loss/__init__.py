"""``pytorch3d.loss`` as train.py:26-30 imports it."""
from .chamfer import chamfer_distance  # noqa: F401


def mesh_normal_consistency(meshes):
    """1 - cos between the normals of every two faces sharing an edge, mean over the pairs (and the meshes of the batch,
    which share one topology here) — train.py:149.  Differentiable torch on the device of the mesh."""
    from gomavatar_b200 import regularizers as RG
    if meshes.isempty():
        return 0
    faces = meshes._faces
    cache = getattr(mesh_normal_consistency, "_cache", None)
    if cache is None or cache[0]() is not faces or cache[1] != faces._version:
        import weakref
        pairs = RG.all_face_pairs(faces, meshes.verts_padded().shape[1])
        cache = (weakref.ref(faces), faces._version, RG.normal_consistency_indices(faces, pairs))
        mesh_normal_consistency._cache = cache
    if cache[2][0].numel() == 0:
        return 0
    return RG.normal_consistency(meshes.verts_padded(), faces, None, cache[2])


def mesh_edge_loss(meshes, target_length=0.0):
    """mean over the edges of (|e| - target_length)^2, meshes weighted equally."""
    if meshes.isempty():
        return 0
    e = meshes.edges_packed()
    v = meshes.verts_packed()
    return (((v[e[:, 0]] - v[e[:, 1]]).norm(dim=1, p=2) - target_length) ** 2.0).mean()


def mesh_laplacian_smoothing(meshes, method="uniform"):
    from gomavatar_b200 import regularizers as RG
    if method != "uniform":
        raise NotImplementedError("pytorch3d stand-in: uniform Laplacian only")
    return RG.laplacian_smoothing(meshes.verts_padded(), meshes._faces)

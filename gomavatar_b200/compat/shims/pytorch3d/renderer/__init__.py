raise ImportError(
    "pytorch3d.renderer is not provided by the gomavatar_b200.compat stand-in: the reference's mesh normal renderer "
    "(models/modules/renderer/mesh.py) is replaced by gomavatar_b200.mesh_renderer.Renderer, which "
    "gomavatar_b200.model.Model builds from the same cfg node (compat.install(b200_model=True), the default)")

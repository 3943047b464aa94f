from gomavatar_b200.meshes import Meshes  # noqa: F401

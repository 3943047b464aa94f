"""Stand-in for ``trimesh`` (requirements.txt:9, absent offline): exactly what utils/pc_util.py:4-5,166-172 uses."""
import numpy as np

from . import grouping, remesh  # noqa: F401


class Trimesh:
    """``Trimesh(vertices, faces, vertex_attributes=...)`` with ``process=True`` semantics reduced to their effect on a
    clean mesh (none): coincident vertices, which trimesh would merge and renumber, are refused."""

    def __init__(self, vertices=None, faces=None, vertex_attributes=None, process=True, **kwargs):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)
        self.vertex_attributes = dict(vertex_attributes or {})
        if process and len(np.unique(self.vertices, axis=0)) != len(self.vertices):
            raise NotImplementedError("trimesh stand-in: mesh has coincident vertices (trimesh would merge them)")

    @property
    def edges(self):
        return remesh.faces_to_edges(self.faces)

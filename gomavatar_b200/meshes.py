"""``Meshes``: the slice of ``pytorch3d.structures.Meshes`` that GoMAvatar touches, in plain torch.

The reference hands PyTorch3D ``Meshes`` objects from ``Model.forward`` to its loss function
(``outputs['mesh']``, ``outputs['mesh_canonical']``: models/model.py:223-224,295-296 -> train.py:123-152,
utils/network_util.py:748-792) and builds its edge tables from them (models/model.py:115-134).  PyTorch3D 0.7.0 is a
third-party dependency absent here (README.md:21); the accessors below restate its published semantics
for a batch of N meshes that SHARE one face list — the only case the reference creates:

* packed tensors concatenate the meshes, faces offset by ``n * V``;
* ``edges_packed``: unique undirected edges ``(min, max)`` sorted by ``min * V_total + max``;
  ``faces_packed_to_edges_packed [F,3]``: column k = the edge OPPOSITE corner k (edges (1,2), (2,0), (0,1));
* ``verts_normals_*``: area-weighted face normals accumulated on the vertices, normalised with eps 1e-6;
* ``laplacian_packed``: sparse ``L = D^-1 A - I`` (uniform weights).

Topology (edges, the face->edge map, the Laplacian pattern) depends on ``faces`` only and is cached per face tensor, so a
training step pays for it once per (sub)division, not per call.  Everything is differentiable torch and runs wherever
the tensors live; the hot path never needs it (``regularizers.compute_loss`` uses ``gom_mesh_regularizers``) — it exists
so that reference code written against PyTorch3D keeps working (``gomavatar_b200.compat``)."""
from __future__ import annotations

import weakref

import torch

_TOPOLOGY = {}            # id(faces) -> (weakref(faces), version, n_verts, dict)


def _topology(faces, n_verts):
    """{'edges' [E,2], 'face_to_edge' [F,3]} of ONE mesh, cached for as long as the ``faces`` tensor lives unchanged."""
    key = id(faces)
    hit = _TOPOLOGY.get(key)
    if hit is not None and hit[0]() is faces and hit[1] == faces._version and hit[2] == n_verts:
        return hit[3]
    f = faces.long()
    e = torch.cat([f[:, [1, 2]], f[:, [2, 0]], f[:, [0, 1]]], dim=0)
    e, _ = torch.sort(e, dim=1)
    uniq, inv = torch.unique(e[:, 0] * n_verts + e[:, 1], sorted=True, return_inverse=True)
    topo = {"edges": torch.stack([uniq // n_verts, uniq % n_verts], dim=1), "face_to_edge": inv.reshape(3, -1).t().contiguous()}
    for k in [k for k, v in _TOPOLOGY.items() if v[0]() is None]:           # drop entries of dead tensors
        del _TOPOLOGY[k]
    _TOPOLOGY[key] = (weakref.ref(faces), faces._version, n_verts, topo)
    return topo


class Meshes:
    """``Meshes(verts, faces)``: verts ``[N,V,3]`` (or a list of N ``[V,3]``), faces ``[F,3]``, ``[N,F,3]`` or a list —
    all meshes must share the face list (checked only by shape; the reference passes ``faces[None]``)."""

    def __init__(self, verts, faces, **unused):
        if isinstance(verts, (list, tuple)):
            verts = torch.stack(list(verts))
        if isinstance(faces, (list, tuple)):
            faces = faces[0]
        if verts.dim() == 2:
            verts = verts[None]
        if faces.dim() == 3:
            faces = faces[0]
        if verts.dim() != 3 or verts.shape[-1] != 3 or faces.dim() != 2 or faces.shape[-1] != 3:
            raise ValueError("Meshes expects verts [N,V,3] and faces [F,3] / [N,F,3]")
        self._verts, self._faces = verts, faces
        self.device = verts.device

    # ------------------------------------------------------------------ sizes
    def __len__(self):
        return self._verts.shape[0]

    def isempty(self):
        return self._verts.shape[0] == 0 or self._verts.shape[1] == 0

    def num_verts_per_mesh(self):
        N, V = self._verts.shape[:2]
        return torch.full((N,), V, dtype=torch.int64, device=self.device)

    def num_faces_per_mesh(self):
        return torch.full((len(self),), self._faces.shape[0], dtype=torch.int64, device=self.device)

    def verts_packed_to_mesh_idx(self):
        N, V = self._verts.shape[:2]
        return torch.arange(N, device=self.device).repeat_interleave(V)

    # ------------------------------------------------------------------ geometry
    def verts_padded(self):
        return self._verts

    def verts_list(self):
        return list(self._verts)

    def verts_packed(self):
        return self._verts.reshape(-1, 3)

    def faces_padded(self):
        return self._faces.long()[None].expand(len(self), -1, -1)

    def faces_list(self):
        return [self._faces.long() for _ in range(len(self))]

    def _offsets(self, per_mesh):
        N, V = self._verts.shape[:2]
        off = (torch.arange(N, device=per_mesh.device) * V)[:, None, None]
        return (per_mesh.long()[None] + off).reshape(-1, per_mesh.shape[-1])

    def faces_packed(self):
        return self._offsets(self._faces)

    def edges_packed(self):
        return self._offsets(_topology(self._faces, self._verts.shape[1])["edges"])

    def faces_packed_to_edges_packed(self):
        t = _topology(self._faces, self._verts.shape[1])
        E = t["edges"].shape[0]
        off = (torch.arange(len(self), device=self._faces.device) * E)[:, None, None]
        return (t["face_to_edge"][None] + off).reshape(-1, 3)

    def verts_normals_padded(self):
        from .mesh_renderer import vertex_normals
        return vertex_normals(self._verts, self._faces)

    def verts_normals_packed(self):
        return self.verts_normals_padded().reshape(-1, 3)

    def laplacian_packed(self):
        """Sparse [NV,NV] uniform Laplacian: L_ij = 1/deg(i) for every edge (i,j), L_ii = -1 (pytorch3d.ops.laplacian)."""
        e = self.edges_packed()
        n = self._verts.shape[0] * self._verts.shape[1]
        idx = torch.cat([e.t(), e.flip(1).t()], dim=1)
        ones = torch.ones(idx.shape[1], dtype=torch.float32, device=self.device)
        deg = torch.zeros(n, dtype=torch.float32, device=self.device).index_add(0, idx[0], ones)
        w = torch.where(deg[idx[0]] > 0, 1.0 / deg[idx[0]], torch.zeros_like(ones))
        diag = torch.arange(n, device=self.device)
        L = torch.sparse_coo_tensor(torch.cat([idx, torch.stack([diag, diag])], dim=1),
                                    torch.cat([w, -torch.ones(n, dtype=torch.float32, device=self.device)]), (n, n))
        return L.coalesce()

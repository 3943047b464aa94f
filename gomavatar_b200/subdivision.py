"""Midpoint (1 -> 4) mesh subdivision, the host-side event of the reference's training schedule.

Mirrors reference ``utils/pc_util.py::subdivide`` / ``_subdivide`` (:49-172, adapted there from trimesh's
``remesh.subdivide``) as ``models/model.py::Model.subdivide`` (:136-179) uses it at ``train.py:341-346`` (iterations
``cfg.model.subdivide_iters``: 50 001 for ZJU-MoCap, 10 001 for PeopleSnapshot), ``train.py:275-279`` (replay on resume)
and ``eval.py:302-305``.  It runs a handful of times per training run on identical replicas, so it is plain numpy on the
host — not part of the per-frame hot path and no kernel.

Exact layout of the result (what parity means here; every rank and a resumed run must produce the same arrays):

* one new vertex per unique undirected edge, appended after the old vertices.  The reference numbers them in the order of
  ``trimesh.grouping.unique_rows(sorted_edges)`` — ``np.unique`` over the row hash ``min_vertex | max_vertex << 32`` —
  i.e. ascending by (max vertex, min vertex).  trimesh itself is a third-party dependency absent from the reference tree
  (``requirements.txt:9``, unpinned); its two helpers (``faces_to_edges``, ``unique_rows``) are restated from its
  published source, while the reference's own ``_subdivide`` has been executed on top of them to pin this file
  (``tests/golden/golden_subdivide.npz``, ``oracle/make_golden.py::subdivide_golden``).
* face ``f = (a, b, c)`` with midpoints ``m0 = ab, m1 = bc, m2 = ca`` becomes rows ``4f .. 4f+3`` =
  ``(a, m0, m2), (m0, b, m1), (m2, m1, c), (m0, m1, m2)`` — same winding — so per-face parameters are repeated 4x in place
  (``x[..., None].repeat(1, 1, 4)``, model.py:159-171).
* midpoint position = mean of the two end points (float64 mean of fp32 values, rounded once to fp32 — bit-identical to the
  reference's path through ``trimesh.Trimesh.vertices``); midpoint vertex attributes ('weights': LBS rows) = mean of the
  two end points in the attribute's own dtype; the reference's special cases 'so3' (zeros) and 'scale' (edge length of
  the attribute values) are kept although ``Model.subdivide`` only passes 'weights'.

``trimesh.Trimesh(vertices, faces)`` is constructed with ``process=True`` in the reference, which merges vertices with
identical positions before subdividing; a canonical mesh never has any, and one that does is refused here instead of
silently changing the vertex count.
"""
from __future__ import annotations

import numpy as np


def faces_to_edges(faces):
    """trimesh.geometry.faces_to_edges: the 3 directed edges (0->1, 1->2, 2->0) of every face, face-major: [3F, 2]."""
    faces = np.asarray(faces)
    return faces[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)


def unique_rows(rows):
    """trimesh.grouping.unique_rows for a 2-column non-negative integer array: (index of the first occurrence of every
    distinct row, inverse), distinct rows ordered by the 64-bit hash ``col0 ^ (col1 << 32)``."""
    rows = np.asarray(rows).astype(np.int64)
    if rows.size and (rows.min() < 0 or rows.max() >= 2 ** 31):
        raise ValueError("vertex indices must be in [0, 2^31)")
    key = np.bitwise_xor(rows[:, 0], rows[:, 1] << 32)
    _, unique, inverse = np.unique(key, return_index=True, return_inverse=True)
    return unique, inverse.reshape(-1)


def subdivide_mesh(vertices, faces, attributes=None, return_edges=False):
    """vertices [V,3], faces [F,3] int, attributes {name: [V,d]} -> (new_vertices [V+E,3] float64, new_faces [4F,3],
    new_attributes, (edges [12F,2] if return_edges), index_dict {old face: its 4 new faces}) — the return value of the
    reference's ``subdivide`` (pc_util.py:166-172)."""
    vertices = np.asarray(vertices, dtype=np.float64)
    faces = np.asarray(faces)
    if faces.ndim != 2 or faces.shape[1] != 3 or vertices.ndim != 2 or vertices.shape[1] != 3:
        raise ValueError("subdivide_mesh expects vertices [V,3] and faces [F,3]")
    if faces.size and (faces.min() < 0 or faces.max() >= len(vertices)):
        raise ValueError("faces index vertices out of range")
    if len(np.unique(vertices, axis=0)) != len(vertices):
        raise ValueError("mesh has duplicate vertex positions: the reference (trimesh, process=True) would merge them and "
                         "renumber the vertices; clean the canonical mesh first")
    edges = np.sort(faces_to_edges(faces), axis=1)
    unique, inverse = unique_rows(edges)
    mid = vertices[edges[unique]].mean(axis=1)
    mid_idx = inverse.reshape(-1, 3) + len(vertices)
    new_faces = np.column_stack([faces[:, 0], mid_idx[:, 0], mid_idx[:, 2],
                                 mid_idx[:, 0], faces[:, 1], mid_idx[:, 1],
                                 mid_idx[:, 2], mid_idx[:, 1], faces[:, 2],
                                 mid_idx[:, 0], mid_idx[:, 1], mid_idx[:, 2]]).reshape(-1, 3)
    new_vertices = np.vstack((vertices, mid))
    index_dict = {int(k): v for k, v in zip(range(len(faces)), np.arange(4 * len(faces)).reshape(-1, 4))}

    new_attributes = {}
    for key, values in (attributes or {}).items():
        values = np.asarray(values)
        if len(values) != len(vertices):
            raise ValueError(f"attribute {key!r} has {len(values)} rows for {len(vertices)} vertices")
        if key == "so3":
            attr_mid = np.zeros([unique.shape[0], 3], values.dtype)
        elif key == "scale":
            edge_len = np.linalg.norm(values[edges[unique][:, 1]] - values[edges[unique][:, 0]], axis=-1)
            attr_mid = np.ones([unique.shape[0], 3], values.dtype) * edge_len[..., None]
        else:
            attr_mid = values[edges[unique]].mean(axis=1)
        new_attributes[key] = np.vstack((values, attr_mid))
    if return_edges:
        return new_vertices, new_faces, new_attributes, faces_to_edges(new_faces), index_dict
    return new_vertices, new_faces, new_attributes, index_dict


def subdivided_sizes(n_verts, n_faces, levels=1):
    """(V, F) after ``levels`` subdivisions of a closed manifold mesh: V' = V + E, F' = 4F, E = 3F/2."""
    for _ in range(levels):
        n_verts, n_faces = n_verts + 3 * n_faces // 2, 4 * n_faces
    return n_verts, n_faces

"""Deterministic synthetic scenes in the reference's input contract.

No SMPL model files, ZJU-MoCap frames or checkpoints exist offline (probe table), so the
hot path is exercised on a seeded SMPL-topology-like humanoid: closed triangle mesh with exactly ``F``
faces (= Gaussians), a 24-joint skeleton on the SMPL kinematic tree, <=4-nnz skinning weights laid out
``[25, V]`` like ``Model.lbs_weights`` (reference ``models/model.py:62-72``, row 24 = background), body
poses as 72-d axis-angle vectors and ZJU-like pinhole cameras.

Everything here is host-side numpy (the reference does the same work in its DataLoader worker:
``dataset/train.py:209-287``); the per-frame dictionaries use the reference's keys and shapes
(``K[3,3] E[4,4] cnl_gtfms[24,4,4] dst_Rs[24,3,3] dst_Ts[24,3] dst_posevec[69] bgcolor[3]``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# SMPL kinematic tree (same topology as reference utils/body_util.py:36-39; parent[i] < i).
SMPL_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int32)
N_JOINTS = 24


def make_skeleton() -> np.ndarray:
    """Anthropometric T-pose joints ``[24,3]`` (metres; +Y up, +X subject-left, +Z front), height ~1.7 m."""
    j = np.zeros((N_JOINTS, 3), dtype=np.float64)
    j[0] = (0.0, 0.0, 0.0)            # pelvis
    j[1] = (0.07, -0.09, 0.0)         # l hip
    j[2] = (-0.07, -0.09, 0.0)        # r hip
    j[3] = (0.0, 0.11, -0.02)         # spine1
    j[4] = (0.10, -0.47, 0.0)         # l knee
    j[5] = (-0.10, -0.47, 0.0)        # r knee
    j[6] = (0.0, 0.25, 0.0)           # spine2
    j[7] = (0.09, -0.87, -0.03)       # l ankle
    j[8] = (-0.09, -0.87, -0.03)      # r ankle
    j[9] = (0.0, 0.31, 0.02)          # spine3
    j[10] = (0.10, -0.93, 0.09)       # l foot
    j[11] = (-0.10, -0.93, 0.09)      # r foot
    j[12] = (0.0, 0.52, -0.02)        # neck
    j[13] = (0.08, 0.43, -0.01)       # l collar
    j[14] = (-0.08, 0.43, -0.01)      # r collar
    j[15] = (0.0, 0.60, 0.02)         # head
    j[16] = (0.19, 0.46, -0.02)       # l shoulder
    j[17] = (-0.19, 0.46, -0.02)      # r shoulder
    j[18] = (0.45, 0.46, -0.03)       # l elbow
    j[19] = (-0.45, 0.46, -0.03)      # r elbow
    j[20] = (0.70, 0.46, -0.03)       # l wrist
    j[21] = (-0.70, 0.46, -0.03)      # r wrist
    j[22] = (0.78, 0.46, -0.03)       # l hand
    j[23] = (-0.78, 0.46, -0.03)      # r hand
    return j.astype(np.float32)


# Body parts: (point a, point b, radius across, radius depth). Points are given in terms of joints.
def _parts(j: np.ndarray):
    J = j.astype(np.float64)
    up = np.array([0.0, 1.0, 0.0])
    fwd = np.array([0.0, 0.0, 1.0])
    return [
        ("head", J[15] + 0.02 * up, J[15] + 0.16 * up, 0.085, 0.095),
        ("neck", J[12] - 0.02 * up, J[15] + 0.02 * up, 0.050, 0.050),
        ("chest", J[6] - 0.02 * up, J[9] + 0.12 * up, 0.175, 0.110),
        ("belly", J[0] - 0.02 * up, J[6] + 0.02 * up, 0.150, 0.105),
        ("hips", J[0] - 0.10 * up - 0.08 * np.array([1.0, 0, 0]), J[0] - 0.10 * up + 0.08 * np.array([1.0, 0, 0]), 0.100, 0.105),
        ("l_upper_arm", J[16], J[18], 0.047, 0.047),
        ("r_upper_arm", J[17], J[19], 0.047, 0.047),
        ("l_forearm", J[18], J[20], 0.038, 0.038),
        ("r_forearm", J[19], J[21], 0.038, 0.038),
        ("l_hand", J[20] + 0.03 * np.array([1.0, 0, 0]), J[22] + 0.06 * np.array([1.0, 0, 0]), 0.040, 0.020),
        ("r_hand", J[21] - 0.03 * np.array([1.0, 0, 0]), J[23] - 0.06 * np.array([1.0, 0, 0]), 0.040, 0.020),
        ("l_thigh", J[1], J[4], 0.078, 0.078),
        ("r_thigh", J[2], J[5], 0.078, 0.078),
        ("l_calf", J[4], J[7], 0.052, 0.052),
        ("r_calf", J[5], J[8], 0.052, 0.052),
        ("l_foot", J[7] - 0.04 * up - 0.03 * fwd, J[10] - 0.01 * up + 0.09 * fwd, 0.045, 0.035),
        ("r_foot", J[8] - 0.04 * up - 0.03 * fwd, J[11] - 0.01 * up + 0.09 * fwd, 0.045, 0.035),
    ]


def _capsule(a, b, r0, r1, S, R):
    """Closed lat-long capsule from a to b with elliptical section (r0 across, r1 depth).

    R rings of S vertices + 2 poles  ->  V = R*S+2, F = 2*S*R (genus 0, consistently outward-wound).
    """
    axis = b - a
    L = float(np.linalg.norm(axis))
    w = axis / L
    ref = np.array([0.0, 0.0, 1.0]) if abs(w[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    u = np.cross(ref, w)
    u /= np.linalg.norm(u)          # "across" direction
    v = np.cross(w, u)              # "depth" direction
    rc = 0.5 * (r0 + r1)            # cap radius along the axis
    arc = math.pi * rc + L          # profile arc length pole-to-pole
    verts = [a - w * rc]
    for i in range(1, R + 1):
        s = arc * i / (R + 1)
        if s < 0.5 * math.pi * rc:                       # bottom cap
            th = s / rc
            ax, rad = -rc * math.cos(th), math.sin(th)
        elif s < 0.5 * math.pi * rc + L:                 # cylinder
            ax, rad = s - 0.5 * math.pi * rc, 1.0
        else:                                            # top cap
            th = (s - 0.5 * math.pi * rc - L) / rc
            ax, rad = L + rc * math.sin(th), math.cos(th)
        # half-segment twist on alternate rings gives near-isotropic triangles
        ph = (np.arange(S) + 0.5 * (i & 1)) * (2.0 * math.pi / S)
        ring = a[None] + w[None] * ax + rad * (r0 * np.cos(ph)[:, None] * u[None] + r1 * np.sin(ph)[:, None] * v[None])
        verts.extend(ring)
    verts.append(b + w * rc)
    verts = np.asarray(verts, dtype=np.float64)
    top = R * S + 1
    faces = []
    ring0 = lambda i: 1 + (i - 1) * S
    for k in range(S):
        faces.append((0, ring0(1) + (k + 1) % S, ring0(1) + k))
        faces.append((top, ring0(R) + k, ring0(R) + (k + 1) % S))
    for i in range(1, R):
        lo, hi = ring0(i), ring0(i + 1)
        for k in range(S):
            k1 = (k + 1) % S
            if i & 1:   # lower ring is twisted by half a segment
                faces.append((lo + k, lo + k1, hi + k1))
                faces.append((lo + k, hi + k1, hi + k))
            else:
                faces.append((lo + k, lo + k1, hi + k))
                faces.append((lo + k1, hi + k1, hi + k))
    return verts, np.asarray(faces, dtype=np.int64)


def _plan_resolution(parts, n_faces):
    """Choose (S_i, R_i) per part: near-uniform edge length and sum(2 S_i R_i) == n_faces exactly."""
    assert n_faces % 2 == 0 and n_faces >= 2000, "n_faces must be even and >= 2000"
    dims = []
    for _, a, b, r0, r1 in parts:
        L = float(np.linalg.norm(b - a))
        rc = 0.5 * (r0 + r1)
        per = math.pi * (3 * (r0 + r1) - math.sqrt((3 * r0 + r1) * (r0 + 3 * r1)))  # ellipse perimeter
        dims.append((per, math.pi * rc + L))
    area = sum(p * l for p, l in dims)
    e = math.sqrt(area / (n_faces / 2.0) / 0.866)  # rows are ~0.866 e apart
    for attempt in range(400):
        ee = e * (1.0 + 0.0005 * ((attempt + 1) // 2) * (1 if attempt & 1 else -1))
        S = [max(6, int(round(p / ee))) for p, _ in dims]
        R = [max(3, int(round(l / (0.866 * ee)))) for _, l in dims]
        D = n_faces // 2 - sum(s * r for s, r in zip(S, R))
        order = np.argsort([-s * r for s, r in zip(S, R)])[:6]
        best = None
        rng = range(-4, 5)
        import itertools
        for dl in itertools.product(rng, repeat=len(order)):
            if sum(S[i] * d for i, d in zip(order, dl)) == D:
                cost = sum(abs(d) for d in dl)
                if best is None or cost < best[0]:
                    best = (cost, dl)
        if best is not None:
            for i, d in zip(order, best[1]):
                R[i] += d
            assert 2 * sum(s * r for s, r in zip(S, R)) == n_faces
            return S, R
    raise RuntimeError(f"could not plan a mesh with exactly {n_faces} faces")


@dataclass
class Scene:
    vertices: np.ndarray      # [V,3] f32 canonical (T-pose)
    faces: np.ndarray         # [F,3] i64
    lbs_weights: np.ndarray   # [25,V] f32 (row 24 = background, zeros)
    joints: np.ndarray        # [24,3] f32 canonical joints
    cnl_gtfms: np.ndarray     # [24,4,4] f32

    @property
    def n_vertices(self):
        return self.vertices.shape[0]

    @property
    def n_faces(self):
        return self.faces.shape[0]

    def canonical_info(self):
        """Same keys as the reference's ``dataset.get_canonical_info()`` (dataset/train.py:289-302)."""
        return {
            "canonical_vertex": self.vertices,
            "canonical_lbs_weights": self.lbs_weights[:-1].T.copy(),   # [V,24]; the model appends the bg row
            "faces": self.faces,
            "canonical_joints": self.joints,
        }


def _segment_dist(p, a, b):
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-12), 0.0, 1.0)
    return np.linalg.norm(p - (a[None] + t[:, None] * ab[None]), axis=1)


def make_lbs_weights(vertices, joints, k=4, sigma=0.05):
    """[25,V] weights: k nearest bones, exp(-d^2/2 sigma^2), normalised; row 24 (background) = 0."""
    V = vertices.shape[0]
    P = vertices.astype(np.float64)
    J = joints.astype(np.float64)
    d = np.full((N_JOINTS, V), 1e9)
    children = {i: [c for c in range(N_JOINTS) if SMPL_PARENTS[c] == i] for i in range(N_JOINTS)}
    for jn in range(N_JOINTS):
        if children[jn]:
            for c in children[jn]:
                d[jn] = np.minimum(d[jn], _segment_dist(P, J[jn], J[c]))
        else:  # leaf: short stub continuing the parent bone
            dirn = J[jn] - J[SMPL_PARENTS[jn]]
            dirn = dirn / max(np.linalg.norm(dirn), 1e-9)
            d[jn] = _segment_dist(P, J[jn], J[jn] + 0.08 * dirn)
    dmin = d.min(axis=0, keepdims=True)
    w = np.exp(-(d ** 2 - dmin ** 2) / (2 * sigma ** 2))
    kth = np.sort(w, axis=0)[-k][None]
    w = np.where(w >= kth, w, 0.0)
    w /= w.sum(axis=0, keepdims=True)
    out = np.zeros((N_JOINTS + 1, V), dtype=np.float32)
    out[:N_JOINTS] = w.astype(np.float32)
    return out


def rvec_to_rmtx(rvec):
    """Rodrigues with the reference's regularised axis (utils/body_util.py:288-307): r = v/(|v|+1e-5)."""
    v = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(v))
    r = v / (th + 1e-5)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return math.cos(th) * np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * np.outer(r, r)


def body_pose_to_body_RTs(pose72, tpose_joints):
    """Per-joint local rotation / translation (reference utils/body_util.py:332-363 semantics)."""
    ja = np.asarray(pose72, dtype=np.float64).reshape(-1, 3)
    Rs = np.zeros((N_JOINTS, 3, 3), dtype=np.float32)
    Ts = np.zeros((N_JOINTS, 3), dtype=np.float32)
    for i in range(N_JOINTS):
        Rs[i] = rvec_to_rmtx(ja[i]).astype(np.float32)
        Ts[i] = tpose_joints[i] if i == 0 else tpose_joints[i] - tpose_joints[SMPL_PARENTS[i]]
    return Rs, Ts


def canonical_global_tfms(joints):
    """4x4 canonical joint transforms (pure translations; reference utils/body_util.py:400-424)."""
    joints = np.asarray(joints, dtype=np.float32)
    g = np.zeros((N_JOINTS, 4, 4), dtype=np.float32)
    for i in range(N_JOINTS):
        step = np.eye(4, dtype=np.float32)
        step[:3, 3] = joints[i] if i == 0 else joints[i] - joints[SMPL_PARENTS[i]]
        g[i] = step if i == 0 else g[SMPL_PARENTS[i]] @ step   # fp32 chain, like the reference
    return g


def make_humanoid(n_faces=13776, seed=0, jitter=0.05) -> Scene:
    joints = make_skeleton()
    parts = _parts(joints)
    S, R = _plan_resolution(parts, n_faces)
    vs, fs, off = [], [], 0
    for (name, a, b, r0, r1), s, r in zip(parts, S, R):
        v, f = _capsule(a, b, r0, r1, s, r)
        vs.append(v)
        fs.append(f + off)
        off += v.shape[0]
    verts = np.concatenate(vs, 0)
    faces = np.concatenate(fs, 0)
    assert faces.shape[0] == n_faces
    # jitter so that no face is exactly equilateral / degenerate (SURVEY.md §7 gradient singularity)
    rng = np.random.default_rng(seed)
    e = np.linalg.norm(verts[faces[:, 0]] - verts[faces[:, 1]], axis=1).mean()
    verts = verts + rng.normal(0.0, jitter * e, size=verts.shape)
    verts = verts.astype(np.float32)
    w = make_lbs_weights(verts, joints)
    return Scene(verts, faces, w, joints, canonical_global_tfms(joints))


def make_poses(n, seed=0):
    """[n,72] axis-angle body poses, root = 0; magnitudes follow the reference's PeopleSnapshot fits
    (data/snapshot/poses/*/anim_nerf_train.npz: ~0.05-0.15 rad per axis, shoulders ~0.8 rad about z)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)[:, None, None] / max(n, 1)
    base = np.zeros((1, N_JOINTS, 3))
    base[0, 16, 2] = -0.8   # lower the arms from the T-pose
    base[0, 17, 2] = 0.8
    amp = rng.uniform(0.03, 0.15, size=(1, N_JOINTS, 3))
    amp[0, 16:20] *= 1.5
    phase = rng.uniform(0, 2 * math.pi, size=(1, N_JOINTS, 3))
    freq = rng.integers(1, 4, size=(1, N_JOINTS, 3))
    pose = base + amp * np.sin(2 * math.pi * freq * t + phase) + rng.normal(0, 0.02, size=(n, N_JOINTS, 3))
    pose[:, 0] = 0.0
    return pose.reshape(n, 72).astype(np.float32)


def make_camera(azimuth, img_size=512, distance=3.5, height=0.1, focal=537.0, base_size=512, target=(0.0, -0.1, 0.0)):
    """ZJU-like pinhole camera: K [3,3], E [4,4] world->camera (x right, y down, z forward)."""
    W, H = (img_size, img_size) if np.isscalar(img_size) else img_size
    s = W / float(base_size)
    K = np.array([[focal * s, 0, 0.5 * W], [0, focal * s, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    C = np.array([distance * math.sin(azimuth), height, distance * math.cos(azimuth)])
    f = np.asarray(target, dtype=np.float64) - C
    f /= np.linalg.norm(f)
    x = np.cross(f, np.array([0.0, 1.0, 0.0]))
    x /= np.linalg.norm(x)
    y = np.cross(f, x)
    Rm = np.stack([x, y, f], 0)
    E = np.eye(4)
    E[:3, :3] = Rm
    E[:3, 3] = -Rm @ C
    return K, E.astype(np.float32)


def make_frames(scene: Scene, n_frames, img_size=512, seed=0, focal=537.0, distance=3.5, base_size=512):
    """Batched ``Model.forward`` inputs (reference models/model.py:184-188) for n_frames frames."""
    poses = make_poses(n_frames, seed=seed)
    rng = np.random.default_rng(seed + 1)
    out = {k: [] for k in ("K", "E", "cnl_gtfms", "dst_Rs", "dst_Ts", "dst_posevec", "bgcolor")}
    for i in range(n_frames):
        az = 2 * math.pi * ((i * 7) % 23) / 23.0     # 23 azimuths (ZJU has 23 cameras)
        K, E = make_camera(az, img_size=img_size, focal=focal, distance=distance, base_size=base_size)
        Rs, Ts = body_pose_to_body_RTs(poses[i], scene.joints)
        out["K"].append(K)
        out["E"].append(E)
        out["cnl_gtfms"].append(scene.cnl_gtfms)
        out["dst_Rs"].append(Rs)
        out["dst_Ts"].append(Ts)
        out["dst_posevec"].append(poses[i, 3:] + 1e-2)   # dataset/train.py:277
        out["bgcolor"].append(rng.uniform(0, 1, size=3).astype(np.float32))
    return {k: np.stack(v, 0).astype(np.float32) for k, v in out.items()}


def make_params(scene: Scene, seed=1, reference_init=False):
    """Learnable parameters in the reference's SoA layouts (models/model.py:74-85): vertices[3,V], so3[3,F],
    scale[3,F], appearance[3,F]."""
    F = scene.n_faces
    rng = np.random.default_rng(seed)
    if reference_init:
        so3 = np.zeros((3, F), np.float32)
        scale = np.ones((3, F), np.float32)
        app = np.full((3, F), 0.5, np.float32)
    else:
        so3 = rng.normal(0, 0.1, size=(3, F)).astype(np.float32)
        scale = rng.uniform(0.7, 1.3, size=(3, F)).astype(np.float32)
        app = rng.uniform(0, 1, size=(3, F)).astype(np.float32)
    return {"vertices": scene.vertices.T.copy(), "so3": so3, "scale": scale, "appearance": app}

"""Run the reference's ``train.py`` / ``eval.py`` byte-unchanged on the B200 path.

    import gomavatar_b200.compat as compat
    compat.install()                       # then, with the reference checkout on sys.path:  import train; train.main(args)
or  python -m gomavatar_b200.compat /path/to/GoMAvatar train.py --cfg exps/zju-mocap_377.yaml

``install()`` does three things, each only where needed:

1. third-party modules the reference imports at module level and that are NOT installed are provided by small stand-ins
   under ``compat/shims`` (appended to the END of ``sys.path``: a real installation always wins):
   ``pytorch3d`` (``structures.Meshes`` = ``gomavatar_b200.meshes.Meshes``, ``transforms.so3.so3_exp_map``,
   ``loss.mesh_normal_consistency`` / ``mesh_edge_loss``, import-only names for the rest), ``trimesh`` (what
   ``utils/pc_util.py::subdivide`` needs), ``skimage.metrics.structural_similarity`` (-> ``gom_eval_metrics``, the
   skimage-0.18 definition eval.py:106-108 relies on), ``torchmetrics`` (PSNR / SSIM of ``Evaluator_snapshot``),
   ``seaborn.color_palette``, ``termcolor.colored``, ``matplotlib`` (import-only);
2. ``diff_gaussian_rasterization`` resolves to this repository's drop-in (repo root on ``sys.path``; INTEGRATION.md level 0);
3. with ``b200_model=True`` (default) ``models.model`` — the module ``train.py:18`` / ``eval.py:18`` import ``Model`` from —
   is this package's ``gomavatar_b200.model`` (INTEGRATION.md level 2): same constructor, ``forward``, ``subdivide``,
   ``get_param_groups`` and state dict, so the rest of the reference's loop (datasets, ``compute_loss``, Adam, checkpoints,
   ``Evaluator``) runs as written.

Nothing here is on the measured hot path and nothing here falls back to a CPU implementation of it: the stand-ins are either
bookkeeping (mesh edge tables, colour palettes) or thin adapters onto ``libgom_b200.so``.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHIMMED = ("pytorch3d", "trimesh", "skimage", "torchmetrics", "seaborn", "termcolor", "matplotlib")


def _provided_by_shim(name):
    spec = importlib.util.find_spec(name)
    origin = getattr(spec, "origin", None) or ""
    return spec is not None and os.path.abspath(origin).startswith(SHIM_DIR)


def install(reference_root=None, b200_model=True):
    """Returns {'shims': [names served by the stand-ins], 'model': bool}.  Idempotent."""
    if SHIM_DIR not in sys.path:
        sys.path.append(SHIM_DIR)                      # last: installed packages take precedence
    if REPO_ROOT not in sys.path:
        sys.path.append(REPO_ROOT)                     # diff_gaussian_rasterization/ (and gomavatar_b200 itself)
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    importlib.invalidate_caches()
    served = [n for n in SHIMMED if _provided_by_shim(n)]
    if b200_model:
        from .. import model as b200_model_module
        sys.modules["models.model"] = b200_model_module
    return {"shims": served, "model": bool(b200_model)}


def uninstall():
    """Undo ``install`` (tests): drop the path entries, the model override and any shim module already imported."""
    for p in (SHIM_DIR,):
        while p in sys.path:
            sys.path.remove(p)
    mod = sys.modules.get("models.model")
    if mod is not None and getattr(mod, "__name__", "") == "gomavatar_b200.model":
        del sys.modules["models.model"]
    for name in list(sys.modules):
        m = sys.modules[name]
        f = getattr(m, "__file__", None) or ""
        if name.split(".")[0] in SHIMMED and os.path.abspath(f).startswith(SHIM_DIR):
            del sys.modules[name]
    importlib.invalidate_caches()


def run(reference_root, script, argv):
    """``python <script> <argv>`` inside ``reference_root`` with the stand-ins installed (the reference opens its configs
    by relative path: configs/__init__.py:14)."""
    import runpy
    reference_root = os.path.abspath(reference_root)
    install(reference_root)
    os.chdir(reference_root)
    sys.argv = [script] + list(argv)
    runpy.run_path(os.path.join(reference_root, script), run_name="__main__")

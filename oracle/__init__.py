"""CPU oracle for the GoMAvatar hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package; nothing under ``gomavatar_b200/`` does (a test enforces it).

What it restates, and how each piece is pinned:

* ``geometry.py``  — ``utils/body_util.py:591-644`` (joint chain, LBS), ``utils/network_util.py:66-92``
  (Rodrigues), ``models/model.py:27-41,212-250`` (face gather, Steiner frame, covariance) in plain torch.
  PINNED: ``oracle/make_golden.py`` runs the reference's own ``get_global_RTs``/``apply_lbs``, its
  ``get_transformation_from_triangle_steiner`` and its unmodified ``Model.forward`` + ``Renderer.forward``
  (third-party imports stubbed) in the build container and commits the vectors under ``tests/golden/``.
  ``so3_exp_map`` is PyTorch3D 0.7.0 (not in the reference tree, not installable offline): restated from its
  published semantics — parity unpinned for that one function.
* ``camera.py``    — ``models/modules/renderer/gaussian.py:30-66`` host math (pinned through the golden above).
* ``raster_oracle.c`` / ``raster.py`` — the splat rasterizer.  The algorithm lives in the third-party package
  ``diff_gaussian_rasterization`` (graphdeco-inria, UNPINNED ``pip install git+https://...`` at reference
  ``README.md:36``; API shape implies ``main`` ~ 59f5f77), whose source is absent from ``/root/reference`` and
  from this machine.  The oracle restates its published algorithm.  The reference has
  no tests, golden images or known-answer vectors for it:  **PARITY UNPINNED** at the rasterizer boundary.
  Mitigation: ``raster_torch.py`` is an independent float64 autograd restatement used to prove that the C
  oracle's hand-written backward is the derivative of its forward.
* ``losses.py``    — ``train.py:53-55,98-121`` (unpack, L1, LPIPS glue), ``utils/lpips/*`` (LPIPS-VGG v0.1; pinned
  against the in-tree reference LPIPS with a seeded random trunk because ImageNet weights cannot be
  downloaded), ``eval.py:101-108`` PSNR / skimage-0.18 SSIM semantics (skimage absent: parity unpinned).
"""

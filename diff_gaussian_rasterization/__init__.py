"""Import-name shim: lets the reference's ``from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer`` (models/modules/renderer/gaussian.py:9) resolve to the B200-native rasterizer when the repo
root is on PYTHONPATH.  See INTEGRATION.md."""
from gomavatar_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]

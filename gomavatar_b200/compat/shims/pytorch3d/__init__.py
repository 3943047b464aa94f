"""Stand-in for PyTorch3D 0.7.0 (README.md:21; absent offline, PARITY UNPINNED — semantics restated from its published
source): the pieces GoMAvatar's training / evaluation loop touches.  See gomavatar_b200/compat."""
__version__ = "0.7.0+gomavatar_b200.compat"

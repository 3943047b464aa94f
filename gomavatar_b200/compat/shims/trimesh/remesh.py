from gomavatar_b200.subdivision import faces_to_edges  # noqa: F401

from . import grouping  # noqa: F401

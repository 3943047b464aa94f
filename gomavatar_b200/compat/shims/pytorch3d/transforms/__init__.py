from .so3 import so3_exp_map, so3_log_map  # noqa: F401

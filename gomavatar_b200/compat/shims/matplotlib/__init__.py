"""Import-only stand-in for ``matplotlib`` (utils/tb_util.py:8-10 imports it at module level; only its
``summ_*`` heat-map helpers draw with it, none of which train.py's loop calls)."""


def use(*args, **kwargs):
    return None

"""Import-only stand-in (see matplotlib/__init__.py): any plotting call says what is missing."""


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise NotImplementedError(f"matplotlib.pyplot.{name}: matplotlib is not installed (gomavatar_b200.compat stand-in)")

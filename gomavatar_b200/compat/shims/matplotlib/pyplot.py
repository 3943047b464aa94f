def __getattr__(name):
    raise NotImplementedError(f"matplotlib.pyplot.{name}: matplotlib is not installed (gomavatar_b200.compat stand-in)")

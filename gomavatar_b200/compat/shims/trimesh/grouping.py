from gomavatar_b200.subdivision import unique_rows as _unique_rows


def unique_rows(data, digits=None, keep_order=False):
    if keep_order:
        raise NotImplementedError("trimesh stand-in: unique_rows(keep_order=True)")
    return _unique_rows(data)

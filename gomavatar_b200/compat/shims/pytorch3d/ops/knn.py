def knn_points(*args, **kwargs):
    raise NotImplementedError("pytorch3d stand-in: ops.knn_points is not provided (unused by GoMAvatar's mesh path)")

"""ctypes binding of libgom_b200.so (C ABI: include/gom_b200.h).

There is deliberately NO fallback: if the library has not been built (``python -m gomavatar_b200.build``) every op
raises.  The structs below mirror include/gom_b200.h field by field; their sizes are checked against the
``gom_sizeof_*`` exports at load time.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_size_t, c_void_p, c_float, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgom_b200.so")

ABI_VERSION = 11
STATUS_OVERFLOW = 1
STATUS_TIMEOUT = 2


class GomCameraArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("height", c_int32), ("width", c_int32), ("_pad", c_int32),
                ("K", c_void_p), ("E", c_void_p), ("viewmatrix", c_void_p), ("projmatrix", c_void_p),
                ("tanfov", c_void_p), ("campos", c_void_p)]


class GomRasterFwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_gauss", c_int32), ("height", c_int32), ("width", c_int32),
                ("n_channels", c_int32), ("interleaved", c_int32), ("inst_capacity", c_int64),
                ("means3D", c_void_p), ("means3D_stride", c_int64),
                ("cov3D", c_void_p), ("cov3D_stride", c_int64),
                ("colors", c_void_p), ("colors_stride", c_int64),
                ("opacities", c_void_p), ("opacities_stride", c_int64),
                ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("tanfov", c_void_p), ("bg", c_void_p),
                ("out_color", c_void_p), ("final_T", c_void_p), ("n_contrib", c_void_p), ("radii", c_void_p),
                ("depth", c_void_p), ("xy", c_void_p), ("conic_opacity", c_void_p), ("rect", c_void_p),
                ("tile_count", c_void_p), ("tile_offset", c_void_p), ("tile_cursor", c_void_p),
                ("inst_keys", c_void_p), ("point_list", c_void_p), ("status", c_void_p), ("worklist", c_void_p), ("point_mask", c_void_p)]


class GomRasterBwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_gauss", c_int32), ("height", c_int32), ("width", c_int32),
                ("n_channels", c_int32), ("interleaved", c_int32), ("color_grad_channels", c_int32), ("_pad", c_int32),
                ("inst_capacity", c_int64),
                ("means3D", c_void_p), ("means3D_stride", c_int64),
                ("cov3D", c_void_p), ("cov3D_stride", c_int64),
                ("colors", c_void_p), ("colors_stride", c_int64),
                ("viewmatrix", c_void_p), ("projmatrix", c_void_p), ("tanfov", c_void_p), ("bg", c_void_p),
                ("final_T", c_void_p), ("n_contrib", c_void_p), ("radii", c_void_p), ("xy", c_void_p),
                ("conic_opacity", c_void_p), ("tile_offset", c_void_p), ("point_list", c_void_p), ("worklist", c_void_p), ("point_mask", c_void_p),
                ("dL_dout", c_void_p),
                ("dL_dmeans3D", c_void_p), ("dL_dcov3D", c_void_p),
                ("dL_dcolors", c_void_p), ("dL_dcolors_stride", c_int64),
                ("dL_dopacity", c_void_p), ("dL_dmeans2D", c_void_p), ("dL_dconic", c_void_p)]


class GomJointFwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_joints", c_int32), ("parents", c_void_p), ("cnl_gtfms", c_void_p),
                ("dst_Rs", c_void_p), ("dst_Ts", c_void_p), ("global_Rs", c_void_p), ("global_Ts", c_void_p),
                ("chain_G", c_void_p), ("cnl_inv", c_void_p)]


class GomJointBwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_joints", c_int32), ("parents", c_void_p), ("dst_Rs", c_void_p),
                ("dst_Ts", c_void_p), ("chain_G", c_void_p), ("cnl_inv", c_void_p), ("dL_dglobal_Rs", c_void_p),
                ("dL_dglobal_Ts", c_void_p), ("dL_ddst_Rs", c_void_p), ("dL_ddst_Ts", c_void_p)]


class GomLbsFwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_joints", c_int32), ("n_verts", c_int32), ("_pad", c_int32),
                ("xyz", c_void_p), ("xyz_stride", c_int64), ("lbs_weights", c_void_p), ("global_Rs", c_void_p),
                ("global_Ts", c_void_p), ("out", c_void_p)]


class GomLbsBwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_joints", c_int32), ("n_verts", c_int32), ("_pad", c_int32),
                ("xyz", c_void_p), ("xyz_stride", c_int64), ("lbs_weights", c_void_p), ("global_Rs", c_void_p),
                ("global_Ts", c_void_p), ("dL_dout", c_void_p), ("dL_dxyz", c_void_p), ("dL_dxyz_stride", c_int64),
                ("dL_dglobal_Rs", c_void_p), ("dL_dglobal_Ts", c_void_p)]


class GomFaceFwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_faces", c_int32), ("n_verts", c_int32), ("faces_int64", c_int32),
                ("sigma", c_float), ("_pad", c_int32), ("verts", c_void_p), ("faces", c_void_p), ("so3", c_void_p),
                ("scale", c_void_p), ("means3D", c_void_p), ("cov3D", c_void_p)]


class GomFaceBwdArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_faces", c_int32), ("n_verts", c_int32), ("faces_int64", c_int32),
                ("sigma", c_float), ("_pad", c_int32), ("verts", c_void_p), ("faces", c_void_p), ("so3", c_void_p),
                ("scale", c_void_p), ("dL_dmeans3D", c_void_p), ("dL_dcov3D", c_void_p), ("dL_dverts", c_void_p),
                ("dL_dso3", c_void_p), ("dL_dscale", c_void_p)]


class GomPhotoArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("height", c_int32), ("width", c_int32), ("_pad", c_int32),
                ("rgb", c_void_p), ("rgb_pixel_stride", c_int64), ("mask", c_void_p), ("mask_pixel_stride", c_int64),
                ("bgcolor", c_void_p), ("gt_rgb", c_void_p), ("gt_mask", c_void_p), ("unpacked", c_void_p),
                ("loss_sums", c_void_p), ("dL_dunpacked", c_void_p), ("dL_dlosses", c_void_p),
                ("dL_drgb", c_void_p), ("dL_drgb_pixel_stride", c_int64),
                ("dL_dmask", c_void_p), ("dL_dmask_pixel_stride", c_int64)]


class GomShadeArgs(ctypes.Structure):
    _fields_ = [("n_pixels", c_int64), ("rgba", c_void_p), ("shading", c_void_p), ("rgbs", c_void_p), ("masks", c_void_p),
                ("dL_drgbs", c_void_p), ("dL_dmasks", c_void_p), ("dL_drgba", c_void_p), ("dL_dshading", c_void_p)]


class GomLpipsInputArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("height", c_int32), ("width", c_int32), ("from_unit_range", c_int32),
                ("pred", c_void_p), ("gt", c_void_p), ("out", c_void_p), ("dL_dout", c_void_p), ("dL_dpred", c_void_p)]


class GomBiasReluArgs(ctypes.Structure):
    _fields_ = [("n_pixels", c_int64), ("channels", c_int32), ("_pad", c_int32), ("x", c_void_p), ("bias", c_void_p)]


class GomReluBwdArgs(ctypes.Structure):
    _fields_ = [("n", c_int64), ("act", c_void_p), ("grad", c_void_p)]


class GomLpipsTapArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("height", c_int32), ("width", c_int32), ("channels", c_int32),
                ("pool", c_int32), ("_pad", c_int32), ("feats", c_void_p), ("lin", c_void_p), ("layer_sums", c_void_p),
                ("pooled", c_void_p), ("dL_dval", c_void_p), ("dL_dpooled", c_void_p), ("dL_dpre", c_void_p)]


class GomEvalMetricsArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("height", c_int32), ("width", c_int32), ("quantize", c_int32),
                ("pred", c_void_p), ("gt", c_void_p), ("ssim_sum", c_void_p), ("sq_err_sum", c_void_p),
                ("pred_8b", c_void_p)]


class GomConvFirstArgs(ctypes.Structure):
    _fields_ = [("n_images", c_int32), ("height", c_int32), ("width", c_int32), ("use_tensor_cores", c_int32), ("x", c_void_p),
                ("weight", c_void_p), ("bias", c_void_p), ("out", c_void_p), ("dL_dout", c_void_p), ("dL_dx", c_void_p),
                ("scratch", c_void_p), ("act", c_void_p), ("mask_out", c_void_p)]


class GomConvPackArgs(ctypes.Structure):
    _fields_ = [("c_out", c_int32), ("c_in", c_int32), ("transpose", c_int32), ("split", c_int32), ("kernel_size", c_int32),
                ("_pad", c_int32), ("weight", c_void_p), ("packed", c_void_p)]


class GomConv3x3Args(ctypes.Structure):
    _fields_ = [("n_images", c_int32), ("height", c_int32), ("width", c_int32), ("c_in", c_int32), ("c_out", c_int32),
                ("relu", c_int32), ("precision", c_int32), ("tma_round", c_int32), ("kernel_size", c_int32), ("_pad", c_int32),
                ("x", c_void_p), ("x_lo", c_void_p),
                ("w_packed", c_void_p), ("bias", c_void_p), ("mask_in", c_void_p), ("mask_out", c_void_p), ("out", c_void_p),
                ("status", c_void_p)]


class GomTf32SplitArgs(ctypes.Structure):
    _fields_ = [("n", c_int64), ("x", c_void_p), ("hi", c_void_p), ("lo", c_void_p), ("col_sum", c_void_p), ("n_cols", c_int32),
                ("_pad", c_int32)]


class GomLinearWgradArgs(ctypes.Structure):
    _fields_ = [("rows", c_int64), ("m", c_int32), ("n", c_int32), ("zero_first", c_int32), ("_pad", c_int32), ("g", c_void_p),
                ("g_lo", c_void_p), ("x", c_void_p), ("x_lo", c_void_p), ("out", c_void_p), ("status", c_void_p)]


ADAM_MAX_SEGMENTS = 16


class GomAdamArgs(ctypes.Structure):
    _fields_ = [("n", c_int64), ("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("grad_scale", c_float),
                ("lr_decay_rate", c_float), ("lr_decay_steps", c_float), ("n_segments", c_int32), ("_pad", c_int32),
                ("dev_steps", c_void_p), ("iter", c_int64),
                ("seg_end", c_int64 * ADAM_MAX_SEGMENTS), ("seg_step", c_int64 * ADAM_MAX_SEGMENTS),
                ("seg_lr", c_float * ADAM_MAX_SEGMENTS), ("seg_active", c_int32 * ADAM_MAX_SEGMENTS)]


class GomMeshRasterArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_verts", c_int32), ("n_faces", c_int32), ("height", c_int32), ("width", c_int32),
                ("faces_int64", c_int32), ("soft", c_int32), ("faces_per_pixel", c_int32), ("blur_radius", c_float),
                ("_pad", c_int32), ("list_capacity", c_int64), ("verts_ndc", c_void_p), ("faces", c_void_p),
                ("vert_normals", c_void_p), ("tile_count", c_void_p), ("tile_offset", c_void_p), ("tile_cursor", c_void_p),
                ("face_list", c_void_p), ("status", c_void_p), ("worklist", c_void_p), ("pix_to_face", c_void_p), ("normal_map", c_void_p),
                ("alpha", c_void_p), ("zcut", c_void_p), ("idcut", c_void_p), ("dL_dnormal_map", c_void_p),
                ("dL_dalpha", c_void_p), ("dL_dverts_ndc", c_void_p), ("dL_dvert_normals", c_void_p)]


class GomNonRigidInputArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_verts", c_int32), ("xyz_frames", c_int32), ("cond", c_int32), ("multires", c_int32),
                ("cols", c_int32), ("rows_padded", c_int64), ("alpha", c_float), ("_pad", c_int32), ("xyz", c_void_p),
                ("posevec", c_void_p), ("h0", c_void_p), ("enc", c_void_p), ("g_h0", c_void_p), ("g_enc", c_void_p), ("g_xyz", c_void_p)]


class GomRodriguesArgs(ctypes.Structure):
    _fields_ = [("n_rot", c_int32), ("group", c_int32), ("prepend_identity", c_int32), ("eps", c_float), ("rvec", c_void_p),
                ("R", c_void_p), ("g_R", c_void_p), ("g_rvec", c_void_p)]


class GomNarrowLinearArgs(ctypes.Structure):
    _fields_ = [("rows", c_int64), ("c_in", c_int32), ("n_out", c_int32), ("x", c_void_p), ("weight", c_void_p), ("bias", c_void_p),
                ("y", c_void_p), ("g_y", c_void_p), ("g_x", c_void_p), ("g_weight", c_void_p), ("g_bias", c_void_p)]


class GomVertexNormalsArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_verts", c_int32), ("n_faces", c_int32), ("faces_int64", c_int32), ("verts", c_void_p),
                ("faces", c_void_p), ("E", c_void_p), ("acc", c_void_p), ("normals_cam", c_void_p), ("dL_dnormals_cam", c_void_p),
                ("scratch", c_void_p), ("dL_dverts", c_void_p)]


class GomNdcArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_verts", c_int32), ("height", c_int32), ("width", c_int32), ("verts", c_void_p),
                ("K", c_void_p), ("E", c_void_p), ("ndc", c_void_p), ("dL_dndc", c_void_p), ("dL_dverts", c_void_p)]


class GomDilatedMaskL1Args(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("height", c_int32), ("width", c_int32), ("kernel_size", c_int32), ("dilate", c_int32),
                ("grad_scale", c_float), ("pred", c_void_p), ("mask_gt", c_void_p), ("sum", c_void_p), ("grad", c_void_p)]


class GomShadowMlpArgs(ctypes.Structure):
    _fields_ = [("n_pixels", c_int64), ("capacity", c_int64), ("multires", c_int32), ("width", c_int32), ("depth", c_int32),
                ("save_hidden", c_int32), ("normals", c_void_p), ("W_in", c_void_p), ("b_in", c_void_p), ("W_hid", c_void_p),
                ("b_hid", c_void_p), ("W_out", c_void_p), ("b_out", c_void_p), ("block_count", c_void_p),
                ("fg_index", c_void_p), ("n_fg", c_void_p), ("w_images", c_void_p), ("bg_value", c_void_p), ("out", c_void_p),
                ("act_img", c_void_p), ("status", c_void_p), ("g_out", c_void_p), ("dz_img", c_void_p), ("g_normals", c_void_p),
                ("dzo_sums", c_void_p), ("partials", c_void_p), ("g_W_in", c_void_p), ("g_b_in", c_void_p), ("g_W_hid", c_void_p),
                ("g_b_hid", c_void_p), ("g_w_out", c_void_p), ("g_b_out", c_void_p), ("bg_scratch", c_void_p)]


class GomMeshRegArgs(ctypes.Structure):
    _fields_ = [("n_frames", c_int32), ("n_verts", c_int32), ("n_pairs", c_int32), ("n_faces", c_int32), ("do_laplacian", c_int32),
                ("do_normal", c_int32), ("do_color", c_int32), ("n_color_pairs", c_int32), ("verts", c_void_p), ("row_ptr", c_void_p),
                ("col", c_void_p), ("pair_vid", c_void_p), ("pair_face", c_void_p), ("colors", c_void_p), ("lap", c_void_p),
                ("sums", c_void_p), ("g_verts_lap", c_void_p), ("g_verts_nc", c_void_p), ("g_colors", c_void_p)]


# every symbol include/gom_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "gom_abi_version", "gom_last_error", "gom_launch_count", "gom_profile_enable", "gom_profile_num_slots",
    "gom_profile_slot_name", "gom_profile_read", "gom_camera_from_KE", "gom_raster_forward", "gom_raster_backward",
    "gom_joint_transforms_forward", "gom_joint_transforms_backward", "gom_lbs_forward", "gom_lbs_backward",
    "gom_face_gaussians_forward", "gom_face_gaussians_backward", "gom_photometric_forward", "gom_photometric_backward",
    "gom_sizeof_photo_args", "gom_shade_forward", "gom_shade_backward", "gom_sizeof_shade_args", "gom_sizeof_camera_args", "gom_sizeof_raster_fwd_args", "gom_sizeof_raster_bwd_args",
    "gom_sizeof_joint_fwd_args", "gom_sizeof_joint_bwd_args", "gom_sizeof_lbs_fwd_args", "gom_sizeof_lbs_bwd_args",
    "gom_sizeof_face_fwd_args", "gom_sizeof_face_bwd_args",
    "gom_lpips_input_forward", "gom_lpips_input_backward", "gom_bias_relu", "gom_relu_backward",
    "gom_lpips_tap_forward", "gom_lpips_tap_backward", "gom_sizeof_lpips_input_args", "gom_sizeof_bias_relu_args",
    "gom_sizeof_relu_bwd_args", "gom_sizeof_lpips_tap_args", "gom_eval_metrics", "gom_sizeof_eval_metrics_args",
    "gom_conv_first_forward", "gom_conv_first_backward", "gom_sizeof_conv_first_args",
    "gom_conv3x3", "gom_conv3x3_pack_weights", "gom_tf32_split", "gom_sizeof_conv3x3_args", "gom_sizeof_conv_pack_args",
    "gom_sizeof_tf32_split_args", "gom_linear_wgrad", "gom_sizeof_linear_wgrad_args",
    "gom_nonrigid_input_forward", "gom_nonrigid_input_backward", "gom_sizeof_nonrigid_input_args",
    "gom_rodrigues_forward", "gom_rodrigues_backward", "gom_sizeof_rodrigues_args",
    "gom_narrow_linear_forward", "gom_narrow_linear_backward", "gom_sizeof_narrow_linear_args",
    "gom_adam_step", "gom_sizeof_adam_args",
    "gom_mesh_raster_forward", "gom_mesh_raster_backward", "gom_sizeof_mesh_raster_args",
    "gom_vertex_normals_forward", "gom_vertex_normals_backward", "gom_ndc_forward", "gom_ndc_backward", "gom_dilated_mask_l1",
    "gom_sizeof_vertex_normals_args", "gom_sizeof_ndc_args", "gom_sizeof_dilated_mask_l1_args",
    "gom_shadow_mlp_forward", "gom_shadow_mlp_backward", "gom_shadow_mlp_background_prepare", "gom_shadow_mlp_background_apply", "gom_shadow_mlp_weight_image_bytes", "gom_shadow_mlp_tile_words",
    "gom_shadow_mlp_partial_floats", "gom_shadow_mlp_num_ctas", "gom_sizeof_shadow_mlp_args",
    "gom_shadow_mlp_background_prepare", "gom_shadow_mlp_background_apply", "gom_shadow_mlp_bg_scratch_floats",
    "gom_mesh_regularizers", "gom_sizeof_mesh_reg_args",
]

_STRUCTS = {
    "camera": GomCameraArgs, "raster_fwd": GomRasterFwdArgs, "raster_bwd": GomRasterBwdArgs,
    "joint_fwd": GomJointFwdArgs, "joint_bwd": GomJointBwdArgs, "lbs_fwd": GomLbsFwdArgs, "lbs_bwd": GomLbsBwdArgs,
    "face_fwd": GomFaceFwdArgs, "face_bwd": GomFaceBwdArgs, "photo": GomPhotoArgs, "shade": GomShadeArgs,
    "lpips_input": GomLpipsInputArgs, "bias_relu": GomBiasReluArgs, "relu_bwd": GomReluBwdArgs,
    "lpips_tap": GomLpipsTapArgs, "eval_metrics": GomEvalMetricsArgs,
    "conv_first": GomConvFirstArgs, "adam": GomAdamArgs,
    "conv3x3": GomConv3x3Args, "conv_pack": GomConvPackArgs, "tf32_split": GomTf32SplitArgs,
    "linear_wgrad": GomLinearWgradArgs, "nonrigid_input": GomNonRigidInputArgs, "rodrigues": GomRodriguesArgs, "narrow_linear": GomNarrowLinearArgs,
    "mesh_raster": GomMeshRasterArgs, "vertex_normals": GomVertexNormalsArgs, "ndc": GomNdcArgs,
    "dilated_mask_l1": GomDilatedMaskL1Args, "shadow_mlp": GomShadowMlpArgs, "mesh_reg": GomMeshRegArgs,
}
_ENTRY_POINTS = ["gom_camera_from_KE", "gom_raster_forward", "gom_raster_backward", "gom_joint_transforms_forward",
                 "gom_joint_transforms_backward", "gom_lbs_forward", "gom_lbs_backward", "gom_face_gaussians_forward",
                 "gom_face_gaussians_backward", "gom_photometric_forward", "gom_photometric_backward", "gom_shade_forward", "gom_shade_backward",
                 "gom_lpips_input_forward", "gom_lpips_input_backward", "gom_bias_relu", "gom_relu_backward",
                 "gom_lpips_tap_forward", "gom_lpips_tap_backward", "gom_eval_metrics",
                 "gom_conv_first_forward", "gom_conv_first_backward", "gom_adam_step",
                 "gom_conv3x3", "gom_conv3x3_pack_weights", "gom_tf32_split", "gom_linear_wgrad", "gom_nonrigid_input_forward", "gom_nonrigid_input_backward", "gom_rodrigues_forward", "gom_rodrigues_backward", "gom_narrow_linear_forward", "gom_narrow_linear_backward",
                 "gom_mesh_raster_forward", "gom_mesh_raster_backward", "gom_vertex_normals_forward", "gom_vertex_normals_backward",
                 "gom_ndc_forward", "gom_ndc_backward", "gom_dilated_mask_l1", "gom_shadow_mlp_forward", "gom_shadow_mlp_backward", "gom_shadow_mlp_background_prepare", "gom_shadow_mlp_background_apply",
                 "gom_mesh_regularizers"]

_lib = None


class GomError(RuntimeError):
    pass


def lib():
    """Load the library once; raise loudly if it is missing or does not match this binding."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GomError(f"{LIB_PATH} not found: build it with `python -m gomavatar_b200.build` "
                       "(there is no CPU / PyTorch fallback for the hot path)")
    L = ctypes.CDLL(LIB_PATH)
    L.gom_last_error.restype = c_char_p
    L.gom_abi_version.restype = c_int
    if L.gom_abi_version() != ABI_VERSION:
        raise GomError(f"libgom_b200.so ABI {L.gom_abi_version()} != binding {ABI_VERSION}: rebuild")
    for key, cls in _STRUCTS.items():
        fn = getattr(L, f"gom_sizeof_{key}_args")
        fn.restype = c_size_t
        if ctypes.sizeof(cls) != fn():
            raise GomError(f"struct {cls.__name__}: ctypes mirror is {ctypes.sizeof(cls)} B, library says {fn()} B")
    for name in _ENTRY_POINTS:
        f = getattr(L, name)
        f.restype = c_int
        f.argtypes = [c_void_p, c_void_p]
    L.gom_shadow_mlp_weight_image_bytes.restype = c_size_t
    L.gom_shadow_mlp_bg_scratch_floats.restype = c_size_t
    L.gom_shadow_mlp_weight_image_bytes.argtypes = [c_int]
    L.gom_shadow_mlp_tile_words.restype = c_size_t
    L.gom_shadow_mlp_tile_words.argtypes = [c_int, c_int]
    L.gom_shadow_mlp_partial_floats.restype = c_size_t
    L.gom_shadow_mlp_num_ctas.restype = c_int
    L.gom_launch_count.restype = ctypes.c_longlong
    L.gom_profile_slot_name.restype = c_char_p
    L.gom_profile_read.argtypes = [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int)]
    _lib = L
    return L


def launch_count():
    return int(lib().gom_launch_count())


def profile_enable(on=True):
    lib().gom_profile_enable(1 if on else 0)


def profile_read():
    """{kernel name: (total_ms, launches)} for every kernel timed since profile_enable(True)."""
    L = lib()
    out = {}
    for s in range(L.gom_profile_num_slots()):
        ms, n = ctypes.c_double(0), c_int(0)
        check(L.gom_profile_read(s, ctypes.byref(ms), ctypes.byref(n)), "gom_profile_read")
        if n.value:
            out[L.gom_profile_slot_name(s).decode()] = (ms.value, n.value)
    return out


def check(rc, what):
    if rc != 0:
        raise GomError(f"{what} failed ({rc}): {lib().gom_last_error().decode()}")


_get_device = None      # torch._C._cuda_getDevice / _cuda_getCurrentRawStream: the C entry points behind torch.cuda.current_device() /
_get_raw_stream = None  # current_stream().cuda_stream, without their ~15 us of Python (lazy-init checks, Stream object) per launch


def _fast_torch():
    global _get_device, _get_raw_stream
    import torch
    try:
        torch.cuda.init()
    except Exception as e:                               # no driver / no device: there is no CPU path to fall back to
        raise GomError(f"no usable CUDA device ({type(e).__name__}: {e}); libgom_b200 has no CPU path") from e
    _get_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    _get_raw_stream = raw if raw is not None else (lambda dev: torch.cuda.current_stream(dev).cuda_stream)


def stream_ptr():
    """current torch CUDA stream (of the current device) as the gom_stream_t argument"""
    if _get_raw_stream is None:
        _fast_torch()
    return c_void_p(_get_raw_stream(_get_device()))


def call(name, args):
    """invoke an entry point on the current torch stream and raise on a non-zero return code"""
    rc = getattr(lib(), name)(ctypes.byref(args), stream_ptr())
    if rc != 0:
        check(rc, name)


def ptr(t):
    """device pointer of a torch tensor (None -> NULL).  Kernels are launched on the CURRENT device's current stream, so a
    tensor living on another GPU (model.to('cuda:1') without torch.cuda.set_device(1)) is refused instead of being
    dereferenced from the wrong device."""
    if t is None:
        return None
    if t.is_cuda:
        if _get_device is None:
            _fast_torch()
        if t.device.index != _get_device():
            raise GomError(f"tensor on {t.device} but the current CUDA device is cuda:{_get_device()}: "
                           "call torch.cuda.set_device(...) (or use `with torch.cuda.device(...)`) before launching")
    return c_void_p(t.data_ptr())

/*
 * gom_b200.h — C ABI of the B200-native GoMAvatar hot path (libgom_b200.so, sm_100a).
 *
 * This is the drop-in boundary underneath the reference's two Python boundaries:
 *   #1  Model.forward(K, E, cnl_gtfms, dst_Rs, dst_Ts, ...)            reference models/model.py:184-188
 *   #2  GaussianRasterizationSettings / GaussianRasterizer.forward(...) reference models/modules/renderer/gaussian.py:9,20,53-67,83-91
 * The reference's own native layer is the pybind11/ATen module `diff_gaussian_rasterization._C`
 * (rasterize_gaussians / rasterize_gaussians_backward; third-party, not in the reference tree) plus ~60 small
 * torch ops; this header is what replaces them.
 *
 * Conventions (every entry point):
 *   - plain C structs of raw DEVICE pointers + sizes; no torch / C++ types cross the boundary;
 *   - the CALLER owns every buffer, including scratch and state kept for backward (so a caching allocator and
 *     CUDA graphs see them); the library never calls cudaMalloc, never synchronises the device, never touches the
 *     default stream: all work is enqueued on the `stream` argument (a cudaStream_t passed as void*);
 *   - returns 0 on success or a negative GOM_ERR_*; never throws; gom_last_error() gives the message (per thread);
 *   - all arrays are fp32 / int32 / uint32 / uint64, C-contiguous in the stated shape; "frame stride" arguments
 *     are in ELEMENTS; a stride of 0 shares one array between all frames of the batch;
 *   - B = n_frames (the reference is B = 1), P = n_gauss (= faces F on the fused path), T = tiles per frame.
 */
#ifndef GOM_B200_H
#define GOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOM_ABI_VERSION 11
#define GOM_TILE 16              /* 16x16-pixel tiles, as upstream's BLOCK_X/BLOCK_Y */
#define GOM_MAX_CHANNELS 4
#define GOM_MAX_JOINTS 64

enum {
    GOM_OK = 0,
    GOM_ERR_INVALID = -1,        /* bad argument (null pointer, unsupported size, ...) */
    GOM_ERR_CUDA = -2,           /* a CUDA runtime call failed (message has the cudaError string) */
    GOM_ERR_UNSUPPORTED = -3
};

/* status bits written to device memory by the rasterizer (checked lazily by the host) */
#define GOM_STATUS_OVERFLOW 1u   /* a frame produced more (Gaussian,tile) instances than inst_capacity */
#define GOM_STATUS_TIMEOUT 2u    /* a tcgen05 pipeline wait ran into its bound (protocol error); results are invalid */

typedef void *gom_stream_t;      /* cudaStream_t */

int gom_abi_version(void);
const char *gom_last_error(void);

/* Instrumentation (used by bench.py): number of kernels this library has launched in this process, and optional
 * CUDA-event timers recorded on the launching stream around each kernel (slot = kernel, see gom_profile_slot_name). */
long long gom_launch_count(void);
void gom_profile_enable(int on);                 /* on = 1 also clears previously recorded samples */
int gom_profile_num_slots(void);
const char *gom_profile_slot_name(int slot);
int gom_profile_read(int slot, double *total_ms, int *count);   /* waits for the recorded events */

/* --------------------------------------------------------------------------------------------------------------
 * Camera setup.  Replaces the host math + 4 .item() syncs + H2D of reference gaussian.py:30-47,60-61:
 * K [B,3,3], E [B,4,4]  ->  viewmatrix = E^T, projmatrix = E^T K_ndc^T  (row-major [B,16] each, i.e. exactly the
 * tensors the reference hands to GaussianRasterizationSettings), tanfov [B,2] = (W/2fx, H/2fy), campos [B,3].
 * znear 0.001 / zfar 100 as in the reference.
 */
typedef struct {
    int32_t n_frames, height, width, _pad;
    const float *K;              /* [B,3,3] */
    const float *E;              /* [B,4,4] */
    float *viewmatrix;           /* [B,16] */
    float *projmatrix;           /* [B,16] */
    float *tanfov;               /* [B,2]  */
    float *campos;               /* [B,3] (nullable) */
} GomCameraArgs;
int gom_camera_from_KE(const GomCameraArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Splat rasterizer forward.  Replaces `_C.rasterize_gaussians` (preprocessCUDA -> InclusiveSum -> D2H sync ->
 * duplicateWithKeys -> SortPairs -> identifyTileRanges -> renderCUDA), for the
 * branch GoMAvatar uses: colors_precomp + cov3D_precomp, sh_degree 0, scale_modifier 1.
 * No host sync: instance buffers have a fixed per-frame capacity; overflow sets GOM_STATUS_OVERFLOW in status[b]
 * (the image of that frame is then invalid and the caller re-runs with a larger capacity).
 */
typedef struct {
    int32_t n_frames, n_gauss, height, width;
    int32_t n_channels;          /* 3 (reference pass) or 4 (fused RGB + alpha pass) */
    int32_t interleaved;         /* 0: out_color [B,C,H,W] (reference layout); 1: [B,H,W,C] */
    int64_t inst_capacity;       /* per-frame capacity of inst_keys / point_list */
    /* inputs */
    const float *means3D;   int64_t means3D_stride;     /* [B,P,3] */
    const float *cov3D;     int64_t cov3D_stride;       /* [B,P,6]  xx,xy,xz,yy,yz,zz */
    const float *colors;    int64_t colors_stride;      /* [B,P,C]  (stride 0: one [P,C] for all frames) */
    const float *opacities; int64_t opacities_stride;   /* [B,P] */
    const float *viewmatrix;     /* [B,16] */
    const float *projmatrix;     /* [B,16] */
    const float *tanfov;         /* [B,2]  */
    const float *bg;             /* [B,C]  */
    /* outputs */
    float *out_color;            /* [B,C,H,W] or [B,H,W,C] */
    float *final_T;              /* [B,H,W]  transmittance (mask = 1 - final_T over a zero background) */
    uint32_t *n_contrib;         /* [B,H,W]  index of the last contributing list entry (1-based) */
    int32_t *radii;              /* [B,P]    0 = culled */
    /* per-Gaussian state kept for backward */
    float *depth;                /* [B,P]   view-space z */
    float *xy;                   /* [B,P,2] pixel centre */
    float *conic_opacity;        /* [B,P,4] conic A,B,C + opacity */
    int32_t *rect;               /* [B,P,4] tile rect minx,miny,maxx,maxy (max exclusive) */
    /* binning state */
    uint32_t *tile_count;        /* [B,T]   */
    uint32_t *tile_offset;       /* [B,T+1] exclusive scan; [b][T] = N_dup of frame b */
    uint32_t *tile_cursor;       /* [B,T]   scratch */
    uint64_t *inst_keys;         /* [B,cap] (depth bits << 32 | gaussian id), grouped by tile, unsorted */
    uint32_t *point_list;        /* [B,cap] gaussian ids grouped by tile, sorted by (depth, id) */
    uint32_t *status;            /* [B]     GOM_STATUS_* bits */
    uint32_t *worklist;          /* [B*T]   out, nullable: (frame * T + tile) of every tile, longest list first — the order the
                                    sort / blend kernels (and the backward, if handed over) walk the tiles in */
    uint8_t *point_mask;         /* [B,cap] out, nullable: per point_list entry, bit s = the entry can reach the 8x4-pixel sub-block s
                                    of its tile (conservative alpha >= 1/255 test, done once in the sort instead of per blend warp) */
} GomRasterFwdArgs;
int gom_raster_forward(const GomRasterFwdArgs *a, gom_stream_t stream);

/* Splat rasterizer backward.  Replaces `_C.rasterize_gaussians_backward` (renderCUDA bwd, computeCov2DCUDA,
 * preprocessCUDA bwd; App. A.6-A.7).  Gradient buffers are zeroed by the call itself. */
typedef struct {
    int32_t n_frames, n_gauss, height, width;
    int32_t n_channels, interleaved;
    int32_t color_grad_channels; /* 0 = all; 3 with n_channels 4: the 4th (constant alpha) channel gets no gradient (left 0) */
    int32_t _pad;
    int64_t inst_capacity;
    const float *means3D;   int64_t means3D_stride;
    const float *cov3D;     int64_t cov3D_stride;
    const float *colors;    int64_t colors_stride;
    const float *viewmatrix;
    const float *projmatrix;
    const float *tanfov;
    const float *bg;
    /* saved forward state */
    const float *final_T;
    const uint32_t *n_contrib;
    const int32_t *radii;
    const float *xy;
    const float *conic_opacity;
    const uint32_t *tile_offset;
    const uint32_t *point_list;
    const uint32_t *worklist;    /* [B*T] nullable: the forward's tile order (longest list first) */
    const uint8_t *point_mask;   /* [B,cap] nullable: the forward's sub-block masks */
    /* upstream gradient */
    const float *dL_dout;        /* same layout as out_color */
    /* outputs */
    float *dL_dmeans3D;          /* [B,P,3] */
    float *dL_dcov3D;            /* [B,P,6] */
    float *dL_dcolors;  int64_t dL_dcolors_stride;      /* [B,P,C]; stride 0: [P,C] summed over frames */
    float *dL_dopacity;          /* [B,P] (nullable: not computed) */
    float *dL_dmeans2D;          /* [B,P,2] screen-space gradient (App. A.6), also scratch */
    float *dL_dconic;            /* [B,P,3] scratch */
} GomRasterBwdArgs;
int gom_raster_backward(const GomRasterBwdArgs *a, gom_stream_t stream);


/* --------------------------------------------------------------------------------------------------------------
 * Skeleton: per-joint skinning transforms.  Replaces reference utils/body_util.py:612-638 `get_global_RTs`
 * (+ _construct_G_tensor :591-609): local G_i = [R_i|T_i], chained root-to-leaf along `parents` (parents[0] = -1,
 * parents[i] < i; SMPL_PARENT at body_util.py:36-39), F_i = G_i inv(cnl_gtfms_i); returns F[:3,:3] and F[:3,3].
 * ~50 torch launches (23 sequential matmul+clone, batched LU) become one.
 */
typedef struct {
    int32_t n_frames, n_joints;
    const int32_t *parents;      /* [J] (device) */
    const float *cnl_gtfms;      /* [B,J,4,4] */
    const float *dst_Rs;         /* [B,J,3,3] */
    const float *dst_Ts;         /* [B,J,3]   */
    float *global_Rs;            /* [B,J,3,3] out */
    float *global_Ts;            /* [B,J,3]   out */
    float *chain_G;              /* [B,J,12]  top three rows of the chained transforms (kept for backward) */
    float *cnl_inv;              /* [B,J,16]  inverted canonical transforms (kept for backward) */
} GomJointFwdArgs;
int gom_joint_transforms_forward(const GomJointFwdArgs *a, gom_stream_t stream);

typedef struct {
    int32_t n_frames, n_joints;
    const int32_t *parents;
    const float *dst_Rs, *dst_Ts;    /* forward inputs */
    const float *chain_G, *cnl_inv;  /* forward state */
    const float *dL_dglobal_Rs;      /* [B,J,3,3] */
    const float *dL_dglobal_Ts;      /* [B,J,3]   */
    float *dL_ddst_Rs;               /* [B,J,3,3] out */
    float *dL_ddst_Ts;               /* [B,J,3]   out */
} GomJointBwdArgs;
int gom_joint_transforms_backward(const GomJointBwdArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Linear-blend skinning.  Replaces reference utils/body_util.py:641-644 `apply_lbs`:
 *   out[b,:,v] = sum_{j<J} w[j,v] (R[b,j] xyz[:,v] + T[b,j]),  weights are the model's [J+1,V] buffer whose last
 * (background) row is ignored; no renormalisation.  SoA layouts are the reference's ([3,V], [J+1,V]) so checkpoints
 * load unchanged.  Weight tiles are staged into shared memory with TMA bulk copies (cp.async.bulk) when the buffer
 * is 16-byte aligned, else with plain coalesced loads.
 */
typedef struct {
    int32_t n_frames, n_joints, n_verts, _pad;
    const float *xyz;        int64_t xyz_stride;        /* [B,3,V] or one [3,V] (stride 0) */
    const float *lbs_weights;                            /* [J+1,V] (rows 0..J-1 are read) */
    const float *global_Rs;                              /* [B,J,3,3] */
    const float *global_Ts;                              /* [B,J,3]   */
    float *out;                                          /* [B,3,V]   */
} GomLbsFwdArgs;
int gom_lbs_forward(const GomLbsFwdArgs *a, gom_stream_t stream);

typedef struct {
    int32_t n_frames, n_joints, n_verts, _pad;
    const float *xyz;        int64_t xyz_stride;
    const float *lbs_weights;
    const float *global_Rs, *global_Ts;
    const float *dL_dout;                                /* [B,3,V] */
    float *dL_dxyz;          int64_t dL_dxyz_stride;     /* [B,3,V], or [3,V] summed over frames (stride 0) */
    float *dL_dglobal_Rs;                                /* [B,J,3,3] (nullable together with dL_dglobal_Ts) */
    float *dL_dglobal_Ts;                                /* [B,J,3]   */
} GomLbsBwdArgs;
int gom_lbs_backward(const GomLbsBwdArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Gaussians on mesh: one Gaussian per face, carried by the face's Steiner-ellipse frame.  Replaces reference
 * models/model.py:225-234 (+ get_transformation_from_triangle_steiner :27-41, PyTorch3D so3_exp_map) and the
 * covariance packing of models/modules/renderer/gaussian.py:71-75:
 *   mean_f = centroid(tri_f);  Sigma_f = A_f R(so3_f) S_f S_f^T R^T A_f^T packed (xx,xy,xz,yy,yz,zz).
 * ~35 torch launches and a dozen [F,3,3] temporaries become one kernel.
 */
typedef struct {
    int32_t n_frames, n_faces, n_verts, faces_int64;
    float sigma; int32_t _pad;
    const float *verts;          /* [B,3,V] posed ("observation") vertices, SoA */
    const void *faces;           /* [F,3] int32, or int64 when faces_int64 (the model's registered buffer) */
    const float *so3;            /* [3,F] */
    const float *scale;          /* [3,F] */
    float *means3D;              /* [B,F,3] out */
    float *cov3D;                /* [B,F,6] out */
} GomFaceFwdArgs;
int gom_face_gaussians_forward(const GomFaceFwdArgs *a, gom_stream_t stream);

typedef struct {
    int32_t n_frames, n_faces, n_verts, faces_int64;
    float sigma; int32_t _pad;
    const float *verts;
    const void *faces;
    const float *so3, *scale;
    const float *dL_dmeans3D;    /* [B,F,3] */
    const float *dL_dcov3D;      /* [B,F,6] */
    float *dL_dverts;            /* [B,3,V] out (zeroed by the call, scatter-added) */
    float *dL_dso3;              /* [3,F]   out, summed over frames */
    float *dL_dscale;            /* [3,F]   out, summed over frames */
} GomFaceBwdArgs;
int gom_face_gaussians_backward(const GomFaceBwdArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Photometric losses.  Replaces reference train.py:53-55 (`unpack`: rgb*mask + bg*(1-mask), per-frame background)
 * and train.py:101-111 (mean |rgb - gt|, mean |mask - gt_mask|), forward and backward, one pass each way.
 * rgb / mask may be views into one interleaved [B,H,W,4] render (pixel strides 4 / 4) or separate tensors (3 / 1).
 * forward : unpacked [B,H,W,3]; loss_sums[0] = sum |unpacked - gt_rgb|, loss_sums[1] = sum |mask - gt_mask| (nullable)
 * backward: dL_drgb, dL_dmask from dL_dunpacked (nullable, e.g. LPIPS' gradient) and dL_dlosses[2] (device; gradient
 *           of the two MEAN losses, i.e. the loss coefficients times the upstream scalar).
 * bgcolor == NULL skips the compositing (unpacked = rgb), as the reference does when random_bgcolor is off.
 */
typedef struct {
    int32_t n_frames, height, width, _pad;
    const float *rgb;   int64_t rgb_pixel_stride;
    const float *mask;  int64_t mask_pixel_stride;
    const float *bgcolor;        /* [B,3] nullable */
    const float *gt_rgb;         /* [B,H,W,3] nullable */
    const float *gt_mask;        /* [B,H,W]   nullable */
    float *unpacked;             /* [B,H,W,3] (forward) */
    float *loss_sums;            /* [2]       (forward, nullable) */
    const float *dL_dunpacked;   /* [B,H,W,3] (backward, nullable) */
    const float *dL_dlosses;     /* [2]       (backward, nullable) */
    float *dL_drgb;     int64_t dL_drgb_pixel_stride;
    float *dL_dmask;    int64_t dL_dmask_pixel_stride;
} GomPhotoArgs;
int gom_photometric_forward(const GomPhotoArgs *a, gom_stream_t stream);
int gom_photometric_backward(const GomPhotoArgs *a, gom_stream_t stream);

/* Pseudo-shading of the rasterizer's interleaved output, one launch each way (reference models/model.py:281-287:
 * `rgbs = albedos * shadings` with albedos / masks the channel slices of the rendered RGBA image):
 *   forward   rgbs[p,k] = rgba[p,k] * shading[p] (k < 3),  masks[p] = rgba[p,3]
 *   backward  dL_drgba[p] = (dL_drgbs[p,:] * shading[p], dL_dmasks[p]),  dL_dshading[p] = sum_k dL_drgbs[p,k] * rgba[p,k]
 * (replaces a strided multiply, its two product gradients, a channel reduction and autograd's slice backward: two zero
 * fills, two strided copies and an add over the full image). */
typedef struct {
    int64_t n_pixels;            /* B*H*W */
    const float *rgba;           /* [n,4], 16-byte aligned */
    const float *shading;        /* [n] */
    float *rgbs;                 /* [n,3] out (forward) */
    float *masks;                /* [n]   out (forward) */
    const float *dL_drgbs;       /* [n,3] (backward; NULL = zeros) */
    const float *dL_dmasks;      /* [n]   (backward; NULL = zeros) */
    float *dL_drgba;             /* [n,4] out (backward), 16-byte aligned */
    float *dL_dshading;          /* [n]   out (backward) */
} GomShadeArgs;
int gom_shade_forward(const GomShadeArgs *a, gom_stream_t stream);
int gom_shade_backward(const GomShadeArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * LPIPS-VGG v0.1 perceptual loss — everything that is not a convolution (csrc/lpips.cu).  Replaces the torch ops of
 * reference utils/lpips/lpips.py:81-123 (ScalingLayer :126-133, NetLinLayer :136-146), utils/lpips/__init__.py:40-42
 * (normalize_tensor), and the ReLU / MaxPool2d layers of utils/lpips/pretrained_networks.py:96-134, forward AND
 * backward, as called from train.py:113-121.  Activations are NHWC fp32; a batch is [B predictions | B targets].
 */
typedef struct {
    int32_t n_frames, height, width;
    int32_t from_unit_range;     /* 1: inputs are in [0,1] and the 2x-1 of train.py:114-116 is folded in */
    const float *pred;           /* [B,H,W,3] */
    const float *gt;             /* [B,H,W,3] */
    float *out;                  /* [2B,H,W,3] (forward)  (v - shift_c) / scale_c, ScalingLayer */
    const float *dL_dout;        /* [B,H,W,3]  (backward) gradient wrt the prediction half of `out` */
    float *dL_dpred;             /* [B,H,W,3]  (backward) */
} GomLpipsInputArgs;
int gom_lpips_input_forward(const GomLpipsInputArgs *a, gom_stream_t stream);
int gom_lpips_input_backward(const GomLpipsInputArgs *a, gom_stream_t stream);

/* x = max(x + bias[c], 0) in place over [n_pixels, channels] (the epilogue of each VGG convolution) */
typedef struct {
    int64_t n_pixels; int32_t channels, _pad;
    float *x;
    const float *bias;           /* [channels] */
} GomBiasReluArgs;
int gom_bias_relu(const GomBiasReluArgs *a, gom_stream_t stream);

/* grad *= (act > 0) in place (ReLU backward of the untapped layers) */
typedef struct {
    int64_t n;                   /* elements, multiple of 4 */
    const float *act;            /* post-ReLU activation */
    float *grad;
} GomReluBwdArgs;
int gom_relu_backward(const GomReluBwdArgs *a, gom_stream_t stream);

/* One tapped layer (relu1_2, 2_2, 3_3, 4_3, 5_3).
 * forward : layer_sums[b] += mean_{pixels} sum_c lin_c (f0_c/(|f0|+eps) - f1_c/(|f1|+eps))^2, |f| = sqrt(sum f^2 + eps),
 *           f0 = feats[b], f1 = feats[B+b]; with pool = 1 also writes the 2x2/2 max-pool of all 2B images.
 * backward: dL_dpre = [ dL_dval[b] * d(layer value)/df0 + maxpool-backward(dL_dpooled) ] * (f0 > 0): the gradient
 *           wrt the PRE-ReLU convolution output of the prediction half, ready for the convolution's dgrad.
 * channels in {32, 64, 128, 256, 512}. */
typedef struct {
    int32_t n_frames, height, width, channels;
    int32_t pool, _pad;
    const float *feats;          /* [2B,h,w,C] post-ReLU */
    const float *lin;            /* [C] 1x1 head weights */
    float *layer_sums;           /* [B]  (forward, accumulated into) */
    float *pooled;               /* [2B,h/2,w/2,C] (forward, pool = 1) */
    const float *dL_dval;        /* [B]  (backward) */
    const float *dL_dpooled;     /* [B,h/2,w/2,C] (backward, nullable) */
    float *dL_dpre;              /* [B,h,w,C] (backward) */
} GomLpipsTapArgs;
int gom_lpips_tap_forward(const GomLpipsTapArgs *a, gom_stream_t stream);
int gom_lpips_tap_backward(const GomLpipsTapArgs *a, gom_stream_t stream);

/* First VGG16 convolution of LPIPS (3 -> 64 channels, 3x3, stride 1, zero padding 1) fused with bias + ReLU, and its
 * input gradient, with fp32 accuracy.  Replaces `features[0:2]` of reference utils/lpips/pretrained_networks.py:96-134.
 * NHWC activations; weight is torch's contiguous [64,3,3,3].
 * forward : out = relu(conv(x) + bias)          backward: dL_dx = conv_transpose(dL_dout)  (dL_dout already ReLU-masked)
 * use_tensor_cores = 0: FP32-FMA kernels (csrc/conv_first.cu, FMA-pipe bound); 1: tcgen05 GEMMs over pixel rows with
 * 3xTF32 products (csrc/conv_first_tc.cu, HBM bound); the backward then needs `scratch`. */
typedef struct {
    int32_t n_images, height, width, use_tensor_cores;
    const float *x;              /* [N,H,W,3]  (forward) */
    const float *weight;         /* [64,3,3,3] */
    const float *bias;           /* [64]       (forward) */
    float *out;                  /* [N,H,W,64] (forward) */
    const float *dL_dout;        /* [N,H,W,64] (backward) */
    float *dL_dx;                /* [N,H,W,3]  (backward) */
    float *scratch;              /* [9, N*H*W, 4] floats (backward with use_tensor_cores), 16-byte aligned; else NULL */
    const float *act;            /* optional [N,H,W,64] (backward with use_tensor_cores): this convolution's ReLU output; when given,
                                    dL_dout is the UNMASKED gradient and the ReLU backward (dL_dout * [act > 0]) is fused in */
    uint32_t *mask_out;          /* optional [N,H,W,2] (forward with use_tensor_cores): ReLU bit mask of `out` (bit j of word w =
                                    [channel 32 w + j > 0]) in the layout gom_conv3x3's mask_in reads: the dgrad of conv1_2 then
                                    applies this layer's ReLU backward itself and the backward here needs neither `act` nor a mask */
} GomConvFirstArgs;
int gom_conv_first_forward(const GomConvFirstArgs *a, gom_stream_t stream);
int gom_conv_first_backward(const GomConvFirstArgs *a, gom_stream_t stream);

/* The other twelve VGG16 convolutions of LPIPS (conv1_2 ... conv5_3: 3x3, stride 1, zero padding 1, channel counts that
 * are multiples of 32) as tcgen05 implicit GEMMs: forward with bias + ReLU fused, input gradient with the ReLU backward
 * of the layer below fused.  Replaces `features[2:30]` of reference utils/lpips/pretrained_networks.py:96-134 and their
 * autograd (cuDNN in the reference).  TF32 products, fp32 accumulation — the reference's stock cuDNN setting
 * (torch.backends.cudnn.allow_tf32 is never touched by it); precision = 1 selects 3xTF32 (fp32-GEMM accuracy).
 *
 * Both directions are the same GEMM  out[p, n] = sum_{tap, c} x[p + tap - (1,1), c] * w_packed[tap][n][c]  over NHWC
 * activations; what differs is the packed weight (gom_conv3x3_pack_weights) and the epilogue:
 *   forward : x = layer input  [N,H,W,C],  w_packed = fwd pack [9][K][C],  out = relu(. + bias)          [N,H,W,K];
 *             mask_out (optional) receives one word per pixel and 32 output channels, bit j = [out channel 32 w + j > 0];
 *   dgrad   : x = dL/dout      [N,H,W,K],  w_packed = bwd pack [9][C][K] (taps flipped),  out = dL/dx    [N,H,W,C],
 *             multiplied by the ReLU mask of the activation that was this layer's input when mask_in (the mask_out of
 *             the forward call that produced that activation) is given.
 * The activations travel global -> shared memory as TMA tensor tiles (the 18 x 18-pixel halo of a 16 x 16-pixel output
 * tile per 32-channel block, zero-filled outside the image; the nine taps are shifted views of it: no im2col buffer),
 * the accumulators live in tensor memory, the result leaves through TMA tensor stores. */
typedef struct {
    int32_t c_out, c_in;         /* K, C of the torch weight [K,C,3,3] */
    int32_t transpose;           /* 0: forward pack [9][K][C];  1: dgrad pack [9][C][K] with the taps flipped */
    int32_t split;               /* 0: one TF32-rounded image; 1: two images [hi | lo] for precision = 1 */
    int32_t kernel_size;         /* 3: weight [K,C,3,3], nine taps;  1: weight [K,C] (a Linear layer), one tap */
    int32_t _pad;
    const float *weight;         /* [K,C,ks,ks] contiguous */
    float *packed;               /* ks*ks*K*C floats (x2 when split) */
} GomConvPackArgs;
int gom_conv3x3_pack_weights(const GomConvPackArgs *a, gom_stream_t stream);

typedef struct {
    int32_t n_images, height, width;
    int32_t c_in, c_out;         /* channels of x and of out (dgrad: c_in = K, c_out = C) */
    int32_t relu;                /* forward: apply ReLU after the bias */
    int32_t precision;           /* 0: TF32;  1: 3xTF32 (needs x_lo and a split weight pack) */
    int32_t tma_round;           /* 1: the TMA engine rounds x to TF32 (round-to-nearest) on its way to shared memory;
                                    0: the tensor core truncates the fp32 words it reads */
    int32_t kernel_size;         /* 3 (the convolution described above) or 1: out[p, n] = sum_c x[p, c] w_packed[0][n][c], a plain
                                    GEMM over pixel rows with the same epilogues — the Linear layers of the reference's MLPs
                                    (models/modules/non_rigid_module.py:75-147) with rows laid out as an [N,H,W] "image";
                                    c_out must then be a multiple of 64 */
    int32_t _pad;
    const float *x;              /* [N,H,W,c_in] */
    const float *x_lo;           /* precision = 1: x - trunc_tf32(x), same shape (gom_tf32_split); else NULL */
    const float *w_packed;       /* from gom_conv3x3_pack_weights */
    const float *bias;           /* [c_out] or NULL */
    const uint32_t *mask_in;     /* [N,H,W,c_out/32] or NULL: out is zeroed where the bit is 0 */
    uint32_t *mask_out;          /* [N,H,W,c_out/32] or NULL: bit = [out > 0] */
    float *out;                  /* [N,H,W,c_out] */
    uint32_t *status;            /* [1] nullable: GOM_STATUS_TIMEOUT */
} GomConv3x3Args;
int gom_conv3x3(const GomConv3x3Args *a, gom_stream_t stream);

/* hi = x truncated to TF32 (what the tensor core reads from an fp32 word; nullable), lo = x - hi; n a multiple of 4.
 * col_sum (nullable): x is a row-major [n / n_cols, n_cols] matrix (n_cols a multiple of 4 dividing 1024) and col_sum[n_cols]
 * is INCREMENTED by its column sums (the bias gradient of a Linear layer, formed while the gradient is read anyway). */
typedef struct {
    int64_t n;
    const float *x;
    float *hi, *lo;
    float *col_sum;
    int32_t n_cols, _pad;
} GomTf32SplitArgs;
int gom_tf32_split(const GomTf32SplitArgs *a, gom_stream_t stream);

/* Weight gradient of a Linear layer over a tall batch of rows (reference: autograd of every nn.Linear of
 * models/modules/non_rigid_module.py:75-147):  out[m, n] (+)= sum_r g[r, m] x[r, n]  with 3xTF32 products on the tensor cores
 * (g x + g_lo x + g x_lo; g_lo, x_lo from gom_tf32_split).  g [rows, m], x [rows, n] row-major fp32; m = 128; n a multiple of
 * 32, at most 256.  Split-K over the SMs, partial products meet in `out` through TMA reduce-add stores. */
typedef struct {
    int64_t rows;
    int32_t m, n;
    int32_t zero_first;          /* 1: out is zeroed first (on the stream); 0: the product is added to out */
    int32_t _pad;
    const float *g, *g_lo;       /* [rows, m] */
    const float *x, *x_lo;       /* [rows, n] */
    float *out;                  /* [m, n] */
    uint32_t *status;            /* [1] nullable: GOM_STATUS_TIMEOUT */
} GomLinearWgradArgs;
int gom_linear_wgrad(const GomLinearWgradArgs *a, gom_stream_t stream);

/* Input rows of the non-rigid deformation MLP (reference models/modules/non_rigid_module.py:15-72,128-140), one launch each
 * way: h0[r] = [ posevec[b] (cond) | enc(x_v) (6 multires) | 0 ... ] for r = b V + v, rows >= B V zero, and enc alone;
 * enc[k 6 + s 3 + d] = w_k (s ? cos : sin)(2^k x_d), w_k = (1 - cos(pi clamp(alpha - k, 0, 1))) / 2 (Hann window, :33-43).
 * backward: g_xyz from the encoding columns of g_h0 plus g_enc (either nullable); with xyz_frames = 1 the frames add up. */
typedef struct {
    int32_t n_frames, n_verts;
    int32_t xyz_frames;          /* 1 (one canonical vertex set for all frames) or n_frames */
    int32_t cond, multires;
    int32_t cols;                /* row length of h0 (>= cond + 6 multires; a multiple of 32 for the GEMM) */
    int64_t rows_padded;         /* rows of h0 / enc (>= B V; a multiple of 16 for the GEMM) */
    float alpha;                 /* multires * max(iter - kick_in, 0) / (full_band - kick_in) */
    int32_t _pad;
    const float *xyz;            /* [xyz_frames,3,V] */
    const float *posevec;        /* [B,cond] */
    float *h0;                   /* [rows_padded, cols] out */
    float *enc;                  /* [rows_padded, 6 multires] out */
    const float *g_h0, *g_enc;   /* backward */
    float *g_xyz;                /* [xyz_frames,3,V] backward */
} GomNonRigidInputArgs;
int gom_nonrigid_input_forward(const GomNonRigidInputArgs *a, gom_stream_t stream);
int gom_nonrigid_input_backward(const GomNonRigidInputArgs *a, gom_stream_t stream);

/* Rodrigues formula of reference utils/network_util.py:66-92 (theta = sqrt(eps + |r|^2), eps = 1e-5 there) and its backward.
 * rvec [n_rot,3] in groups of `group` consecutive rotations (one group per frame); R [n_rot / group, group + prepend_identity,
 * 3,3]: prepend_identity = 1 puts the identity in front of every group (the root joint of
 * models/modules/pose_refinement_module.py:39-48).  backward: g_rvec [n_rot,3] from g_R (same shape as R). */
typedef struct {
    int32_t n_rot, group, prepend_identity;
    float eps;
    const float *rvec;
    float *R;
    const float *g_R;
    float *g_rvec;
} GomRodriguesArgs;
int gom_rodrigues_forward(const GomRodriguesArgs *a, gom_stream_t stream);
int gom_rodrigues_backward(const GomRodriguesArgs *a, gom_stream_t stream);

/* Linear layer with 1 .. 4 outputs over a tall batch of rows (the 128 -> 3 output layer of reference
 * models/modules/non_rigid_module.py:112-118): y = x W^T + b; backward: g_x = g_y W (nullable), g_weight = g_y^T x and
 * g_bias = column sums of g_y (both zeroed first).  x [rows, c_in] row-major, c_in a multiple of 4, at most 256. */
typedef struct {
    int64_t rows;
    int32_t c_in, n_out;
    const float *x;              /* [rows, c_in] */
    const float *weight;         /* [n_out, c_in] */
    const float *bias;           /* [n_out] nullable */
    float *y;                    /* [rows, n_out] (forward) */
    const float *g_y;            /* [rows, n_out] (backward) */
    float *g_x;                  /* [rows, c_in] (backward, nullable) */
    float *g_weight;             /* [n_out, c_in] (backward) */
    float *g_bias;               /* [n_out] (backward, nullable) */
} GomNarrowLinearArgs;
int gom_narrow_linear_forward(const GomNarrowLinearArgs *a, gom_stream_t stream);
int gom_narrow_linear_backward(const GomNarrowLinearArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Evaluation metrics.  Replaces reference eval.py:101-108 (Evaluator.psnr_metric / ssim_metric: skimage 0.18
 * structural_similarity defaults — 7x7 uniform window, sample covariance, K1 .01, K2 .03, data_range 2, 3-px crop)
 * and the 8-bit quantisation before them (utils/image_util.py:21-22, eval.py:355-361).
 *   quantize = 1: pred / gt are raw float images, quantised as uint8(255 * clip(v, 0, 1)) (truncation) first;
 *   quantize = 0: pred / gt already are k / 255 (what Evaluator.evaluate receives).
 * ssim_sum[b]   = sum of the per-pixel, per-channel SSIM over the cropped image: ssim = ssim_sum / (3 (H-6) (W-6));
 * sq_err_sum[b] = sum of squared 8-bit differences: mse = sq_err_sum / (65025 * 3 H W), psnr = -10 log10(mse).
 */
typedef struct {
    int32_t n_frames, height, width, quantize;
    const float *pred;           /* [B,H,W,3] */
    const float *gt;             /* [B,H,W,3] */
    double *ssim_sum;            /* [B] out */
    uint64_t *sq_err_sum;        /* [B] out */
    uint8_t *pred_8b;            /* [B,H,W,3] out, nullable: the quantised prediction (what eval.py writes to PNG) */
} GomEvalMetricsArgs;
int gom_eval_metrics(const GomEvalMetricsArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Mesh normal map + soft silhouette.  Replaces what reference models/modules/renderer/mesh.py:66-128 asks of PyTorch3D
 * 0.7.0 on every Model.forward (models/model.py:271-274): MeshRasterizer(faces_per_pixel 1, blur 0) + NormalShader
 * (per pixel: sum of the nearest inside face's three vertex normals, 0 on the background), and, when soft = 1
 * (training), MeshRenderer(SoftSilhouetteShader) with faces_per_pixel = K and blur_radius:
 *   alpha = 1 - prod_k (1 - sigmoid(-d_k / 1e-4)) over the K faces of smallest interpolated z whose signed squared
 *   NDC distance d_k to the pixel centre is negative (inside) or below blur_radius.
 * verts_ndc are the output of the reference's ndc_T_world (utils/pc_util.py:30-46): x, y in NDC (+X left, +Y up, the
 * shorter image side spans [-1,1]), z = camera depth.  Faces are binned to 8x8-pixel tiles (T = ceil(W/8) ceil(H/8)) into caller-owned lists of
 * fixed capacity (overflow -> GOM_STATUS_OVERFLOW in status[b]).
 * backward: dL_dvert_normals from dL_dnormal_map, dL_dverts_ndc (x, y; z gets 0) from dL_dalpha; both zeroed first.
 */
typedef struct {
    int32_t n_frames, n_verts, n_faces, height, width;
    int32_t faces_int64;         /* faces are int64 (the model's registered buffer) instead of int32 */
    int32_t soft;                /* 1: also render the soft silhouette (training) */
    int32_t faces_per_pixel;     /* K of the soft pass (reference: 50) */
    float blur_radius;           /* of the soft pass, squared NDC units (reference: ln(1/1e-4 - 1) * sigma) */
    int32_t _pad;
    int64_t list_capacity;       /* per-frame capacity of face_list */
    const float *verts_ndc;      /* [B,V,3] */
    const void *faces;           /* [F,3] */
    const float *vert_normals;   /* [B,V,3] (already rotated into the camera frame, model.py:271-273) */
    uint32_t *tile_count;        /* [B,T]   */
    uint32_t *tile_offset;       /* [B,T+1] */
    uint32_t *tile_cursor;       /* [B,T]   */
    uint32_t *face_list;         /* [B,cap] face ids grouped by tile */
    uint32_t *status;            /* [B] */
    uint32_t *worklist;          /* [B*T + 4] all tiles by decreasing list length (written by forward, read by backward), then
                                    the number of non-empty tiles and the work counters of the persistent tile kernels */
    int32_t *pix_to_face;        /* [B,H,W] nearest inside face, -1 = none */
    float *normal_map;           /* [B,H,W,3] */
    float *alpha;                /* [B,H,W]   (soft) */
    float *zcut;                 /* [B,H,W]   (soft) K-nearest cut kept for backward: +inf = no truncation */
    int32_t *idcut;              /* [B,H,W]   (soft) */
    const float *dL_dnormal_map; /* [B,H,W,3] (backward, nullable) */
    const float *dL_dalpha;      /* [B,H,W]   (backward, nullable) */
    float *dL_dverts_ndc;        /* [B,V,3]   (backward) */
    float *dL_dvert_normals;     /* [B,V,3]   (backward) */
} GomMeshRasterArgs;
int gom_mesh_raster_forward(const GomMeshRasterArgs *a, gom_stream_t stream);
int gom_mesh_raster_backward(const GomMeshRasterArgs *a, gom_stream_t stream);

/* Inputs of the mesh renderer, one launch each way instead of torch's gather / cross / index_add / normalize chains.
 * gom_vertex_normals_*: PyTorch3D Meshes.verts_normals_padded (per face the three corner cross products accumulated on its
 *   vertices, n / max(|n|, 1e-6)) rotated into the camera frame by E[:3,:3] — reference models/model.py:271-273.
 * gom_ndc_*: reference utils/pc_util.py:30-46 ndc_T_world: (x_ndc, y_ndc, z_cam) per vertex. */
typedef struct {
    int32_t n_frames, n_verts, n_faces;
    int32_t faces_int64;
    const float *verts;          /* [B,3,V] posed vertices (world) */
    const void *faces;           /* [F,3] */
    const float *E;              /* [B,4,4] */
    float *acc;                  /* [B,V,3] unnormalised normals: written by forward, read by backward */
    float *normals_cam;          /* [B,V,3] out */
    const float *dL_dnormals_cam;/* [B,V,3] (backward) */
    float *scratch;              /* [B,V,3] (backward) */
    float *dL_dverts;            /* [B,3,V] (backward; zeroed first) */
} GomVertexNormalsArgs;
int gom_vertex_normals_forward(const GomVertexNormalsArgs *a, gom_stream_t stream);
int gom_vertex_normals_backward(const GomVertexNormalsArgs *a, gom_stream_t stream);

typedef struct {
    int32_t n_frames, n_verts, height, width;
    const float *verts;          /* [B,3,V] */
    const float *K;              /* [B,3,3] */
    const float *E;              /* [B,4,4] */
    float *ndc;                  /* [B,V,3] out */
    const float *dL_dndc;        /* [B,V,3] (backward) */
    float *dL_dverts;            /* [B,3,V] (backward; overwritten) */
} GomNdcArgs;
int gom_ndc_forward(const GomNdcArgs *a, gom_stream_t stream);
int gom_ndc_backward(const GomNdcArgs *a, gom_stream_t stream);

/* reference train.py:137-146: sum |pred - maxpool_k(mask_gt)| over all pixels (stride 1, padding k/2, -inf padding like
 * F.max_pool2d; dilate = 0: plain L1) into sum[0] (zeroed first), and grad = sign(pred - dilated) * grad_scale (nullable). */
typedef struct {
    int32_t n_frames, height, width;
    int32_t kernel_size;         /* odd, <= 15 */
    int32_t dilate;
    float grad_scale;            /* e.g. 1 / (B H W) for the mean */
    const float *pred;           /* [B,H,W] soft silhouette of the mesh renderer */
    const float *mask_gt;        /* [B,H,W] */
    double *sum;                 /* [1] */
    float *grad;                 /* [B,H,W] nullable */
} GomDilatedMaskL1Args;
int gom_dilated_mask_l1(const GomDilatedMaskL1Args *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Adam over the flat parameter arena, one launch.  Replaces `optimizer.step()` of reference train.py:339 for the
 * parameter groups of models/model.py:305-324 (torch.optim.Adam semantics, amsgrad off, weight decay 0).  Segment s
 * covers arena elements [seg_end[s-1], seg_end[s]) with learning rate seg_lr[s]; grad_scale multiplies the gradient
 * first (1 / world_size after the summing all-reduce).
 * torch.optim.Adam keeps one step counter per PARAMETER and skips parameters whose .grad is None (the reference's
 * non-rigid / pose-refinement MLPs before their kick_in_iter): a segment with seg_active[s] = 0 is left untouched
 * (parameter, both moments and its counter), and the bias corrections 1 - beta^t use the segment's own t.
 * Step counters live on the host (dev_steps = NULL: the caller passes seg_step[s] = t of THIS step) or on the device
 * (dev_steps != NULL: int64 [GOM_ADAM_MAX_SEGMENTS + 1], counters of the COMPLETED steps per segment and, last, of the whole
 * optimizer; the call reads t = dev_steps[s] + 1 and a second tiny launch increments the counters of the active segments
 * afterwards) — the second form has no step-dependent kernel argument, so the step can sit inside a CUDA graph.  With
 * lr_decay_steps > 0 the learning rate is seg_lr[s] * lr_decay_rate^(iter / lr_decay_steps), iter = the optimizer-wide
 * counter (reference train.py:166-175: update_lr, exponential decay), evaluated on the device in the second form.
 */
#define GOM_ADAM_MAX_SEGMENTS 16
typedef struct {
    int64_t n;
    float *param;                /* [n] updated in place */
    const float *grad;           /* [n] */
    float *exp_avg;              /* [n] */
    float *exp_avg_sq;           /* [n] */
    float beta1, beta2, eps, grad_scale;
    float lr_decay_rate, lr_decay_steps;     /* lr_decay_steps <= 0: constant learning rates */
    int32_t n_segments, _pad;
    int64_t *dev_steps;          /* nullable, see above */
    int64_t iter;                /* host form: optimizer-wide step index of this step (0-based) for the decay */
    int64_t seg_end[GOM_ADAM_MAX_SEGMENTS];
    int64_t seg_step[GOM_ADAM_MAX_SEGMENTS]; /* host form: t (>= 1) of this step for the segment */
    float seg_lr[GOM_ADAM_MAX_SEGMENTS];
    int32_t seg_active[GOM_ADAM_MAX_SEGMENTS];
} GomAdamArgs;
int gom_adam_step(const GomAdamArgs *a, gom_stream_t stream);

/* --------------------------------------------------------------------------------------------------------------
 * Pseudo-shading MLP on the foreground pixels of the normal map.  Replaces `self.shadow_module(normal.reshape(-1, H*W, 3))`
 * of reference models/model.py:281-283, i.e. models/modules/shadow_module.py:107-117 (positional encoding
 * [x, sin(2^k x), cos(2^k x)]_k<multires of :14-62, then Linear/ReLU x depth, Linear(width,1), sigmoid) for the
 * configurations the reference ships (exps/ all have mlp_width 128, mlp_depth 3, skips beyond the depth, multires 6).
 * One call = pixel compaction (normal != 0; the background gets sigmoid(MLP(posenc(0))) as a constant), weight
 * preparation and one persistent tcgen05 kernel (3xTF32 split products, fp32 accumulation in tensor memory).
 * out[p] is the sigmoid output for EVERY pixel (the caller multiplies by 2, model.py:283).
 * Backward (depth <= 3): call forward with save_hidden = 1 — it then also writes every matrix operand it formed as
 * TF32 hi/lo "activation images" into act_img (private layout, gom_shadow_mlp_tile_words(depth, 0) words per tile of 128
 * foreground rows; rows >= capacity are dropped and GOM_STATUS_OVERFLOW is set in status[0]) — and later
 * gom_shadow_mlp_backward with the SAME struct plus g_out: it writes dL/dnormal of every FOREGROUND pixel into g_normals
 * (background entries are left untouched) and the parameter gradients of the foreground rows into g_W_in ... g_b_out.
 * The background pixels share one row (normal 0): gom_shadow_mlp_background_prepare (BEFORE the backward: g_normals of every
 * pixel — the row's input gradient times g_out on the background, 0 elsewhere — and the sum of g_out over the background) and
 * gom_shadow_mlp_background_apply (AFTER it: parameter gradients += that sum times the row's gradients) add their share.
 * Two tcgen05 kernels (data gradients; split-K weight gradients with per-CTA partials) + a fixed-order reduction.
 */
typedef struct {
    int64_t n_pixels;            /* B*H*W */
    int64_t capacity;            /* foreground rows act_img / dz_img hold (save_hidden / backward): a multiple of 128 */
    int32_t multires;            /* positional-encoding octaves (6): encoding width 3 + 6*multires <= 63 */
    int32_t width;               /* must be 128 */
    int32_t depth;               /* number of Linear+ReLU layers (mlp_depth, 3); 1..8 */
    int32_t save_hidden;
    const float *normals;        /* [n_pixels,3], 16-byte aligned */
    const float *W_in, *b_in;    /* [width, 3+6*multires], [width]            (block_mlps.0) */
    const float *W_hid, *b_hid;  /* [depth-1, width, width], [depth-1, width]  (block_mlps.2, .4, ...) */
    const float *W_out, *b_out;  /* [width], [1]                               (last Linear) */
    uint32_t *block_count;       /* scratch [ceil(n_pixels/1024)] */
    int32_t *fg_index;           /* [n_pixels] out: foreground pixel ids in pixel order (first n_fg entries valid) */
    int32_t *n_fg;               /* [1] out */
    float *w_images;             /* scratch, gom_shadow_mlp_weight_image_bytes(depth) bytes, 128-byte aligned */
    float *bg_value;             /* [1] out: the background constant */
    float *out;                  /* [n_pixels] out */
    void *act_img;               /* [capacity/128 tiles, gom_shadow_mlp_tile_words(depth,0)] u32 out (save_hidden), else NULL */
    uint32_t *status;            /* [1] out: GOM_STATUS_OVERFLOW | GOM_STATUS_TIMEOUT */
    /* ---- backward only */
    const float *g_out;          /* [n_pixels] dL/dout */
    void *dz_img;                /* [capacity/128, gom_shadow_mlp_tile_words(depth,1)] u32 scratch, ZERO-INITIALISED once by the caller */
    float *g_normals;            /* [n_pixels,3]: foreground rows are written */
    float *dzo_sums;             /* scratch [capacity/128 * 4] */
    float *partials;             /* scratch [gom_shadow_mlp_num_ctas(), gom_shadow_mlp_partial_floats()] */
    float *g_W_in, *g_b_in;      /* [width, 3+6*multires], [width] out */
    float *g_W_hid, *g_b_hid;    /* [depth-1, width, width], [depth-1, width] out */
    float *g_w_out, *g_b_out;    /* [width], [1] out */
    float *bg_scratch;           /* [gom_shadow_mlp_bg_scratch_floats()] (gom_shadow_mlp_background_* only) */
} GomShadowMlpArgs;
int gom_shadow_mlp_forward(const GomShadowMlpArgs *a, gom_stream_t stream);
int gom_shadow_mlp_backward(const GomShadowMlpArgs *a, gom_stream_t stream);
int gom_shadow_mlp_background_prepare(const GomShadowMlpArgs *a, gom_stream_t stream);
int gom_shadow_mlp_background_apply(const GomShadowMlpArgs *a, gom_stream_t stream);
size_t gom_shadow_mlp_bg_scratch_floats(void);
size_t gom_shadow_mlp_weight_image_bytes(int depth);
size_t gom_shadow_mlp_tile_words(int depth, int which);     /* which: 0 = act_img, 1 = dz_img */
size_t gom_shadow_mlp_partial_floats(void);
int gom_shadow_mlp_num_ctas(void);                          /* CTAs of the persistent kernels = SMs of the current device */

/* --------------------------------------------------------------------------------------------------------------
 * Mesh regularisers of reference train.py:123-160 for B posed meshes of one topology: uniform Laplacian smoothing
 * (utils/network_util.py:669-792, method "uniform"), normal consistency (pytorch3d.loss.mesh_normal_consistency,
 * train.py:149) and colour consistency (utils/network_util.py:795-799).  One call writes the three SUMS (the caller
 * divides: B*V, B*P, 3*Pc) and the unit gradients of the three MEANS; the caller scales them by the loss coefficients.
 * Topology (CSR adjacency of the unique edges, per pair the shared edge and the two opposite vertices) is static between
 * subdivisions and supplied by the caller (regularizers.py builds it once).
 */
typedef struct {
    int32_t n_frames, n_verts, n_pairs, n_faces;   /* n_pairs = P: rows of pair_vid (normal consistency) */
    int32_t do_laplacian, do_normal, do_color;
    int32_t n_color_pairs;                          /* Pc: rows of pair_face (colour consistency; the reference's
                                                       face_connectivity misses the pair of the last edge, so Pc = P - 1
                                                       on a closed mesh: models/model.py:119-123) */
    const float *verts;          /* [B,3,V] (the model's vertices_observation layout) */
    const int32_t *row_ptr;      /* [V+1] CSR adjacency */
    const int32_t *col;          /* [2E] */
    const int32_t *pair_vid;     /* [P,4] v0, v1 (shared edge, v0 < v1), opposite vertex of face a, of face b */
    const int32_t *pair_face;    /* [Pc,2] the two faces (colour term) */
    const float *colors;         /* [F,3] */
    float *lap;                  /* [B,3,V] scratch */
    double *sums;                /* [3] out: sum |lap|^2, sum (1 - cos), sum |dcol| */
    float *g_verts_lap;          /* [B,3,V] out: d mean|lap|^2 / d verts */
    float *g_verts_nc;           /* [B,3,V] out: d mean(1 - cos) / d verts */
    float *g_colors;             /* [F,3]   out: d mean|dcol| / d colors */
} GomMeshRegArgs;
int gom_mesh_regularizers(const GomMeshRegArgs *a, gom_stream_t stream);

/* struct sizes, so that a foreign-language binding can assert its mirror of the structs */
size_t gom_sizeof_camera_args(void);
size_t gom_sizeof_raster_fwd_args(void);
size_t gom_sizeof_raster_bwd_args(void);
size_t gom_sizeof_joint_fwd_args(void);
size_t gom_sizeof_joint_bwd_args(void);
size_t gom_sizeof_lbs_fwd_args(void);
size_t gom_sizeof_lbs_bwd_args(void);
size_t gom_sizeof_face_fwd_args(void);
size_t gom_sizeof_face_bwd_args(void);
size_t gom_sizeof_photo_args(void);
size_t gom_sizeof_shade_args(void);
size_t gom_sizeof_lpips_input_args(void);
size_t gom_sizeof_bias_relu_args(void);
size_t gom_sizeof_relu_bwd_args(void);
size_t gom_sizeof_lpips_tap_args(void);
size_t gom_sizeof_eval_metrics_args(void);
size_t gom_sizeof_conv_first_args(void);
size_t gom_sizeof_conv3x3_args(void);
size_t gom_sizeof_conv_pack_args(void);
size_t gom_sizeof_tf32_split_args(void);
size_t gom_sizeof_linear_wgrad_args(void);
size_t gom_sizeof_nonrigid_input_args(void);
size_t gom_sizeof_rodrigues_args(void);
size_t gom_sizeof_narrow_linear_args(void);
size_t gom_sizeof_adam_args(void);
size_t gom_sizeof_mesh_raster_args(void);
size_t gom_sizeof_vertex_normals_args(void);
size_t gom_sizeof_ndc_args(void);
size_t gom_sizeof_dilated_mask_l1_args(void);
size_t gom_sizeof_shadow_mlp_args(void);
size_t gom_sizeof_mesh_reg_args(void);

#ifdef __cplusplus
}
#endif
#endif /* GOM_B200_H */

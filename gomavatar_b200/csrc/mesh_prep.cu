// mesh_prep.cu — the small per-vertex / per-face / per-pixel passes around the mesh normal-map renderer and its loss, each a
// single launch instead of the ~20-60 eager elementwise / gather / scatter kernels torch autograd needs for them (at the
// reference's batch of one frame those launches, not their bytes, are the cost: SURVEY.md §8 f-1, f-3):
//   * gom_vertex_normals_*  : PyTorch3D `Meshes.verts_normals_padded` (area-weighted face normals accumulated on the vertices,
//                             normalised with eps 1e-6) rotated into the camera frame — reference models/model.py:271-273;
//   * gom_ndc_*             : reference utils/pc_util.py:30-46 `ndc_T_world` (world -> camera -> pixel -> PyTorch3D NDC, z = depth);
//   * gom_dilated_mask_l1   : reference train.py:137-146 — L1 between the soft mesh silhouette and the k x k max-pooled
//                             (dilated) ground-truth mask, value and gradient in one pass.
// Layouts: vertices [B,3,V] (the model's SoA layout), per-vertex outputs [B,V,3] (what csrc/mesh_raster.cu reads).
#include <math.h>

#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;

struct VnDev {
    int B, V, F, faces_int64;
    const float *verts;            // [B,3,V]
    const void *faces;             // [F,3]
    const float *E;                // [B,4,4]
    float *acc;                    // [B,V,3] unnormalised vertex normals
    float *out;                    // [B,V,3] normalised, camera frame
    const float *g_out;            // backward
    float *g_acc;                  // [B,V,3] scratch
    float *g_verts;                // [B,3,V]
};

__device__ __forceinline__ int3 vn_face(const VnDev &a, int f) {
    if (a.faces_int64) {
        const long long *p = reinterpret_cast<const long long *>(a.faces) + 3LL * f;
        return make_int3((int)p[0], (int)p[1], (int)p[2]);
    }
    const int *p = reinterpret_cast<const int *>(a.faces) + 3LL * f;
    return make_int3(p[0], p[1], p[2]);
}
__device__ __forceinline__ float3 ld3(const float *v, int V, int i) { return make_float3(v[i], v[V + i], v[2 * V + i]); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ void atomic_add3(float *p, float3 v) { atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); }

// per (frame, face): the three corner cross products of PyTorch3D (each is twice the area times the face normal)
__global__ void __launch_bounds__(kThreads) k_vn_faces(VnDev a) {
    const int b = blockIdx.y, f = blockIdx.x * kThreads + threadIdx.x;
    if (f >= a.F) return;
    const int3 id = vn_face(a, f);
    const float *v = a.verts + (long long)b * 3 * a.V;
    const float3 v0 = ld3(v, a.V, id.x), v1 = ld3(v, a.V, id.y), v2 = ld3(v, a.V, id.z);
    float *acc = a.acc + (long long)b * a.V * 3;
    atomic_add3(acc + 3 * id.y, cross3(sub3(v2, v1), sub3(v0, v1)));
    atomic_add3(acc + 3 * id.z, cross3(sub3(v0, v2), sub3(v1, v2)));
    atomic_add3(acc + 3 * id.x, cross3(sub3(v1, v0), sub3(v2, v0)));
}

// per (frame, vertex): n / max(|n|, 1e-6) (torch.nn.functional.normalize), then R n with R = E[:3,:3]
__global__ void __launch_bounds__(kThreads) k_vn_finish(VnDev a) {
    const int b = blockIdx.y, i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.V) return;
    const float *n = a.acc + ((long long)b * a.V + i) * 3;
    const float *E = a.E + b * 16;
    const float nx = n[0], ny = n[1], nz = n[2];
    const float inv = 1.f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-6f);
    const float x = nx * inv, y = ny * inv, z = nz * inv;
    float *o = a.out + ((long long)b * a.V + i) * 3;
    o[0] = E[0] * x + E[1] * y + E[2] * z;
    o[1] = E[4] * x + E[5] * y + E[6] * z;
    o[2] = E[8] * x + E[9] * y + E[10] * z;
}

__global__ void __launch_bounds__(kThreads) k_vn_finish_bwd(VnDev a) {
    const int b = blockIdx.y, i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.V) return;
    const float *n = a.acc + ((long long)b * a.V + i) * 3;
    const float *g = a.g_out + ((long long)b * a.V + i) * 3;
    const float *E = a.E + b * 16;
    // gradient w.r.t. the normalised normal: R^T g
    const float hx = E[0] * g[0] + E[4] * g[1] + E[8] * g[2];
    const float hy = E[1] * g[0] + E[5] * g[1] + E[9] * g[2];
    const float hz = E[2] * g[0] + E[6] * g[1] + E[10] * g[2];
    const float nx = n[0], ny = n[1], nz = n[2];
    const float len = sqrtf(nx * nx + ny * ny + nz * nz);
    float ox, oy, oz;
    if (len > 1e-6f) {                        // d(n / |n|): (h - nhat (nhat . h)) / |n|
        const float inv = 1.f / len, x = nx * inv, y = ny * inv, z = nz * inv, d = x * hx + y * hy + z * hz;
        ox = (hx - x * d) * inv; oy = (hy - y * d) * inv; oz = (hz - z * d) * inv;
    } else {                                  // clamped denominator: n / 1e-6
        ox = hx * 1e6f; oy = hy * 1e6f; oz = hz * 1e6f;
    }
    float *o = a.g_acc + ((long long)b * a.V + i) * 3;
    o[0] = ox; o[1] = oy; o[2] = oz;
}

// g . (p x q): d/dp = q x g, d/dq = g x p
__global__ void __launch_bounds__(kThreads) k_vn_faces_bwd(VnDev a) {
    const int b = blockIdx.y, f = blockIdx.x * kThreads + threadIdx.x;
    if (f >= a.F) return;
    const int3 id = vn_face(a, f);
    const float *v = a.verts + (long long)b * 3 * a.V;
    const float3 v0 = ld3(v, a.V, id.x), v1 = ld3(v, a.V, id.y), v2 = ld3(v, a.V, id.z);
    const float *ga = a.g_acc + (long long)b * a.V * 3;
    const float3 g0 = make_float3(ga[3 * id.x], ga[3 * id.x + 1], ga[3 * id.x + 2]);
    const float3 g1 = make_float3(ga[3 * id.y], ga[3 * id.y + 1], ga[3 * id.y + 2]);
    const float3 g2 = make_float3(ga[3 * id.z], ga[3 * id.z + 1], ga[3 * id.z + 2]);
    float3 d0 = make_float3(0.f, 0.f, 0.f), d1 = d0, d2 = d0;
    {   // normal at v1 += (v2 - v1) x (v0 - v1)
        const float3 p = sub3(v2, v1), q = sub3(v0, v1), dp = cross3(q, g1), dq = cross3(g1, p);
        d2 = add3(d2, dp); d0 = add3(d0, dq); d1 = sub3(d1, add3(dp, dq));
    }
    {   // normal at v2 += (v0 - v2) x (v1 - v2)
        const float3 p = sub3(v0, v2), q = sub3(v1, v2), dp = cross3(q, g2), dq = cross3(g2, p);
        d0 = add3(d0, dp); d1 = add3(d1, dq); d2 = sub3(d2, add3(dp, dq));
    }
    {   // normal at v0 += (v1 - v0) x (v2 - v0)
        const float3 p = sub3(v1, v0), q = sub3(v2, v0), dp = cross3(q, g0), dq = cross3(g0, p);
        d1 = add3(d1, dp); d2 = add3(d2, dq); d0 = sub3(d0, add3(dp, dq));
    }
    float *gv = a.g_verts + (long long)b * 3 * a.V;
    atomicAdd(gv + id.x, d0.x); atomicAdd(gv + a.V + id.x, d0.y); atomicAdd(gv + 2 * a.V + id.x, d0.z);
    atomicAdd(gv + id.y, d1.x); atomicAdd(gv + a.V + id.y, d1.y); atomicAdd(gv + 2 * a.V + id.y, d1.z);
    atomicAdd(gv + id.z, d2.x); atomicAdd(gv + a.V + id.z, d2.y); atomicAdd(gv + 2 * a.V + id.z, d2.z);
}

// ------------------------------------------------------------------------------------------------- ndc_T_world
struct NdcDev {
    int B, V;
    float inv_s2, cx, cy;          // xs = cx - x * inv_s2, ys = cy - y * inv_s2 (inv_s2 = 2 / min(H, W))
    const float *verts, *K, *E;
    float *ndc;
    const float *g_ndc;
    float *g_verts;
};

template <bool BWD> __global__ void __launch_bounds__(kThreads) k_ndc(NdcDev a) {
    const int b = blockIdx.y, i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.V) return;
    const float *v = a.verts + (long long)b * 3 * a.V;
    const float *E = a.E + b * 16, *K = a.K + b * 9;
    const float px = v[i], py = v[a.V + i], pz = v[2 * a.V + i];
    float q[4];
#pragma unroll
    for (int r = 0; r < 4; r++) q[r] = E[4 * r] * px + E[4 * r + 1] * py + E[4 * r + 2] * pz + E[4 * r + 3];
    const float iw = 1.f / q[3];
    const float c0 = q[0] * iw, c1 = q[1] * iw, c2 = q[2] * iw;
    const float u0 = K[0] * c0 + K[1] * c1 + K[2] * c2, u1 = K[3] * c0 + K[4] * c1 + K[5] * c2, u2 = K[6] * c0 + K[7] * c1 + K[8] * c2;
    const float iu = 1.f / u2;
    if (!BWD) {
        float *o = a.ndc + ((long long)b * a.V + i) * 3;
        o[0] = a.cx - (u0 * iu) * a.inv_s2;
        o[1] = a.cy - (u1 * iu) * a.inv_s2;
        o[2] = c2;
    } else {
        const float *g = a.g_ndc + ((long long)b * a.V + i) * 3;
        const float gx = -a.inv_s2 * g[0], gy = -a.inv_s2 * g[1];
        const float gu0 = gx * iu, gu1 = gy * iu, gu2 = -(gx * u0 + gy * u1) * iu * iu;
        const float gc0 = K[0] * gu0 + K[3] * gu1 + K[6] * gu2;
        const float gc1 = K[1] * gu0 + K[4] * gu1 + K[7] * gu2;
        const float gc2 = K[2] * gu0 + K[5] * gu1 + K[8] * gu2 + g[2];
        const float gq0 = gc0 * iw, gq1 = gc1 * iw, gq2 = gc2 * iw, gq3 = -(gc0 * c0 + gc1 * c1 + gc2 * c2) * iw;
        float *o = a.g_verts + (long long)b * 3 * a.V;
        o[i] = E[0] * gq0 + E[4] * gq1 + E[8] * gq2 + E[12] * gq3;
        o[a.V + i] = E[1] * gq0 + E[5] * gq1 + E[9] * gq2 + E[13] * gq3;
        o[2 * a.V + i] = E[2] * gq0 + E[6] * gq1 + E[10] * gq2 + E[14] * gq3;
    }
}

// ------------------------------------------------------------------------------------------- dilated mask L1
constexpr int kDilTileW = 32, kDilTileH = 8, kDilMaxK = 15;

__global__ void __launch_bounds__(kDilTileW * kDilTileH) k_dilated_mask_l1(GomDilatedMaskL1Args a) {
    __shared__ float s_gt[(kDilTileH + kDilMaxK - 1) * (kDilTileW + kDilMaxK - 1)];
    __shared__ float s_part[kDilTileW * kDilTileH / 32];
    const int r = a.dilate ? a.kernel_size / 2 : 0;
    const int pw = kDilTileW + 2 * r, ph = kDilTileH + 2 * r;
    const int b = blockIdx.z, x0 = blockIdx.x * kDilTileW, y0 = blockIdx.y * kDilTileH;
    const int tid = threadIdx.y * kDilTileW + threadIdx.x;
    const float *gt = a.mask_gt + (long long)b * a.height * a.width;
    for (int i = tid; i < pw * ph; i += kDilTileW * kDilTileH) {
        const int yy = y0 - r + i / pw, xx = x0 - r + i % pw;
        s_gt[i] = (yy >= 0 && yy < a.height && xx >= 0 && xx < a.width) ? gt[(long long)yy * a.width + xx] : -INFINITY;   // max_pool2d pads with -inf
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    float d = 0.f;
    if (x < a.width && y < a.height) {
        float m = -INFINITY;
        for (int dy = 0; dy <= 2 * r; dy++)
            for (int dx = 0; dx <= 2 * r; dx++) m = fmaxf(m, s_gt[(threadIdx.y + dy) * pw + threadIdx.x + dx]);
        const long long pix = ((long long)b * a.height + y) * a.width + x;
        const float e = a.pred[pix] - m;
        d = fabsf(e);
        if (a.grad) a.grad[pix] = (e > 0.f ? 1.f : e < 0.f ? -1.f : 0.f) * a.grad_scale;       // torch: d|e|/de = sign(e), 0 at 0
    }
    d = warp_sum(d);
    if ((tid & 31) == 0) s_part[tid >> 5] = d;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < kDilTileW * kDilTileH / 32; i++) s += (double)s_part[i];
        atomicAdd(a.sum, s);
    }
}

int fill_vn(const GomVertexNormalsArgs *p, VnDev &a) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->n_verts > 0 && p->n_faces >= 0, "sizes");
    GOM_REQUIRE(p->verts && p->faces && p->E && p->acc, "null pointer");
    a.B = p->n_frames; a.V = p->n_verts; a.F = p->n_faces; a.faces_int64 = p->faces_int64;
    a.verts = p->verts; a.faces = p->faces; a.E = p->E; a.acc = p->acc; a.out = p->normals_cam;
    a.g_out = p->dL_dnormals_cam; a.g_acc = p->scratch; a.g_verts = p->dL_dverts;
    return GOM_OK;
}

int fill_ndc(const GomNdcArgs *p, NdcDev &a) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->n_verts > 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->verts && p->K && p->E, "null pointer");
    a.B = p->n_frames; a.V = p->n_verts;
    const float H = (float)p->height, W = (float)p->width;
    if (p->height < p->width) { a.inv_s2 = 2.f / H; a.cx = W / H; a.cy = 1.f; }     // reference utils/pc_util.py:40-45
    else { a.inv_s2 = 2.f / W; a.cx = 1.f; a.cy = H / W; }
    a.verts = p->verts; a.K = p->K; a.E = p->E; a.ndc = p->ndc; a.g_ndc = p->dL_dndc; a.g_verts = p->dL_dverts;
    return GOM_OK;
}

}  // namespace

extern "C" int gom_vertex_normals_forward(const GomVertexNormalsArgs *p, gom_stream_t stream_) {
    VnDev a;
    if (int rc = fill_vn(p, a)) return rc;
    GOM_REQUIRE(p->normals_cam, "null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    GOM_CUDA(cudaMemsetAsync(a.acc, 0, sizeof(float) * 3 * (size_t)a.B * a.V, stream));
    if (a.F > 0) {
        k_vn_faces<<<dim3(gom_div_up(a.F, kThreads), a.B), kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
    }
    k_vn_finish<<<dim3(gom_div_up(a.V, kThreads), a.B), kThreads, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_vertex_normals_backward(const GomVertexNormalsArgs *p, gom_stream_t stream_) {
    VnDev a;
    if (int rc = fill_vn(p, a)) return rc;
    GOM_REQUIRE(p->dL_dnormals_cam && p->scratch && p->dL_dverts, "null gradient pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    GOM_CUDA(cudaMemsetAsync(a.g_verts, 0, sizeof(float) * 3 * (size_t)a.B * a.V, stream));
    k_vn_finish_bwd<<<dim3(gom_div_up(a.V, kThreads), a.B), kThreads, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    if (a.F > 0) {
        k_vn_faces_bwd<<<dim3(gom_div_up(a.F, kThreads), a.B), kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
    }
    return GOM_OK;
}

extern "C" int gom_ndc_forward(const GomNdcArgs *p, gom_stream_t stream_) {
    NdcDev a;
    if (int rc = fill_ndc(p, a)) return rc;
    GOM_REQUIRE(p->ndc, "null output");
    k_ndc<false><<<dim3(gom_div_up(a.V, kThreads), a.B), kThreads, 0, (cudaStream_t)stream_>>>(a);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_ndc_backward(const GomNdcArgs *p, gom_stream_t stream_) {
    NdcDev a;
    if (int rc = fill_ndc(p, a)) return rc;
    GOM_REQUIRE(p->dL_dndc && p->dL_dverts, "null gradient pointer");
    k_ndc<true><<<dim3(gom_div_up(a.V, kThreads), a.B), kThreads, 0, (cudaStream_t)stream_>>>(a);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_dilated_mask_l1(const GomDilatedMaskL1Args *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->kernel_size >= 1 && p->kernel_size <= kDilMaxK && (p->kernel_size & 1), "kernel_size must be odd, at most 15");
    GOM_REQUIRE(p->pred && p->mask_gt && p->sum, "null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    GOM_CUDA(cudaMemsetAsync(p->sum, 0, sizeof(double), stream));
    dim3 grid(gom_div_up(p->width, kDilTileW), gom_div_up(p->height, kDilTileH), p->n_frames), block(kDilTileW, kDilTileH);
    k_dilated_mask_l1<<<grid, block, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" size_t gom_sizeof_vertex_normals_args(void) { return sizeof(GomVertexNormalsArgs); }
extern "C" size_t gom_sizeof_ndc_args(void) { return sizeof(GomNdcArgs); }
extern "C" size_t gom_sizeof_dilated_mask_l1_args(void) { return sizeof(GomDilatedMaskL1Args); }

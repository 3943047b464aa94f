// mesh_raster.cu — normal-map rasterisation and soft silhouette of the posed mesh, forward + backward (sm_100a).
//
// Replaces what reference models/modules/renderer/mesh.py:66-128 asks of PyTorch3D 0.7.0 on EVERY Model.forward
// (models/model.py:271-274): `MeshRasterizer` (hard, faces_per_pixel = 1, blur 0) + `NormalShader`, and in training
// `MeshRenderer(SoftSilhouetteShader)` with faces_per_pixel = 50 and blur_radius = ln(1/1e-4 - 1) * sigma — both with
// `bin_size = 0`, i.e. PyTorch3D's NAIVE kernel that tests every pixel against every face (3.6 .. 14 G point-in-triangle
// tests per call).  Semantics follow PyTorch3D's `CheckPixelInsideFace` / geometry_utils (App. B;
// parity unpinned, oracle/mesh_raster.py).  B200-first design:
//   * faces are binned to 8x8-pixel tiles (a face with its 2.4-pixel blur margin spans ~6.5 pixels: 16x16 tiles made every
//     pixel test 4x more faces than necessary) by their blur-padded bounding boxes (count -> scan -> emit, no host sync,
//     fixed capacity + overflow flag exactly like the splat rasterizer), so a pixel only meets the faces of its tile;
//   * ONE pass renders the hard result (nearest inside face -> pix_to_face, summed vertex normals) and the soft
//     silhouette alpha = 1 - prod(1 - sigmoid(-d/1e-4)).  The product does not depend on order, so a pixel with at most
//     K candidates needs no queue; beyond K (common at 30 k faces: the 2.4-pixel blur disc meets ~15 faces per surface
//     layer) the K nearest by (z, face id) are tracked in a thread-local queue (replace-the-maximum with per-group
//     maxima: ~15 comparisons per replacement) and the cut is recorded for the backward;
//   * backward: d alpha / d d_k = -(1 - alpha) p_k / sigma (the (1 - p_k) factor cancels), chained through the
//     squared point-segment distance to the two vertices of the nearest edge; normal-map gradient scattered to the
//     hit face's three vertex normals.  A batch of B frames per launch.
#include <math.h>

#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;                // face-parallel kernels (count, emit)
constexpr int kBin = 8, kBinShift = 3;       // faces are binned to 8x8-pixel tiles; one 64-thread block rasterises one tile
constexpr int kPix = kBin * kBin;
constexpr int kMaxK = 64;                  // thread-local K-nearest queue of the soft silhouette (reference: K = 50)
constexpr float kEpsArea = 1e-8f;          // PyTorch3D kEpsilon
constexpr float kBlendSigma = 1e-4f;       // BlendParams().sigma (SoftSilhouetteShader default, mesh.py:107-112)

struct MeshDev {
    int B, V, F, H, W, gx, gy, T, K, faces_int64, soft;
    long long cap;
    float blur, S, sx, sy;                 // blur radius (NDC^2); S = min(H,W); sx = W/S, sy = H/S
    const float *verts;                    // [B,V,3] NDC x, y and camera z
    const void *faces;                     // [F,3]
    const float *vnormals;                 // [B,V,3]
    uint32_t *tile_count, *tile_offset, *tile_cursor, *face_list, *status;
    int32_t *pix_to_face; float *normal; float *alpha; float *zcut; int32_t *idcut;
    // backward
    const float *d_normal, *d_alpha; float *d_verts, *d_vnormals;
};

__device__ __forceinline__ int3 load_face(const MeshDev &a, int f) {
    if (a.faces_int64) {
        const long long *p = reinterpret_cast<const long long *>(a.faces) + 3LL * f;
        return make_int3((int)p[0], (int)p[1], (int)p[2]);
    }
    const int *p = reinterpret_cast<const int *>(a.faces) + 3LL * f;
    return make_int3(p[0], p[1], p[2]);
}

// pixel index range [lo, hi) whose centres can lie in [vmin, vmax] (NDC), with one pixel of slack: x = s - (2 i + 1)/S
__device__ __forceinline__ void ndc_to_pixel_range(float vmin, float vmax, float s, float S, int n, int &lo, int &hi) {
    const float a = ((s - vmax) * S - 1.0f) * 0.5f, b = ((s - vmin) * S - 1.0f) * 0.5f;
    lo = max(0, (int)floorf(a) - 1);
    hi = min(n, (int)ceilf(b) + 2);
}

// tile rectangle of a face (empty when it can never pass CheckPointOutsideBoundingBox / the zero-area cull)
__device__ __forceinline__ bool face_tiles(const MeshDev &a, int b, int f, int4 &rc) {
    const int3 id = load_face(a, f);
    const float *vb = a.verts + (long long)b * a.V * 3;
    const float ax = vb[3 * id.x], ay = vb[3 * id.x + 1], az = vb[3 * id.x + 2];
    const float bx = vb[3 * id.y], by = vb[3 * id.y + 1], bz = vb[3 * id.y + 2];
    const float cx = vb[3 * id.z], cy = vb[3 * id.z + 1], cz = vb[3 * id.z + 2];
    if (fmaxf(fmaxf(az, bz), cz) < kEpsArea) return false;
    const float area = (cx - ax) * (by - ay) - (cy - ay) * (bx - ax);
    if (area <= kEpsArea && area >= -kEpsArea) return false;
    if (!(isfinite(ax) && isfinite(ay) && isfinite(bx) && isfinite(by) && isfinite(cx) && isfinite(cy))) return false;
    const float br = sqrtf(a.blur);
    int x0, x1, y0, y1;
    ndc_to_pixel_range(fminf(fminf(ax, bx), cx) - br, fmaxf(fmaxf(ax, bx), cx) + br, a.sx, a.S, a.W, x0, x1);
    ndc_to_pixel_range(fminf(fminf(ay, by), cy) - br, fmaxf(fmaxf(ay, by), cy) + br, a.sy, a.S, a.H, y0, y1);
    if (x1 <= x0 || y1 <= y0) return false;
    rc = make_int4(x0 >> kBinShift, y0 >> kBinShift, (x1 + kBin - 1) >> kBinShift, (y1 + kBin - 1) >> kBinShift);
    return true;
}

__global__ void __launch_bounds__(kThreads) k_mesh_count(MeshDev a) {
    const int b = blockIdx.y, f = blockIdx.x * kThreads + threadIdx.x;
    if (f >= a.F) return;
    int4 rc;
    if (!face_tiles(a, b, f, rc)) return;
    uint32_t *cnt = a.tile_count + (long long)b * a.T;
    for (int y = rc.y; y < rc.w; y++)
        for (int x = rc.x; x < rc.z; x++) atomicAdd(cnt + y * a.gx + x, 1u);
}

__global__ void __launch_bounds__(1024) k_mesh_scan(MeshDev a) {
    __shared__ uint32_t wsum[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t *cnt = a.tile_count + (long long)b * a.T;
    uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
    uint32_t *cur = a.tile_cursor + (long long)b * a.T;
    unsigned long long carry = 0;
    for (int base = 0; base < a.T; base += 1024) {
        const int i = base + tid;
        const uint32_t v = i < a.T ? cnt[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const unsigned long long excl = carry + (x - v) + (wid > 0 ? wsum[wid - 1] : 0u);
        if (i < a.T) {
            const uint32_t e = (uint32_t)(excl > 0xffffffffULL ? 0xffffffffULL : excl);
            off[i] = e;
            cur[i] = e;
        }
        carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) {
        off[a.T] = (uint32_t)(carry > 0xffffffffULL ? 0xffffffffULL : carry);
        a.status[b] = carry > (unsigned long long)a.cap ? GOM_STATUS_OVERFLOW : 0u;
    }
}

__global__ void __launch_bounds__(kThreads) k_mesh_emit(MeshDev a) {
    const int b = blockIdx.y, f = blockIdx.x * kThreads + threadIdx.x;
    if (f >= a.F) return;
    int4 rc;
    if (!face_tiles(a, b, f, rc)) return;
    uint32_t *cur = a.tile_cursor + (long long)b * a.T;
    uint32_t *list = a.face_list + (long long)b * a.cap;
    for (int y = rc.y; y < rc.w; y++)
        for (int x = rc.x; x < rc.z; x++) {
            const uint32_t pos = atomicAdd(cur + y * a.gx + x, 1u);
            if ((long long)pos < a.cap) list[pos] = (uint32_t)f;
        }
}

// ------------------------------------------------------------------------------------------- per (pixel, face) test
struct FaceRec { float ax, ay, az, bx, by, bz, cx, cy, cz; };

__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

// squared distance to segment (a, b); also the clamped parameter and whether the segment is degenerate
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by, float &tt, bool &degenerate) {
    const float dx = bx - ax, dy = by - ay;
    const float l2 = dx * dx + dy * dy;
    degenerate = l2 <= kEpsArea;
    if (degenerate) { tt = 1.f; return (px - bx) * (px - bx) + (py - by) * (py - by); }
    const float t = ((px - ax) * dx + (py - ay) * dy) / l2;
    tt = fminf(fmaxf(t, 0.f), 1.f);
    const float qx = ax + tt * dx, qy = ay + tt * dy;
    return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

// PyTorch3D CheckPixelInsideFace up to the queue insertion.  Returns false when the face does not count for the pixel.
__device__ __forceinline__ bool pixel_face(const FaceRec &r, float px, float py, float blur, float br, float &pz, bool &inside,
                                           float &dist, int &edge, float &tt, bool &degenerate) {
    const float xmin = fminf(fminf(r.ax, r.bx), r.cx) - br, xmax = fmaxf(fmaxf(r.ax, r.bx), r.cx) + br;
    const float ymin = fminf(fminf(r.ay, r.by), r.cy) - br, ymax = fmaxf(fmaxf(r.ay, r.by), r.cy) + br;
    if (px > xmax || px < xmin || py > ymax || py < ymin) return false;
    const float area = edge_fn(r.cx, r.cy, r.ax, r.ay, r.bx, r.by) + kEpsArea;
    const float w0 = edge_fn(px, py, r.bx, r.by, r.cx, r.cy) / area;
    const float w1 = edge_fn(px, py, r.cx, r.cy, r.ax, r.ay) / area;
    const float w2 = edge_fn(px, py, r.ax, r.ay, r.bx, r.by) / area;
    pz = w0 * r.az + w1 * r.bz + w2 * r.cz;
    if (!(pz >= 0.f)) return false;
    inside = w0 > 0.f && w1 > 0.f && w2 > 0.f;
    float t01, t02, t12; bool g01, g02, g12;
    const float e01 = seg_dist2(px, py, r.ax, r.ay, r.bx, r.by, t01, g01);
    const float e02 = seg_dist2(px, py, r.ax, r.ay, r.cx, r.cy, t02, g02);
    const float e12 = seg_dist2(px, py, r.bx, r.by, r.cx, r.cy, t12, g12);
    dist = e01; edge = 0; tt = t01; degenerate = g01;
    if (e02 < dist) { dist = e02; edge = 1; tt = t02; degenerate = g02; }
    if (e12 < dist) { dist = e12; edge = 2; tt = t12; degenerate = g12; }
    return inside || dist < blur;
}

__device__ __forceinline__ FaceRec fetch_face(const MeshDev &a, int b, int f, int3 &id) {
    id = load_face(a, f);
    const float *vb = a.verts + (long long)b * a.V * 3;
    FaceRec r;
    r.ax = vb[3 * id.x]; r.ay = vb[3 * id.x + 1]; r.az = vb[3 * id.x + 2];
    r.bx = vb[3 * id.y]; r.by = vb[3 * id.y + 1]; r.bz = vb[3 * id.y + 2];
    r.cx = vb[3 * id.z]; r.cy = vb[3 * id.z + 1]; r.cz = vb[3 * id.z + 2];
    return r;
}

// lexicographic (z, face id) order used by the K-nearest selection
__device__ __forceinline__ bool zid_less(float z0, int f0, float z1, int f1) { return z0 < z1 || (z0 == z1 && f0 < f1); }

// ------------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(kPix) k_mesh_raster_fwd(MeshDev a) {
    __shared__ FaceRec s_rec[kPix];
    __shared__ int s_fid[kPix];
    const int b = blockIdx.z, tile = blockIdx.y * a.gx + blockIdx.x, tid = threadIdx.y * kBin + threadIdx.x;
    const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
    long long start = off[tile], end = off[tile + 1];
    if (start > a.cap) start = a.cap;
    if (end > a.cap) end = a.cap;
    const int n = (int)(end - start);
    const uint32_t *list = a.face_list + (long long)b * a.cap + start;
    const int x = blockIdx.x * kBin + threadIdx.x, y = blockIdx.y * kBin + threadIdx.y;
    const bool in_img = x < a.W && y < a.H;
    const float px = a.sx - (2.0f * x + 1.0f) / a.S, py = a.sy - (2.0f * y + 1.0f) / a.S;
    const float br = sqrtf(a.blur);

    float best_z = INFINITY; int best_f = -1;
    float prod = 1.f; int cand = 0;
    // the K candidates of smallest (z, face id) seen so far (thread-local arrays: local memory, L1-resident) and the
    // largest of them (mz, mf at slot mi); only consulted once a pixel has more than K candidates
    float hz[kMaxK], hp[kMaxK]; int hf[kMaxK];
    float gz[kMaxK / 8]; int gf[kMaxK / 8], gi[kMaxK / 8];       // per group of 8 slots: its largest key and where it sits
    float mz = -INFINITY; int mf = -1, mi = 0;
    const int K = a.K;
    const bool keep = a.soft && K <= kMaxK;
    for (int base = 0; base < n; base += kPix) {
        __syncthreads();
        if (base + tid < n) {
            int3 id;
            const int f = (int)list[base + tid];
            s_rec[tid] = fetch_face(a, b, f, id);
            s_fid[tid] = f;
        }
        __syncthreads();
        const int m = min(kPix, n - base);
        if (!in_img) continue;
        for (int j = 0; j < m; j++) {
            float pz, dist, tt; bool inside, deg; int edge;
            if (!pixel_face(s_rec[j], px, py, a.blur, br, pz, inside, dist, edge, tt, deg)) continue;
            const int f = s_fid[j];
            if (inside && zid_less(pz, f, best_z, best_f < 0 ? 0x7fffffff : best_f)) { best_z = pz; best_f = f; }
            if (a.soft) {
                const float p = 1.f / (1.f + __expf((inside ? -dist : dist) / kBlendSigma));     // sigmoid(-d / sigma)
                prod *= 1.f - p;
                if (keep) {
                    if (cand < K) {
                        hz[cand] = pz; hf[cand] = f; hp[cand] = p;
                        const int g = cand >> 3;
                        if ((cand & 7) == 0 || zid_less(gz[g], gf[g], pz, f)) { gz[g] = pz; gf[g] = f; gi[g] = cand; }
                        if (zid_less(mz, mf, pz, f)) { mz = pz; mf = f; mi = cand; }
                    } else if (zid_less(pz, f, mz, mf)) {                       // replaces the current K-th nearest
                        hz[mi] = pz; hf[mi] = f; hp[mi] = p;
                        const int g = mi >> 3, lo = g << 3, hi = min(lo + 8, K);
                        float tz = hz[lo]; int tf = hf[lo], ti = lo;            // new maximum of the touched group ...
                        for (int i = lo + 1; i < hi; i++)
                            if (zid_less(tz, tf, hz[i], hf[i])) { tz = hz[i]; tf = hf[i]; ti = i; }
                        gz[g] = tz; gf[g] = tf; gi[g] = ti;
                        mz = gz[0]; mf = gf[0]; mi = gi[0];                     // ... then of the group maxima
                        for (int gg = 1; gg < ((K + 7) >> 3); gg++)
                            if (zid_less(mz, mf, gz[gg], gf[gg])) { mz = gz[gg]; mf = gf[gg]; mi = gi[gg]; }
                    }
                }
                cand++;
            }
        }
    }
    if (!in_img) return;
    const long long pix = ((long long)b * a.H + y) * a.W + x;
    a.pix_to_face[pix] = best_f;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (best_f >= 0) {
        const int3 id = load_face(a, best_f);
        const float *vn = a.vnormals + (long long)b * a.V * 3;
        nx = vn[3 * id.x] + vn[3 * id.y] + vn[3 * id.z];
        ny = vn[3 * id.x + 1] + vn[3 * id.y + 1] + vn[3 * id.z + 1];
        nz = vn[3 * id.x + 2] + vn[3 * id.y + 2] + vn[3 * id.z + 2];
    }
    a.normal[3 * pix] = nx; a.normal[3 * pix + 1] = ny; a.normal[3 * pix + 2] = nz;
    if (!a.soft) return;
    float zc = INFINITY; int ic = 0x7fffffff;
    if (cand > K && keep) {
        // more than K candidates: the silhouette is the product over the K nearest only (PyTorch3D's per-pixel queue);
        // the cut (z, id) of the K-th nearest is recorded for the backward
        prod = 1.f;
        for (int i = 0; i < K; i++) prod *= 1.f - hp[i];
        zc = mz; ic = mf;
    } else if (cand > K) {
        // K beyond the thread-local queue: exact selection by K passes over the list, each taking the smallest key
        // above the previous one
        float lz = -INFINITY; int lf = -1;
        prod = 1.f;
        for (int k = 0; k < K; k++) {
            float sz = INFINITY; int sf = 0x7fffffff; float sp = 0.f;
            for (int i = 0; i < n; i++) {
                int3 id;
                const int f = (int)list[i];
                const FaceRec r = fetch_face(a, b, f, id);
                float pz, dist, tt; bool inside, deg; int edge;
                if (!pixel_face(r, px, py, a.blur, br, pz, inside, dist, edge, tt, deg)) continue;
                if (!zid_less(lz, lf, pz, f)) continue;                       // already taken
                if (zid_less(pz, f, sz, sf)) { sz = pz; sf = f; sp = 1.f / (1.f + __expf((inside ? -dist : dist) / kBlendSigma)); }
            }
            prod *= 1.f - sp;
            lz = sz; lf = sf;
        }
        zc = lz; ic = lf;
    }
    a.alpha[pix] = 1.f - prod;
    a.zcut[pix] = zc;
    a.idcut[pix] = ic;
}

// ------------------------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(kPix) k_mesh_raster_bwd(MeshDev a) {
    __shared__ FaceRec s_rec[kPix];
    __shared__ int s_fid[kPix];
    __shared__ int3 s_vid[kPix];
    __shared__ int s_any;
    const int b = blockIdx.z, tile = blockIdx.y * a.gx + blockIdx.x, tid = threadIdx.y * kBin + threadIdx.x;
    const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
    long long start = off[tile], end = off[tile + 1];
    if (start > a.cap) start = a.cap;
    if (end > a.cap) end = a.cap;
    const int n = (int)(end - start);
    const uint32_t *list = a.face_list + (long long)b * a.cap + start;
    const int x = blockIdx.x * kBin + threadIdx.x, y = blockIdx.y * kBin + threadIdx.y;
    const bool in_img = x < a.W && y < a.H;
    const long long pix = ((long long)b * a.H + y) * a.W + x;
    const float px = a.sx - (2.0f * x + 1.0f) / a.S, py = a.sy - (2.0f * y + 1.0f) / a.S;
    const float br = sqrtf(a.blur);

    // normal map: scatter dL/dn to the hit face's three vertex normals
    if (in_img && a.d_normal && a.d_vnormals) {
        const int f = a.pix_to_face[pix];
        if (f >= 0) {
            const int3 id = load_face(a, f);
            float *g = a.d_vnormals + (long long)b * a.V * 3;
            const float gx = a.d_normal[3 * pix], gy = a.d_normal[3 * pix + 1], gz = a.d_normal[3 * pix + 2];
            const int v[3] = {id.x, id.y, id.z};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (gx != 0.f) atomicAdd(g + 3 * v[k], gx);
                if (gy != 0.f) atomicAdd(g + 3 * v[k] + 1, gy);
                if (gz != 0.f) atomicAdd(g + 3 * v[k] + 2, gz);
            }
        }
    }
    if (!a.soft || !a.d_alpha || !a.d_verts) return;
    // soft silhouette: d alpha / d d_k = -(1 - alpha) p_k / sigma
    float coef = 0.f, zc = INFINITY; int ic = 0x7fffffff;
    if (in_img) {
        coef = -a.d_alpha[pix] * (1.f - a.alpha[pix]) / kBlendSigma;
        zc = a.zcut[pix]; ic = a.idcut[pix];
    }
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (coef != 0.f) s_any = 1;
    __syncthreads();
    if (!s_any) return;                                        // interior / empty tiles: alpha saturated, nothing flows
    float *gv = a.d_verts + (long long)b * a.V * 3;
    for (int base = 0; base < n; base += kPix) {
        __syncthreads();
        if (base + tid < n) {
            int3 id;
            const int f = (int)list[base + tid];
            s_rec[tid] = fetch_face(a, b, f, id);
            s_fid[tid] = f;
            s_vid[tid] = id;
        }
        __syncthreads();
        const int m = min(kPix, n - base);
        if (coef == 0.f) continue;
        for (int j = 0; j < m; j++) {
            float pz, dist, tt; bool inside, deg; int edge;
            const FaceRec &r = s_rec[j];
            if (!pixel_face(r, px, py, a.blur, br, pz, inside, dist, edge, tt, deg)) continue;
            if (zid_less(zc, ic, pz, s_fid[j])) continue;                   // beyond the K nearest of this pixel
            const float p = 1.f / (1.f + __expf((inside ? -dist : dist) / kBlendSigma));
            const float g_abs = (inside ? -1.f : 1.f) * coef * p;            // dL / d(unsigned squared distance)
            if (g_abs == 0.f) continue;
            // PointLineDistanceBackward on the nearest edge (v_a, v_b): grad_va = g (1 - tt) 2 (q - p), grad_vb = g tt 2 (q - p)
            const int3 id = s_vid[j];
            float ax, ay, bx, by; int va, vb;
            if (edge == 0) { ax = r.ax; ay = r.ay; bx = r.bx; by = r.by; va = id.x; vb = id.y; }
            else if (edge == 1) { ax = r.ax; ay = r.ay; bx = r.cx; by = r.cy; va = id.x; vb = id.z; }
            else { ax = r.bx; ay = r.by; bx = r.cx; by = r.cy; va = id.y; vb = id.z; }
            if (deg) {                                                       // degenerate edge: distance to v_b only
                atomicAdd(gv + 3 * vb, -2.f * (px - bx) * g_abs);
                atomicAdd(gv + 3 * vb + 1, -2.f * (py - by) * g_abs);
            } else {
                const float qx = ax + tt * (bx - ax), qy = ay + tt * (by - ay);
                const float ux = 2.f * (qx - px) * g_abs, uy = 2.f * (qy - py) * g_abs;
                if (tt < 1.f) { atomicAdd(gv + 3 * va, (1.f - tt) * ux); atomicAdd(gv + 3 * va + 1, (1.f - tt) * uy); }
                if (tt > 0.f) { atomicAdd(gv + 3 * vb, tt * ux); atomicAdd(gv + 3 * vb + 1, tt * uy); }
            }
        }
    }
}

int fill_dev(const GomMeshRasterArgs *p, MeshDev &a) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->n_verts > 0 && p->n_faces >= 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->faces_per_pixel >= 1 && p->blur_radius >= 0.f, "faces_per_pixel / blur_radius");
    GOM_REQUIRE(p->list_capacity > 0 && p->list_capacity < 0xffffffffLL, "list_capacity");
    GOM_REQUIRE(p->verts_ndc && p->faces && p->vert_normals, "null input");
    GOM_REQUIRE(p->tile_count && p->tile_offset && p->tile_cursor && p->face_list && p->status, "null binning state");
    GOM_REQUIRE(p->pix_to_face && p->normal_map, "null output");
    GOM_REQUIRE(!p->soft || (p->alpha && p->zcut && p->idcut), "soft silhouette outputs");
    a.B = p->n_frames; a.V = p->n_verts; a.F = p->n_faces; a.H = p->height; a.W = p->width;
    a.gx = (a.W + kBin - 1) / kBin; a.gy = (a.H + kBin - 1) / kBin; a.T = a.gx * a.gy;
    GOM_REQUIRE(a.gy <= 65535, "image too tall");
    a.K = p->faces_per_pixel; a.faces_int64 = p->faces_int64; a.soft = p->soft; a.cap = p->list_capacity;
    a.blur = p->soft ? p->blur_radius : 0.f;
    a.S = (float)(a.H < a.W ? a.H : a.W); a.sx = a.W / a.S; a.sy = a.H / a.S;
    a.verts = p->verts_ndc; a.faces = p->faces; a.vnormals = p->vert_normals;
    a.tile_count = p->tile_count; a.tile_offset = p->tile_offset; a.tile_cursor = p->tile_cursor;
    a.face_list = p->face_list; a.status = p->status;
    a.pix_to_face = p->pix_to_face; a.normal = p->normal_map; a.alpha = p->alpha; a.zcut = p->zcut; a.idcut = p->idcut;
    a.d_normal = p->dL_dnormal_map; a.d_alpha = p->dL_dalpha; a.d_verts = p->dL_dverts_ndc; a.d_vnormals = p->dL_dvert_normals;
    return GOM_OK;
}

}  // namespace

extern "C" int gom_mesh_raster_forward(const GomMeshRasterArgs *p, gom_stream_t stream_) {
    MeshDev a;
    if (int rc = fill_dev(p, a)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    GOM_CUDA(cudaMemsetAsync(a.tile_count, 0, sizeof(uint32_t) * (size_t)a.B * a.T, stream));
    gom_prof_begin(GOM_PROF_MESH_BIN, stream);
    if (a.F > 0) {
        dim3 grid(gom_div_up(a.F, kThreads), a.B);
        k_mesh_count<<<grid, kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
    }
    k_mesh_scan<<<a.B, 1024, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    if (a.F > 0) {
        dim3 grid(gom_div_up(a.F, kThreads), a.B);
        k_mesh_emit<<<grid, kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
    }
    gom_prof_end(GOM_PROF_MESH_BIN, stream);
    dim3 bgrid(a.gx, a.gy, a.B), bblock(kBin, kBin);
    gom_prof_begin(GOM_PROF_MESH_FWD, stream);
    k_mesh_raster_fwd<<<bgrid, bblock, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_MESH_FWD, stream);
    return GOM_OK;
}

extern "C" int gom_mesh_raster_backward(const GomMeshRasterArgs *p, gom_stream_t stream_) {
    MeshDev a;
    if (int rc = fill_dev(p, a)) return rc;
    GOM_REQUIRE(p->dL_dverts_ndc && p->dL_dvert_normals, "null gradient output");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = sizeof(float) * 3 * (size_t)a.B * a.V;
    GOM_CUDA(cudaMemsetAsync(a.d_verts, 0, n, stream));
    GOM_CUDA(cudaMemsetAsync(a.d_vnormals, 0, n, stream));
    dim3 bgrid(a.gx, a.gy, a.B), bblock(kBin, kBin);
    gom_prof_begin(GOM_PROF_MESH_BWD, stream);
    k_mesh_raster_bwd<<<bgrid, bblock, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_MESH_BWD, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_mesh_raster_args(void) { return sizeof(GomMeshRasterArgs); }

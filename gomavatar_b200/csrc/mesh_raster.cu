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
// Work distribution (round 2): the posed body covers only ~300-400 of the 4 096 tiles of a 512 x 512 frame, with lists of
// 300 (median) to 2 000 faces, and the reference trains ONE frame per step: one block of two warps per tile left 148 SMs
// waiting for the single warp of the longest tile (1.29 ms per frame).  Now
//   * k_mesh_worklist orders all B * T tiles by list length (longest first, empty tiles last); persistent blocks pull tiles
//     from an atomic counter, the empty tiles are filled with defaults by a static split;
//   * a tile is rasterised by 512 threads = 64 pixels x S = 8 SLICES of its face list (the 8 lanes of a pixel are neighbours in
//     a warp): every lane tests every 8th face of the staged chunk against its pixel, first only the blur-padded bounding box
//     (4 compares on a record prepared once per face and tile: box, reciprocal area, reciprocal edge lengths), compacting
//     the survivors into a per-lane list in shared memory, then evaluates the survivors with all lanes busy;
//   * each lane keeps the K nearest of ITS candidates; the K nearest of the pixel are selected at the end: the K-th smallest
//     depth by bisection over the bits in which the pixel's candidate depths differ (counts summed over the 8 lanes by
//     shuffles), ties in depth by face id.  With 8 slices a lane rarely sees more than K candidates, so the
//     replace-the-maximum path is rare.
#include <math.h>
#include <stdlib.h>

#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;                // face-parallel kernels (count, emit)
constexpr int kBin = 8, kBinShift = 3;       // faces are binned to 8x8-pixel tiles; one 64-thread block rasterises one tile
constexpr int kPix = kBin * kBin;
constexpr int kMaxK = 64;                  // thread-local K-nearest queue of the soft silhouette (reference: K = 50)
constexpr int kSlices = 8;                 // lanes per pixel: each takes every kSlices-th face of the tile's list
constexpr int kTileThreads = kPix * kSlices;
constexpr int kChunk = 256;                // faces staged per round (one per thread of the first 256)
constexpr int kWorkClasses = 256;          // worklist order: list length / 8, longest first
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
    uint32_t *worklist;                    // [B*T] tiles by decreasing list length, then [B*T] = non-empty tiles, [B*T+1], [B*T+2] = work counters (fwd, bwd)
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

// pixel index range [lo, hi) whose centres can lie in [vmin, vmax] (NDC): x_i = s - (2 i + 1)/S, so i in [a, b]; 0.01 pixel of
// slack covers the rounding of a, b and of the pixel centres (~1e-4 pixel at S = 4 096)
__device__ __forceinline__ void ndc_to_pixel_range(float vmin, float vmax, float s, float S, int n, int &lo, int &hi) {
    const float a = ((s - vmax) * S - 1.0f) * 0.5f, b = ((s - vmin) * S - 1.0f) * 0.5f;
    lo = max(0, (int)ceilf(a - 0.01f));
    hi = min(n, (int)floorf(b + 0.01f) + 1);
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

// Every kernel of this file must compute the per-(pixel, face) quantities with the SAME roundings: the backward replays the
// forward's K-nearest cut by comparing recomputed depths with the stored one, and re-decides `dist < blur`.  nvcc contracts
// a * b + c into FMAs per inlining context, so the shared arithmetic is written with explicit rounding intrinsics (measured: with
// plain expressions one pixel in 16 000 picked a different cut face in the two forward kernels at 120 000 faces).
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return __fmaf_rn(__fsub_rn(px, ax), __fsub_rn(by, ay), -__fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}

// reciprocal squared length of edge (a, b), or -1 for a degenerate edge (PyTorch3D PointLineDistanceForward: l2 <= kEpsilon)
__device__ __forceinline__ float inv_len2(float ax, float ay, float bx, float by) {
    const float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay), l2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
    return l2 <= kEpsArea ? -1.f : __fdiv_rn(1.f, l2);
}

// squared distance to segment (a, b) with the reciprocal squared length prepared (il < 0: degenerate -> distance to b)
__device__ __forceinline__ float seg_dist2_pre(float px, float py, float ax, float ay, float bx, float by, float il, float &tt) {
    if (il < 0.f) {
        const float ex = __fsub_rn(px, bx), ey = __fsub_rn(py, by);
        tt = 1.f;
        return __fmaf_rn(ex, ex, __fmul_rn(ey, ey));
    }
    const float dx = __fsub_rn(bx, ax), dy = __fsub_rn(by, ay);
    const float dot = __fmaf_rn(__fsub_rn(px, ax), dx, __fmul_rn(__fsub_rn(py, ay), dy));
    tt = fminf(fmaxf(__fmul_rn(dot, il), 0.f), 1.f);
    const float ex = __fsub_rn(px, __fmaf_rn(tt, dx, ax)), ey = __fsub_rn(py, __fmaf_rn(tt, dy, ay));
    return __fmaf_rn(ex, ex, __fmul_rn(ey, ey));
}

// PyTorch3D CheckPixelInsideFace after the bounding-box test, on a face with its reciprocals prepared (ia = 1 / (area + eps),
// il** = inv_len2 of the three edges).  ONE definition for every kernel of this file: the backward replays the forward's
// K-nearest cut by comparing recomputed depths with the stored one, so the arithmetic must be the same instruction for
// instruction.  Returns false when the face does not count for the pixel.
__device__ __forceinline__ bool pixel_face_core(float px, float py, float ax, float ay, float az, float bx, float by, float bz, float cx,
                                                float cy, float cz, float ia, float il01, float il02, float il12, float blur, float &pz,
                                                bool &inside, float &dist, int &edge, float &tt) {
    const float w0 = __fmul_rn(edge_fn(px, py, bx, by, cx, cy), ia);
    const float w1 = __fmul_rn(edge_fn(px, py, cx, cy, ax, ay), ia);
    const float w2 = __fmul_rn(edge_fn(px, py, ax, ay, bx, by), ia);
    pz = __fmaf_rn(w2, cz, __fmaf_rn(w1, bz, __fmul_rn(w0, az)));
    if (!(pz >= 0.f)) return false;
    pz = fabsf(pz);                                            // -0 -> +0: depth keys are compared through their bit patterns
    inside = w0 > 0.f && w1 > 0.f && w2 > 0.f;
    float t01, t02, t12;
    const float e01 = seg_dist2_pre(px, py, ax, ay, bx, by, il01, t01);
    const float e02 = seg_dist2_pre(px, py, ax, ay, cx, cy, il02, t02);
    const float e12 = seg_dist2_pre(px, py, bx, by, cx, cy, il12, t12);
    dist = e01; edge = 0; tt = t01;
    if (e02 < dist) { dist = e02; edge = 1; tt = t02; }
    if (e12 < dist) { dist = e12; edge = 2; tt = t12; }
    return inside || dist < blur;
}

// the one-block-per-tile kernel's entry: box test + reciprocals on the fly
__device__ __forceinline__ bool pixel_face(const FaceRec &r, float px, float py, float blur, float br, float &pz, bool &inside,
                                           float &dist, int &edge, float &tt, bool &degenerate) {
    const float xmin = fminf(fminf(r.ax, r.bx), r.cx) - br, xmax = fmaxf(fmaxf(r.ax, r.bx), r.cx) + br;
    const float ymin = fminf(fminf(r.ay, r.by), r.cy) - br, ymax = fmaxf(fmaxf(r.ay, r.by), r.cy) + br;
    if (px > xmax || px < xmin || py > ymax || py < ymin) return false;
    const float ia = __fdiv_rn(1.f, __fadd_rn(edge_fn(r.cx, r.cy, r.ax, r.ay, r.bx, r.by), kEpsArea));
    const float il01 = inv_len2(r.ax, r.ay, r.bx, r.by), il02 = inv_len2(r.ax, r.ay, r.cx, r.cy), il12 = inv_len2(r.bx, r.by, r.cx, r.cy);
    const bool ok = pixel_face_core(px, py, r.ax, r.ay, r.az, r.bx, r.by, r.bz, r.cx, r.cy, r.cz, ia, il01, il02, il12, blur, pz, inside, dist, edge, tt);
    degenerate = (edge == 0 ? il01 : edge == 1 ? il02 : il12) < 0.f;
    return ok;
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

// ------------------------------------------------------------------------------------------ worklist (one block)
__global__ void __launch_bounds__(1024) k_mesh_worklist(MeshDev a) {
    __shared__ uint32_t hist[kWorkClasses], cursor[kWorkClasses];
    const int tid = threadIdx.x, total = a.B * a.T;
    if (tid < kWorkClasses) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < total; i += 1024) {
        const uint32_t c = a.tile_count[i];
        atomicAdd(&hist[c ? min((uint32_t)kWorkClasses - 1, (c >> 3) + 1) : 0u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int c = kWorkClasses - 1; c >= 0; c--) { cursor[c] = run; run += hist[c]; }
        a.worklist[total] = (uint32_t)total - hist[0];
        a.worklist[total + 1] = 0u;
        a.worklist[total + 2] = 0u;
    }
    __syncthreads();
    for (int i = tid; i < total; i += 1024) {
        const uint32_t c = a.tile_count[i];
        const uint32_t pos = atomicAdd(&cursor[c ? min((uint32_t)kWorkClasses - 1, (c >> 3) + 1) : 0u], 1u);
        a.worklist[pos] = (uint32_t)i;
    }
}

// ------------------------------------------------------------------------ sliced tile kernels (forward, backward)
// Per-face record prepared once per (tile, face) by the staging thread, structure-of-arrays in shared memory (the 4 lanes of
// a pixel read 4 neighbouring faces: 4 banks, each broadcast to the 8 pixels of the warp).
enum { R_XMIN, R_XMAX, R_YMIN, R_YMAX, R_AX, R_AY, R_AZ, R_BX, R_BY, R_BZ, R_CX, R_CY, R_CZ, R_IAREA, R_IL01, R_IL02, R_IL12, R_FIELDS };

struct TileSmem {
    float rec[R_FIELDS][kChunk];
    int fid[kChunk];
    int3 vid[kChunk];                                 // backward only
    uint8_t idx[kChunk / kSlices][kTileThreads];      // per-lane compacted survivors of the box test (index / kSlices)
    int item, any;
};

__device__ __forceinline__ void stage_face(const MeshDev &a, TileSmem &sm, int b, int f, int slot, float br, bool with_ids) {
    int3 id;
    const FaceRec r = fetch_face(a, b, f, id);
    sm.rec[R_XMIN][slot] = fminf(fminf(r.ax, r.bx), r.cx) - br; sm.rec[R_XMAX][slot] = fmaxf(fmaxf(r.ax, r.bx), r.cx) + br;
    sm.rec[R_YMIN][slot] = fminf(fminf(r.ay, r.by), r.cy) - br; sm.rec[R_YMAX][slot] = fmaxf(fmaxf(r.ay, r.by), r.cy) + br;
    sm.rec[R_AX][slot] = r.ax; sm.rec[R_AY][slot] = r.ay; sm.rec[R_AZ][slot] = r.az;
    sm.rec[R_BX][slot] = r.bx; sm.rec[R_BY][slot] = r.by; sm.rec[R_BZ][slot] = r.bz;
    sm.rec[R_CX][slot] = r.cx; sm.rec[R_CY][slot] = r.cy; sm.rec[R_CZ][slot] = r.cz;
    sm.rec[R_IAREA][slot] = __fdiv_rn(1.f, __fadd_rn(edge_fn(r.cx, r.cy, r.ax, r.ay, r.bx, r.by), kEpsArea));
    sm.rec[R_IL01][slot] = inv_len2(r.ax, r.ay, r.bx, r.by);
    sm.rec[R_IL02][slot] = inv_len2(r.ax, r.ay, r.cx, r.cy);
    sm.rec[R_IL12][slot] = inv_len2(r.bx, r.by, r.cx, r.cy);
    sm.fid[slot] = f;
    if (with_ids) sm.vid[slot] = id;
}

// CheckPixelInsideFace on the staged record `j` (the box test has already passed)
__device__ __forceinline__ bool pixel_face_pre(const TileSmem &sm, int j, float px, float py, float blur, float &pz, bool &inside,
                                               float &dist, int &edge, float &tt) {
    return pixel_face_core(px, py, sm.rec[R_AX][j], sm.rec[R_AY][j], sm.rec[R_AZ][j], sm.rec[R_BX][j], sm.rec[R_BY][j], sm.rec[R_BZ][j],
                           sm.rec[R_CX][j], sm.rec[R_CY][j], sm.rec[R_CZ][j], sm.rec[R_IAREA][j], sm.rec[R_IL01][j], sm.rec[R_IL02][j],
                           sm.rec[R_IL12][j], blur, pz, inside, dist, edge, tt);
}

// box test of this lane's faces of the staged chunk (every kSlices-th, starting at `slice`); survivors -> sm.idx[.][tid]
__device__ __forceinline__ int box_survivors(TileSmem &sm, int m, int slice, int tid, float px, float py) {
    int cnt = 0;
    for (int j = slice, i = 0; j < m; j += kSlices, i++) {
        const bool out = px > sm.rec[R_XMAX][j] || px < sm.rec[R_XMIN][j] || py > sm.rec[R_YMAX][j] || py < sm.rec[R_YMIN][j];
        if (!out) sm.idx[cnt++][tid] = (uint8_t)i;
    }
    return cnt;
}

template <typename T> __device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int d = 1; d < kSlices; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ uint32_t group_min(uint32_t v) {
#pragma unroll
    for (int d = 1; d < kSlices; d <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ uint32_t group_max(uint32_t v) {
#pragma unroll
    for (int d = 1; d < kSlices; d <<= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// largest (z, id) key among slots [8 g, 8 g + 8) of a lane's queue (slots >= K do not exist); the 16 loads are independent
__device__ __forceinline__ void queue_group_max(const float *hz, const int *hf, int g, int K, float &tz, int &tf, int &ti) {
    float vz[8]; int vf[8];
    const int lo = g << 3;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const bool valid = lo + i < K;
        vz[i] = valid ? hz[lo + i] : -INFINITY;
        vf[i] = valid ? hf[lo + i] : -1;
    }
    tz = vz[0]; tf = vf[0]; ti = lo;
#pragma unroll
    for (int i = 1; i < 8; i++)
        if (zid_less(tz, tf, vz[i], vf[i])) { tz = vz[i]; tf = vf[i]; ti = lo + i; }
}
__device__ __forceinline__ void queue_max(const float (&gz)[kMaxK / 8], const int (&gf)[kMaxK / 8], const int (&gi)[kMaxK / 8], float &mz, int &mf, int &mi) {
    mz = gz[0]; mf = gf[0]; mi = gi[0];
#pragma unroll
    for (int g = 1; g < kMaxK / 8; g++)
        if (zid_less(mz, mf, gz[g], gf[g])) { mz = gz[g]; mf = gf[g]; mi = gi[g]; }
}

// thread -> pixel of the 8 x 8 tile and slice: a warp holds a small block of pixels (2 x 2 at 8 slices), the kSlices lanes of a
// pixel are neighbours
constexpr int kWarpPixX = kSlices == 4 ? 4 : 2, kWarpPixY = 32 / kSlices / kWarpPixX;
static_assert(kWarpPixX * kWarpPixY * kSlices == 32 && kBin % kWarpPixX == 0 && kBin % kWarpPixY == 0, "warp = block of pixels x slices");
__device__ __forceinline__ void tile_thread(int tid, int &tx, int &ty, int &slice) {
    const int warp = tid >> 5, lane = tid & 31, pl = lane / kSlices;
    constexpr int kWarpsX = kBin / kWarpPixX;
    slice = lane % kSlices;
    tx = (warp % kWarpsX) * kWarpPixX + (pl % kWarpPixX);
    ty = (warp / kWarpsX) * kWarpPixY + (pl / kWarpPixX);
}

__global__ void __launch_bounds__(kTileThreads, 1) k_mesh_tiles_fwd(MeshDev a) {
    __shared__ TileSmem sm;
    const int tid = threadIdx.x, total = a.B * a.T;
    const int n_work = (int)a.worklist[total];
    int tx, ty, slice;
    tile_thread(tid, tx, ty, slice);
    const float br = sqrtf(a.blur);
    const int K = a.K;
    float hz[kMaxK], hp[kMaxK]; int hf[kMaxK];                    // this lane's K nearest candidates (local memory)
    float gz[kMaxK / 8]; int gf[kMaxK / 8], gi[kMaxK / 8];       // per group of 8 slots: its largest key and where it sits

    for (;;) {
        __syncthreads();
        if (tid == 0) sm.item = (int)atomicAdd(a.worklist + total + 1, 1u);
        __syncthreads();
        const int item = sm.item;
        if (item >= n_work) break;
        const int gt = (int)a.worklist[item], b = gt / a.T, tile = gt - b * a.T;
        const int bx_ = tile % a.gx, by_ = tile / a.gx;
        const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
        long long start = off[tile], end = off[tile + 1];
        if (start > a.cap) start = a.cap;
        if (end > a.cap) end = a.cap;
        const int n = (int)(end - start);
        const uint32_t *list = a.face_list + (long long)b * a.cap + start;
        const int x = bx_ * kBin + tx, y = by_ * kBin + ty;
        const bool in_img = x < a.W && y < a.H;
        const float px = a.sx - (2.0f * x + 1.0f) / a.S, py = a.sy - (2.0f * y + 1.0f) / a.S;

        float best_z = INFINITY; int best_f = 0x7fffffff;
        float prod = 1.f; int cand = 0;
        float mz = -INFINITY; int mf = -1, mi = 0;
        for (int base = 0; base < n; base += kChunk) {
            __syncthreads();
            if (tid < kChunk && base + tid < n) stage_face(a, sm, b, (int)list[base + tid], tid, br, false);
            __syncthreads();
            if (!in_img) continue;
            const int m = min(kChunk, n - base);
            const int cnt = box_survivors(sm, m, slice, tid, px, py);
            for (int c = 0; c < cnt; c++) {
                const int j = (int)sm.idx[c][tid] * kSlices + slice;
                float pz, dist, tt; bool inside; int edge;
                if (!pixel_face_pre(sm, j, px, py, a.blur, pz, inside, dist, edge, tt)) continue;
                const int f = sm.fid[j];
                if (inside && zid_less(pz, f, best_z, best_f)) { best_z = pz; best_f = f; }
                if (!a.soft) continue;
                const float p = 1.f / (1.f + __expf((inside ? -dist : dist) / kBlendSigma));     // sigmoid(-d / sigma)
                prod *= 1.f - p;
                if (cand < K) {
                    hz[cand] = pz; hf[cand] = f; hp[cand] = p;
                } else {
                    // this lane alone has seen more than K candidates (rare with 8 slices): replace its current K-th nearest.
                    // The largest key of every group of 8 slots is kept in registers (all indexing unrolled), and the 8 slots of
                    // a group are read by independent loads, so a replacement costs one local-memory round trip, not thirty.
                    if (cand == K) {
#pragma unroll
                        for (int g = 0; g < kMaxK / 8; g++) {
                            gz[g] = -INFINITY; gf[g] = -1; gi[g] = 0;
                            if (g * 8 < K) queue_group_max(hz, hf, g, K, gz[g], gf[g], gi[g]);
                        }
                        queue_max(gz, gf, gi, mz, mf, mi);
                    }
                    if (zid_less(pz, f, mz, mf)) {
                        hz[mi] = pz; hf[mi] = f; hp[mi] = p;
                        const int g = mi >> 3;
                        float tz; int tf, ti;
                        queue_group_max(hz, hf, g, K, tz, tf, ti);
#pragma unroll
                        for (int gg = 0; gg < kMaxK / 8; gg++)
                            if (gg == g) { gz[gg] = tz; gf[gg] = tf; gi[gg] = ti; }
                        queue_max(gz, gf, gi, mz, mf, mi);
                    }
                }
                cand++;
            }
        }

        // ---------------------------------------------------------------- merge the kSlices lanes of each pixel
        // nearest inside face: smallest (z, id)
#pragma unroll
        for (int d = 1; d < kSlices; d <<= 1) {
            const float oz = __shfl_xor_sync(0xffffffffu, best_z, d);
            const int of = __shfl_xor_sync(0xffffffffu, best_f, d);
            if (zid_less(oz, of, best_z, best_f)) { best_z = oz; best_f = of; }
        }
        const int tot = group_sum(cand);
        float all_prod = prod;
#pragma unroll
        for (int d = 1; d < kSlices; d <<= 1) all_prod *= __shfl_xor_sync(0xffffffffu, all_prod, d);
        float zc = INFINITY; int ic = 0x7fffffff;
        const bool sel = a.soft && tot > K;
        if (__any_sync(0xffffffffu, sel)) {
            // K-th smallest (z, id) of the union of the lanes' queues.  Depths are non-negative floats: their bit patterns order
            // like the values.  Bisection over the bits below the common prefix of the pixel's smallest and largest depth.
            const int qn = sel ? min(cand, K) : 0;
            const int qn_max = __reduce_max_sync(0xffffffffu, qn);
            // the lane's depth keys move from local memory (a round trip to L2 per access once six blocks share an SM) into
            // registers for the ~25 counting passes; 0xffffffff = empty slot (never below a trial value, never equal to a depth)
            uint32_t zr[kMaxK];
#pragma unroll
            for (int i0 = 0; i0 < kMaxK; i0 += 8) {
                if (i0 < qn_max) {
#pragma unroll
                    for (int i = i0; i < i0 + 8; i++) zr[i] = i < qn ? __float_as_uint(hz[i]) : 0xffffffffu;
                } else {
#pragma unroll
                    for (int i = i0; i < i0 + 8; i++) zr[i] = 0xffffffffu;
                }
            }
            uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
            for (int i0 = 0; i0 < kMaxK; i0 += 8)
                if (i0 < qn_max) {
#pragma unroll
                    for (int i = i0; i < i0 + 8; i++) { lo = min(lo, zr[i]); hi = max(hi, zr[i] == 0xffffffffu ? 0u : zr[i]); }
                }
            lo = group_min(lo); hi = group_max(hi);
            const int nb = sel ? 32 - __clz((int)(lo ^ hi)) : 0;          // bits [0, nb) differ inside this pixel (nb <= 31)
            const int nb_max = __reduce_max_sync(0xffffffffu, nb);
            uint32_t Z = sel ? (hi >> nb) << nb : 0u;
            for (int bit = nb_max - 1; bit >= 0; bit--) {
                const uint32_t trial = Z | (1u << bit);
                int c = 0;
#pragma unroll
                for (int i0 = 0; i0 < kMaxK; i0 += 8)
                    if (i0 < qn_max) {
#pragma unroll
                        for (int i = i0; i < i0 + 8; i++) c += zr[i] < trial ? 1 : 0;
                    }
                c = group_sum(c);
                if (bit < nb && c < K) Z = trial;
            }
            int c_less = 0, c_eq = 0; uint32_t f_max = 0u;
#pragma unroll
            for (int i0 = 0; i0 < kMaxK; i0 += 8)
                if (i0 < qn_max) {
#pragma unroll
                    for (int i = i0; i < i0 + 8; i++) {
                        c_less += zr[i] < Z ? 1 : 0;
                        if (zr[i] == Z) { c_eq++; f_max = max(f_max, (uint32_t)hf[i]); }
                    }
                }
            c_less = group_sum(c_less); c_eq = group_sum(c_eq); f_max = group_max(f_max);
            const int need = K - c_less;                                  // how many of the faces at depth Z belong to the K nearest
            uint32_t fcut = f_max;
            const bool tie = sel && need < c_eq;
            if (__any_sync(0xffffffffu, tie)) {                           // several faces at exactly the cut depth: smallest ids first
                uint32_t Fc = 0u;
                for (int bit = 30; bit >= 0; bit--) {
                    const uint32_t trial = Fc | (1u << bit);
                    int c = 0;
                    for (int i = 0; i < qn; i++) c += (__float_as_uint(hz[i]) == Z && (uint32_t)hf[i] < trial) ? 1 : 0;
                    c = group_sum(c);
                    if (c < need) Fc = trial;
                }
                if (tie) fcut = Fc;
            }
            float pr = 1.f;
#pragma unroll
            for (int i0 = 0; i0 < kMaxK; i0 += 8)
                if (i0 < qn_max) {
#pragma unroll
                    for (int i = i0; i < i0 + 8; i++) {
                        const bool in = zr[i] < Z || (zr[i] == Z && (uint32_t)hf[i] <= fcut);
                        const float q = 1.f - hp[i < qn ? i : 0];          // independent loads: their latencies overlap
                        if (in) pr *= q;
                    }
                }
#pragma unroll
            for (int d = 1; d < kSlices; d <<= 1) pr *= __shfl_xor_sync(0xffffffffu, pr, d);
            if (sel) { all_prod = pr; zc = __uint_as_float(Z); ic = (int)fcut; }
        }
        if (!in_img) continue;
        const long long pix = ((long long)b * a.H + y) * a.W + x;
        const bool hit = best_f != 0x7fffffff;
        if (slice < 3) {                                                  // slices 0..2: one component of the normal each
            float nv = 0.f;
            if (hit) {
                const int3 id = load_face(a, best_f);
                const float *vn = a.vnormals + (long long)b * a.V * 3 + slice;
                nv = vn[3 * id.x] + vn[3 * id.y] + vn[3 * id.z];
            }
            a.normal[3 * pix + slice] = nv;
        }
        if (slice == 3) a.pix_to_face[pix] = hit ? best_f : -1;
        if (a.soft) {
            if (slice == 0) a.alpha[pix] = 1.f - all_prod;
            if (slice == 1) a.zcut[pix] = zc;
            if (slice == 2) a.idcut[pix] = ic;
        }
    }

    // ------------------------------------------------------------------------ empty tiles: defaults, static split
    for (int e = n_work + blockIdx.x * (kTileThreads / kPix) + (tid >> 6); e < total; e += gridDim.x * (kTileThreads / kPix)) {
        const int gt = (int)a.worklist[e], b = gt / a.T, tile = gt - b * a.T;
        const int x = (tile % a.gx) * kBin + (tid & 7), y = (tile / a.gx) * kBin + ((tid >> 3) & 7);
        if (x >= a.W || y >= a.H) continue;
        const long long pix = ((long long)b * a.H + y) * a.W + x;
        a.pix_to_face[pix] = -1;
        a.normal[3 * pix] = 0.f; a.normal[3 * pix + 1] = 0.f; a.normal[3 * pix + 2] = 0.f;
        if (a.soft) { a.alpha[pix] = 0.f; a.zcut[pix] = INFINITY; a.idcut[pix] = 0x7fffffff; }
    }
}

__global__ void __launch_bounds__(kTileThreads) k_mesh_tiles_bwd(MeshDev a) {
    __shared__ TileSmem sm;
    const int tid = threadIdx.x, total = a.B * a.T;
    const int n_work = (int)a.worklist[total];
    int tx, ty, slice;
    tile_thread(tid, tx, ty, slice);
    const float br = sqrtf(a.blur);
    const bool do_normal = a.d_normal && a.d_vnormals, do_soft = a.soft && a.d_alpha && a.d_verts;
    for (;;) {
        __syncthreads();
        if (tid == 0) { sm.item = (int)atomicAdd(a.worklist + total + 2, 1u); sm.any = 0; }
        __syncthreads();
        const int item = sm.item;
        if (item >= n_work) break;
        const int gt = (int)a.worklist[item], b = gt / a.T, tile = gt - b * a.T;
        const int x = (tile % a.gx) * kBin + tx, y = (tile / a.gx) * kBin + ty;
        const bool in_img = x < a.W && y < a.H;
        const long long pix = ((long long)b * a.H + y) * a.W + x;
        const float px = a.sx - (2.0f * x + 1.0f) / a.S, py = a.sy - (2.0f * y + 1.0f) / a.S;

        // normal map: scatter dL/dn to the hit face's three vertex normals (slice k < 3 takes component k)
        if (in_img && do_normal && slice < 3) {
            const int f = a.pix_to_face[pix];
            const float g = f >= 0 ? a.d_normal[3 * pix + slice] : 0.f;
            if (g != 0.f) {
                const int3 id = load_face(a, f);
                float *gn = a.d_vnormals + (long long)b * a.V * 3 + slice;
                atomicAdd(gn + 3 * id.x, g); atomicAdd(gn + 3 * id.y, g); atomicAdd(gn + 3 * id.z, g);
            }
        }
        if (!do_soft) continue;
        // soft silhouette: d alpha / d d_k = -(1 - alpha) p_k / sigma
        float coef = 0.f, zc = INFINITY; int ic = 0x7fffffff;
        if (in_img) {
            coef = -a.d_alpha[pix] * (1.f - a.alpha[pix]) / kBlendSigma;
            zc = a.zcut[pix]; ic = a.idcut[pix];
        }
        if (coef != 0.f) sm.any = 1;
        __syncthreads();
        if (!sm.any) continue;                                     // interior tiles: alpha saturated, nothing flows
        const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
        long long start = off[tile], end = off[tile + 1];
        if (start > a.cap) start = a.cap;
        if (end > a.cap) end = a.cap;
        const int n = (int)(end - start);
        const uint32_t *list = a.face_list + (long long)b * a.cap + start;
        float *gv = a.d_verts + (long long)b * a.V * 3;
        for (int base = 0; base < n; base += kChunk) {
            __syncthreads();
            if (tid < kChunk && base + tid < n) stage_face(a, sm, b, (int)list[base + tid], tid, br, true);
            __syncthreads();
            if (coef == 0.f) continue;
            const int m = min(kChunk, n - base);
            const int cnt = box_survivors(sm, m, slice, tid, px, py);
            for (int c = 0; c < cnt; c++) {
                const int j = (int)sm.idx[c][tid] * kSlices + slice;
                float pz, dist, tt; bool inside; int edge;
                if (!pixel_face_pre(sm, j, px, py, a.blur, pz, inside, dist, edge, tt)) continue;
                if (zid_less(zc, ic, pz, sm.fid[j])) continue;                  // beyond the K nearest of this pixel
                const float p = 1.f / (1.f + __expf((inside ? -dist : dist) / kBlendSigma));
                const float g_abs = (inside ? -1.f : 1.f) * coef * p;            // dL / d(unsigned squared distance)
                if (g_abs == 0.f) continue;
                // PointLineDistanceBackward on the nearest edge (v_a, v_b): grad_va = g (1 - tt) 2 (q - p), grad_vb = g tt 2 (q - p)
                const int3 id = sm.vid[j];
                float ax, ay, bx, by, il; int va, vb;
                if (edge == 0) { ax = sm.rec[R_AX][j]; ay = sm.rec[R_AY][j]; bx = sm.rec[R_BX][j]; by = sm.rec[R_BY][j]; il = sm.rec[R_IL01][j]; va = id.x; vb = id.y; }
                else if (edge == 1) { ax = sm.rec[R_AX][j]; ay = sm.rec[R_AY][j]; bx = sm.rec[R_CX][j]; by = sm.rec[R_CY][j]; il = sm.rec[R_IL02][j]; va = id.x; vb = id.z; }
                else { ax = sm.rec[R_BX][j]; ay = sm.rec[R_BY][j]; bx = sm.rec[R_CX][j]; by = sm.rec[R_CY][j]; il = sm.rec[R_IL12][j]; va = id.y; vb = id.z; }
                if (il < 0.f) {                                                  // degenerate edge: distance to v_b only
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
}

// persistent blocks: as many as are resident at once, never more than there are tiles
template <typename Kern> int tile_grid(const MeshDev &a, Kern kern) {
    int dev = 0, sms = 148, per_sm = 2;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    const long long want = (long long)per_sm * sms, total = (long long)a.B * a.T;
    return (int)(want < total ? want : total);
}

int fill_dev(const GomMeshRasterArgs *p, MeshDev &a) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_frames <= 65535 && p->n_verts > 0 && p->n_faces >= 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->faces_per_pixel >= 1 && p->blur_radius >= 0.f, "faces_per_pixel / blur_radius");
    GOM_REQUIRE(p->list_capacity > 0 && p->list_capacity < 0xffffffffLL, "list_capacity");
    GOM_REQUIRE(p->verts_ndc && p->faces && p->vert_normals, "null input");
    GOM_REQUIRE(p->tile_count && p->tile_offset && p->tile_cursor && p->face_list && p->status && p->worklist, "null binning state");
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
    a.face_list = p->face_list; a.status = p->status; a.worklist = p->worklist;
    GOM_REQUIRE((long long)a.B * a.T < (1ll << 30), "too many tiles");
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
    k_mesh_worklist<<<1, 1024, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_MESH_BIN, stream);
    gom_prof_begin(GOM_PROF_MESH_FWD, stream);
    const char *legacy = getenv("GOM_MESH_LEGACY");           // tests: A/B against the one-block-per-tile kernel
    if ((a.soft && a.K > kMaxK) || (legacy && legacy[0] == '1')) {     // K beyond the per-lane queue: that kernel's exact slow path
        dim3 bgrid(a.gx, a.gy, a.B), bblock(kBin, kBin);
        k_mesh_raster_fwd<<<bgrid, bblock, 0, stream>>>(a);
    } else {
        k_mesh_tiles_fwd<<<tile_grid(a, k_mesh_tiles_fwd), kTileThreads, 0, stream>>>(a);
    }
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
    GOM_CUDA(cudaMemsetAsync(a.worklist + (size_t)a.B * a.T + 2, 0, sizeof(uint32_t), stream));      // work counter of this launch
    gom_prof_begin(GOM_PROF_MESH_BWD, stream);
    k_mesh_tiles_bwd<<<tile_grid(a, k_mesh_tiles_bwd), kTileThreads, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_MESH_BWD, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_mesh_raster_args(void) { return sizeof(GomMeshRasterArgs); }

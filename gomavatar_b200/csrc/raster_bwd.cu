// raster_bwd.cu — tile-based 3D-Gaussian splat rasterizer, backward (sm_100a).
//
// Replaces `_C.rasterize_gaussians_backward` of the third-party diff_gaussian_rasterization (renderCUDA backward,
// computeCov2DCUDA, preprocessCUDA backward).  Upstream issues ~10 global fp32 atomicAdds per
// (pixel, Gaussian) hit; here each warp first reduces its 32 pixels with a transposing shuffle butterfly (skipped
// entirely when no lane of the warp is hit) and one global RED per (8x4 sub-block, Gaussian, component) remains.
#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;      // per-Gaussian kernels

struct BwdDev {
    int B, P, H, W, C, interleaved, gx, gy, T;
    long long cap;
    const float *means3D; long long means3D_stride;
    const float *cov3D; long long cov3D_stride;
    const float *colors; long long colors_stride;
    const float *view, *proj, *tanfov, *bg;
    const float *final_T; const uint32_t *n_contrib; const int32_t *radii;
    const float2 *xy; const float4 *conic_opacity; const uint32_t *tile_offset, *point_list, *worklist; const uint8_t *point_mask;
    const float *dL_dout;
    float *dL_dmeans3D, *dL_dcov3D, *dL_dcolors; long long dL_dcolors_stride;
    float *dL_dopacity; float *dL_dmean2D; float *dL_dconic;
};

// ------------------------------------------------------------------------------------------ App. A.6 blend backward
// Every warp is independent (no block barrier, no shared-memory atomics): it owns one 8x4-pixel sub-block of one tile,
// walks that tile's sorted list back to front from ITS OWN last contributor, fetches 32 entries per chunk (one per
// lane, next chunk prefetched), culls them against the sub-block (entry_reaches_rect) and visits the survivors through
// the ballot mask.  The 8 per-Gaussian gradient components (mean2D.xy, conic A B C, colour rgb) of the 32 pixels are
// summed with a transposing butterfly — 4+2+1+1+1 = 9 shuffles instead of 8 x 5 — that leaves component i's total on
// lane group i, and 8 lanes issue ONE fire-and-forget RED.ADD.F32 each.
constexpr int kBwdThreads = 128;
constexpr int kBwdWarps = kBwdThreads / 32;

// sum each of v[0..7] over the 32 lanes; on return lanes with (lane & 3) == 0 hold the total of component
// ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).
__device__ __forceinline__ float butterfly8(const float (&v)[8], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float w[4], x[2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float keep = h16 ? v[i + 4] : v[i], send = h16 ? v[i] : v[i + 4];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float keep = h8 ? w[i + 2] : w[i], send = h8 ? w[i] : w[i + 2];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float y = (h4 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, h4 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;
}

// C: rendered channels; CG: leading colour channels that need a gradient (GoMAvatar's 4th channel is the constant 1 that
// renders alpha, gaussian.py:49); OPAC: dL/dopacity wanted (GoMAvatar: opacity == 1 without gradient, model.py:242).
template <int C, int CG, bool OPAC>
__global__ void __launch_bounds__(kBwdThreads) k_blend_bwd(BwdDev a) {
    constexpr int NG = 5 + CG + (OPAC ? 1 : 0);            // mean2D.xy, conic A B C, colour[CG], (opacity)
    __shared__ float2 s_xy[kBwdWarps][32];
    __shared__ float4 s_co[kBwdWarps][32];
    __shared__ __align__(16) float s_col[kBwdWarps][32 * C];
    __shared__ uint32_t s_id[kBwdWarps][32];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kBwdWarps + wib;      // global warp = (frame, tile, sub-block)
    const int sub = (int)(gw & 7);
    const long long wi = gw >> 3;
    if (wi >= (long long)a.B * a.T) return;
    const long long ft = a.worklist ? (long long)a.worklist[wi] : wi;      // longest lists first (the forward's order)
    const int tile = (int)(ft % a.T), b = (int)(ft / a.T);
    const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
    long long start = off[tile], end = off[tile + 1];
    if (start > a.cap) start = a.cap;
    if (end > a.cap) end = a.cap;
    const int n = (int)(end - start);
    if (n == 0) return;

    const int tx = tile % a.gx, ty = tile / a.gx;
    const int x0 = tx * 16 + (sub & 1) * 8, y0 = ty * 16 + (sub >> 1) * 4;
    const int x = x0 + (lane & 7), y = y0 + (lane >> 3);
    const bool inside = x < a.W && y < a.H;
    const long long pix = ((long long)b * a.H + y) * a.W + x;
    const float T_final = inside ? a.final_T[pix] : 0.f;
    const uint32_t last_contributor = inside ? a.n_contrib[pix] : 0u;
    const int n_eff = min(n, (int)__reduce_max_sync(0xffffffffu, last_contributor));
    if (n_eff == 0) return;                    // entries behind every pixel's last contributor are never visited

    float dpix[C], accum_rec[C], last_color[C];
    float bg_dot = 0.f;
    const float *bg = a.bg + b * C;
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
        float v = 0.f;
        if (inside) {
            v = a.interleaved ? a.dL_dout[pix * C + ch]
                              : a.dL_dout[((long long)b * C + ch) * a.H * a.W + (long long)y * a.W + x];
        }
        dpix[ch] = v;
        accum_rec[ch] = 0.f;
        last_color[ch] = 0.f;
        bg_dot += bg[ch] * v;
    }
    float T = T_final, last_alpha = 0.f;
    const float pxf = (float)x, pyf = (float)y;
    const float rcx = (float)x0 + 3.5f, rcy = (float)y0 + 1.5f;
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

    const uint32_t *plist = a.point_list + (long long)b * a.cap + start;
    const float2 *gxy = a.xy + (long long)b * a.P;
    const float4 *gco = a.conic_opacity + (long long)b * a.P;
    const float *gcol = a.colors + b * a.colors_stride;
    float *o_mean2D = a.dL_dmean2D + (long long)b * a.P * 2;
    float *o_conic = a.dL_dconic + (long long)b * a.P * 3;
    float *o_opac = a.dL_dopacity ? a.dL_dopacity + (long long)b * a.P : nullptr;
    float *o_col = a.dL_dcolors + b * a.dL_dcolors_stride;

    // destination of the component this lane ends up holding after butterfly8 (NG <= 8 path)
    const int comp = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    float *red_base = nullptr;
    int red_stride = 0;
    if ((lane & 3) == 0 && comp < NG) {
        if (comp < 2) { red_base = o_mean2D + comp; red_stride = 2; }
        else if (comp < 5) { red_base = o_conic + (comp - 2); red_stride = 3; }
        else if (comp < 5 + CG) { red_base = o_col + (comp - 5); red_stride = C; }
        else { red_base = o_opac; red_stride = 1; }
    }

    const uint8_t *pmask = a.point_mask ? a.point_mask + (long long)b * a.cap + start : nullptr;
    // ids (+ sub-block masks) two chunks ahead, records of the reaching entries one chunk ahead (see k_blend)
    uint32_t id_c = 0, id_n = 0; bool rel_c = false, rel_n = false;
    float2 xy_c = make_float2(0.f, 0.f); float4 co_c = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_id = [&](int idx, uint32_t &id, bool &rel) {
        rel = false;
        if (idx >= 0) { id = plist[idx]; rel = pmask ? ((pmask[idx] >> sub) & 1u) != 0u : true; }
    };
    fetch_id(n_eff - 1 - lane, id_c, rel_c);
    fetch_id(n_eff - 33 - lane, id_n, rel_n);
    if (rel_c) { xy_c = __ldg(gxy + id_c); co_c = __ldg(gco + id_c); }
    for (int hi = n_eff; hi > 0; hi -= 32) {               // lane j of a chunk holds list entry hi-1-j (back to front)
        const bool rel = rel_c && (pmask != nullptr || entry_reaches_rect(xy_c, co_c, rcx, rcy, 3.5f, 1.5f));
        unsigned mask = __ballot_sync(0xffffffffu, rel);
        if (rel) {
            s_id[wib][lane] = id_c;
            s_xy[wib][lane] = xy_c;
            s_co[wib][lane] = co_c;
            if constexpr (C == 4) {
                reinterpret_cast<float4 *>(s_col[wib])[lane] = __ldg(reinterpret_cast<const float4 *>(gcol) + id_c);
            } else {
#pragma unroll
                for (int ch = 0; ch < C; ch++) s_col[wib][lane * C + ch] = __ldg(gcol + (long long)id_c * C + ch);
            }
        }
        id_c = id_n; rel_c = rel_n;
        if (rel_c) { xy_c = __ldg(gxy + id_c); co_c = __ldg(gco + id_c); }
        fetch_id(hi - 65 - lane, id_n, rel_n);
        __syncwarp();
        // Two survivors per round: their Gaussian values (shared-memory reads, exp) and, afterwards, their two butterfly
        // reductions are independent and overlap; only the T / accumulated-colour recurrence between them is sequential.
        constexpr int kU = 2;
        constexpr int kNG = NG > 8 ? NG : 8;
        while (mask) {
            int j[kU];
            float g[kU][kNG];
            bool valid[kU];
            float G[kU], alpha[kU], dx[kU], dy[kU];
            float4 co[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const bool have = mask != 0u;
                j[u] = have ? __ffs(mask) - 1 : 0;
                mask &= mask - 1;
                const uint32_t k = (uint32_t)(hi - 1 - j[u]);
                const float2 c = s_xy[wib][j[u]];
                co[u] = s_co[wib][j[u]];
                dx[u] = c.x - pxf; dy[u] = c.y - pyf;
                const float power = -0.5f * (co[u].x * dx[u] * dx[u] + co[u].z * dy[u] * dy[u]) - co[u].y * dx[u] * dy[u];
                G[u] = __expf(power);
                alpha[u] = fminf(0.99f, co[u].w * G[u]);
                valid[u] = have && k < last_contributor && (power <= 0.0f) && (alpha[u] >= 1.0f / 255.0f);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
#pragma unroll
                for (int c = 0; c < kNG; c++) g[u][c] = 0.f;
                if (valid[u]) {
                    const float inv1ma = __fdividef(1.f, 1.f - alpha[u]);
                    T = T * inv1ma;
                    const float dchannel_dcolor = alpha[u] * T;
                    float dL_dalpha = 0.f;
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        const float col = s_col[wib][j[u] * C + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = col;
                        dL_dalpha += (col - accum_rec[ch]) * dpix[ch];
                        if (ch < CG) g[u][5 + ch] = dchannel_dcolor * dpix[ch];
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha[u];
                    dL_dalpha += (-T_final * inv1ma) * bg_dot;
                    const float dL_dG = co[u].w * dL_dalpha;          // the 0.99 clamp is ignored, as upstream
                    const float gdx = G[u] * dx[u], gdy = G[u] * dy[u];
                    const float dG_ddelx = -gdx * co[u].x - gdy * co[u].y;
                    const float dG_ddely = -gdy * co[u].z - gdx * co[u].y;
                    g[u][0] = dL_dG * dG_ddelx * ddelx_dx;
                    g[u][1] = dL_dG * dG_ddely * ddely_dy;
                    g[u][2] = -0.5f * gdx * dx[u] * dL_dG;
                    g[u][3] = -0.5f * gdx * dy[u] * dL_dG;
                    g[u][4] = -0.5f * gdy * dy[u] * dL_dG;
                    if (OPAC) g[u][5 + CG] = G[u] * dL_dalpha;
                }
            }
            bool any[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) any[u] = __any_sync(0xffffffffu, valid[u]);
            if constexpr (NG <= 8) {
                float tot[kU];
#pragma unroll
                for (int u = 0; u < kU; u++) tot[u] = any[u] ? butterfly8(g[u], lane) : 0.f;
#pragma unroll
                for (int u = 0; u < kU; u++)
                    if (any[u] && red_base && tot[u] != 0.f) atomicAdd(red_base + (long long)s_id[wib][j[u]] * red_stride, tot[u]);
            } else {
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    if (!any[u]) continue;
                    const uint32_t id = s_id[wib][j[u]];
#pragma unroll
                    for (int c = 0; c < NG; c++) {
                        const float tot = warp_sum(g[u][c]);
                        if (lane == 0 && tot != 0.f) {
                            float *dst = c < 2 ? o_mean2D + 2LL * id + c
                                       : c < 5 ? o_conic + 3LL * id + (c - 2)
                                       : c < 5 + CG ? o_col + (long long)id * C + (c - 5) : o_opac + id;
                            atomicAdd(dst, tot);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------- App. A.7 preprocess backward
__global__ void __launch_bounds__(kThreads) k_preprocess_bwd(BwdDev a) {
    __shared__ float cam[34];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) cam[threadIdx.x] = a.view[b * 16 + threadIdx.x];
    else if (threadIdx.x < 32) cam[threadIdx.x] = a.proj[b * 16 + threadIdx.x - 16];
    else if (threadIdx.x < 34) cam[threadIdx.x] = a.tanfov[b * 2 + threadIdx.x - 32];
    __syncthreads();
    const int g = blockIdx.x * kThreads + threadIdx.x;
    if (g >= a.P) return;
    const float *view = cam, *proj = cam + 16;
    const float tanfovx = cam[32], tanfovy = cam[33];
    const long long o = (long long)b * a.P + g;
    float dmean[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.radii[o] > 0) {
        const float *mp = a.means3D + b * a.means3D_stride + 3LL * g;
        const float p[3] = {mp[0], mp[1], mp[2]};
        const float2 *cp = reinterpret_cast<const float2 *>(a.cov3D + b * a.cov3D_stride + 6LL * g);
        const float2 c01 = cp[0], c23 = cp[1], c45 = cp[2];
        const float s[6] = {c01.x, c01.y, c23.x, c23.y, c45.x, c45.y};
        const float fx = xdiv((float)a.W, xmul(2.0f, tanfovx)), fy = xdiv((float)a.H, xmul(2.0f, tanfovy));
        Cov2D q;
        cov2d_exact(p, s, view, fx, fy, tanfovx, tanfovy, q);
        const float ca = q.a, cb = q.b, cc = q.c;
        const float gA = a.dL_dconic[3 * o], gB = a.dL_dconic[3 * o + 1], gC = a.dL_dconic[3 * o + 2];
        const float den = ca * cc - cb * cb;
        const float k2 = 1.0f / ((den * den) + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (k2 != 0.f) {
            dL_da = k2 * (-cc * cc * gA + 2 * cb * cc * gB + (den - ca * cc) * gC);
            dL_dc = k2 * (-ca * ca * gC + 2 * ca * cb * gB + (den - ca * cc) * gA);
            dL_db = k2 * 2 * (cb * cc * gA - (den + 2 * cb * cb) * gB + ca * cb * gC);
            const float *M0 = q.M0, *M1 = q.M1;
            dcov[0] = M0[0] * M0[0] * dL_da + M0[0] * M1[0] * dL_db + M1[0] * M1[0] * dL_dc;
            dcov[3] = M0[1] * M0[1] * dL_da + M0[1] * M1[1] * dL_db + M1[1] * M1[1] * dL_dc;
            dcov[5] = M0[2] * M0[2] * dL_da + M0[2] * M1[2] * dL_db + M1[2] * M1[2] * dL_dc;
            dcov[1] = 2 * M0[0] * M0[1] * dL_da + (M0[0] * M1[1] + M0[1] * M1[0]) * dL_db + 2 * M1[0] * M1[1] * dL_dc;
            dcov[2] = 2 * M0[0] * M0[2] * dL_da + (M0[0] * M1[2] + M0[2] * M1[0]) * dL_db + 2 * M1[0] * M1[2] * dL_dc;
            dcov[4] = 2 * M0[2] * M0[1] * dL_da + (M0[1] * M1[2] + M0[2] * M1[1]) * dL_db + 2 * M1[1] * M1[2] * dL_dc;
        }
        // (iii) cov2D -> mean through M = J(t)·R
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float dM0 = 2 * q.v0[k] * dL_da + q.v1[k] * dL_db;
            const float dM1 = 2 * q.v1[k] * dL_dc + q.v0[k] * dL_db;
            dJ00 += view[4 * k + 0] * dM0;
            dJ02 += view[4 * k + 2] * dM0;
            dJ11 += view[4 * k + 1] * dM1;
            dJ12 += view[4 * k + 2] * dM1;
        }
        const float tz = 1.f / q.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = q.xm * -fx * tz2 * dJ02;
        const float dty = q.ym * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * q.t[0]) * tz3 * dJ02 + (2 * fy * q.t[1]) * tz3 * dJ12;
#pragma unroll
        for (int k = 0; k < 3; k++) dmean[k] = view[4 * k + 0] * dtx + view[4 * k + 1] * dty + view[4 * k + 2] * dtz;
        // (iv) mean2D -> mean through the projection
        const float mhx = proj[0] * p[0] + proj[4] * p[1] + proj[8] * p[2] + proj[12];
        const float mhy = proj[1] * p[0] + proj[5] * p[1] + proj[9] * p[2] + proj[13];
        const float mhw = proj[3] * p[0] + proj[7] * p[1] + proj[11] * p[2] + proj[15];
        const float mw = 1.0f / (mhw + 0.0000001f);
        const float mul1 = mhx * mw * mw, mul2 = mhy * mw * mw;
        const float g2x = a.dL_dmean2D[2 * o], g2y = a.dL_dmean2D[2 * o + 1];
#pragma unroll
        for (int k = 0; k < 3; k++)
            dmean[k] += (proj[4 * k + 0] * mw - proj[4 * k + 3] * mul1) * g2x + (proj[4 * k + 1] * mw - proj[4 * k + 3] * mul2) * g2y;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) a.dL_dmeans3D[3 * o + k] = dmean[k];
#pragma unroll
    for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * o + k] = dcov[k];
}

}  // namespace

extern "C" int gom_raster_backward(const GomRasterBwdArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_gauss >= 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->n_channels == 3 || p->n_channels == 4, "n_channels must be 3 or 4");
    GOM_REQUIRE(p->n_frames <= 65535 && p->inst_capacity > 0, "n_frames / inst_capacity");
    GOM_REQUIRE(p->means3D && p->cov3D && p->colors && p->viewmatrix && p->projmatrix && p->tanfov && p->bg, "null input");
    GOM_REQUIRE(p->final_T && p->n_contrib && p->radii && p->xy && p->conic_opacity && p->tile_offset && p->point_list &&
                    p->dL_dout, "null saved state");
    GOM_REQUIRE(p->dL_dmeans3D && p->dL_dcov3D && p->dL_dcolors && p->dL_dmeans2D && p->dL_dconic, "null output");
    GOM_REQUIRE(((uintptr_t)p->cov3D % 8) == 0 && (p->cov3D_stride % 2) == 0, "cov3D must be 8-byte aligned");
    GOM_REQUIRE(p->n_channels != 4 || (((uintptr_t)p->colors % 16) == 0 && (p->colors_stride % 4) == 0), "4-channel colors must be 16-byte aligned");
    GOM_REQUIRE(p->color_grad_channels == 0 || p->color_grad_channels == p->n_channels ||
                    (p->n_channels == 4 && p->color_grad_channels == 3), "color_grad_channels must be 0, n_channels, or 3 of 4");
    GOM_REQUIRE((long long)p->n_frames * (((p->width + 15) / 16) * ((p->height + 15) / 16)) * 8 / kBwdWarps < 0x7fffffffLL, "grid too large");
    cudaStream_t stream = (cudaStream_t)stream_;
    BwdDev a;
    a.B = p->n_frames; a.P = p->n_gauss; a.H = p->height; a.W = p->width; a.C = p->n_channels;
    a.interleaved = p->interleaved;
    a.gx = (a.W + GOM_TILE - 1) / GOM_TILE; a.gy = (a.H + GOM_TILE - 1) / GOM_TILE; a.T = a.gx * a.gy;
    a.cap = p->inst_capacity;
    a.means3D = p->means3D; a.means3D_stride = p->means3D_stride;
    a.cov3D = p->cov3D; a.cov3D_stride = p->cov3D_stride;
    a.colors = p->colors; a.colors_stride = p->colors_stride;
    a.view = p->viewmatrix; a.proj = p->projmatrix; a.tanfov = p->tanfov; a.bg = p->bg;
    a.final_T = p->final_T; a.n_contrib = p->n_contrib; a.radii = p->radii;
    a.xy = reinterpret_cast<const float2 *>(p->xy);
    a.conic_opacity = reinterpret_cast<const float4 *>(p->conic_opacity);
    a.tile_offset = p->tile_offset; a.point_list = p->point_list; a.worklist = p->worklist; a.point_mask = p->point_mask; a.dL_dout = p->dL_dout;
    a.dL_dmeans3D = p->dL_dmeans3D; a.dL_dcov3D = p->dL_dcov3D;
    a.dL_dcolors = p->dL_dcolors; a.dL_dcolors_stride = p->dL_dcolors_stride;
    a.dL_dopacity = p->dL_dopacity; a.dL_dmean2D = p->dL_dmeans2D; a.dL_dconic = p->dL_dconic;

    const size_t BP = (size_t)a.B * a.P;
    const size_t ncol = (p->dL_dcolors_stride == 0 ? (size_t)a.P : BP) * a.C;
    if (a.P > 0) {
        GOM_CUDA(cudaMemsetAsync(a.dL_dmean2D, 0, sizeof(float) * 2 * BP, stream));
        GOM_CUDA(cudaMemsetAsync(a.dL_dconic, 0, sizeof(float) * 3 * BP, stream));
        GOM_CUDA(cudaMemsetAsync(a.dL_dcolors, 0, sizeof(float) * ncol, stream));
        if (a.dL_dopacity) GOM_CUDA(cudaMemsetAsync(a.dL_dopacity, 0, sizeof(float) * BP, stream));
        const long long n_warps = (long long)a.B * a.T * 8;
        const unsigned bgrid = (unsigned)((n_warps + kBwdWarps - 1) / kBwdWarps);
        const int cg = p->color_grad_channels > 0 ? p->color_grad_channels : a.C;
        const bool opac = a.dL_dopacity != nullptr;
        gom_prof_begin(GOM_PROF_BLEND_BWD, stream);
        if (a.C == 3 && !opac) k_blend_bwd<3, 3, false><<<bgrid, kBwdThreads, 0, stream>>>(a);
        else if (a.C == 3) k_blend_bwd<3, 3, true><<<bgrid, kBwdThreads, 0, stream>>>(a);
        else if (cg == 3 && !opac) k_blend_bwd<4, 3, false><<<bgrid, kBwdThreads, 0, stream>>>(a);
        else if (!opac) k_blend_bwd<4, 4, false><<<bgrid, kBwdThreads, 0, stream>>>(a);
        else k_blend_bwd<4, 4, true><<<bgrid, kBwdThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
        gom_prof_end(GOM_PROF_BLEND_BWD, stream);
        dim3 grid(gom_div_up(a.P, kThreads), a.B);
        gom_prof_begin(GOM_PROF_PREPROCESS_BWD, stream);
        k_preprocess_bwd<<<grid, kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
        gom_prof_end(GOM_PROF_PREPROCESS_BWD, stream);
    }
    return GOM_OK;
}

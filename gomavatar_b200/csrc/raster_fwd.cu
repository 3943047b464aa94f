// raster_fwd.cu — tile-based 3D-Gaussian splat rasterizer, forward (sm_100a).
//
// Replaces the forward half of the third-party `diff_gaussian_rasterization._C` that the reference calls at
// models/modules/renderer/gaussian.py:83-91 (algorithm: SURVEY.md App. A.3-A.5).  B200-first design:
//   * a batch of B frames per launch (grid.z / grid.y = frame) so that the ~150 non-empty tiles of one 512^2 frame
//     do not leave most of the 148 SMs idle;
//   * no host synchronisation: upstream's D2H read of num_rendered is replaced by fixed-capacity instance buffers
//     and a device status flag, which also makes the whole forward CUDA-graph capturable;
//   * binning is count -> scan -> scatter into per-tile segments, and the depth sort is a per-tile COUNTING sort in shared
//     memory (k_tile_sort: depth bits bucketed over the tile's own depth range, exact (depth, id) order restored inside the
//     few multi-entry buckets) instead of a global 64-bit radix sort (6+ passes over HBM);
//   * tiles are processed longest list first (k_worklist orders all B*T tiles by list length on the device), so the launch
//     tail is made of the short lists, and the blend runs as independent warps (8x4-pixel sub-blocks) in small blocks;
//   * one fused pass renders 3 (reference layout) or 4 (RGB + alpha) channels.
#include "gom_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSortCap = 4096;      // entries of one counting-sort pass (32 KB of keys + 16 KB of bucket counters in shared memory);
                                    // longer lists are sorted in several passes over consecutive bucket ranges
constexpr int kBuckets = 4096;      // depth buckets of the counting sort
constexpr int kLogBuckets = 12;
constexpr int kMaxRun = 48;         // longest same-bucket run a single thread orders by insertion
constexpr int kBlendThreads = 128;  // 4 independent warps per blend block

struct FwdDev {
    int B, P, H, W, C, interleaved, gx, gy, T;
    long long cap;
    const float *means3D; long long means3D_stride;
    const float *cov3D; long long cov3D_stride;
    const float *colors; long long colors_stride;
    const float *opac; long long opac_stride;
    const float *view, *proj, *tanfov, *bg;
    float *out_color, *final_T; uint32_t *n_contrib; int32_t *radii;
    float *depth; float2 *xy; float4 *conic_opacity; int4 *rect;
    uint32_t *tile_count, *tile_offset, *tile_cursor; unsigned long long *inst_keys; uint32_t *point_list;
    uint32_t *status;
    uint32_t *worklist;              // [B*T] (frame * T + tile), longest list first; nullptr: identity order
    uint8_t *point_mask;             // [B,cap] per list entry: bit s = the entry can reach 8x4 sub-block s of its tile; nullptr: test in the blend
};

// ------------------------------------------------------------------------------------------- App. A.3 preprocess
// Tile counters: the Gaussians of a block are neighbours on the mesh and touch the same few dozen tiles, so the block counts in
// a shared-memory histogram of the tile grid and adds each non-zero bin to the global counter once (one global atomic per block
// and tile instead of one per Gaussian and tile: at 220 416 Gaussians the contended global atomics were the whole kernel).
constexpr int kMaxSmemTiles = 4096;         // tile grids up to 1024 x 1024 pixels; larger images count in global memory directly

__global__ void __launch_bounds__(kThreads) k_preprocess(FwdDev a) {
    __shared__ float cam[34];
    __shared__ uint32_t s_cnt[kMaxSmemTiles];
    const int b = blockIdx.y;
    const bool smem_hist = a.T <= kMaxSmemTiles;
    if (threadIdx.x < 16) cam[threadIdx.x] = a.view[b * 16 + threadIdx.x];
    else if (threadIdx.x < 32) cam[threadIdx.x] = a.proj[b * 16 + threadIdx.x - 16];
    else if (threadIdx.x < 34) cam[threadIdx.x] = a.tanfov[b * 2 + threadIdx.x - 32];
    if (smem_hist)
        for (int t = threadIdx.x; t < a.T; t += kThreads) s_cnt[t] = 0u;
    __syncthreads();
    const int g = blockIdx.x * kThreads + threadIdx.x;
    if (g < a.P) {
    const float *view = cam, *proj = cam + 16;
    const float tanfovx = cam[32], tanfovy = cam[33];
    const long long o = (long long)b * a.P + g;

    int radius = 0;
    float depth = 0.f;
    float2 pxy = make_float2(0.f, 0.f);
    float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 rc = make_int4(0, 0, 0, 0);

    const float *mp = a.means3D + b * a.means3D_stride + 3LL * g;
    const float p[3] = {mp[0], mp[1], mp[2]};
    float pv[3];
    xform4x3(view, p, pv);
    if (pv[2] > 0.2f) {                                      // near cull: z_view <= 0.2 is dropped
        const float2 *cp = reinterpret_cast<const float2 *>(a.cov3D + b * a.cov3D_stride + 6LL * g);
        const float2 c01 = cp[0], c23 = cp[1], c45 = cp[2];
        const float s[6] = {c01.x, c01.y, c23.x, c23.y, c45.x, c45.y};
        float ph[3];
        xform4x3(proj, p, ph);
        const float pw = xdiv(1.0f, xadd(xform_w(proj, p), 0.0000001f));
        const float ppx = xmul(ph[0], pw), ppy = xmul(ph[1], pw);
        const float fx = xdiv((float)a.W, xmul(2.0f, tanfovx)), fy = xdiv((float)a.H, xmul(2.0f, tanfovy));
        Cov2D q;
        cov2d_exact(p, s, view, fx, fy, tanfovx, tanfovy, q);
        const float det = xsub(xmul(q.a, q.c), xmul(q.b, q.b));
        if (det != 0.0f) {
            const float det_inv = xdiv(1.f, det);
            const float mid = xmul(0.5f, xadd(q.a, q.c));
            const float disc = xsqrt(fmaxf(0.1f, xsub(xmul(mid, mid), det)));
            const float lam1 = xadd(mid, disc), lam2 = xsub(mid, disc);
            const int r = (int)ceilf(xmul(3.f, xsqrt(fmaxf(lam1, lam2))));
            const float px = ndc2pix(ppx, a.W), py = ndc2pix(ppy, a.H);
            const float rf = (float)r;
            const int minx = min(a.gx, max(0, (int)xdiv(xsub(px, rf), 16.f)));
            const int miny = min(a.gy, max(0, (int)xdiv(xsub(py, rf), 16.f)));
            const int maxx = min(a.gx, max(0, (int)xdiv(xsub(xadd(xadd(px, rf), 16.f), 1.0f), 16.f)));
            const int maxy = min(a.gy, max(0, (int)xdiv(xsub(xadd(xadd(py, rf), 16.f), 1.0f), 16.f)));
            if ((maxx - minx) * (maxy - miny) > 0) {
                radius = r;
                depth = pv[2];
                pxy = make_float2(px, py);
                co = make_float4(xmul(q.c, det_inv), xmul(-q.b, det_inv), xmul(q.a, det_inv),
                                 a.opac[b * a.opac_stride + g]);
                rc = make_int4(minx, miny, maxx, maxy);
                uint32_t *cnt = smem_hist ? s_cnt : a.tile_count + (long long)b * a.T;
                for (int y = miny; y < maxy; y++)
                    for (int x = minx; x < maxx; x++) atomicAdd(cnt + y * a.gx + x, 1u);
            }
        }
    }
    a.radii[o] = radius;
    a.depth[o] = depth;
    a.xy[o] = pxy;
    a.conic_opacity[o] = co;
    a.rect[o] = rc;
    }
    if (smem_hist) {
        __syncthreads();
        uint32_t *cnt = a.tile_count + (long long)b * a.T;
        for (int t = threadIdx.x; t < a.T; t += kThreads) {
            const uint32_t c = s_cnt[t];
            if (c) atomicAdd(cnt + t, c);
        }
    }
}

// ------------------------------------------------------------------------- per-frame exclusive scan of tile counts
__global__ void __launch_bounds__(1024) k_scan_tiles(FwdDev a) {
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

// ----------------------------------------------------- App. A.4: one (depth,id) key per touched tile, into its segment
__global__ void __launch_bounds__(kThreads) k_emit(FwdDev a) {
    __shared__ uint32_t s_cnt[kMaxSmemTiles], s_base[kMaxSmemTiles];
    const int b = blockIdx.y;
    const int g = blockIdx.x * kThreads + threadIdx.x;
    const bool smem_hist = a.T <= kMaxSmemTiles;
    const long long o = (long long)b * a.P + g;
    int4 rc = make_int4(0, 0, 0, 0);
    unsigned long long key = 0ull;
    if (g < a.P) {
        rc = a.rect[o];
        key = ((unsigned long long)__float_as_uint(a.depth[o]) << 32) | (uint32_t)g;
    }
    const bool live = rc.z > rc.x && rc.w > rc.y;
    uint32_t *cur = a.tile_cursor + (long long)b * a.T;
    unsigned long long *keys = a.inst_keys + (long long)b * a.cap;
    if (!smem_hist) {                                     // huge tile grid: one global atomic per instance
        if (live)
            for (int y = rc.y; y < rc.w; y++)
                for (int x = rc.x; x < rc.z; x++) {
                    const uint32_t pos = atomicAdd(cur + y * a.gx + x, 1u);
                    if ((long long)pos < a.cap) keys[pos] = key;
                }
        return;
    }
    // block-level reservation: count per tile in shared memory, reserve each tile's range with ONE global atomic, then hand
    // out the slots of the range with shared-memory atomics (order inside a tile's segment is arbitrary anyway: it is sorted next)
    for (int t = threadIdx.x; t < a.T; t += kThreads) s_cnt[t] = 0u;
    __syncthreads();
    if (live)
        for (int y = rc.y; y < rc.w; y++)
            for (int x = rc.x; x < rc.z; x++) atomicAdd(&s_cnt[y * a.gx + x], 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < a.T; t += kThreads) {
        const uint32_t c = s_cnt[t];
        if (c) { s_base[t] = atomicAdd(cur + t, c); s_cnt[t] = 0u; }
    }
    __syncthreads();
    if (live)
        for (int y = rc.y; y < rc.w; y++)
            for (int x = rc.x; x < rc.z; x++) {
                const int t = y * a.gx + x;
                const uint32_t pos = s_base[t] + atomicAdd(&s_cnt[t], 1u);
                if ((long long)pos < a.cap) keys[pos] = key;
            }
}

// --------------------------------------------------------------------- bitonic network for arbitrary n (in place)
// All compare-exchanges are ascending (min to the lower index), so virtual +inf padding above n never moves and
// pairs that reach beyond n are simply skipped.  Works on shared or global memory (block-wide, kThreads threads).
__device__ __forceinline__ void cmpxchg(unsigned long long *k, int i, int j) {
    const unsigned long long x = k[i], y = k[j];
    if (x > y) { k[i] = y; k[j] = x; }
}

// One stage = `half` independent compare-exchanges; a caller walks a range of them.
__device__ __forceinline__ void flip_pairs(unsigned long long *k, int n, int lsize, int t_begin, int t_end, int t_step) {
    const int size = 1 << lsize, hs = size >> 1, lhs = lsize - 1;
    for (int t = t_begin; t < t_end; t += t_step) {
        const int blk = t >> lhs, w = t & (hs - 1);
        const int i = (blk << lsize) + w, j = (blk << lsize) + (size - 1 - w);
        if (j < n) cmpxchg(k, i, j);
    }
}
__device__ __forceinline__ void disperse_pairs(unsigned long long *k, int n, int lstep, int t_begin, int t_end, int t_step) {
    const int step = 1 << lstep;
    for (int t = t_begin; t < t_end; t += t_step) {
        const int i = ((t >> lstep) << (lstep + 1)) + (t & (step - 1)), j = i + step;
        if (j < n) cmpxchg(k, i, j);
    }
}

// Block-wide sort (kThreads = 8 warps).  Warp w owns the `chunk` = npad / 8 consecutive elements [w chunk, (w+1) chunk):
// every stage whose pairs stay inside a chunk — merges of size <= chunk entirely, and the disperse steps <= chunk / 2 of
// the larger merges — is done by that warp alone behind __syncwarp(); only the stages that cross chunks pay a block
// barrier (10 instead of 66 at 2 048 entries).  Same compare-exchange network as before, so the result is identical.
__device__ void block_sort(unsigned long long *k, int n, int tid) {
    if (n < 2) return;
    int lpad = 1;
    while ((1 << lpad) < n) lpad++;
    const int half = 1 << (lpad - 1);                                  // pairs per stage
    const int lane = tid & 31, warp = tid >> 5;
    if (lpad <= 6) {                                                   // <= 64 entries: one warp, one pair per lane
        if (warp == 0) {
            for (int lsize = 1; lsize <= lpad; lsize++) {
                flip_pairs(k, n, lsize, lane, half, 32);
                __syncwarp();
                for (int lstep = lsize - 2; lstep >= 0; lstep--) {
                    disperse_pairs(k, n, lstep, lane, half, 32);
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        return;
    }
    const int lchunk = lpad - 3;                                       // log2(chunk), >= 4
    const int cp = 1 << (lchunk - 1);                                  // pairs per chunk
    const int w_begin = warp * cp + lane, w_end = (warp + 1) * cp;
    for (int lsize = 1; lsize <= lpad; lsize++) {
        if (lsize <= lchunk) {                                         // the whole merge stays inside the chunks
            flip_pairs(k, n, lsize, w_begin, w_end, 32);
            __syncwarp();
            for (int lstep = lsize - 2; lstep >= 0; lstep--) {
                disperse_pairs(k, n, lstep, w_begin, w_end, 32);
                __syncwarp();
            }
            if (lsize == lchunk) __syncthreads();                      // the next merge crosses chunks
        } else {
            flip_pairs(k, n, lsize, tid, half, kThreads);
            __syncthreads();
            int lstep = lsize - 2;
            for (; lstep >= lchunk; lstep--) {                         // pairs (i, i + step) with 2 step > chunk
                disperse_pairs(k, n, lstep, tid, half, kThreads);
                __syncthreads();
            }
            for (; lstep >= 0; lstep--) {
                disperse_pairs(k, n, lstep, w_begin, w_end, 32);
                __syncwarp();
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ all B*T tiles ordered by list length, longest first
// One block.  Key = bit length of the list length and its next two bits (132 classes); order inside a class is arbitrary
// (it only decides which SM runs which tile, never a result).
__global__ void __launch_bounds__(1024) k_worklist(FwdDev a) {
    constexpr int kBins = 33 * 4;
    __shared__ uint32_t hist[kBins];
    const int total = a.B * a.T;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    auto bin_of = [&](int e) {
        const int b = e / a.T, t = e - b * a.T;
        const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
        const uint32_t n = off[t + 1] - off[t];
        if (n == 0) return kBins - 1;
        const int len = 32 - __clz(n);                                   // 1 .. 32
        const uint32_t frac = len >= 3 ? (n >> (len - 3)) & 3u : (n << (3 - len)) & 3u;
        return (32 - len) * 4 + (3 - (int)frac);
    };
    for (int e = threadIdx.x; e < total; e += blockDim.x) atomicAdd(&hist[bin_of(e)], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < kBins; i++) { const uint32_t c = hist[i]; hist[i] = run; run += c; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) a.worklist[atomicAdd(&hist[bin_of(e)], 1u)] = (uint32_t)e;
}

// ------------------------------------------------------------------------------------ App. A.4: per-tile depth sort
// Sorts one tile's (depth bits << 32 | id) keys ascending — exactly the order of upstream's stable radix sort of
// (tile | depth) keys over instances emitted in id order — and writes the ids to point_list.
//   1. min / max of the depth bits over the list; bucket = (bits - min) >> shift, shift chosen so that the tile's own depth
//      range spans the kBuckets buckets (positive floats order like their bit patterns);
//   2. histogram + exclusive scan of the bucket counters;
//   3. scatter into shared memory at bucket base + arrival order (arrival order is arbitrary, hence:)
//   4. every bucket holding more than one key is put in exact key order by one thread (insertion; buckets are a fraction of a
//      millimetre deep, so runs are 1-3 keys).  A bucket with more than kMaxRun keys (flat geometry, degenerate range) sends the
//      pass to the block-wide bitonic network instead.
// Lists longer than kSortCap are processed as consecutive bucket ranges of at most kSortCap keys each, re-reading the list
// once per range; only a single bucket with more than kSortCap keys falls back to the bitonic network in global memory.
// bit s of the result: entry (xy, conic/opacity) passes the conservative alpha >= 1/255 test somewhere in sub-block s
// (8 columns x 4 rows; s & 1 = left / right half, s >> 1 = row band) of the tile whose top-left pixel is (tx0, ty0)
__device__ __forceinline__ uint32_t sub_block_mask(float2 c, float4 co, int tx0, int ty0) {
    uint32_t m = 0;
#pragma unroll
    for (int sb = 0; sb < 8; sb++) {
        const float rcx = (float)(tx0 + (sb & 1) * 8) + 3.5f, rcy = (float)(ty0 + (sb >> 1) * 4) + 1.5f;
        m |= entry_reaches_rect(c, co, rcx, rcy, 3.5f, 1.5f) ? (1u << sb) : 0u;
    }
    return m;
}

__global__ void __launch_bounds__(kThreads) k_tile_sort(FwdDev a) {
    extern __shared__ __align__(16) unsigned long long skeys[];          // kSortCap keys, then kBuckets counters
    uint32_t *cnt = reinterpret_cast<uint32_t *>(skeys + kSortCap);
    __shared__ uint32_t s_red[2][kThreads / 32];
    __shared__ uint32_t s_scan[kThreads / 32];
    __shared__ int s_flag;

    const uint32_t entry = a.worklist ? a.worklist[blockIdx.x] : blockIdx.x;
    const int b = (int)(entry / a.T), tile = (int)(entry % a.T);
    const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
    long long start = off[tile], end = off[tile + 1];
    if (start > a.cap) start = a.cap;                      // overflowed frame: stay in bounds, result is flagged
    if (end > a.cap) end = a.cap;
    const int n = (int)(end - start);
    if (n == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long *gkeys = a.inst_keys + (long long)b * a.cap + start;
    uint32_t *plist = a.point_list + (long long)b * a.cap + start;
    // the sorted ids leave together with their sub-block masks: the cull test runs once per entry here (all lanes busy)
    // instead of once per entry and sub-block warp in the forward AND the backward blend
    uint8_t *pmask = a.point_mask ? a.point_mask + (long long)b * a.cap + start : nullptr;
    const float2 *gxy = a.xy + (long long)b * a.P;
    const float4 *gco = a.conic_opacity + (long long)b * a.P;
    const int tx0 = (tile % a.gx) * 16, ty0 = (tile / a.gx) * 16;
    auto emit = [&](int pos, uint32_t id) {
        plist[pos] = id;
        if (pmask) pmask[pos] = (uint8_t)sub_block_mask(__ldg(gxy + id), __ldg(gco + id), tx0, ty0);
    };
    if (n == 1) { if (tid == 0) emit(0, (uint32_t)gkeys[0]); return; }

    // 1. depth range
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int i = tid; i < n; i += kThreads) { const uint32_t d = (uint32_t)(gkeys[i] >> 32); lo = min(lo, d); hi = max(hi, d); }
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
    for (int i = tid; i < kBuckets; i += kThreads) cnt[i] = 0;
    if (tid == 0) s_flag = 0;
    __syncthreads();
    lo = s_red[0][0]; hi = s_red[1][0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; w++) { lo = min(lo, s_red[0][w]); hi = max(hi, s_red[1][w]); }
    const uint32_t range = hi - lo;
    const int shift = range ? max(0, (32 - __clz(range)) - kLogBuckets) : 0;
    // 2. histogram, exclusive scan (16 consecutive buckets per thread)
    for (int i = tid; i < n; i += kThreads) atomicAdd(&cnt[((uint32_t)(gkeys[i] >> 32) - lo) >> shift], 1u);
    __syncthreads();
    {
        constexpr int kPer = kBuckets / kThreads;
        uint32_t v[kPer], sum = 0;
#pragma unroll
        for (int k = 0; k < kPer; k++) { v[k] = cnt[tid * kPer + k]; sum += v[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t base = incl - sum;
        for (int w = 0; w < warp; w++) base += s_scan[w];
#pragma unroll
        for (int k = 0; k < kPer; k++) { cnt[tid * kPer + k] = base; base += v[k]; }
    }
    __syncthreads();
    // 3./4. one pass per bucket range of at most kSortCap keys.  cnt[k] is the exclusive base of bucket k until the bucket has
    // been scattered, its end afterwards (= the base of bucket k + 1).
    int lo_b = 0;
    while (lo_b < kBuckets) {
        const uint32_t base_lo = cnt[lo_b];
        if (base_lo >= (uint32_t)n) break;                               // only empty buckets are left
        int hi_b;                                                        // largest hi_b with base(hi_b) - base_lo <= kSortCap
        if ((uint32_t)n - base_lo <= (uint32_t)kSortCap) hi_b = kBuckets;
        else {
            int l = lo_b, r = kBuckets - 1;                              // invariant: base(l) fits, base(r + 1) may not
            while (l < r) { const int m = (l + r + 1) >> 1; if (cnt[m] - base_lo <= (uint32_t)kSortCap) l = m; else r = m - 1; }
            hi_b = l;
        }
        if (hi_b == lo_b) {                                              // one bucket alone exceeds a pass: global-memory bitonic
            __syncthreads();
            block_sort(gkeys, n, tid);
            for (int i = tid; i < n; i += kThreads) emit(i, (uint32_t)gkeys[i]);
            return;
        }
        const uint32_t count = (hi_b == kBuckets ? (uint32_t)n : cnt[hi_b]) - base_lo;
        __syncthreads();                                                 // everyone has read the bases of this range
        for (int i = tid; i < n; i += kThreads) {
            const unsigned long long key = gkeys[i];
            const int bk = (int)(((uint32_t)(key >> 32) - lo) >> shift);
            if (bk >= lo_b && bk < hi_b) skeys[atomicAdd(&cnt[bk], 1u) - base_lo] = key;
        }
        __syncthreads();
        for (int bk = lo_b + tid; bk < hi_b; bk += kThreads) {
            const int rb = (int)((bk == lo_b ? base_lo : cnt[bk - 1]) - base_lo), re = (int)(cnt[bk] - base_lo);
            if (re - rb > kMaxRun) { s_flag = 1; continue; }
            for (int i = rb + 1; i < re; i++) {
                const unsigned long long key = skeys[i];
                int j = i - 1;
                while (j >= rb && skeys[j] > key) { skeys[j + 1] = skeys[j]; j--; }
                skeys[j + 1] = key;
            }
        }
        __syncthreads();
        if (s_flag) {                                                    // a long run: order the whole pass with the network
            block_sort(skeys, (int)count, tid);
            __syncthreads();
            if (tid == 0) s_flag = 0;
        }
        for (int i = tid; i < (int)count; i += kThreads) emit((int)base_lo + i, (uint32_t)skeys[i]);
        __syncthreads();
        lo_b = hi_b;
    }
}

// ------------------------------------------------------------------------------------------ App. A.5: alpha blending
// Every warp is independent: it owns one 8x4-pixel sub-block of one tile (global warp index -> worklist entry, sub-block) and
// walks the tile's sorted list front to back in chunks of 32 entries — one entry per lane is fetched and tested against the
// sub-block (entry_reaches_rect), survivors are staged in the warp's own shared-memory slots and visited through the ballot
// mask.  A warp whose 32 pixels are saturated stops; the next chunk's records are prefetched while the current one is blended.
template <int C>
__global__ void __launch_bounds__(kBlendThreads) k_blend(FwdDev a) {
    constexpr int kWarps = kBlendThreads / 32;
    __shared__ float2 s_xy[kWarps][32];
    __shared__ float4 s_co[kWarps][32];
    __shared__ __align__(16) float s_col[kWarps][32 * C];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kWarps + wib;
    const int sub = (int)(gw & 7);
    const long long wi = gw >> 3;
    if (wi >= (long long)a.B * a.T) return;
    const uint32_t entry = a.worklist ? a.worklist[wi] : (uint32_t)wi;
    const int b = (int)(entry / a.T), tile = (int)(entry % a.T);
    const uint32_t *off = a.tile_offset + (long long)b * (a.T + 1);
    long long start = off[tile], end = off[tile + 1];
    if (start > a.cap) start = a.cap;
    if (end > a.cap) end = a.cap;
    const int n = (int)(end - start);
    const uint32_t *plist = a.point_list + (long long)b * a.cap + start;

    const int tx = tile % a.gx, ty = tile / a.gx;
    const int x0 = tx * 16 + (sub & 1) * 8, y0 = ty * 16 + (sub >> 1) * 4;
    const int x = x0 + (lane & 7), y = y0 + (lane >> 3);
    const bool inside = x < a.W && y < a.H;
    const float pxf = (float)x, pyf = (float)y;
    const float rcx = (float)x0 + 3.5f, rcy = (float)y0 + 1.5f;
    const float2 *gxy = a.xy + (long long)b * a.P;
    const float4 *gco = a.conic_opacity + (long long)b * a.P;
    const float *gcol = a.colors + b * a.colors_stride;

    bool done = !inside;
    float T = 1.0f;
    float acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) acc[ch] = 0.f;
    uint32_t last = 0;

    const uint8_t *pmask = a.point_mask ? a.point_mask + (long long)b * a.cap + start : nullptr;
    // software pipeline: ids (+ masks) are fetched two chunks ahead, the records of the entries that reach this sub-block one
    // chunk ahead, so neither latency sits in front of the blend of the current chunk
    uint32_t id_c = 0, id_n = 0; bool rel_c = false, rel_n = false;
    float2 xy_c = make_float2(0.f, 0.f); float4 co_c = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_id = [&](int idx, uint32_t &id, bool &rel) {
        rel = false;
        if (idx < n) { id = plist[idx]; rel = pmask ? ((pmask[idx] >> sub) & 1u) != 0u : true; }
    };
    fetch_id(lane, id_c, rel_c);
    fetch_id(32 + lane, id_n, rel_n);
    if (rel_c) { xy_c = __ldg(gxy + id_c); co_c = __ldg(gco + id_c); }
    for (int base = 0; base < n; base += 32) {
        if (__all_sync(0xffffffffu, done)) break;
        const bool rel = rel_c && (pmask != nullptr || entry_reaches_rect(xy_c, co_c, rcx, rcy, 3.5f, 1.5f));
        unsigned mask = __ballot_sync(0xffffffffu, rel);
        if (rel) {
            s_xy[wib][lane] = xy_c;
            s_co[wib][lane] = co_c;
            if constexpr (C == 4) {
                reinterpret_cast<float4 *>(s_col[wib])[lane] = __ldg(reinterpret_cast<const float4 *>(gcol) + id_c);
            } else {
#pragma unroll
                for (int ch = 0; ch < C; ch++) s_col[wib][lane * C + ch] = __ldg(gcol + (long long)id_c * C + ch);
            }
        }
        id_c = id_n; rel_c = rel_n;                        // advance the pipeline
        if (rel_c) { xy_c = __ldg(gxy + id_c); co_c = __ldg(gco + id_c); }
        fetch_id(base + 64 + lane, id_n, rel_n);
        __syncwarp();
        // The transmittance chain T <- T (1 - alpha) is the only true dependency between entries: alpha of the next kU
        // survivors is evaluated first (independent shared-memory reads, FMAs and exp's that overlap), then the chain is
        // walked.  Same operations in the same order per pixel as a one-by-one loop.
        constexpr int kU = 4;
        while (mask) {
            int j[kU];
            float alpha[kU];
            bool ok[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const bool have = mask != 0u;
                j[u] = have ? __ffs(mask) - 1 : 0;
                mask &= mask - 1;
                const float2 c = s_xy[wib][j[u]];
                const float4 co = s_co[wib][j[u]];
                const float dx = c.x - pxf, dy = c.y - pyf;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                alpha[u] = fminf(0.99f, co.w * __expf(power));
                ok[u] = have && power <= 0.0f && alpha[u] >= 1.0f / 255.0f;
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                if (!ok[u] || done) continue;
                const float test_T = T * (1.f - alpha[u]);
                if (test_T < 0.0001f) { done = true; continue; }
                const float w = alpha[u] * T;
#pragma unroll
                for (int ch = 0; ch < C; ch++) acc[ch] += s_col[wib][j[u] * C + ch] * w;
                T = test_T;
                last = (uint32_t)(base + j[u] + 1);
            }
        }
        __syncwarp();
    }
    if (inside) {
        const long long pix = ((long long)b * a.H + y) * a.W + x;
        a.final_T[pix] = T;
        a.n_contrib[pix] = last;
        const float *bg = a.bg + b * C;
        if (a.interleaved) {
            if (C == 4) {
                reinterpret_cast<float4 *>(a.out_color)[pix] =
                    make_float4(acc[0] + T * bg[0], acc[1] + T * bg[1], acc[2] + T * bg[2], acc[C - 1] + T * bg[C - 1]);
            } else {
#pragma unroll
                for (int ch = 0; ch < C; ch++) a.out_color[pix * C + ch] = acc[ch] + T * bg[ch];
            }
        } else {
            const long long hw = (long long)a.H * a.W;
#pragma unroll
            for (int ch = 0; ch < C; ch++)
                a.out_color[((long long)b * C + ch) * hw + (long long)y * a.W + x] = acc[ch] + T * bg[ch];
        }
    }
}

}  // namespace

extern "C" int gom_raster_forward(const GomRasterFwdArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_gauss >= 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->n_channels == 3 || p->n_channels == 4, "n_channels must be 3 or 4");
    GOM_REQUIRE(p->inst_capacity > 0 && p->inst_capacity < 0xffffffffLL, "inst_capacity");
    GOM_REQUIRE(p->n_frames <= 65535, "n_frames");
    GOM_REQUIRE(p->means3D && p->cov3D && p->colors && p->opacities && p->viewmatrix && p->projmatrix && p->tanfov &&
                    p->bg, "null input");
    GOM_REQUIRE(p->out_color && p->final_T && p->n_contrib && p->radii && p->depth && p->xy && p->conic_opacity &&
                    p->rect && p->tile_count && p->tile_offset && p->tile_cursor && p->inst_keys && p->point_list &&
                    p->status, "null output/state");
    GOM_REQUIRE(((uintptr_t)p->cov3D % 8) == 0 && (p->cov3D_stride % 2) == 0, "cov3D must be 8-byte aligned");
    GOM_REQUIRE(!(p->interleaved && p->n_channels == 4) || ((uintptr_t)p->out_color % 16) == 0, "out_color alignment");
    GOM_REQUIRE(p->n_channels != 4 || (((uintptr_t)p->colors % 16) == 0 && (p->colors_stride % 4) == 0), "4-channel colors must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    FwdDev a;
    a.B = p->n_frames; a.P = p->n_gauss; a.H = p->height; a.W = p->width; a.C = p->n_channels;
    a.interleaved = p->interleaved;
    a.gx = (a.W + GOM_TILE - 1) / GOM_TILE; a.gy = (a.H + GOM_TILE - 1) / GOM_TILE; a.T = a.gx * a.gy;
    GOM_REQUIRE((long long)a.B * a.T * 8 < 0x7fffffffLL, "too many tiles");
    a.cap = p->inst_capacity;
    a.means3D = p->means3D; a.means3D_stride = p->means3D_stride;
    a.cov3D = p->cov3D; a.cov3D_stride = p->cov3D_stride;
    a.colors = p->colors; a.colors_stride = p->colors_stride;
    a.opac = p->opacities; a.opac_stride = p->opacities_stride;
    a.view = p->viewmatrix; a.proj = p->projmatrix; a.tanfov = p->tanfov; a.bg = p->bg;
    a.out_color = p->out_color; a.final_T = p->final_T; a.n_contrib = p->n_contrib; a.radii = p->radii;
    a.depth = p->depth; a.xy = reinterpret_cast<float2 *>(p->xy);
    a.conic_opacity = reinterpret_cast<float4 *>(p->conic_opacity); a.rect = reinterpret_cast<int4 *>(p->rect);
    a.tile_count = p->tile_count; a.tile_offset = p->tile_offset; a.tile_cursor = p->tile_cursor;
    a.inst_keys = reinterpret_cast<unsigned long long *>(p->inst_keys); a.point_list = p->point_list;
    a.status = p->status;
    a.worklist = p->worklist;
    a.point_mask = p->point_mask;
    GOM_REQUIRE(((uintptr_t)p->xy % 8) == 0 && ((uintptr_t)p->conic_opacity % 16) == 0 && ((uintptr_t)p->rect % 16) == 0 &&
                    ((uintptr_t)p->inst_keys % 8) == 0, "state alignment");

    GOM_CUDA(cudaMemsetAsync(a.tile_count, 0, sizeof(uint32_t) * (size_t)a.B * a.T, stream));
    if (a.P > 0) {
        dim3 grid(gom_div_up(a.P, kThreads), a.B);
        gom_prof_begin(GOM_PROF_PREPROCESS, stream);
        k_preprocess<<<grid, kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
        gom_prof_end(GOM_PROF_PREPROCESS, stream);
    }
    gom_prof_begin(GOM_PROF_SCAN, stream);
    k_scan_tiles<<<a.B, 1024, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_SCAN, stream);
    if (a.P > 0) {
        dim3 grid(gom_div_up(a.P, kThreads), a.B);
        gom_prof_begin(GOM_PROF_EMIT, stream);
        k_emit<<<grid, kThreads, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
        gom_prof_end(GOM_PROF_EMIT, stream);
    }
    const int n_tiles = a.B * a.T;
    if (a.worklist) {
        gom_prof_begin(GOM_PROF_WORKLIST, stream);
        k_worklist<<<1, 1024, 0, stream>>>(a);
        GOM_LAUNCH_CHECK();
        gom_prof_end(GOM_PROF_WORKLIST, stream);
    }
    constexpr size_t kSortSmem = sizeof(unsigned long long) * kSortCap + sizeof(uint32_t) * kBuckets;
    {
        // > 48 KB of dynamic shared memory is an opt-in per function AND per device (a process may drive several)
        static bool big_smem_enabled[64] = {};
        int dev = 0;
        GOM_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !big_smem_enabled[dev]) {
            GOM_CUDA(cudaFuncSetAttribute(k_tile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
            if (dev >= 0 && dev < 64) big_smem_enabled[dev] = true;
        }
    }
    gom_prof_begin(GOM_PROF_TILE_SORT, stream);
    k_tile_sort<<<n_tiles, kThreads, kSortSmem, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_TILE_SORT, stream);
    const unsigned bgrid = (unsigned)(((long long)n_tiles * 8 + kBlendThreads / 32 - 1) / (kBlendThreads / 32));
    gom_prof_begin(GOM_PROF_BLEND_FWD, stream);
    if (a.C == 3) k_blend<3><<<bgrid, kBlendThreads, 0, stream>>>(a);
    else k_blend<4><<<bgrid, kBlendThreads, 0, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_BLEND_FWD, stream);
    return GOM_OK;
}

// shadow_mlp.cu — the pseudo-shading MLP of reference models/modules/shadow_module.py:67-117 as one
// tcgen05 kernel over the FOREGROUND pixels of the normal map.
//
//   reference:  posenc(normal[B*H*W,3]) -> Linear(39,128) ReLU -> Linear(128,128) ReLU -> Linear(128,128) ReLU
//               -> Linear(128,1) -> sigmoid, on every pixel, ~25 torch launches, every activation through HBM.
//   here:       1. compaction: pixels whose normal is not exactly (0,0,0) (the mesh renderer writes exact zeros on the
//                  background, mesh.py:103-112) are listed in pixel order (count -> scan -> scatter, deterministic);
//                  background pixels receive the one constant sigmoid(MLP(posenc(0))), computed once per call;
//               2. weights are split w = hi + lo into two TF32 numbers and written as 128B-swizzled K-major shared
//                  memory images (k_shadow_prep), so the main kernel streams them with plain 1-D TMA bulk copies;
//               3. k_shadow_fwd: persistent, one CTA per SM, 128 pixels per tile.  The activations NEVER leave tensor
//                  memory: the epilogue warps read the fp32 accumulator with tcgen05.ld, apply bias + ReLU, split the
//                  result into TF32 hi/lo and write it back with tcgen05.st as the A operand of the next layer's
//                  tcgen05.mma (A from TMEM, B = weight image in shared memory).  Every product is formed as
//                  lo*hi + hi*lo + hi*hi (3xTF32, fp32 accumulate): error ~2^-21 like an fp32 GEMM, at a third of the
//                  TF32 tensor rate instead of the FFMA rate.  The last layer (128 -> 1) and the sigmoid are done in
//                  registers.
//   backward:   4. when a backward pass will follow, the forward also writes every operand it formed (encoding and
//                  post-ReLU activations, fp32) as ready-made K-major SWIZZLE_128B "activation images"
//                  [feature][pixel row] per 32-row block, so that later kernels consume them with 1-D TMA;
//               5. k_shadow_bwd_data: the same pipeline run backwards (dZ_l = (dZ_{l+1} W_{l+1}) * [H_l > 0], transposed
//                  weight images), ending in the positional-encoding backward -> dL/dnormal; it writes dZ_l images;
//               6. k_shadow_bwd_weights: split-K GEMMs dW_l = dZ_l^T H_{l-1} (both operands are the saved images, both
//                  from shared memory), accumulated over all tiles of a CTA in tensor memory; a row of ones appended
//                  to the B operand yields the bias gradients in the same MMAs; per-CTA partials are summed in a fixed
//                  order by k_shadow_bwd_reduce (deterministic).
//
// TMEM map of k_shadow_fwd / k_shadow_bwd_data (512 columns): [0,128) and [128,256) accumulators D0/D1 (alternating per
// layer), [256,384) A_hi, [384,512) A_lo; lane = pixel row of the tile.
// Warp roles (192 threads): warps 0-3 epilogue (warp w owns TMEM lanes 32w..32w+31), warp 4 producer (TMA), warp 5 MMA
// issuer.  The next layer's MMAs start k-block by k-block while the epilogue is still producing the rest of its operand;
// the weight stream (6 x 32 KB stages) runs ahead of both.
#include "gom_common.cuh"
#include "gom_tcgen05.cuh"

namespace {

using namespace gomtc;

constexpr int kWidth = 128;             // hidden width (the only one the reference configs use)
constexpr int kStageBytes = 32768;      // one K-block of 32 columns: [hi 128x32 fp32 | lo 128x32 fp32]
constexpr int kHalfStage = 16384;
constexpr int kStages = 6;
constexpr int kMaxDepth = 8;
constexpr int kEncPad = 64;             // encoding columns in TMEM (3 + 6*multires <= 63)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 0, kColAhi = 256, kColAlo = 384;   // D is double-buffered: [0,128), [128,256)
constexpr int kThreads = 192;
constexpr uint32_t kStatusTimeout = 2u;

struct ShadowDev {
    long long n_pixels, capacity;          // capacity: rows (a multiple of 128) the image buffers hold
    int multires, enc, depth, save_hidden, n_ctas;
    const float *normals;
    const float *W_in, *b_in, *W_hid, *b_hid, *W_out, *b_out;
    uint32_t *block_count;
    int *fg_index, *n_fg;
    float *w_images, *bg_value, *out;
    uint32_t *act_img, *dz_img;
    uint32_t *status;
    const float *g_out;
    float *g_normals, *dzo_sums, *partials;
    float *g_W_in, *g_b_in, *g_W_hid, *g_b_hid, *g_w_out, *g_b_out;
};

// ------------------------------------------------------------------------------------------- activation images
// One tile (128 pixel rows) of saved operands, in 32-bit words.  Row block kb = rows 32 kb .. 32 kb + 31 of the tile
// (= the epilogue warp kb); inside an image, element (feature j, row-in-block q) sits at word swz(j, q): row j of a
// K-major SWIZZLE_128B matrix whose K index is the pixel row.  Images hold the fp32 VALUES (one copy): the tensor core
// ignores the 13 low mantissa bits of a TF32 operand, so an image is its own "hi" part and k_shadow_bwd_weights forms
// lo = v - trunc(v) in shared memory next to it (storing hi and lo separately doubled the HBM bytes of three kernels).
//   act tile: [slot 0 (encoding, 64 features): 4 blocks x 2048] [slot s = 1..depth (H_s, 128 features): 4 blocks x 4096]
//   dz tile:  [layer l = 0..depth-1 (dZ_l, 128 features): 4 blocks x 4096] [dz_out, 16 rows of which row 0 is used: 4 x 512]
__host__ __device__ __forceinline__ size_t act_tile_words(int depth) { return 8192 + (size_t)depth * 16384; }
__host__ __device__ __forceinline__ size_t dz_tile_words(int depth) { return (size_t)depth * 16384 + 2048; }
__host__ __device__ __forceinline__ size_t act_slot_block(int slot, int kb) {       // word offset of (slot, kb) in an act tile
    return slot == 0 ? (size_t)kb * 2048 : 8192 + (size_t)(slot - 1) * 16384 + (size_t)kb * 4096;
}
__host__ __device__ __forceinline__ size_t act_slot_words(int slot) { return slot == 0 ? 2048 : 4096; }   // words of one image
__device__ __forceinline__ int swz(int j, int q) { return j * 32 + ((((q >> 2) ^ (j & 7)) << 2) | (q & 3)); }

// number of 32-column K-blocks of layer l and the images before it
__host__ __device__ __forceinline__ int layer_kblocks(int l) { return l == 0 ? 2 : 4; }
__host__ __device__ __forceinline__ int images_total(int depth) { return 2 + 4 * (depth - 1); }
__host__ __device__ __forceinline__ int images_total_with_backward(int depth) { return images_total(depth) + 4 * depth; }

// ------------------------------------------------------------------------------------------------- weight images
// Image (layer l, K-block kb) = 32 KB: hi part then lo part, each the canonical K-major SWIZZLE_128B layout of a
// [128 rows (output feature n)] x [32 fp32 (input feature k)] tile: row n at byte n*128, its 16-byte chunk c stored at
// chunk position c ^ (n & 7).  Columns beyond the layer's input width are zero.  The last block evaluates the MLP on
// posenc(0) in plain fp32 -> bg_value.
__global__ void k_shadow_prep(ShadowDev a) {
    const int n_img = images_total(a.depth);
    if ((int)blockIdx.x == (int)gridDim.x - 1) {
        __shared__ float h0[kWidth], h1[kWidth];
        const int t = threadIdx.x;
        float *cur = h0, *nxt = h1;
        if (t < kWidth) {
            float s = a.b_in[t];
            for (int k = 0; k < a.enc; k++) {
                const float e = (k >= 3 && ((k - 3) % 6) >= 3) ? 1.f : 0.f;      // cos(0) = 1, sin(0) = 0, x = 0
                s += a.W_in[t * a.enc + k] * e;
            }
            cur[t] = fmaxf(s, 0.f);
        }
        __syncthreads();
        for (int l = 1; l < a.depth; l++) {
            if (t < kWidth) {
                const float *W = a.W_hid + (size_t)(l - 1) * kWidth * kWidth + (size_t)t * kWidth;
                float s = a.b_hid[(l - 1) * kWidth + t];
                for (int k = 0; k < kWidth; k++) s += W[k] * cur[k];
                nxt[t] = fmaxf(s, 0.f);
            }
            __syncthreads();
            float *tmp = cur; cur = nxt; nxt = tmp;
        }
        if (t == 0) {
            float s = a.b_out[0];
            for (int k = 0; k < kWidth; k++) s += a.W_out[k] * cur[k];
            a.bg_value[0] = 1.f / (1.f + expf(-s));
        }
        return;
    }
    const int e = blockIdx.x * blockDim.x + threadIdx.x;            // one element of one image
    const int n_img_all = n_img + (a.save_hidden ? 4 * a.depth : 0);
    if (e >= n_img_all * 4096) return;
    const int img = e >> 12, n = (e >> 5) & 127, kl = e & 31;
    float w = 0.f;
    if (img < n_img) {                                              // forward: B[n = output feature][k = input feature]
        int l, kb;
        if (img < 2) { l = 0; kb = img; } else { l = 1 + (img - 2) / 4; kb = (img - 2) % 4; }
        const int k = kb * 32 + kl;
        if (l == 0) { if (k < a.enc) w = a.W_in[n * a.enc + k]; }
        else w = a.W_hid[(size_t)(l - 1) * kWidth * kWidth + (size_t)n * kWidth + k];
    } else {                                                        // backward step s: layer l = depth-1-s, B[n = input][k = output]
        const int s_ = (img - n_img) >> 2, kb = (img - n_img) & 3, l = a.depth - 1 - s_;
        const int k = kb * 32 + kl;
        if (l == 0) { if (n < a.enc) w = a.W_in[k * a.enc + n]; }
        else w = a.W_hid[(size_t)(l - 1) * kWidth * kWidth + (size_t)k * kWidth + n];
    }
    uint32_t hi, lo;
    split_tf32(w, hi, lo);
    uint32_t *dst = reinterpret_cast<uint32_t *>(a.w_images) + (size_t)img * (kStageBytes / 4);
    const int pos = n * 32 + ((((kl >> 2) ^ (n & 7))) << 2) + (kl & 3);
    dst[pos] = hi;
    dst[kHalfStage / 4 + pos] = lo;
}

// ----------------------------------------------------------------------------------------------------- compaction
// 1024 pixels per block, 4 consecutive pixels per thread.
__device__ __forceinline__ void load_fg4(const float *normals, long long p0, long long n, bool fg[4]) {
    if (p0 + 3 < n) {
        const float4 *q = reinterpret_cast<const float4 *>(normals + 3 * p0);       // 48-byte aligned
        const float4 u = __ldg(q), v = __ldg(q + 1), w = __ldg(q + 2);
        fg[0] = (u.x != 0.f) | (u.y != 0.f) | (u.z != 0.f);
        fg[1] = (u.w != 0.f) | (v.x != 0.f) | (v.y != 0.f);
        fg[2] = (v.z != 0.f) | (v.w != 0.f) | (w.x != 0.f);
        fg[3] = (w.y != 0.f) | (w.z != 0.f) | (w.w != 0.f);
    } else {
        for (int i = 0; i < 4; i++) {
            fg[i] = false;
            if (p0 + i < n) {
                const float *q = normals + 3 * (p0 + i);
                fg[i] = (q[0] != 0.f) | (q[1] != 0.f) | (q[2] != 0.f);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_shadow_count(ShadowDev a) {
    __shared__ int wsum[8];
    const long long p0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    bool fg[4];
    load_fg4(a.normals, p0, a.n_pixels, fg);
    int c = (int)fg[0] + (int)fg[1] + (int)fg[2] + (int)fg[3];
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < 8; i++) s += wsum[i];
        a.block_count[blockIdx.x] = (uint32_t)s;
    }
}

// exclusive scan of block_count[0..nblk) in place (one block of 1024 threads), total -> n_fg, overflow -> status
__global__ void __launch_bounds__(1024) k_shadow_scan(ShadowDev a, int nblk) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    const int per = (nblk + 1023) / 1024;
    const int lo = t * per, hi = min(lo + per, nblk);
    uint32_t s = 0;
    for (int i = lo; i < hi; i++) s += a.block_count[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const uint32_t v = (t >= d) ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (int i = lo; i < hi; i++) {
        const uint32_t c = a.block_count[i];
        a.block_count[i] = run;
        run += c;
    }
    if (t == 1023) {
        a.n_fg[0] = (int)part[1023];
        a.status[0] = (a.save_hidden && (long long)part[1023] > a.capacity) ? GOM_STATUS_OVERFLOW : 0u;
    }
}

__global__ void __launch_bounds__(256) k_shadow_scatter(ShadowDev a) {
    __shared__ int wsum[8];
    const long long p0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    bool fg[4];
    load_fg4(a.normals, p0, a.n_pixels, fg);
    const int c = (int)fg[0] + (int)fg[1] + (int)fg[2] + (int)fg[3];
    int inc = c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int base = (int)a.block_count[blockIdx.x];
    for (int w = 0; w < warp; w++) base += wsum[w];
    int pos = base + inc - c;
    const float bg = a.bg_value[0];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (p0 + i >= a.n_pixels) break;
        if (fg[i]) a.fg_index[pos++] = (int)(p0 + i);
        else a.out[p0 + i] = bg;
    }
}

// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kWidth >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);

// positional encoding of one normal into 64 zero-padded columns: [n, sin(2^k n), cos(2^k n)]_k
__device__ __forceinline__ void encode_normal(float nx, float ny, float nz, int multires, float enc[kEncPad]) {
#pragma unroll
    for (int c = 0; c < kEncPad; c++) enc[c] = 0.f;
    enc[0] = nx; enc[1] = ny; enc[2] = nz;
    // sincosf at every third octave, two angle doublings in between (sin 2a = 2 s c, cos 2a = 1 - 2 s^2): the absolute
    // error at most quadruples (~2.5e-7), a third of the transcendental work
    float s[3], c[3];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        if (k < multires) {
            if (k % 3 == 0) {
                const float f = (float)(1 << k);
                sincosf(nx * f, &s[0], &c[0]);
                sincosf(ny * f, &s[1], &c[1]);
                sincosf(nz * f, &s[2], &c[2]);
            } else {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float s2 = 2.f * s[j] * c[j], c2 = 1.f - 2.f * s[j] * s[j];
                    s[j] = s2; c[j] = c2;
                }
            }
#pragma unroll
            for (int j = 0; j < 3; j++) { enc[3 + 6 * k + j] = s[j]; enc[3 + 6 * k + 3 + j] = c[j]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ the MLP kernel
// Pipeline (per CTA, g = running layer counter over all its tiles):
//   epilogue(g)  reads accumulator D[g&1] 32 columns at a time, writes the next layer's A operand k-block by k-block and
//                signals a_ready[kb] after each one;
//   MMA(g+1)     starts on k-block kb as soon as a_ready[kb] fires and accumulates into the OTHER buffer D[(g+1)&1], so
//                it overlaps the rest of epilogue(g); one commit -> z_ready when the layer is complete.
//   The last layer's epilogue first writes the NEXT tile's encoding (prefetched and encoded while the tensor core was
//   busy) so that MMA(next tile, layer 0) overlaps the final 128 -> 1 dot product.
template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1) k_shadow_fwd(ShadowDev a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], a_ready[4], z_ready;
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;
    __shared__ float s_bias[kMaxDepth * kWidth], s_wout[kWidth], s_bout;

    uint8_t *stages = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int kb = 0; kb < 4; kb++) mbar_init(&a_ready[kb], kTileRows);
        mbar_init(&z_ready, 1);
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < a.depth * kWidth; i += kThreads)
        s_bias[i] = (i < kWidth) ? a.b_in[i] : a.b_hid[i - kWidth];
    for (int i = threadIdx.x; i < kWidth; i += kThreads) s_wout[i] = a.W_out[i];
    if (threadIdx.x == 0) s_bout = a.b_out[0];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int n_fg = a.n_fg[0];
    const int n_tiles = (n_fg + kTileRows - 1) / kTileRows;
    const int depth = a.depth, n_img = images_total(depth);
    const int ksteps0 = (a.enc + 7) >> 3;
    volatile int *ab = &abort_flag;
    bool ok = true;

    if (warp == 4) {
        // ===================================================== weight producer: 32 KB images, in consumption order
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x)
                for (int img = 0; img < n_img; img++, it++) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    if (!mbar_wait(&empty_bar[s], ph ^ 1u, ab)) { ok = false; break; }
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    tma_bulk_g2s(smem_u32(stages + (size_t)s * kStageBytes),
                                 reinterpret_cast<const uint8_t *>(a.w_images) + (size_t)img * kStageBytes, kStageBytes, &full_bar[s]);
                }
        }
    } else if (warp == 5) {
        // ================================================================================ MMA issuer (one thread)
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x)
                for (int l = 0; l < depth && ok; l++, g++) {
                    const uint32_t dcol = tmem + kColD + (g & 1u) * kWidth;
                    const int ksteps = (l == 0) ? ksteps0 : (kWidth / 8);
                    for (int kb = 0; kb < 4 && ok; kb++) {
                        if (!mbar_wait(&a_ready[kb], g & 1u, ab)) { ok = false; break; }
                        if (kb >= layer_kblocks(l)) continue;
                        tc_fence_after();
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                        if (!mbar_wait(&full_bar[s], ph, ab)) { ok = false; break; }
                        tc_fence_after();
                        const uint32_t sb = smem_u32(stages + (size_t)s * kStageBytes);
                        const int ks_n = min(4, ksteps - kb * 4);
                        for (int ks = 0; ks < ks_n; ks++) {
                            const uint32_t kcol = (uint32_t)(kb * 32 + ks * 8);
                            const uint64_t b_hi = make_b_desc(sb + ks * 32), b_lo = make_b_desc(sb + kHalfStage + ks * 32);
                            mma_tf32_ts(dcol, tmem + kColAlo + kcol, b_hi, kInstrDesc, (kb | ks) != 0);
                            mma_tf32_ts(dcol, tmem + kColAhi + kcol, b_lo, kInstrDesc, 1u);
                            mma_tf32_ts(dcol, tmem + kColAhi + kcol, b_hi, kInstrDesc, 1u);
                        }
                        tc_commit(&empty_bar[s]);          // frees the stage once these MMAs have read it
                        it++;
                    }
                    if (ok) tc_commit(&z_ready);           // accumulator of layer g complete
                }
        }
    } else {
        // ============================================ epilogue warps: thread r <-> pixel row r of the tile <-> TMEM lane r
        const int r = threadIdx.x;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        const int cap_tiles = (int)(a.capacity / kTileRows);
        const size_t act_words = act_tile_words(depth);
        uint32_t g = 0;

        // Split one feature block (32 columns) of `v` into TF32 hi/lo and store it as A operand columns [32 fb, 32 fb + 32)
        // (to_tmem); when a backward pass will follow, also store the fp32 values into the activation image `img` of
        // (slot, row block = this warp).
        auto store_block = [&](const float *v, int fb, bool to_tmem, uint32_t *img) {
            if (to_tmem) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 32; j++) split_tf32(v[j], hi[j], lo[j]);
                tmem_st16(tl + kColAhi + fb * 32, hi);
                tmem_st16(tl + kColAhi + fb * 32 + 16, hi + 16);
                tmem_st16(tl + kColAlo + fb * 32, lo);
                tmem_st16(tl + kColAlo + fb * 32 + 16, lo + 16);
            }
            if (SAVE && img) {
#pragma unroll
                for (int j = 0; j < 32; j++) img[swz(fb * 32 + j, lane)] = __float_as_uint(v[j]);
            }
        };
        auto act_img_ptr = [&](int t, int slot) -> uint32_t * {        // image of (tile t, slot, row block = this warp)
            if (!SAVE || t >= cap_tiles) return nullptr;
            return a.act_img + (size_t)t * act_words + act_slot_block(slot, warp);
        };
        auto publish_encoding = [&](const float *enc, int t) {          // layer-0 operand of tile t, then release all four
            uint32_t *img = act_img_ptr(t, 0);
            store_block(enc, 0, true, img);
            store_block(enc + 32, 1, true, img);
            tmem_wait_st();
            tc_fence_before();
#pragma unroll
            for (int kb = 0; kb < 4; kb++) mbar_arrive(&a_ready[kb]);
        };

        int tile = blockIdx.x;
        long long row = (long long)tile * kTileRows + r;
        bool valid = tile < n_tiles && row < n_fg;
        int pix = valid ? a.fg_index[row] : -1;
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (valid) { const float *q = a.normals + 3ll * pix; nx = q[0]; ny = q[1]; nz = q[2]; }
        if (tile < n_tiles) {
            float enc[kEncPad];
            encode_normal(nx, ny, nz, a.multires, enc);
            publish_encoding(enc, tile);
        }
        for (; tile < n_tiles && ok; tile += gridDim.x) {
            // prefetch the next tile's pixel (index, then normal: two dependent global loads, consumed a whole tile later)
            const int tile_n = tile + gridDim.x;
            const long long row_n = (long long)tile_n * kTileRows + r;
            const bool valid_n = tile_n < n_tiles && row_n < n_fg;
            const int pix_n = valid_n ? a.fg_index[row_n] : -1;
            float nxn = 0.f, nyn = 0.f, nzn = 0.f;
            if (valid_n) { const float *q = a.normals + 3ll * pix_n; nxn = q[0]; nyn = q[1]; nzn = q[2]; }

            for (int l = 0; l < depth; l++, g++) {
                const bool last = (l == depth - 1);
                const uint32_t dcol = tl + kColD + (g & 1u) * kWidth;
                uint32_t *img = act_img_ptr(tile, l + 1);
                if (!last) {
                    if (!mbar_wait(&z_ready, g & 1u, ab)) { ok = false; break; }
                    tc_fence_after();
                    uint32_t z[32];
                    tmem_ld32(dcol, z);
#pragma unroll 1
                    for (int fb = 0; fb < 4; fb++) {
                        float h[32];
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; j++) h[j] = fmaxf(__uint_as_float(z[j]) + s_bias[l * kWidth + fb * 32 + j], 0.f);
                        if (fb < 3) tmem_ld32(dcol + (fb + 1) * 32, z);      // next block's accumulator behind this block's work
                        store_block(h, fb, true, img);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(&a_ready[fb]);
                    }
                } else {
                    float encn[kEncPad];
                    if (tile_n < n_tiles) encode_normal(nxn, nyn, nzn, a.multires, encn);   // while the tensor core is busy
                    if (!mbar_wait(&z_ready, g & 1u, ab)) { ok = false; break; }
                    tc_fence_after();
                    if (tile_n < n_tiles) publish_encoding(encn, tile_n);   // every MMA that read A has completed: start the next tile
                    float acc = 0.f;
                    uint32_t z[32];
                    tmem_ld32(dcol, z);
#pragma unroll 1
                    for (int fb = 0; fb < 4; fb++) {
                        float h[32];
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            h[j] = fmaxf(__uint_as_float(z[j]) + s_bias[l * kWidth + fb * 32 + j], 0.f);
                            acc = fmaf(h[j], s_wout[fb * 32 + j], acc);
                        }
                        if (fb < 3) tmem_ld32(dcol + (fb + 1) * 32, z);
                        if (SAVE) store_block(h, fb, false, img);
                    }
                    if (valid) a.out[pix] = 1.f / (1.f + expf(-(acc + s_bout)));
                    tc_fence_before();      // orders these tcgen05.ld before the a_ready arrivals of the next tile's layer 0
                }
            }
            row = row_n; valid = valid_n; pix = pix_n;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag) atomicOr(a.status, kStatusTimeout);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// =========================================================================================== backward: data gradients
// Same pipeline as k_shadow_fwd, run from the output back to the encoding.  Step s = 0..depth-1 multiplies dZ_l (l = depth-1-s,
// in TMEM as hi/lo) with the transposed weight image of layer l; the epilogue masks the product with [H_l > 0] (read from
// the forward's activation images) to get dZ_{l-1}, writes it back as the next operand and into the dZ images for
// k_shadow_bwd_weights.  The last product is dL/d(encoding); its epilogue applies the positional-encoding Jacobian.
__global__ void __launch_bounds__(kThreads, 1) k_shadow_bwd_data(ShadowDev a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], a_ready[4], z_ready;
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;
    __shared__ float s_wout[kWidth];

    uint8_t *stages = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int kb = 0; kb < 4; kb++) mbar_init(&a_ready[kb], kTileRows);
        mbar_init(&z_ready, 1);
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kWidth; i += kThreads) s_wout[i] = a.W_out[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int n_fg = a.n_fg[0];
    const int cap_tiles = (int)(a.capacity / kTileRows);
    const int n_tiles = min((n_fg + kTileRows - 1) / kTileRows, cap_tiles);
    const int depth = a.depth, img0 = images_total(depth);
    volatile int *ab = &abort_flag;
    bool ok = true;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x)
                for (int img = 0; img < 4 * depth; img++, it++) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    if (!mbar_wait(&empty_bar[s], ph ^ 1u, ab)) { ok = false; break; }
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    tma_bulk_g2s(smem_u32(stages + (size_t)s * kStageBytes),
                                 reinterpret_cast<const uint8_t *>(a.w_images) + (size_t)(img0 + img) * kStageBytes, kStageBytes, &full_bar[s]);
                }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x)
                for (int st = 0; st < depth && ok; st++, g++) {
                    const uint32_t dcol = tmem + kColD + (g & 1u) * kWidth;
                    for (int kb = 0; kb < 4 && ok; kb++, it++) {
                        if (!mbar_wait(&a_ready[kb], g & 1u, ab)) { ok = false; break; }
                        tc_fence_after();
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                        if (!mbar_wait(&full_bar[s], ph, ab)) { ok = false; break; }
                        tc_fence_after();
                        const uint32_t sb = smem_u32(stages + (size_t)s * kStageBytes);
                        for (int ks = 0; ks < 4; ks++) {
                            const uint32_t kcol = (uint32_t)(kb * 32 + ks * 8);
                            const uint64_t b_hi = make_b_desc(sb + ks * 32), b_lo = make_b_desc(sb + kHalfStage + ks * 32);
                            mma_tf32_ts(dcol, tmem + kColAlo + kcol, b_hi, kInstrDesc, (kb | ks) != 0);
                            mma_tf32_ts(dcol, tmem + kColAhi + kcol, b_lo, kInstrDesc, 1u);
                            mma_tf32_ts(dcol, tmem + kColAhi + kcol, b_hi, kInstrDesc, 1u);
                        }
                        tc_commit(&empty_bar[s]);
                    }
                    if (ok) tc_commit(&z_ready);
                }
        }
    } else {
        const int r = threadIdx.x;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        const size_t act_words = act_tile_words(depth), dz_words = dz_tile_words(depth);
        uint32_t g = 0;

        // v (32 features of this row) -> TF32 hi/lo -> A operand columns [32 fb, 32 fb + 32) and the dZ image of `layer`
        auto store_block = [&](const float *v, int fb, int t, int layer) {
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; j++) split_tf32(v[j], hi[j], lo[j]);
            tmem_st16(tl + kColAhi + fb * 32, hi);
            tmem_st16(tl + kColAhi + fb * 32 + 16, hi + 16);
            tmem_st16(tl + kColAlo + fb * 32, lo);
            tmem_st16(tl + kColAlo + fb * 32 + 16, lo + 16);
            uint32_t *img = a.dz_img + (size_t)t * dz_words + (size_t)layer * 16384 + (size_t)warp * 4096;
#pragma unroll
            for (int j = 0; j < 32; j++) img[swz(fb * 32 + j, lane)] = __float_as_uint(v[j]);
        };
        // [H_slot > 0] of this row, features [32 fb, 32 fb + 32), from the forward's activation image
        auto load_mask = [&](int t, int slot, int fb, uint32_t m[32]) {
            const uint32_t *img = a.act_img + (size_t)t * act_words + act_slot_block(slot, warp);
#pragma unroll
            for (int j = 0; j < 32; j++) m[j] = __ldg(img + swz(fb * 32 + j, lane));
        };
        // first operand of a tile: dZ_{depth-1} = dz_out * w_out * [H_depth > 0]; also the dz_out image + its warp sum
        auto publish_first = [&](int t, float dzo) {
            uint32_t *dimg = a.dz_img + (size_t)t * dz_words + (size_t)depth * 16384 + (size_t)warp * 512;
            dimg[lane] = __float_as_uint(dzo);                 // row 0 of a 16-row image: swz(0, lane) = lane
            const float ws = warp_sum(dzo);
            if (lane == 0) a.dzo_sums[(size_t)t * 4 + warp] = ws;
#pragma unroll 1
            for (int fb = 0; fb < 4; fb++) {
                uint32_t m[32];
                float v[32];
                load_mask(t, depth, fb, m);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = (__uint_as_float(m[j]) > 0.f) ? dzo * s_wout[fb * 32 + j] : 0.f;
                store_block(v, fb, t, depth - 1);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&a_ready[fb]);
            }
        };
        auto row_inputs = [&](int t, int &pix, float &dzo, float &nx, float &ny, float &nz) {
            const long long row = (long long)t * kTileRows + r;
            pix = (t < n_tiles && row < n_fg) ? a.fg_index[row] : -1;
            dzo = 0.f; nx = ny = nz = 0.f;
            if (pix >= 0) {
                const float y = a.out[pix];
                dzo = a.g_out[pix] * y * (1.f - y);
                const float *q = a.normals + 3ll * pix;
                nx = q[0]; ny = q[1]; nz = q[2];
            }
        };

        int tile = blockIdx.x, pix;
        float dzo, nx, ny, nz;
        row_inputs(tile, pix, dzo, nx, ny, nz);
        if (tile < n_tiles) publish_first(tile, dzo);
        for (; tile < n_tiles && ok; tile += gridDim.x) {
            const int tile_n = tile + gridDim.x;
            int pix_n;
            float dzo_n, nxn, nyn, nzn;
            row_inputs(tile_n, pix_n, dzo_n, nxn, nyn, nzn);           // consumed a whole tile later

            for (int st = 0; st < depth; st++, g++) {
                const bool last = (st == depth - 1);
                const uint32_t dcol = tl + kColD + (g & 1u) * kWidth;
                const int slot = depth - 1 - st;                        // (not last) the product is dL/dH_slot
                uint32_t m[32];
                if (!last) load_mask(tile, slot, 0, m);                 // in flight while the MMAs of this step run
                if (!mbar_wait(&z_ready, g & 1u, ab)) { ok = false; break; }
                tc_fence_after();
                if (!last) {
#pragma unroll 1
                    for (int fb = 0; fb < 4; fb++) {
                        uint32_t z[32], mn[32];
                        float v[32];
                        tmem_ld32(dcol + fb * 32, z);
                        if (fb < 3) load_mask(tile, slot, fb + 1, mn);  // next block's mask behind this block's work
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = (__uint_as_float(m[j]) > 0.f) ? __uint_as_float(z[j]) : 0.f;
                        store_block(v, fb, tile, slot - 1);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(&a_ready[fb]);
#pragma unroll
                        for (int j = 0; j < 32; j++) m[j] = mn[j];
                    }
                } else {
                    if (tile_n < n_tiles) publish_first(tile_n, dzo_n);     // A is free: start the next tile
                    // dL/d(encoding) of this row -> dL/dnormal = dX[0:3] + sum_k 2^k (cos(2^k n) dXsin_k - sin(2^k n) dXcos_k)
                    uint32_t z0[32], z1[32];
                    tmem_ld32(dcol, z0);
                    tmem_ld32(dcol + 32, z1);
                    tmem_wait_ld();
                    float dx[kEncPad];
#pragma unroll
                    for (int j = 0; j < 32; j++) { dx[j] = __uint_as_float(z0[j]); dx[32 + j] = __uint_as_float(z1[j]); }
                    float gn[3] = {dx[0], dx[1], dx[2]};
                    const float nn[3] = {nx, ny, nz};
                    float sn[3], cs[3];
#pragma unroll
                    for (int k = 0; k < 10; k++) {
                        if (k < a.multires) {
                            const float f = (float)(1 << k);
#pragma unroll
                            for (int j = 0; j < 3; j++) {
                                if (k % 3 == 0) sincosf(nn[j] * f, &sn[j], &cs[j]);          // same octave scheme as encode_normal
                                else { const float s2 = 2.f * sn[j] * cs[j], c2 = 1.f - 2.f * sn[j] * sn[j]; sn[j] = s2; cs[j] = c2; }
                                gn[j] += f * (cs[j] * dx[3 + 6 * k + j] - sn[j] * dx[3 + 6 * k + 3 + j]);
                            }
                        }
                    }
                    const float gx = gn[0], gy = gn[1], gz = gn[2];
                    if (pix >= 0) { float *q = a.g_normals + 3ll * pix; q[0] = gx; q[1] = gy; q[2] = gz; }
                    tc_fence_before();
                }
            }
            pix = pix_n; dzo = dzo_n; nx = nxn; ny = nyn; nz = nzn;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag) atomicOr(a.status, kStatusTimeout);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ========================================================================================= backward: weight gradients
// Job j < depth:  D_j[f, 16 + i] += sum_r dZ_j[r, f] * A_j[r, i]   (A_0 = encoding, A_j = H_j), D_j[f, 0] += sum_r dZ_j[r, f]
// Job depth:      D[f, 16]       += sum_r H_depth[r, f] * dz_out[r]                                  (-> dL/dw_out)
// Both operands of every job are saved fp32 images (K = pixel row), streamed with 1-D TMA; the four otherwise idle epilogue
// warps form the TF32 lo parts in shared memory; 16 constant rows in front of the B image (row 0 = ones) give the bias
// gradients.  One accumulator per job lives in tensor memory for the whole kernel.
constexpr int kB2StageBytes = 32768 + 36864;     // A v 16 KB | A lo 16 KB | [ones 2 KB | B v 16 KB | zeros 2 KB | B lo 16 KB]
constexpr int kB2Stages = 3;
constexpr int kB2BOff = 32768, kB2BLoOff = 32768 + 18432;
constexpr int kPartialCols = 512;

__host__ __device__ __forceinline__ int job_cols(int job, int depth) { return job == 0 ? 80 : (job < depth ? 144 : 32); }
__host__ __device__ __forceinline__ int job_col0(int job, int depth) { return job == 0 ? 0 : 80 + (job - 1) * 144; }
__global__ void __launch_bounds__(kThreads, 1) k_shadow_bwd_weights(ShadowDev a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kB2Stages], split_bar[kB2Stages], empty_bar[kB2Stages], done_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;

    uint8_t *stages = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < kB2Stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&split_bar[s], kTileRows); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the constant rows in front of every B image: hi part = [ones; 15 zero rows], lo part = zeros
    for (int s = 0; s < kB2Stages; s++) {
        float *bh = reinterpret_cast<float *>(stages + (size_t)s * kB2StageBytes + kB2BOff);
        float *bl = reinterpret_cast<float *>(stages + (size_t)s * kB2StageBytes + kB2BLoOff);
        for (int i = threadIdx.x; i < 512; i += kThreads) { bh[i] = (i < 32) ? 1.f : 0.f; bl[i] = 0.f; }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int n_fg = a.n_fg[0];
    const int cap_tiles = (int)(a.capacity / kTileRows);
    const int n_tiles = min((n_fg + kTileRows - 1) / kTileRows, cap_tiles);
    const int depth = a.depth, n_jobs = depth + 1;
    const size_t act_words = act_tile_words(depth), dz_words = dz_tile_words(depth);
    const bool has_work = (int)blockIdx.x < n_tiles;
    volatile int *ab = &abort_flag;
    bool ok = true;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
                const uint32_t *act = a.act_img + (size_t)tile * act_words, *dz = a.dz_img + (size_t)tile * dz_words;
                for (int job = 0; job < n_jobs && ok; job++)
                    for (int kb = 0; kb < 4; kb++, it++) {
                        const uint32_t s = it % kB2Stages, ph = (it / kB2Stages) & 1u;
                        if (!mbar_wait(&empty_bar[s], ph ^ 1u, ab)) { ok = false; break; }
                        const uint32_t *A = (job < depth) ? dz + (size_t)job * 16384 + (size_t)kb * 4096
                                                          : act + act_slot_block(depth, kb);
                        const uint32_t *B;
                        uint32_t b_words;                    // words of the B image
                        if (job < depth) { B = act + act_slot_block(job, kb); b_words = (uint32_t)act_slot_words(job); }
                        else { B = dz + (size_t)depth * 16384 + (size_t)kb * 512; b_words = 512; }
                        const uint32_t sb = smem_u32(stages + (size_t)s * kB2StageBytes);
                        mbar_expect_tx(&full_bar[s], 16384 + 4 * b_words);
                        tma_bulk_g2s(sb, A, 16384, &full_bar[s]);
                        tma_bulk_g2s(sb + kB2BOff + 2048, B, 4 * b_words, &full_bar[s]);
                    }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            uint32_t it = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, first = false)
                for (int job = 0; job < n_jobs && ok; job++) {
                    const uint32_t dcol = tmem + (uint32_t)job_col0(job, depth);
                    const uint32_t idesc = instr_desc_n(job_cols(job, depth));
                    for (int kb = 0; kb < 4; kb++, it++) {
                        const uint32_t s = it % kB2Stages, ph = (it / kB2Stages) & 1u;
                        if (!mbar_wait(&split_bar[s], ph, ab)) { ok = false; break; }      // images landed AND lo parts formed
                        tc_fence_after();
                        const uint32_t sb = smem_u32(stages + (size_t)s * kB2StageBytes);
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t a_hi = make_b_desc(sb + ks * 32), a_lo = make_b_desc(sb + kHalfStage + ks * 32);
                            const uint64_t b_hi = make_b_desc(sb + kB2BOff + ks * 32), b_lo = make_b_desc(sb + kB2BLoOff + ks * 32);
                            mma_tf32_ss(dcol, a_lo, b_hi, idesc, !(first && kb == 0 && ks == 0));
                            mma_tf32_ss(dcol, a_hi, b_lo, idesc, 1u);
                            mma_tf32_ss(dcol, a_hi, b_hi, idesc, 1u);
                        }
                        tc_commit(&empty_bar[s]);
                    }
                }
            if (ok && has_work) tc_commit(&done_bar);
        }
    } else {
        // splitters: the TMA delivered fp32 images, which the tensor core reads as their own TF32 "hi" part (it ignores the
        // 13 low mantissa bits); lo = v - trunc(v) is formed here, element for element at the same swizzled position, in
        // the lo halves of the stage (exact: <= 13 significant bits)
        {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x)
                for (int job = 0; job < n_jobs && ok; job++) {
                    const int b_vec = (job < depth ? (int)act_slot_words(job) : 512) / 4;      // float4 of the B image
                    for (int kb = 0; kb < 4; kb++, it++) {
                        const uint32_t s = it % kB2Stages, ph = (it / kB2Stages) & 1u;
                        if (!mbar_wait(&full_bar[s], ph, ab)) { ok = false; break; }
                        uint8_t *st = stages + (size_t)s * kB2StageBytes;
                        const uint4 *av = reinterpret_cast<const uint4 *>(st);
                        uint4 *al = reinterpret_cast<uint4 *>(st + kHalfStage);
                        const uint4 *bv = reinterpret_cast<const uint4 *>(st + kB2BOff + 2048);
                        uint4 *bl = reinterpret_cast<uint4 *>(st + kB2BLoOff + 2048);
                        auto lo_of = [](uint32_t v) { return __float_as_uint(__uint_as_float(v) - __uint_as_float(v & 0xFFFFE000u)); };
#pragma unroll
                        for (int i = 0; i < 8; i++) {                                           // 1024 float4 of A
                            const uint4 v = av[i * kTileRows + threadIdx.x];
                            al[i * kTileRows + threadIdx.x] = make_uint4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
                        }
                        for (int i = threadIdx.x; i < b_vec; i += kTileRows) {
                            const uint4 v = bv[i];
                            bl[i] = make_uint4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive(&split_bar[s]);
                    }
                }
        }
        // dump the accumulators: partials[cta][column][row f], zeros for a CTA that had no tile
        const int f = threadIdx.x;
        float *dst = a.partials + (size_t)blockIdx.x * kPartialCols * kTileRows + f;
        const int total = job_col0(depth, depth) + job_cols(depth, depth);
        if (has_work) ok = mbar_wait(&done_bar, 0u, ab);
        tc_fence_after();
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < total; c0 += 32) {
            uint32_t z[32];
            if (has_work && ok) { tmem_ld32(tl + c0, z); tmem_wait_ld(); }
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (c0 + j < total) dst[(size_t)(c0 + j) * kTileRows] = (has_work && ok) ? __uint_as_float(z[j]) : 0.f;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag) atomicOr(a.status, kStatusTimeout);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// sum the per-CTA partials in CTA order (deterministic) and scatter them into the gradient arrays; block 0 also sums dz_out
__global__ void __launch_bounds__(128) k_shadow_bwd_reduce(ShadowDev a) {
    const int depth = a.depth, f = threadIdx.x;
    // blockIdx.x enumerates the accumulator columns that matter
    int col = -1;
    float *out = nullptr;
    int b = blockIdx.x;
    if (b < 1 + a.enc) {                                             // job 0: column 0 = db_in, 16 + e = dW_in[:, e]
        col = (b == 0) ? 0 : 16 + (b - 1);
        out = (b == 0) ? a.g_b_in + f : a.g_W_in + (size_t)f * a.enc + (b - 1);
    } else {
        b -= 1 + a.enc;
        if (b < (depth - 1) * 129) {                                 // job l: column 0 = db, 16 + i = dW_l[:, i]
            const int l = 1 + b / 129, c = b % 129;
            col = job_col0(l, depth) + ((c == 0) ? 0 : 16 + (c - 1));
            out = (c == 0) ? a.g_b_hid + (size_t)(l - 1) * kWidth + f
                           : a.g_W_hid + (size_t)(l - 1) * kWidth * kWidth + (size_t)f * kWidth + (c - 1);
        } else {                                                     // output job: column 16 = dw_out
            col = job_col0(depth, depth) + 16;
            out = a.g_w_out + f;
        }
    }
    float s = 0.f;
    for (int c = 0; c < a.n_ctas; c++) s += a.partials[((size_t)c * kPartialCols + col) * kTileRows + f];
    *out = s;
    if (blockIdx.x == 0 && f < 32) {                                 // db_out = sum of dz_out, fixed order
        const int n_fg = a.n_fg[0];
        const int n = min((n_fg + kTileRows - 1) / kTileRows, (int)(a.capacity / kTileRows)) * 4;
        float t = 0.f;
        for (int i = f; i < n; i += 32) t += a.dzo_sums[i];
        t = warp_sum(t);
        if (f == 0) a.g_b_out[0] = t;
    }
}

}  // namespace

extern "C" size_t gom_sizeof_shadow_mlp_args(void) { return sizeof(GomShadowMlpArgs); }
extern "C" size_t gom_shadow_mlp_weight_image_bytes(int depth) {
    return (depth >= 1 && depth <= kMaxDepth) ? (size_t)images_total_with_backward(depth) * kStageBytes : 0;
}
extern "C" size_t gom_shadow_mlp_tile_words(int depth, int which) {
    if (depth < 1 || depth > kMaxDepth) return 0;
    return which == 0 ? act_tile_words(depth) : dz_tile_words(depth);
}
extern "C" size_t gom_shadow_mlp_partial_floats(void) { return (size_t)kPartialCols * kTileRows; }

static int g_sm_count = 0;
static int shadow_setup(void) {
    if (g_sm_count) return GOM_OK;
    int dev = 0, sms = 0;
    GOM_CUDA(cudaGetDevice(&dev));
    GOM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int dyn = kStages * kStageBytes + 1024, dyn2 = kB2Stages * kB2StageBytes + 1024;
    GOM_CUDA(cudaFuncSetAttribute(k_shadow_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    GOM_CUDA(cudaFuncSetAttribute(k_shadow_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    GOM_CUDA(cudaFuncSetAttribute(k_shadow_bwd_data, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    GOM_CUDA(cudaFuncSetAttribute(k_shadow_bwd_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn2));
    g_sm_count = sms;
    return GOM_OK;
}
extern "C" int gom_shadow_mlp_num_ctas(void) { return shadow_setup() == GOM_OK ? g_sm_count : 0; }

static int shadow_fill(const GomShadowMlpArgs *p, ShadowDev &d, bool backward) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_pixels > 0 && p->n_pixels < (1ll << 31) - 4096, "n_pixels");
    GOM_REQUIRE(p->width == kWidth, "width must be 128");
    GOM_REQUIRE(p->depth >= 1 && p->depth <= kMaxDepth, "depth must be in [1, 8]");
    GOM_REQUIRE(p->multires >= 0 && p->multires <= 10, "multires must be in [0, 10]");
    GOM_REQUIRE(p->normals && p->W_in && p->b_in && p->W_out && p->b_out, "null input");
    GOM_REQUIRE(p->depth == 1 || (p->W_hid && p->b_hid), "null hidden weights");
    GOM_REQUIRE(p->block_count && p->fg_index && p->n_fg && p->w_images && p->bg_value && p->out && p->status, "null buffer");
    if (p->save_hidden || backward) {
        GOM_REQUIRE(p->depth <= 3, "the backward pass supports depth <= 3 (tensor-memory columns of the weight-gradient accumulators)");
        GOM_REQUIRE(p->act_img && p->capacity >= kTileRows && p->capacity % kTileRows == 0, "save_hidden needs act_img and a capacity that is a multiple of 128");
    }
    if (backward)
        GOM_REQUIRE(p->g_out && p->dz_img && p->g_normals && p->dzo_sums && p->partials && p->g_W_in && p->g_b_in && p->g_w_out &&
                    p->g_b_out && (p->depth == 1 || (p->g_W_hid && p->g_b_hid)), "null backward buffer");
    GOM_REQUIRE((reinterpret_cast<uintptr_t>(p->normals) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w_images) & 127) == 0 &&
                (reinterpret_cast<uintptr_t>(p->act_img) & 127) == 0 && (reinterpret_cast<uintptr_t>(p->dz_img) & 127) == 0,
                "normals must be 16-byte, w_images / act_img / dz_img 128-byte aligned");
    if (int rc = shadow_setup()) return rc;
    d.n_pixels = p->n_pixels; d.capacity = p->capacity;
    d.multires = p->multires; d.enc = 3 + 6 * p->multires; d.depth = p->depth; d.save_hidden = p->save_hidden; d.n_ctas = g_sm_count;
    d.normals = p->normals;
    d.W_in = p->W_in; d.b_in = p->b_in; d.W_hid = p->W_hid; d.b_hid = p->b_hid; d.W_out = p->W_out; d.b_out = p->b_out;
    d.block_count = p->block_count; d.fg_index = p->fg_index; d.n_fg = p->n_fg;
    d.w_images = p->w_images; d.bg_value = p->bg_value; d.out = p->out; d.status = p->status;
    d.act_img = reinterpret_cast<uint32_t *>(p->act_img); d.dz_img = reinterpret_cast<uint32_t *>(p->dz_img);
    d.g_out = p->g_out; d.g_normals = p->g_normals; d.dzo_sums = p->dzo_sums; d.partials = p->partials;
    d.g_W_in = p->g_W_in; d.g_b_in = p->g_b_in; d.g_W_hid = p->g_W_hid; d.g_b_hid = p->g_b_hid; d.g_w_out = p->g_w_out; d.g_b_out = p->g_b_out;
    return GOM_OK;
}

extern "C" int gom_shadow_mlp_forward(const GomShadowMlpArgs *p, gom_stream_t stream_) {
    ShadowDev d;
    if (int rc = shadow_fill(p, d, false)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t dyn_smem = (size_t)kStages * kStageBytes + 1024;
    const int nblk = gom_div_up(p->n_pixels, 1024);
    const int n_img = p->save_hidden ? images_total_with_backward(p->depth) : images_total(p->depth);

    gom_prof_begin(GOM_PROF_SHADOW_COMPACT, stream);
    k_shadow_prep<<<gom_div_up((int64_t)n_img * 4096, 256) + 1, 256, 0, stream>>>(d);
    GOM_LAUNCH_CHECK();
    k_shadow_count<<<nblk, 256, 0, stream>>>(d);
    GOM_LAUNCH_CHECK();
    k_shadow_scan<<<1, 1024, 0, stream>>>(d, nblk);
    GOM_LAUNCH_CHECK();
    k_shadow_scatter<<<nblk, 256, 0, stream>>>(d);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_SHADOW_COMPACT, stream);

    gom_prof_begin(GOM_PROF_SHADOW_FWD, stream);
    if (p->save_hidden) k_shadow_fwd<true><<<g_sm_count, kThreads, dyn_smem, stream>>>(d);
    else k_shadow_fwd<false><<<g_sm_count, kThreads, dyn_smem, stream>>>(d);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_SHADOW_FWD, stream);
    return GOM_OK;
}

// Backward of a forward call made with save_hidden = 1 on the SAME buffers (fg_index, n_fg, out, w_images, act_img).
extern "C" int gom_shadow_mlp_backward(const GomShadowMlpArgs *p, gom_stream_t stream_) {
    ShadowDev d;
    if (int rc = shadow_fill(p, d, true)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    gom_prof_begin(GOM_PROF_SHADOW_BWD_DATA, stream);
    k_shadow_bwd_data<<<g_sm_count, kThreads, (size_t)kStages * kStageBytes + 1024, stream>>>(d);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_SHADOW_BWD_DATA, stream);
    gom_prof_begin(GOM_PROF_SHADOW_BWD_WEIGHTS, stream);
    k_shadow_bwd_weights<<<g_sm_count, kThreads, (size_t)kB2Stages * kB2StageBytes + 1024, stream>>>(d);
    GOM_LAUNCH_CHECK();
    k_shadow_bwd_reduce<<<1 + d.enc + (d.depth - 1) * 129 + 1, 128, 0, stream>>>(d);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_SHADOW_BWD_WEIGHTS, stream);
    return GOM_OK;
}

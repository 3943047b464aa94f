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
//                  registers.  When the caller wants a backward pass the post-ReLU activations are written
//                  feature-major ([layer][128][capacity], coalesced) for it.
//
// TMEM map (512 columns allocated): [0,128) accumulator D, [128,256) A_hi, [256,384) A_lo; lane = pixel row of the tile.
// Warp roles (192 threads): warps 0-3 epilogue (warp w owns TMEM lanes 32w..32w+31), warp 4 weight producer (TMA),
// warp 5 MMA issuer.  Per tile and layer the chain a_ready -> MMAs -> z_ready -> epilogue is serial; the weight stream
// (6 x 32 KB stages) runs ahead of it.
#include "gom_common.cuh"

namespace {

constexpr int kWidth = 128;             // hidden width (the only one the reference configs use)
constexpr int kTileRows = 128;          // pixels per tile = UMMA M
constexpr int kStageBytes = 32768;      // one K-block of 32 columns: [hi 128x32 fp32 | lo 128x32 fp32]
constexpr int kHalfStage = 16384;
constexpr int kStages = 6;
constexpr int kMaxDepth = 8;
constexpr int kEncPad = 64;             // encoding columns in TMEM (3 + 6*multires <= 63)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 0, kColAhi = 128, kColAlo = 256;
constexpr int kThreads = 192;
constexpr uint32_t kStatusTimeout = 2u;

struct ShadowDev {
    long long n_pixels, capacity;
    int multires, enc, depth, save_hidden;
    const float *normals;
    const float *W_in, *b_in, *W_hid, *b_hid, *W_out, *b_out;
    uint32_t *block_count;
    int *fg_index, *n_fg;
    float *w_images, *bg_value, *out, *hidden;
    uint32_t *status;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

// number of 32-column K-blocks of layer l and the images before it
__host__ __device__ __forceinline__ int layer_kblocks(int l) { return l == 0 ? 2 : 4; }
__host__ __device__ __forceinline__ int images_total(int depth) { return 2 + 4 * (depth - 1); }

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
    if (e >= n_img * 4096) return;
    const int img = e >> 12, n = (e >> 5) & 127, kl = e & 31;
    int l, kb;
    if (img < 2) { l = 0; kb = img; } else { l = 1 + (img - 2) / 4; kb = (img - 2) % 4; }
    const int k = kb * 32 + kl;
    float w = 0.f;
    if (l == 0) { if (k < a.enc) w = a.W_in[n * a.enc + k]; }
    else w = a.W_hid[(size_t)(l - 1) * kWidth * kWidth + (size_t)n * kWidth + k];
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

// ------------------------------------------------------------------------------------- mbarrier / TMA / tcgen05 PTX
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error must end the kernel with a status bit instead of hanging the GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, volatile int *abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    for (uint32_t spin = 1;; ++spin) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spin & 255u) == 0u) {
            if (*abort_flag) return false;
            if (clock64() - t0 > 400000000ll) { *abort_flag = 1; return false; }
        }
    }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128, N = 128, K = 8
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1): start address >> 4,
// LBO (ignored for swizzled K-major) = 1, SBO = 1024 B between 8-row groups, layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kWidth >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ the MLP kernel
__global__ void __launch_bounds__(kThreads, 1) k_shadow_fwd(ShadowDev a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], a_ready, z_ready;
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
        mbar_init(&a_ready, kTileRows);
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
            uint32_t it = 0, a_cnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x)
                for (int l = 0; l < depth && ok; l++) {
                    if (!mbar_wait(&a_ready, a_cnt & 1u, ab)) { ok = false; break; }
                    a_cnt++;
                    tc_fence_after();
                    const int ksteps = (l == 0) ? ksteps0 : (kWidth / 8);
                    for (int kb = 0; kb < layer_kblocks(l); kb++, it++) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                        if (!mbar_wait(&full_bar[s], ph, ab)) { ok = false; break; }
                        tc_fence_after();
                        const uint32_t sb = smem_u32(stages + (size_t)s * kStageBytes);
                        const int ks_n = min(4, ksteps - kb * 4);
                        for (int ks = 0; ks < ks_n; ks++) {
                            const uint32_t kcol = (uint32_t)(kb * 32 + ks * 8);
                            const uint64_t b_hi = make_b_desc(sb + ks * 32), b_lo = make_b_desc(sb + kHalfStage + ks * 32);
                            mma_tf32_ts(tmem + kColD, tmem + kColAlo + kcol, b_hi, kInstrDesc, (kb | ks) != 0);
                            mma_tf32_ts(tmem + kColD, tmem + kColAhi + kcol, b_lo, kInstrDesc, 1u);
                            mma_tf32_ts(tmem + kColD, tmem + kColAhi + kcol, b_hi, kInstrDesc, 1u);
                        }
                        tc_commit(&empty_bar[s]);          // frees the stage once these MMAs have read it
                    }
                    if (ok) tc_commit(&z_ready);           // accumulator complete
                }
        }
    } else {
        // ============================================ epilogue warps: thread r <-> pixel row r of the tile <-> TMEM lane r
        const int r = threadIdx.x;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t z_cnt = 0;
        for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
            const long long row = (long long)tile * kTileRows + r;
            const bool valid = row < n_fg;
            const int pix = valid ? a.fg_index[row] : -1;
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (valid) { const float *q = a.normals + 3ll * pix; nx = q[0]; ny = q[1]; nz = q[2]; }
            {   // positional encoding: [n, sin(2^k n), cos(2^k n)]_k, zero padded to 64 columns
                float enc[kEncPad];
#pragma unroll
                for (int c = 0; c < kEncPad; c++) enc[c] = 0.f;
                enc[0] = nx; enc[1] = ny; enc[2] = nz;
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    if (k < a.multires) {
                        const float f = (float)(1 << k);
                        float s, c;
                        sincosf(nx * f, &s, &c); enc[3 + 6 * k + 0] = s; enc[3 + 6 * k + 3] = c;
                        sincosf(ny * f, &s, &c); enc[3 + 6 * k + 1] = s; enc[3 + 6 * k + 4] = c;
                        sincosf(nz * f, &s, &c); enc[3 + 6 * k + 2] = s; enc[3 + 6 * k + 5] = c;
                    }
                }
#pragma unroll
                for (int c = 0; c < kEncPad / 16; c++) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) split_tf32(enc[c * 16 + j], hi[j], lo[j]);
                    tmem_st16(tl + kColAhi + c * 16, hi);
                    tmem_st16(tl + kColAlo + c * 16, lo);
                }
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&a_ready);

            for (int l = 0; l < depth; l++) {
                if (!mbar_wait(&z_ready, z_cnt & 1u, ab)) { ok = false; break; }
                z_cnt++;
                tc_fence_after();
                const bool last = (l == depth - 1);
                const bool save = a.save_hidden && valid && row < a.capacity;
                float *hsave = save ? a.hidden + (size_t)l * kWidth * (size_t)a.capacity + (size_t)row : nullptr;
                float acc = 0.f;
#pragma unroll 1
                for (int c = 0; c < kWidth / 16; c++) {
                    uint32_t z[16], hi[16], lo[16];
                    tmem_ld16(tl + kColD + c * 16, z);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float h = fmaxf(__uint_as_float(z[j]) + s_bias[l * kWidth + c * 16 + j], 0.f);
                        if (save) hsave[(size_t)(c * 16 + j) * (size_t)a.capacity] = h;
                        acc = fmaf(h, s_wout[c * 16 + j], acc);
                        split_tf32(h, hi[j], lo[j]);
                    }
                    if (!last) {
                        tmem_st16(tl + kColAhi + c * 16, hi);
                        tmem_st16(tl + kColAlo + c * 16, lo);
                    }
                }
                if (!last) {
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(&a_ready);
                } else if (valid) {
                    a.out[pix] = 1.f / (1.f + expf(-(acc + s_bout)));
                }
            }
            tc_fence_before();      // orders this tile's tcgen05.ld before the next tile's a_ready arrive
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag) atomicOr(a.status, kStatusTimeout);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

extern "C" size_t gom_sizeof_shadow_mlp_args(void) { return sizeof(GomShadowMlpArgs); }
extern "C" size_t gom_shadow_mlp_weight_image_bytes(int depth) {
    return (depth >= 1 && depth <= kMaxDepth) ? (size_t)images_total(depth) * kStageBytes : 0;
}

extern "C" int gom_shadow_mlp_forward(const GomShadowMlpArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_pixels > 0 && p->n_pixels < (1ll << 31) - 4096, "n_pixels");
    GOM_REQUIRE(p->width == kWidth, "width must be 128");
    GOM_REQUIRE(p->depth >= 1 && p->depth <= kMaxDepth, "depth must be in [1, 8]");
    GOM_REQUIRE(p->multires >= 0 && p->multires <= 10, "multires must be in [0, 10]");
    GOM_REQUIRE(p->normals && p->W_in && p->b_in && p->W_out && p->b_out, "null input");
    GOM_REQUIRE(p->depth == 1 || (p->W_hid && p->b_hid), "null hidden weights");
    GOM_REQUIRE(p->block_count && p->fg_index && p->n_fg && p->w_images && p->bg_value && p->out && p->status, "null buffer");
    GOM_REQUIRE(!p->save_hidden || (p->hidden && p->capacity > 0), "save_hidden needs hidden and capacity");
    GOM_REQUIRE((reinterpret_cast<uintptr_t>(p->normals) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w_images) & 127) == 0,
                "normals must be 16-byte and w_images 128-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;

    ShadowDev d;
    d.n_pixels = p->n_pixels; d.capacity = p->capacity;
    d.multires = p->multires; d.enc = 3 + 6 * p->multires; d.depth = p->depth; d.save_hidden = p->save_hidden;
    d.normals = p->normals;
    d.W_in = p->W_in; d.b_in = p->b_in; d.W_hid = p->W_hid; d.b_hid = p->b_hid; d.W_out = p->W_out; d.b_out = p->b_out;
    d.block_count = p->block_count; d.fg_index = p->fg_index; d.n_fg = p->n_fg;
    d.w_images = p->w_images; d.bg_value = p->bg_value; d.out = p->out; d.hidden = p->hidden; d.status = p->status;

    static int sm_count = 0;
    const size_t dyn_smem = (size_t)kStages * kStageBytes + 1024;
    if (sm_count == 0) {
        int dev = 0;
        GOM_CUDA(cudaGetDevice(&dev));
        GOM_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        GOM_CUDA(cudaFuncSetAttribute(k_shadow_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    }
    const int nblk = gom_div_up(p->n_pixels, 1024);
    const int n_img = images_total(p->depth);

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
    k_shadow_fwd<<<sm_count, kThreads, dyn_smem, stream>>>(d);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_SHADOW_FWD, stream);
    return GOM_OK;
}

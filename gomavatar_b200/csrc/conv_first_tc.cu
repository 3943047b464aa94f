// conv_first_tc.cu — the first VGG16 convolution of LPIPS (3 -> 64, 3x3, pad 1, + bias + ReLU) and its input gradient as
// tcgen05 GEMMs over pixel rows, behind the same entry points as the FFMA kernels of conv_first.cu
// (gom_conv_first_forward / gom_conv_first_backward with use_tensor_cores = 1).
//
// Both are far too thin for the tensor core to matter (K = 27 resp. N = 27): the point is to take the 1728 FMAs per
// pixel off the FP32 pipe, which bound the FFMA kernels at 0.38 ms each, so that the kernels run at the speed of their
// HBM traffic (forward: write 256 B per pixel; backward: read 256 B per pixel).  Products are formed as 3xTF32
// (lo*hi + hi*lo + hi*hi, fp32 accumulate), i.e. with fp32-GEMM accuracy like the FFMA kernels they replace.
//
//   forward   out[p, k]  = relu( sum_{r,s,c} x[p + (r-1, s-1), c] W[k,c,r,s] + b[k] )
//             A[p, (r,s,c)] = the 27 neighbourhood values of pixel p (zero padded, gathered by the pixel's thread straight
//             into tensor memory), B[k, (r,s,c)] = W, one 128 x 64 x 32 product per 128 pixels.
//   backward  T[p, (r,s,c)] = sum_k dY[p,k] W[k,c,r,s]      (128 x 32 x 64 product, A = the pixel's 64 gradients)
//             dX[q, c] = sum_{r,s} T[q - (r-1, s-1), (r,s,c)]: the sum over s (neighbours in the image row = neighbouring
//             lanes) is taken in the GEMM epilogue with shuffles, which writes 3 row-sum planes H[r] (48 B per pixel) instead
//             of the 9 tap planes (144 B); k_conv1_stencil takes the sum over r: 3 coalesced float4 loads per pixel
//
// One persistent CTA per SM running several independent producer/MMA/epilogue groups (see k_conv1_gemm).
// The weight images (TF32 hi/lo, K-major SWIZZLE_128B) are built by each CTA in its own shared memory at start-up.
#include "gom_common.cuh"
#include "gom_tcgen05.cuh"

namespace {

using namespace gomtc;

constexpr uint32_t kTmemCols = 512;

struct Conv1Dev {
    int N, H, W;
    long long rows;                      // N*H*W
    const float *x, *weight, *bias, *dY, *act;
    float *out, *T, *dX;
    uint32_t *status;
    uint32_t *mask_out;                  // forward: ReLU bit mask of out, [rows, 2] (nullable)
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// pixel index -> (image, row, column); 32-bit divisions whenever the index space allows (a 64-bit division is ~100 instructions)
__device__ __forceinline__ void pixel_coords(long long p, long long rows, int H, int W, long long &n, int &y, int &x) {
    if (rows < (1ll << 31)) {
        const uint32_t hw = (uint32_t)H * (uint32_t)W, pu = (uint32_t)p, nu = pu / hw, yx = pu - nu * hw, yu = yx / (uint32_t)W;
        n = nu; y = (int)yu; x = (int)(yx - yu * (uint32_t)W);
    } else {
        const long long HW = (long long)H * W, yx = p - (p / HW) * HW;
        n = p / HW; y = (int)(yx / W); x = (int)(yx - (long long)y * W);
    }
}
__device__ __forceinline__ int swz(int j, int q) { return j * 32 + ((((q >> 2) ^ (j & 7)) << 2) | (q & 3)); }

// MODE 0: forward (K = 32 = one k-block, N = 64).  MODE 1: backward GEMM (K = 64 = two k-blocks, N = 32).
// A CTA runs G independent groups (4 producer/epilogue warps + 1 MMA warp each, one tile in flight per group, own TMEM
// columns and barriers): the per-tile chain gather -> split -> tcgen05.st -> MMA -> tcgen05.ld -> store is latency bound,
// so several chains per SM are interleaved instead of deepening one.  Warps 0..4G-1 are the producers (group = warp / 4;
// warp % 4 is the TMEM lane quarter a warp may access), warps 4G..5G-1 the MMA issuers.
template <int MODE> struct Conv1Cfg {
    static constexpr int KB = MODE == 0 ? 1 : 2;          // k-blocks of 32 columns
    static constexpr int NOUT = MODE == 0 ? 64 : 32;      // accumulator columns
    static constexpr int G = MODE == 0 ? 4 : 3;           // groups per CTA: G * (NOUT + 2 * KB * 32) <= 512 TMEM columns
    static constexpr int COLS = NOUT + 2 * KB * 32;
    static constexpr int THREADS = G * 160;
    static constexpr int STAGE = MODE == 0 ? 512 : 1024;  // float4 of staging per producer warp (backward: gradient rows + ReLU rows)
    static constexpr int DYN_SMEM = G * 4 * STAGE * 16;
};

template <int MODE>
__global__ void __launch_bounds__(Conv1Cfg<MODE>::THREADS, 1) k_conv1_gemm(Conv1Dev a) {
    using Cfg = Conv1Cfg<MODE>;
    constexpr int KB = Cfg::KB, NOUT = Cfg::NOUT, G = Cfg::G, kThreads = Cfg::THREADS;
    __shared__ __align__(1024) uint32_t s_b[KB][2][NOUT * 32];    // [k-block][hi|lo][row n][32 swizzled words]
    extern __shared__ float4 s_stage[];                           // [producer warp][32 rows][16 float4]: row <-> lane transposition
    __shared__ uint64_t a_ready[G], z_ready[G];
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;
    __shared__ __align__(16) float s_bias[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int i = 0; i < G; i++) { mbar_init(&a_ready[i], kTileRows); mbar_init(&z_ready[i], 1); }
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // weight images.  forward: row n = output channel k, column kk = (r*3+s)*3+c.  backward: row n = (r*3+s)*3+c, column = k.
    for (int e = threadIdx.x; e < KB * NOUT * 32; e += kThreads) {
        const int kb = e / (NOUT * 32), n = (e / 32) % NOUT, kl = e % 32;
        float w = 0.f;
        if (MODE == 0) {
            if (kl < 27) { const int tap = kl / 3, c = kl % 3; w = a.weight[(n * 3 + c) * 9 + tap]; }
        } else {
            if (n < 27) { const int tap = n / 3, c = n % 3, k = kb * 32 + kl; w = a.weight[(k * 3 + c) * 9 + tap]; }
        }
        uint32_t hi, lo;
        split_tf32(w, hi, lo);
        s_b[kb][0][swz(n, kl)] = hi;
        s_b[kb][1][swz(n, kl)] = lo;
    }
    if (MODE == 0) for (int i = threadIdx.x; i < 64; i += kThreads) s_bias[i] = a.bias[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gom_pdl_trigger();            // programmatic dependent launch (gom_common.cuh): the set-up above only read weights
    gom_pdl_wait();
    const uint32_t tmem = tmem_slot;
    const long long n_tiles = (a.rows + kTileRows - 1) / kTileRows;
    const long long stride = (long long)gridDim.x * G;
    volatile int *ab = &abort_flag;
    bool ok = true;

    if (warp >= 4 * G) {
        const int g = warp - 4 * G;
        if (lane == 0) {
            const uint32_t idesc = instr_desc_n(NOUT);
            const uint32_t d = tmem + g * Cfg::COLS, a_hi = d + NOUT, a_lo = a_hi + KB * 32;
            uint32_t it = 0;
            for (long long tile = (long long)blockIdx.x * G + g; tile < n_tiles && ok; tile += stride, it++) {
                if (!mbar_wait(&a_ready[g], it & 1u, ab)) { ok = false; break; }
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < KB; kb++)
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint32_t kcol = (uint32_t)(kb * 32 + ks * 8);
                        const uint64_t b_hi = make_b_desc(smem_u32(&s_b[kb][0][0]) + ks * 32);
                        const uint64_t b_lo = make_b_desc(smem_u32(&s_b[kb][1][0]) + ks * 32);
                        mma_tf32_ts(d, a_lo + kcol, b_hi, idesc, (kb | ks) != 0);
                        mma_tf32_ts(d, a_hi + kcol, b_lo, idesc, 1u);
                        mma_tf32_ts(d, a_hi + kcol, b_hi, idesc, 1u);
                    }
                tc_commit(&z_ready[g]);
            }
        }
    } else {
        const int g = warp >> 2, quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t d = tmem + ((uint32_t)(quarter * 32) << 16) + g * Cfg::COLS, a_hi = d + NOUT, a_lo = a_hi + KB * 32;
        const long long HW = (long long)a.H * a.W;
        // Global rows of 64 floats are read / written by the WARP, 512 contiguous bytes per instruction (lane l: float4
        // l & 15 of row 2 i + (l >> 4)), and cross to / from the row-per-thread view through this warp's 8 KB of shared
        // memory; float4 j of row q sits at column j ^ (q & 15), so both views are bank-conflict free per quarter warp.
        // (One thread writing its own 256-byte row costs 32 half-filled sectors per instruction: 2.5 TB/s instead of 6.)
        float4 *stg = s_stage + (size_t)warp * Cfg::STAGE;
        uint32_t it = 0;
        for (long long tile = (long long)blockIdx.x * G + g; tile < n_tiles && ok; tile += stride, it++) {
            const long long p = tile * kTileRows + r;
            const long long p_warp = tile * kTileRows + quarter * 32;       // first row of this warp
            {   // this thread's row of the A operand -> tensor memory
                float v[KB * 32];
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = 0.f;
                    if (p < a.rows) {
                        long long n; int y, x;
                        pixel_coords(p, a.rows, a.H, a.W, n, y, x);
                        const float *img = a.x + n * HW * 3;
#pragma unroll
                        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                            for (int dx = -1; dx <= 1; dx++) {
                                const int yy = y + dy, xx = x + dx;
                                if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
                                    const float *q = img + ((long long)yy * a.W + xx) * 3;
                                    const int t = ((dy + 1) * 3 + (dx + 1)) * 3;
                                    v[t] = __ldg(q); v[t + 1] = __ldg(q + 1); v[t + 2] = __ldg(q + 2);
                                }
                            }
                    }
                } else {
                    // the warp's 32 gradient rows (and, for the fused ReLU backward, the 32 rows of this convolution's own
                    // output) go global -> shared memory with cp.async: no registers are tied up by the 32 loads in flight.
                    // Without the ReLU rows the second half of the staging area is free: the rows of the NEXT tile are then
                    // requested before this tile's MMAs / epilogue, whose latency hides their trip from HBM.
                    const bool pipelined = a.act == nullptr;
                    auto request_rows = [&](long long first_row, float4 *buf) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int q = 2 * i + (lane >> 4), j = lane & 15;
                            float4 *dst = &buf[q * 16 + (j ^ (q & 15))];
                            if (first_row + q < a.rows) {
                                cp_async16(dst, reinterpret_cast<const float4 *>(a.dY + (first_row + q) * 64) + j);
                                if (a.act) cp_async16(dst + 512, reinterpret_cast<const float4 *>(a.act + (first_row + q) * 64) + j);
                            } else {
                                *dst = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (a.act) dst[512] = make_float4(1.f, 1.f, 1.f, 1.f);
                            }
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    };
                    float4 *cur = stg + (pipelined ? (it & 1u) * 512 : 0);
                    __syncwarp();
                    if (!pipelined || it == 0) request_rows(p_warp, cur);
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float4 f = cur[lane * 16 + (i ^ (lane & 15))];
                        if (a.act) {                                // fused ReLU backward of this convolution's own output
                            const float4 y = cur[512 + lane * 16 + (i ^ (lane & 15))];
                            f.x = y.x > 0.f ? f.x : 0.f; f.y = y.y > 0.f ? f.y : 0.f;
                            f.z = y.z > 0.f ? f.z : 0.f; f.w = y.w > 0.f ? f.w : 0.f;
                        }
                        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
                    }
                    // the other half was last read one iteration ago, before the __syncwarp above
                    if (pipelined && tile + stride < n_tiles) request_rows(p_warp + stride * kTileRows, stg + ((it & 1u) ^ 1u) * 512);
                }
#pragma unroll
                for (int c = 0; c < KB * 2; c++) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) split_tf32(v[c * 16 + j], hi[j], lo[j]);
                    tmem_st16(a_hi + c * 16, hi);
                    tmem_st16(a_lo + c * 16, lo);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&a_ready[g]);
            }
            if (!mbar_wait(&z_ready[g], it & 1u, ab)) { ok = false; break; }
            tc_fence_after();
            if (MODE == 0) {
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    uint32_t z[32];
                    tmem_ld32(d + c * 32, z);
                    tmem_wait_ld();
                    uint32_t w = 0u;                              // this pixel's ReLU bits of channels 32 c .. 32 c + 31
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 b = *reinterpret_cast<const float4 *>(&s_bias[c * 32 + 4 * j]);
                        const float v0 = __uint_as_float(z[4 * j]) + b.x, v1 = __uint_as_float(z[4 * j + 1]) + b.y;
                        const float v2 = __uint_as_float(z[4 * j + 2]) + b.z, v3 = __uint_as_float(z[4 * j + 3]) + b.w;
                        w |= ((v0 > 0.f ? 1u : 0u) | (v1 > 0.f ? 2u : 0u) | (v2 > 0.f ? 4u : 0u) | (v3 > 0.f ? 8u : 0u)) << (4 * j);
                        stg[lane * 16 + ((c * 8 + j) ^ (lane & 15))] = make_float4(fmaxf(v0, 0.f), fmaxf(v1, 0.f), fmaxf(v2, 0.f), fmaxf(v3, 0.f));
                    }
                    if (a.mask_out && p < a.rows) a.mask_out[p * 2 + c] = w;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int q = 2 * i + (lane >> 4), j = lane & 15;
                    if (p_warp + q < a.rows)
                        reinterpret_cast<float4 *>(a.out + (p_warp + q) * 64)[j] = stg[q * 16 + (j ^ (q & 15))];
                }
            } else {
                // This pixel's 27 products T[p][(r,s)][c] never reach global memory as such: the three taps of a kernel row r
                // meet at pixel p from p itself (s = 1) and its two neighbours in the image row (T[p-1][(r,2)], T[p+1][(r,0)]),
                // which are the neighbouring lanes — the horizontal half of the stencil is two shuffles per (r, c).  What is
                // stored is H[r][p] = that row sum, 3 planes of float4 (48 B per pixel instead of 144; a warp's 32 consecutive
                // pixels are 512 contiguous bytes per plane); the terms that cross the warp's 32-pixel segment go to two
                // small edge arrays (lane 31's (r,2) products, lane 0's (r,0) products) that k_conv1_stencil adds back.
                uint32_t z[32];
                tmem_ld32(d, z);
                tmem_wait_ld();
                long long n_; int y_, x;
                pixel_coords(p < a.rows ? p : 0, a.rows, a.H, a.W, n_, y_, x);
                const bool has_l = lane > 0 && x > 0, has_r = lane < 31 && x < a.W - 1;
                float4 *hp = reinterpret_cast<float4 *>(a.T);
                const long long n_seg = (a.rows + 31) >> 5, seg = p_warp >> 5;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    float h[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float from_l = __shfl_up_sync(0xFFFFFFFFu, __uint_as_float(z[(r * 3 + 2) * 3 + c]), 1);
                        const float from_r = __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(z[(r * 3 + 0) * 3 + c]), 1);
                        h[c] = __uint_as_float(z[(r * 3 + 1) * 3 + c]);
                        if (has_l) h[c] += from_l;
                        if (has_r) h[c] += from_r;
                    }
                    if (p < a.rows) {
                        hp[(size_t)r * a.rows + p] = make_float4(h[0], h[1], h[2], 0.f);
                        if (lane == 31)
                            hp[3 * a.rows + seg * 3 + r] = make_float4(__uint_as_float(z[(r * 3 + 2) * 3]), __uint_as_float(z[(r * 3 + 2) * 3 + 1]),
                                                                       __uint_as_float(z[(r * 3 + 2) * 3 + 2]), 0.f);
                        if (lane == 0)
                            hp[3 * a.rows + (n_seg + seg) * 3 + r] = make_float4(__uint_as_float(z[r * 9]), __uint_as_float(z[r * 9 + 1]),
                                                                                 __uint_as_float(z[r * 9 + 2]), 0.f);
                    }
                }
            }
            tc_fence_before();      // these tcgen05.ld are ordered before the next a_ready arrive (the next MMA overwrites D)
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag && a.status) atomicOr(a.status, GOM_STATUS_TIMEOUT);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// dX[q, c] = sum over the kernel rows r of H[r][q - (r-1) W] (+ the edge terms of the 32-pixel segments): the vertical half of
// the stencil, 3 coalesced float4 loads per pixel.  Scratch layout (float4): H [3][rows] | E_right [segments][3] | E_left [segments][3].
__global__ void __launch_bounds__(256) k_conv1_stencil(Conv1Dev a) {
    gom_pdl_trigger();
    gom_pdl_wait();
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    if (q >= a.rows) return;
    const long long HW = (long long)a.H * a.W;
    long long n; int y, x;
    pixel_coords(q, a.rows, a.H, a.W, n, y, x);
    const float4 *Hp = reinterpret_cast<const float4 *>(a.T);
    const long long n_seg = (a.rows + 31) >> 5;
    const float4 *Er = Hp + 3 * a.rows, *El = Er + 3 * n_seg;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int yy = y - (r - 1);
        if (yy >= 0 && yy < a.H) {
            const long long pp = n * HW + (long long)yy * a.W + x;
            float4 f = __ldg(Hp + (size_t)r * a.rows + pp);
            if ((pp & 31) == 0 && x > 0) {                         // the left neighbour is lane 31 of the previous segment
                const float4 e = __ldg(Er + ((pp >> 5) - 1) * 3 + r);
                f.x += e.x; f.y += e.y; f.z += e.z;
            }
            if ((pp & 31) == 31 && x < a.W - 1) {                  // the right neighbour is lane 0 of the next segment
                const float4 e = __ldg(El + ((pp >> 5) + 1) * 3 + r);
                f.x += e.x; f.y += e.y; f.z += e.z;
            }
            g0 += f.x; g1 += f.y; g2 += f.z;
        }
    }
    float *o = a.dX + q * 3;
    o[0] = g0; o[1] = g1; o[2] = g2;
}

int g_conv1_sms = 0;
int conv1_setup(void) {
    if (g_conv1_sms) return GOM_OK;
    int dev = 0, sms = 0;
    GOM_CUDA(cudaGetDevice(&dev));
    GOM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GOM_CUDA(cudaFuncSetAttribute(k_conv1_gemm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1Cfg<0>::DYN_SMEM));
    GOM_CUDA(cudaFuncSetAttribute(k_conv1_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1Cfg<1>::DYN_SMEM));
    g_conv1_sms = sms;
    return GOM_OK;
}

}  // namespace

// called by gom_conv_first_forward / gom_conv_first_backward (conv_first.cu) when use_tensor_cores is set
int gom_conv_first_forward_tc(const GomConvFirstArgs *p, cudaStream_t stream) {
    GOM_REQUIRE(((uintptr_t)p->out % 16) == 0, "out must be 16-byte aligned");
    if (int rc = conv1_setup()) return rc;
    Conv1Dev a{};
    a.N = p->n_images; a.H = p->height; a.W = p->width; a.rows = (long long)a.N * a.H * a.W;
    a.x = p->x; a.weight = p->weight; a.bias = p->bias; a.out = p->out; a.status = nullptr; a.mask_out = p->mask_out;
    gom_prof_begin(GOM_PROF_CONV_FIRST_FWD, stream);
    GOM_CUDA(gom_launch_pdl(k_conv1_gemm<0>, dim3(g_conv1_sms), dim3(Conv1Cfg<0>::THREADS), Conv1Cfg<0>::DYN_SMEM, stream, a));
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_CONV_FIRST_FWD, stream);
    return GOM_OK;
}

int gom_conv_first_backward_tc(const GomConvFirstArgs *p, cudaStream_t stream) {
    GOM_REQUIRE(p->scratch != nullptr, "the tensor-core backward needs scratch [9, N*H*W, 4] floats");
    GOM_REQUIRE(((uintptr_t)p->dL_dout % 16) == 0 && ((uintptr_t)p->scratch % 16) == 0, "dL_dout / scratch must be 16-byte aligned");
    if (int rc = conv1_setup()) return rc;
    Conv1Dev a{};
    a.N = p->n_images; a.H = p->height; a.W = p->width; a.rows = (long long)a.N * a.H * a.W;
    a.weight = p->weight; a.dY = p->dL_dout; a.act = p->act; a.T = p->scratch; a.dX = p->dL_dx; a.status = nullptr;
    GOM_REQUIRE(((uintptr_t)p->act % 16) == 0, "act must be 16-byte aligned");
    gom_prof_begin(GOM_PROF_CONV_FIRST_BWD, stream);
    GOM_CUDA(gom_launch_pdl(k_conv1_gemm<1>, dim3(g_conv1_sms), dim3(Conv1Cfg<1>::THREADS), Conv1Cfg<1>::DYN_SMEM, stream, a));
    GOM_LAUNCH_CHECK();
    GOM_CUDA(gom_launch_pdl(k_conv1_stencil, dim3(gom_div_up(a.rows, 256)), dim3(256), 0, stream, a));
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_CONV_FIRST_BWD, stream);
    return GOM_OK;
}

// conv3x3_tc.cu — 3x3 / stride 1 / pad 1 convolutions over NHWC fp32 activations as tcgen05 implicit GEMMs
// (gom_conv3x3, include/gom_b200.h): the twelve VGG16 layers conv1_2 ... conv5_3 of LPIPS, forward (bias + ReLU fused)
// and input gradient (ReLU backward of the layer below fused).  Reference: utils/lpips/pretrained_networks.py:96-134
// (torchvision VGG16 `features`, executed by cuDNN there).
//
// GEMM view:  out[p, n] = sum_{tap = (r,s)} sum_c x[p + (r-1, s-1), c] * Wp[tap][n][c],   M = pixels, N = c_out, K = 9 c_in.
//   * M tile = 8 x 16 pixels of one image = 128 rows = the 128 lanes of tensor memory.
//   * A operand of one k-step (tap, 32-channel block): the 8 x 16 x 32 box of x shifted by the tap, fetched by ONE TMA tensor
//     copy (cp.async.bulk.tensor.4d) — the tensor map's bounds check zero-fills the padding ring, its 128-byte swizzle writes
//     the K-major SWIZZLE_128B layout tcgen05.mma reads.  No im2col buffer exists anywhere.
//   * B operand: the [n0 : n0 + NT] x 32 slab of the packed weight of that tap (3-D tensor map), same layout.
//   * D: NT fp32 columns of tensor memory, double buffered (2 NT <= 512) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * epilogue: tcgen05.ld 32 columns at a time, bias / ReLU / mask in registers, 32 rows x 128 B per warp into swizzled shared
//     memory, one TMA tensor store per warp and chunk (the store clips at the image border).
// Warp roles (192 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5
// epilogue (TMEM lane quarter = warp % 4).  Every mbarrier wait is bounded (gomtc::mbar_wait) and reports GOM_STATUS_TIMEOUT.
#include <cuda.h>

#include "gom_common.cuh"
#include "gom_tcgen05.cuh"

namespace {

using namespace gomtc;

constexpr int kTileH = 8, kTileW = 16;
constexpr int kABytes = 128 * 128;                 // one A k-block: 128 pixel rows x 32 fp32 channels
constexpr int kEpiWarpBytes = 4 * 4096;            // per epilogue warp: 2 store staging + 2 mask buffers of 32 rows x 128 B
constexpr int kThreads = 192;

template <int NT, int KBLK, int STAGES> struct ConvCfg {
    static constexpr int B_BYTES = NT * 128;
    static constexpr int STAGE_BYTES = KBLK * (kABytes + B_BYTES);
    static constexpr int SMEM = STAGES * STAGE_BYTES + 4 * kEpiWarpBytes + 1024;
    static constexpr uint32_t TMEM_COLS = 2 * NT <= 32 ? 32 : 2 * NT <= 64 ? 64 : 2 * NT <= 128 ? 128 : 2 * NT <= 256 ? 256 : 512;
    static_assert(2 * NT <= 512, "two accumulators must fit tensor memory");
    static_assert(SMEM <= 232448, "shared memory budget");
};

struct ConvDev {
    int n_tiles, n_tiles_n, tiles_w, tiles_h;
    int c_blocks;                // c_in / (32 * KBLK)
    int n_pass;                  // 1 (TF32) or 3 (3xTF32: x*w_hi, x_lo*w_hi, x*w_lo)
    int relu, has_act;
    const float *bias;
    uint32_t *status;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// tcgen05.mma kind::tf32, A and B from shared memory; descriptors are passed as their low words (start address, LBO): the high
// word (SBO = 1024 B, descriptor version, SWIZZLE_128B) is the constant 0x40004040 for every operand tile of this kernel
constexpr uint32_t kDescHi = 0x40004040u;
__device__ __forceinline__ void mma_tf32_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss_acc(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.b32 p, 0, 0;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHi) : "memory");
}

struct TileCoord { int n0, w0, h0, img; };
__device__ __forceinline__ TileCoord decode_tile(const ConvDev &p, int tile, int nt) {
    TileCoord t;
    const int ni = tile % p.n_tiles_n;
    int m = tile / p.n_tiles_n;
    t.n0 = ni * nt;
    t.w0 = (m % p.tiles_w) * kTileW;
    m /= p.tiles_w;
    t.h0 = (m % p.tiles_h) * kTileH;
    t.img = m / p.tiles_h;
    return t;
}

template <int NT, int KBLK, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
k_conv3x3(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
          const __grid_constant__ CUtensorMap map_act, const ConvDev p) {
    using Cfg = ConvCfg<NT, KBLK, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2], act_bar[4][2];
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        for (int i = 0; i < 4; i++) { mbar_init(&act_bar[i][0], 1); mbar_init(&act_bar[i][1], 1); }
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    volatile int *ab = &abort_flag;
    const int k_steps = p.n_pass * 9 * p.c_blocks;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------- TMA producer
        // The whole warp walks the loop (warp-uniform control flow keeps addresses in uniform registers); one elected lane
        // issues.  The barrier of the NEXT stage is probed before the copies of the current one are issued, so its ~90-cycle
        // answer arrives while they are being queued.
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        }
        uint32_t st = 0, ph = 0;
        bool ok = true;
        bool ready = mbar_test_wait(&empty_bar[0], 1u);
        for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile, NT);
            for (int pass = 0; pass < p.n_pass && ok; pass++) {
                const CUtensorMap *ma = pass == 1 ? &map_a_lo : &map_a;
                const int tap_off = pass == 2 ? 9 : 0;                  // the lo weight image follows the hi image
                for (int tap = 0; tap < 9 && ok; tap++) {
                    const int r = tap / 3, s = tap - 3 * r;
                    for (int cb = 0; cb < p.c_blocks; cb++) {
                        if (!ready && !mbar_wait(&empty_bar[st], ph ^ 1u, ab)) { ok = false; break; }
                        const uint32_t a_dst = smem_base + st * Cfg::STAGE_BYTES, b_dst = a_dst + KBLK * kABytes;
                        uint64_t *fb = &full_bar[st];
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                        ready = mbar_test_wait(&empty_bar[st], ph ^ 1u);
                        if (elect_one()) {
                            mbar_expect_tx(fb, Cfg::STAGE_BYTES);
#pragma unroll
                            for (int kb = 0; kb < KBLK; kb++) {
                                const int c0 = (cb * KBLK + kb) * 32;
                                tma_load_4d(a_dst + kb * kABytes, ma, c0, t.w0 + s - 1, t.h0 + r - 1, t.img, fb);
                                tma_load_3d(b_dst + kb * Cfg::B_BYTES, &map_b, c0, t.n0, tap_off + tap, fb);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // --------------------------------------------------------------------------------------------- MMA issuer
        const uint32_t idesc = instr_desc_n(NT);
        const uint32_t desc_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);      // low word of the descriptor of stage 0
        uint32_t st = 0, ph = 0, tcount = 0;
        bool ok = true;
        bool ready = mbar_test_wait(&full_bar[0], 0u);
        for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, tcount++) {
            const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
            if (!mbar_wait(&tempty_bar[acc], aph ^ 1u, ab)) { ok = false; break; }
            tc_fence_after();
            const uint32_t d = tmem + acc * NT;
            for (int ks = 0; ks < k_steps; ks++) {
                if (!ready && !mbar_wait(&full_bar[st], ph, ab)) { ok = false; break; }
                const uint32_t a_lo = desc_base + st * (Cfg::STAGE_BYTES >> 4), b_lo = a_lo + ((KBLK * kABytes) >> 4);
                uint64_t *eb = &empty_bar[st];
                if (++st == STAGES) { st = 0; ph ^= 1u; }
                ready = mbar_test_wait(&full_bar[st], ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < KBLK; kb++)
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t al = a_lo + ((kb * kABytes + k * 32) >> 4), bl = b_lo + ((kb * Cfg::B_BYTES + k * 32) >> 4);
                            if (kb == 0 && k == 0) mma_tf32_ss_lo(d, al, bl, idesc, (uint32_t)(ks != 0));
                            else mma_tf32_ss_acc(d, al, bl, idesc);
                        }
                    tc_commit(eb);
                    if (ks == k_steps - 1) tc_commit(&tfull_bar[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ----------------------------------------------------------------------------------------------- epilogue
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const uint32_t wbase = epi_base + (warp - 2) * kEpiWarpBytes;   // [stage 0 | stage 1 | act 0 | act 1], 4 KB each
        const uint32_t row_off = lane * 128;
        const int sw = lane & 7;
        uint32_t tcount = 0, g = 0;                               // g: running 32-column chunk counter (buffer / parity selector)
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, tcount++) {
            const TileCoord t = decode_tile(p, tile, NT);
            const int hq = t.h0 + 2 * q;                          // this warp's two pixel rows of the tile
            const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
            if (p.has_act && lane == 0) {
                mbar_expect_tx(&act_bar[q][g & 1u], 4096);
                tma_load_4d(wbase + 8192 + (g & 1u) * 4096, &map_act, t.n0, t.w0, hq, t.img, &act_bar[q][g & 1u]);
            }
            if (!mbar_wait(&tfull_bar[acc], aph, ab)) { ok = false; break; }
            tc_fence_after();
            constexpr int kChunks = NT / 32;
#pragma unroll 1
            for (int ch = 0; ch < kChunks; ch++, g++) {
                if (p.has_act && ch + 1 < kChunks && lane == 0) {
                    const uint32_t nb = (g + 1) & 1u;
                    mbar_expect_tx(&act_bar[q][nb], 4096);
                    tma_load_4d(wbase + 8192 + nb * 4096, &map_act, t.n0 + (ch + 1) * 32, t.w0, hq, t.img, &act_bar[q][nb]);
                }
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * NT + ch * 32, v);
                tmem_wait_ld();
                if (ch == kChunks - 1) {                          // accumulator drained: hand it back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
                if (p.bias) {
                    const float4 *bp = reinterpret_cast<const float4 *>(p.bias + t.n0 + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 b = __ldg(bp + j);
                        f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.f);
                }
                if (p.has_act) {
                    const uint32_t b = g & 1u;
                    if (!mbar_wait(&act_bar[q][b], (g >> 1) & 1u, ab)) { ok = false; break; }
                    const uint32_t abuf = wbase + 8192 + b * 4096 + row_off;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 y;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "r"(abuf + ((j ^ sw) << 4)));
                        f[4 * j] = y.x > 0.f ? f[4 * j] : 0.f;
                        f[4 * j + 1] = y.y > 0.f ? f[4 * j + 1] : 0.f;
                        f[4 * j + 2] = y.z > 0.f ? f[4 * j + 2] : 0.f;
                        f[4 * j + 3] = y.w > 0.f ? f[4 * j + 3] : 0.f;
                    }
                }
                // the staging buffer about to be written was handed to a TMA store two chunks ago: wait until it has been read
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
                const uint32_t sbuf = wbase + (g & 1u) * 4096;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + row_off + ((j ^ sw) << 4)), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                                 "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&map_out, sbuf, t.n0 + ch * 32, t.w0, hq, t.img);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag && p.status) atomicOr(p.status, GOM_STATUS_TIMEOUT);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------ helpers
__global__ void k_pack_weights(GomConvPackArgs a) {
    const int K = a.c_out, C = a.c_in;
    const long long total = 9ll * K * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        float w;
        if (!a.transpose) {          // packed[tap][k][c] = W[k][c][tap]
            const int c = (int)(e % C), k = (int)((e / C) % K), tap = (int)(e / ((long long)C * K));
            w = a.weight[((long long)k * C + c) * 9 + tap];
        } else {                     // packed[tap][c][k] = W[k][c][8 - tap]
            const int k = (int)(e % K), c = (int)((e / K) % C), tap = (int)(e / ((long long)C * K));
            w = a.weight[((long long)k * C + c) * 9 + (8 - tap)];
        }
        uint32_t hi, lo;
        split_tf32(w, hi, lo);
        a.packed[e] = __uint_as_float(hi);
        if (a.split) a.packed[total + e] = __uint_as_float(lo);
    }
}

// lo = x - trunc_tf32(x): the part of x the tensor core drops when it reads the fp32 word x as a TF32 operand
__global__ void k_tf32_split(GomTf32SplitArgs a) {
    const long long n4 = a.n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(a.x) + i);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
        if (a.hi) reinterpret_cast<float4 *>(a.hi)[i] = h;
        reinterpret_cast<float4 *>(a.lo)[i] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_sms = 0;

int conv_setup(void) {
    if (g_encode && g_sms) return GOM_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GOM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        gom_set_error("gom_conv3x3: cuTensorMapEncodeTiled is not available from this driver");
        return GOM_ERR_UNSUPPORTED;
    }
    int dev = 0, sms = 0;
    GOM_CUDA(cudaGetDevice(&dev));
    GOM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_encode = (EncodeTiledFn)fn;
    g_sms = sms;
    return GOM_OK;
}

// NHWC activation [N,H,W,C] as a 4-D tensor (C, W, H, N) with a (32, 16, box_h, 1) box, 128-byte swizzle, zero fill
int make_act_map(CUtensorMap *m, const float *base, int N, int H, int W, int C, int box_h, bool tf32_round) {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)kTileW, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = g_encode(m, tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gom_set_error("gom_conv3x3: cuTensorMapEncodeTiled (activation) failed: %d", (int)r); return GOM_ERR_CUDA; }
    return GOM_OK;
}
// packed weight [taps][c_out][c_in] as a 3-D tensor (c_in, c_out, taps) with a (32, nt, 1) box
int make_weight_map(CUtensorMap *m, const float *base, int taps, int c_out, int c_in, int nt) {
    const cuuint64_t dims[3] = {(cuuint64_t)c_in, (cuuint64_t)c_out, (cuuint64_t)taps};
    const cuuint64_t strides[2] = {(cuuint64_t)c_in * 4, (cuuint64_t)c_out * c_in * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)nt, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gom_set_error("gom_conv3x3: cuTensorMapEncodeTiled (weights) failed: %d", (int)r); return GOM_ERR_CUDA; }
    return GOM_OK;
}

template <int NT, int KBLK, int STAGES>
int launch_conv(const CUtensorMap &ma, const CUtensorMap &malo, const CUtensorMap &mb, const CUtensorMap &mo, const CUtensorMap &mact,
                const ConvDev &d, cudaStream_t stream) {
    using Cfg = ConvCfg<NT, KBLK, STAGES>;
    static bool configured = false;
    if (!configured) {
        GOM_CUDA(cudaFuncSetAttribute(k_conv3x3<NT, KBLK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        configured = true;
    }
    const int grid = d.n_tiles < g_sms ? d.n_tiles : g_sms;
    k_conv3x3<NT, KBLK, STAGES><<<grid, kThreads, Cfg::SMEM, stream>>>(ma, malo, mb, mo, mact, d);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

}  // namespace

extern "C" int gom_conv3x3_pack_weights(const GomConvPackArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->c_out > 0 && p->c_in > 0, "sizes");
    GOM_REQUIRE(p->weight && p->packed, "null pointer");
    const long long total = 9ll * p->c_out * p->c_in;
    int blocks = gom_div_up(total, 256);
    if (blocks > 2048) blocks = 2048;
    k_pack_weights<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_tf32_split(const GomTf32SplitArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n > 0 && p->n % 4 == 0, "n must be a positive multiple of 4");
    GOM_REQUIRE(p->x && p->lo, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->x % 16) == 0 && ((uintptr_t)p->lo % 16) == 0 && ((uintptr_t)p->hi % 16) == 0, "16-byte alignment");
    int blocks = gom_div_up(p->n / 4, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_tf32_split<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_conv3x3(const GomConv3x3Args *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_images > 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->c_in > 0 && p->c_in % 32 == 0 && p->c_out > 0 && p->c_out % 32 == 0, "channel counts must be multiples of 32");
    GOM_REQUIRE(p->x && p->w_packed && p->out, "null pointer");
    GOM_REQUIRE(p->precision == 0 || (p->precision == 1 && p->x_lo), "precision = 1 needs x_lo");
    GOM_REQUIRE(((uintptr_t)p->x % 16) == 0 && ((uintptr_t)p->out % 16) == 0 && ((uintptr_t)p->w_packed % 16) == 0 &&
                ((uintptr_t)p->act % 16) == 0 && ((uintptr_t)p->x_lo % 16) == 0 && ((uintptr_t)p->bias % 16) == 0, "16-byte alignment");
    if (int rc = conv_setup()) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nt = p->c_out % 256 == 0 ? 256 : p->c_out % 128 == 0 ? 128 : p->c_out % 64 == 0 ? 64 : 32;
    const int kblk = (nt <= 64 && p->c_in % 64 == 0) ? 2 : 1;

    ConvDev d{};
    d.tiles_w = gom_div_up(p->width, kTileW);
    d.tiles_h = gom_div_up(p->height, kTileH);
    d.n_tiles_n = p->c_out / nt;
    const long long n_tiles = (long long)p->n_images * d.tiles_h * d.tiles_w * d.n_tiles_n;
    GOM_REQUIRE(n_tiles < (1ll << 30), "too many tiles");
    d.n_tiles = (int)n_tiles;
    d.c_blocks = p->c_in / (32 * kblk);
    d.n_pass = p->precision == 1 ? 3 : 1;
    d.relu = p->relu;
    d.has_act = p->act != nullptr;
    d.bias = p->bias;
    d.status = p->status;

    CUtensorMap ma, malo, mb, mo, mact;
    const bool round = p->tma_round && p->precision == 0;
    if (int rc = make_act_map(&ma, p->x, p->n_images, p->height, p->width, p->c_in, kTileH, round)) return rc;
    if (int rc = make_act_map(&malo, p->precision == 1 ? p->x_lo : p->x, p->n_images, p->height, p->width, p->c_in, kTileH, false)) return rc;
    if (int rc = make_weight_map(&mb, p->w_packed, p->precision == 1 ? 18 : 9, p->c_out, p->c_in, nt)) return rc;
    if (int rc = make_act_map(&mo, p->out, p->n_images, p->height, p->width, p->c_out, 2, false)) return rc;
    if (int rc = make_act_map(&mact, p->act ? p->act : p->out, p->n_images, p->height, p->width, p->c_out, 2, false)) return rc;

    gom_prof_begin(p->act || !p->relu ? GOM_PROF_CONV3X3_DGRAD : GOM_PROF_CONV3X3_FWD, stream);
    int rc;
    if (nt == 256) rc = launch_conv<256, 1, 3>(ma, malo, mb, mo, mact, d, stream);
    else if (nt == 128) rc = launch_conv<128, 1, 4>(ma, malo, mb, mo, mact, d, stream);
    else if (nt == 64 && kblk == 2) rc = launch_conv<64, 2, 3>(ma, malo, mb, mo, mact, d, stream);
    else if (nt == 64) rc = launch_conv<64, 1, 6>(ma, malo, mb, mo, mact, d, stream);
    else rc = launch_conv<32, 1, 6>(ma, malo, mb, mo, mact, d, stream);
    if (rc) return rc;
    gom_prof_end(p->act || !p->relu ? GOM_PROF_CONV3X3_DGRAD : GOM_PROF_CONV3X3_FWD, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_conv3x3_args(void) { return sizeof(GomConv3x3Args); }
extern "C" size_t gom_sizeof_conv_pack_args(void) { return sizeof(GomConvPackArgs); }
extern "C" size_t gom_sizeof_tf32_split_args(void) { return sizeof(GomTf32SplitArgs); }

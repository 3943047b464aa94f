// conv3x3_tc.cu — 3x3 / stride 1 / pad 1 convolutions over NHWC fp32 activations as tcgen05 implicit GEMMs
// (gom_conv3x3, include/gom_b200.h): the twelve VGG16 layers conv1_2 ... conv5_3 of LPIPS, forward (bias + ReLU fused, ReLU
// bit mask emitted) and input gradient (ReLU backward of the layer below fused through that bit mask).
// Reference: utils/lpips/pretrained_networks.py:96-134 (torchvision VGG16 `features`, executed by cuDNN there).
//
// GEMM view:  out[p, n] = sum_{tap = (r,s)} sum_c x[p + (r-1, s-1), c] * Wp[tap][n][c],   M = pixels, N = c_out, K = 9 c_in.
// The kernel is bound by operand traffic L2 -> shared memory (measured: ~58 B/clk/SM whatever the tile shape), so the tiling
// is chosen to move as few bytes per MMA as possible:
//   * A CTA tile is 16 x 16 output pixels of one image = TWO M = 128 sub-tiles (left / right 8 columns; tensor-memory lane =
//     16 rows x 8 columns) sharing every weight tile.  (Layers with too few such tiles to fill the 148 SMs — the deep layers
//     at the reference's batch size of one frame — use ONE sub-tile and / or 64 output channels per tile: SUB, NT.)
//   * A operand: ONE TMA tensor copy per 32-channel block fetches the 18 x PITCH-pixel halo of the CTA tile (zero-filled outside
//     the image by the tensor map's bounds check, 128-byte swizzle).  The nine taps are nine shifted VIEWS of that halo: the
//     shared-memory matrix descriptor of tap (r,s), sub-tile m starts at pixel (r, s + 8 m) of the halo, its 8-row groups (one
//     image row of 8 pixels each) are PITCH * 128 B apart (SBO).  9x fewer activation bytes than one shifted box per tap, no
//     im2col buffer anywhere.
//   * B operand: the [n0 : n0 + NT] x 32 slab of the packed weight of (tap, channel block), a ring of small stages.
//   * D: 2 sub-tiles x NT fp32 columns of tensor memory, double buffered (4 NT <= 512): the epilogue of tile i overlaps the MMAs
//     of tile i + 1.
//   * epilogue (8 warps, two per TMEM lane quarter, one sub-tile each): tcgen05.ld 32 columns at a time, bias / ReLU / mask in
//     registers, 32 pixels x 128 B per warp into swizzled shared memory, one TMA tensor store per warp and chunk (clipped at the
//     image border); the forward also writes one ReLU mask word per pixel and chunk, which the dgrad of the layer above reads
//     back instead of the activation itself (1/32 of the bytes, no shared-memory staging).
// Warp roles (320 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-9
// epilogue.  Every mbarrier wait is bounded (gomtc::mbar_wait) and reports GOM_STATUS_TIMEOUT.
// Two kernels share this design: k_conv3x3 (one CTA per tile; also the K-split shapes and the 1x1 GEMM) and k_conv3x3_pair
// (the default: two CTAs of a cluster on one M = 256 tile with tcgen05 cta_group::2, each holding half of every weight tile).
#include <cuda.h>
#include <stdlib.h>

#include "gom_common.cuh"
#include "gom_tcgen05.cuh"

namespace {

using namespace gomtc;

constexpr int kTileH = 16;                         // CTA tile: 16 rows x (8 SUB) columns of output pixels
constexpr int kEpiWarpBytes = 4096;                // store staging of one epilogue warp: 32 pixels x 32 channels

// NT: output channels per tile; SUB: M = 128 sub-tiles per tile (1 or 2); TPS: taps per weight stage (1 or 3 — a stage must
// hold enough MMA work, >= ~500 cycles, to cover the issuing warp's per-stage barrier round trip); BST: weight ring depth
// TAPS: 9 = 3x3 convolution (halo + shifted views); 1 = 1x1 convolution = plain GEMM over pixel rows (the MLP layers): no halo,
// and the activation tile travels in the same stage ring as its weight tile (one barrier pair per k-block).
template <int NT, int SUB, int TPS, int BST, int TAPS = 9> struct ConvCfg {
    static constexpr int TILE_W = 8 * SUB;
    static constexpr int HALO = TAPS == 9 ? 2 : 0;
    static constexpr int PITCH = TILE_W + HALO;     // halo pitch in pixels: exactly the columns a tile needs (see gom_conv3x3)
    static constexpr int HALO_H = kTileH + HALO;
    static constexpr int A_TX_BYTES = HALO_H * PITCH * 128;
    static constexpr int A_BYTES = ((A_TX_BYTES + 1023) / 1024) * 1024;
    // halo ring: a halo must be requested a full L2 round trip (~2 500 cycles) before its first MMA; one halo of the smallest
    // shape (NT 64, one sub-tile) is only ~1 700 cycles of MMA work, so that shape keeps three halos in flight
    static constexpr int A_BUFS = TAPS == 1 ? BST : (SUB == 1 && NT <= 64) ? 4 : 2;
    static constexpr int B_TAP_BYTES = NT * 128;
    static constexpr int B_BYTES = TPS * B_TAP_BYTES;
    static constexpr int GROUPS = TAPS / TPS;        // weight stages per halo
    static constexpr int HALO_AT = TPS == 1 ? 5 : 1; // the next halo is requested before this weight stage of the current one
    static constexpr int EPI_WARPS = 4 * SUB;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int ACC_COLS = SUB * NT;       // one accumulator stage
    static constexpr int SMEM = A_BUFS * A_BYTES + BST * B_BYTES + EPI_WARPS * kEpiWarpBytes + 1024;
    static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
    static_assert(2 * ACC_COLS <= 512, "two accumulator stages must fit tensor memory");
    static_assert(SMEM <= 232448, "shared memory budget");
    static_assert(TPS == 1 || TPS == 3, "taps per stage");
    static_assert(TAPS == 9 || (TAPS == 1 && TPS == 1), "kernel size");
};

struct ConvDev {
    int n_tiles, n_tiles_n, tiles_w, tiles_h;
    int H, W;
    int k_splits;                // > 1: the channel blocks of a tile are shared out over k_splits CTAs whose partial sums meet in
                                 // global memory (TMA reduce-add into a zeroed output; bias / ReLU / masks in k_conv_finish)
    int c_blocks;                // channel blocks (of 32) PER SPLIT
    int n_pass;                  // 1 (TF32) or 3 (3xTF32: x*w_hi, x_lo*w_hi, x*w_lo)
    int relu;
    int mask_words;              // c_out / 32: mask words per pixel
    const float *bias;
    const uint32_t *mask_in;
    uint32_t *mask_out;
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
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
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
// tcgen05.mma kind::tf32, A and B from shared memory, descriptors given as (low word, high word) pairs.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_acc(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.eq.b32 p, 0, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
// high word of a K-major SWIZZLE_128B descriptor: SBO (bytes >> 4) | version 1 (bit 46) | base offset (bits 49-51) | layout 2 (bits 61-63)
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t base_offset) {
    return (sbo_bytes >> 4) | (1u << 14) | (base_offset << 17) | (2u << 29);
}

struct TileCoord { int n0, w0, h0, img, cb0; };
__device__ __forceinline__ TileCoord decode_tile(const ConvDev &p, int tile, int nt, int tile_w) {
    TileCoord t;
    t.cb0 = (tile % p.k_splits) * p.c_blocks;                 // splits of one tile are neighbours: they run at the same time
    tile /= p.k_splits;
    const int ni = tile % p.n_tiles_n;
    int m = tile / p.n_tiles_n;
    t.n0 = ni * nt;
    t.w0 = (m % p.tiles_w) * tile_w;
    m /= p.tiles_w;
    t.h0 = (m % p.tiles_h) * kTileH;
    t.img = m / p.tiles_h;
    return t;
}

template <int NT, int SUB, int TPS, int BST, int TAPS>
__global__ void __launch_bounds__(ConvCfg<NT, SUB, TPS, BST, TAPS>::THREADS, 1)
k_conv3x3(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out, const ConvDev p) {
    using Cfg = ConvCfg<NT, SUB, TPS, BST, TAPS>;
    constexpr int PITCH = Cfg::PITCH;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t afull_bar[Cfg::A_BUFS], aempty_bar[Cfg::A_BUFS], bfull_bar[BST], bempty_bar[BST], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + Cfg::A_BUFS * Cfg::A_BYTES;
    const uint32_t epi_base = b_base + BST * Cfg::B_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::A_BUFS; i++) { mbar_init(&afull_bar[i], 1); mbar_init(&aempty_bar[i], 1); }
        for (int i = 0; i < BST; i++) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], Cfg::EPI_WARPS); }
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
    const int items_per_tile = p.n_pass * p.c_blocks;             // halo loads per tile; each is followed by 9 weight tiles
    // programmatic dependent launch: this CTA may have become resident (and done the set-up above) while the previous kernel
    // of the stream was still draining; nothing below may touch global memory before that kernel has completed
    gom_pdl_trigger();
    gom_pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------- TMA producer
        // One elected lane issues; the loop itself is warp-uniform.  Halo i + 1 is requested in the middle of the weight stages
        // of halo i: the MMA warp is then about to release (or has released) the buffer it goes into, and the weight ring
        // still holds work for it while the producer waits for that.
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        }
        uint32_t bst = 0, bph = 0, a_count = 0;
        bool ok = true;
        if constexpr (TAPS == 1) {
            // plain GEMM: k-block i = activation tile + weight tile in stage i % BST, one full / empty barrier pair
            bool ready = mbar_test_wait(&bempty_bar[0], 1u);
            for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile, NT, Cfg::TILE_W);
                for (int item = 0; item < items_per_tile; item++) {
                    const int pass = item / p.c_blocks, cb = item - pass * p.c_blocks;
                    if (!ready && !mbar_wait(&bempty_bar[bst], bph ^ 1u, ab)) { ok = false; break; }
                    const uint32_t a_dst = smem_base + bst * Cfg::A_BYTES, b_dst = b_base + bst * Cfg::B_BYTES;
                    uint64_t *fb = &bfull_bar[bst];
                    if (++bst == BST) { bst = 0; bph ^= 1u; }
                    ready = mbar_test_wait(&bempty_bar[bst], bph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(fb, Cfg::A_TX_BYTES + Cfg::B_BYTES);
                        tma_load_4d(a_dst, pass == 1 ? &map_a_lo : &map_a, (t.cb0 + cb) * 32, t.w0, t.h0, t.img, fb);
                        tma_load_3d(b_dst, &map_b, (t.cb0 + cb) * 32, t.n0, pass == 2 ? 1 : 0, fb);
                    }
                    __syncwarp();
                }
            }
        } else {
        // iterator over halo items (tile, pass, channel block), one ahead of the weight stream
        int a_tile = blockIdx.x, a_item = 0;
        auto issue_halo = [&]() -> bool {
            if (a_tile >= p.n_tiles) return true;
            const TileCoord t = decode_tile(p, a_tile, NT, Cfg::TILE_W);
            const int pass = a_item / p.c_blocks, cb = a_item - pass * p.c_blocks;
            const uint32_t buf = a_count % Cfg::A_BUFS, ph = (a_count / Cfg::A_BUFS) & 1u;
            if (!mbar_wait(&aempty_bar[buf], ph ^ 1u, ab)) return false;
            if (elect_one()) {
                mbar_expect_tx(&afull_bar[buf], Cfg::A_TX_BYTES);
                tma_load_4d(smem_base + buf * Cfg::A_BYTES, pass == 1 ? &map_a_lo : &map_a, (t.cb0 + cb) * 32, t.w0 - 1, t.h0 - 1, t.img, &afull_bar[buf]);
            }
            __syncwarp();
            a_count++;
            if (++a_item == items_per_tile) { a_item = 0; a_tile += gridDim.x; }
            return true;
        };
        for (int i = 0; i < Cfg::A_BUFS - 1 && ok; i++) ok = issue_halo();       // halos run A_BUFS - 1 items ahead of the weights
        bool ready = mbar_test_wait(&bempty_bar[0], 1u);
        for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile, NT, Cfg::TILE_W);
            for (int item = 0; item < items_per_tile && ok; item++) {
                const int pass = item / p.c_blocks, cb = item - pass * p.c_blocks;
                const int tap_off = pass == 2 ? 9 : 0;                  // the lo weight image follows the hi image
                for (int grp = 0; grp < Cfg::GROUPS; grp++) {
                    if (grp == Cfg::HALO_AT && !issue_halo()) { ok = false; break; }
                    if (!ready && !mbar_wait(&bempty_bar[bst], bph ^ 1u, ab)) { ok = false; break; }
                    const uint32_t dst = b_base + bst * Cfg::B_BYTES;
                    uint64_t *fb = &bfull_bar[bst];
                    if (++bst == BST) { bst = 0; bph ^= 1u; }
                    ready = mbar_test_wait(&bempty_bar[bst], bph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(fb, Cfg::B_BYTES);
                        tma_load_3d(dst, &map_b, (t.cb0 + cb) * 32, t.n0, tap_off + grp * TPS, fb);       // TPS taps in one box
                    }
                    __syncwarp();
                }
            }
        }
        }
    } else if (warp == 1) {
        // --------------------------------------------------------------------------------------------- MMA issuer
        const uint32_t idesc = instr_desc_n(NT);
        const uint32_t a_desc0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);        // low descriptor word of halo buffer 0, pixel (0,0)
        const uint32_t b_desc0 = ((b_base & 0x3FFFFu) >> 4) | (1u << 16);
        constexpr uint32_t kBHi = desc_hi(1024, 0);
        uint32_t bst = 0, bph = 0, a_count = 0, tcount = 0;
        bool ok = true;
        bool ready = mbar_test_wait(&bfull_bar[0], 0u);
        for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, tcount++) {
            const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
            if (!mbar_wait(&tempty_bar[acc], aph ^ 1u, ab)) { ok = false; break; }
            const uint32_t d0 = tmem + acc * Cfg::ACC_COLS;
            for (int item = 0; item < items_per_tile && ok; item++, a_count++) {
                const uint32_t abuf = TAPS == 1 ? bst : (a_count % Cfg::A_BUFS);      // GEMM: the activation tile shares the weight stage
                if (TAPS == 9 && !mbar_wait(&afull_bar[abuf], (a_count / Cfg::A_BUFS) & 1u, ab)) { ok = false; break; }
                const uint32_t a_buf_lo = a_desc0 + abuf * (Cfg::A_BYTES >> 4);
#pragma unroll
                for (int grp = 0; grp < Cfg::GROUPS; grp++) {
                    if (!ready && !mbar_wait(&bfull_bar[bst], bph, ab)) { ok = false; break; }
                    const uint32_t b_lo = b_desc0 + bst * (Cfg::B_BYTES >> 4);
                    uint64_t *eb = &bempty_bar[bst];
                    if (++bst == BST) { bst = 0; bph ^= 1u; }
                    ready = mbar_test_wait(&bfull_bar[bst], bph);
                    tc_fence_after();
                    if (elect_one()) {
                        constexpr uint32_t a_hi = desc_hi(PITCH * 128, 0);
#pragma unroll
                        for (int ti = 0; ti < TPS; ti++) {
                            const int tap = grp * TPS + ti, r = TAPS == 9 ? tap / 3 : 0, s = TAPS == 9 ? tap % 3 : 0;
#pragma unroll
                            for (int m = 0; m < SUB; m++)
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    const uint32_t al = a_buf_lo + (((r * PITCH + s + 8 * m) * 128 + k * 32) >> 4);
                                    const uint32_t bl = b_lo + ((ti * Cfg::B_TAP_BYTES + k * 32) >> 4);
                                    if (k == 0 && tap == 0) mma_tf32(d0 + m * NT, al, a_hi, bl, kBHi, idesc, (uint32_t)(item != 0));
                                    else mma_tf32_acc(d0 + m * NT, al, a_hi, bl, kBHi, idesc);
                                }
                        }
                        tc_commit(eb);
                        if (grp == Cfg::GROUPS - 1) {
                            if (TAPS == 9) tc_commit(&aempty_bar[abuf]);
                            if (item == items_per_tile - 1) tc_commit(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ----------------------------------------------------------------------------------------------- epilogue
        const int e = warp - 2;
        const int m = e >> 2;                                     // sub-tile (left / right 8 columns)
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access = tile rows 4q .. 4q + 3
        const uint32_t sbuf = epi_base + e * kEpiWarpBytes;
        const uint32_t row_off = lane * 128;
        const int sw = lane & 7;
        uint32_t tcount = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x, tcount++) {
            const TileCoord t = decode_tile(p, tile, NT, Cfg::TILE_W);
            const int hw = t.h0 + 4 * q, ww = t.w0 + 8 * m;       // this warp's 4 x 8 pixels
            const int ph_ = hw + (lane >> 3), pw = ww + (lane & 7);
            const bool inside = ph_ < p.H && pw < p.W;
            const long long mask_idx = (((long long)t.img * p.H + ph_) * p.W + pw) * p.mask_words + (t.n0 >> 5);
            const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
            if (!mbar_wait(&tfull_bar[acc], aph, ab)) { ok = false; break; }
            tc_fence_after();
            constexpr int kChunks = NT / 32;
#pragma unroll 1
            for (int ch = 0; ch < kChunks; ch++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + m * NT + ch * 32, v);
                uint32_t mword = 0xFFFFFFFFu;
                if (p.mask_in && inside && p.k_splits == 1) mword = __ldg(p.mask_in + mask_idx + ch);
                tmem_wait_ld();
                if (ch == kChunks - 1) {                          // accumulator drained: hand it back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
                const bool partial = p.k_splits > 1;                  // a K-split's partial sum: epilogue math happens in k_conv_finish
                if (p.bias && !partial) {
                    const float4 *bp = reinterpret_cast<const float4 *>(p.bias + t.n0 + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 b = __ldg(bp + j);
                        f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                    }
                }
                if (p.relu && !partial) {
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.f);
                }
                if (p.mask_in && !partial) {
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = (mword >> j) & 1u ? f[j] : 0.f;
                }
                if (p.mask_out && !partial) {
                    uint32_t w = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) w |= (f[j] > 0.f ? 1u : 0u) << j;
                    if (inside) p.mask_out[mask_idx + ch] = w;
                }
                // the staging buffer was handed to a TMA store one chunk ago: wait until it has been read
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + row_off + ((j ^ sw) << 4)), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                                 "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (partial) tma_reduce_add_4d(&map_out, sbuf, t.n0 + ch * 32, ww, hw, t.img);
                    else tma_store_4d(&map_out, sbuf, t.n0 + ch * 32, ww, hw, t.img);
                    bulk_commit();
                }
            }
        }
        // the staging buffer must outlive the store's READ of it; the writes themselves are complete (and visible) at kernel end
        if (lane == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag && p.status) atomicOr(p.status, GOM_STATUS_TIMEOUT);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}


// ===================================================================================== CTA pairs (tcgen05 cta_group::2)
// The same convolution with TWO CTAs (a thread-block cluster of 2 = the two SMs of a TPC) on one M = 256 tile: CTA r of the
// pair owns the pixel tile at columns w0 + r TILE_W (its own halo, its own 128 accumulator rows per sub-tile in its own tensor
// memory) and HALF of every weight tile (rows n0 + r NT/2 ...).  One tcgen05.mma.cta_group::2 issued by the leader (rank 0)
// multiplies both CTAs' activation views with the weight tile that is spread over both shared memories: per MMA a CTA reads
// 4 KB of activations + NT/2 rows of weights instead of NT rows — the shared-memory operand reads that bind cta_group::1
// (128 B/clk: 8 KB per 64-cycle MMA at NT = 128, 6 KB per 32-cycle MMA at NT = 64) drop below the tensor rate, and every
// weight byte crosses L2 -> shared memory once per PAIR.
// Protocol: all TMA loads of both CTAs complete on the LEADER's full barriers (cp.async.bulk.tensor.cta_group::2 with the
// barrier address of rank 0; the leader's producer expects the bytes of both); the leader's MMA warp releases stages /
// publishes accumulators with multicast commits that arrive on the barrier of the same name in BOTH CTAs; the epilogue warps of
// both CTAs hand an accumulator stage back on the leader's barrier (remote arrive).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;        // shared::cluster address of the same offset in the pair's rank-0 CTA
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *leader_bar) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(leader_bar) & kLeaderMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *leader_bar) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(leader_bar) & kLeaderMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {          // arrives on `bar` of both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kLeaderMask) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t instr_desc_pair(int n) {           // M = 256 over the pair
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int NT, int SUB, int TPS, int BST> struct PairCfg {
    using Base = ConvCfg<NT, SUB, TPS, BST, 9>;
    static constexpr int B_TAP_BYTES = (NT / 2) * 128;          // this CTA's half of a tap's weight tile
    static constexpr int B_BYTES = TPS * B_TAP_BYTES;
    static constexpr int SMEM = Base::A_BUFS * Base::A_BYTES + BST * B_BYTES + Base::EPI_WARPS * kEpiWarpBytes + 1024;
    static_assert(NT % 16 == 0 && (NT / 2) % 8 == 0, "half weight tiles are whole 8-row groups");
};

template <int NT, int SUB, int TPS, int BST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ConvCfg<NT, SUB, TPS, BST, 9>::THREADS, 1)
k_conv3x3_pair(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out, const ConvDev p) {
    using Cfg = ConvCfg<NT, SUB, TPS, BST, 9>;
    using PC = PairCfg<NT, SUB, TPS, BST>;
    constexpr int PITCH = Cfg::PITCH;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t afull_bar[Cfg::A_BUFS], aempty_bar[Cfg::A_BUFS], bfull_bar[BST], bempty_bar[BST], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + Cfg::A_BUFS * Cfg::A_BYTES;
    const uint32_t epi_base = b_base + BST * PC::B_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::A_BUFS; i++) { mbar_init(&afull_bar[i], 1); mbar_init(&aempty_bar[i], 1); }
        for (int i = 0; i < BST; i++) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 2 * Cfg::EPI_WARPS); }
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                       // the same warp of both CTAs allocates the pair's tensor memory (same columns in both)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // both CTAs' barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    volatile int *ab = &abort_flag;
    const int items_per_tile = p.n_pass * p.c_blocks;
    gom_pdl_trigger();
    gom_pdl_wait();

    auto tile_of = [&](int tile) {         // this CTA's half of pair tile `tile`
        TileCoord t = decode_tile(p, tile, NT, 2 * Cfg::TILE_W);
        t.w0 += rank * Cfg::TILE_W;
        return t;
    };

    if (warp == 0) {
        // --------------------------------------------------------------------------- TMA producer (both CTAs, own operands)
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        }
        uint32_t bst = 0, bph = 0, a_count = 0;
        bool ok = true;
        int a_tile = pair, a_item = 0;
        auto issue_halo = [&]() -> bool {
            if (a_tile >= p.n_tiles) return true;
            const TileCoord t = tile_of(a_tile);
            const int pass = a_item / p.c_blocks, cb = a_item - pass * p.c_blocks;
            const uint32_t buf = a_count % Cfg::A_BUFS, ph = (a_count / Cfg::A_BUFS) & 1u;
            if (!mbar_wait(&aempty_bar[buf], ph ^ 1u, ab)) return false;
            if (elect_one()) {
                if (leader) mbar_expect_tx(&afull_bar[buf], 2 * Cfg::A_TX_BYTES);
                tma_load_4d_pair(smem_base + buf * Cfg::A_BYTES, pass == 1 ? &map_a_lo : &map_a, cb * 32, t.w0 - 1, t.h0 - 1, t.img, &afull_bar[buf]);
            }
            __syncwarp();
            a_count++;
            if (++a_item == items_per_tile) { a_item = 0; a_tile += n_pairs; }
            return true;
        };
        for (int i = 0; i < Cfg::A_BUFS - 1 && ok; i++) ok = issue_halo();
        bool ready = mbar_test_wait(&bempty_bar[0], 1u);
        for (int tile = pair; tile < p.n_tiles && ok; tile += n_pairs) {
            const TileCoord t = tile_of(tile);
            for (int item = 0; item < items_per_tile && ok; item++) {
                const int pass = item / p.c_blocks, cb = item - pass * p.c_blocks;
                const int tap_off = pass == 2 ? 9 : 0;
                for (int grp = 0; grp < Cfg::GROUPS; grp++) {
                    if (grp == Cfg::HALO_AT && !issue_halo()) { ok = false; break; }
                    if (!ready && !mbar_wait(&bempty_bar[bst], bph ^ 1u, ab)) { ok = false; break; }
                    const uint32_t dst = b_base + bst * PC::B_BYTES;
                    uint64_t *fb = &bfull_bar[bst];
                    if (++bst == BST) { bst = 0; bph ^= 1u; }
                    ready = mbar_test_wait(&bempty_bar[bst], bph ^ 1u);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(fb, 2 * PC::B_BYTES);
                        tma_load_3d_pair(dst, &map_b, cb * 32, t.n0 + rank * (NT / 2), tap_off + grp * TPS, fb);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
        const uint32_t idesc = instr_desc_pair(NT);
        const uint32_t a_desc0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_desc0 = ((b_base & 0x3FFFFu) >> 4) | (1u << 16);
        constexpr uint32_t kBHi = desc_hi(1024, 0);
        uint32_t bst = 0, bph = 0, a_count = 0, tcount = 0;
        bool ok = true;
        bool ready = mbar_test_wait(&bfull_bar[0], 0u);
        for (int tile = pair; tile < p.n_tiles && ok; tile += n_pairs, tcount++) {
            const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
            if (!mbar_wait(&tempty_bar[acc], aph ^ 1u, ab)) { ok = false; break; }
            const uint32_t d0 = tmem + acc * Cfg::ACC_COLS;
            for (int item = 0; item < items_per_tile && ok; item++, a_count++) {
                const uint32_t abuf = a_count % Cfg::A_BUFS;
                if (!mbar_wait(&afull_bar[abuf], (a_count / Cfg::A_BUFS) & 1u, ab)) { ok = false; break; }
                const uint32_t a_buf_lo = a_desc0 + abuf * (Cfg::A_BYTES >> 4);
#pragma unroll
                for (int grp = 0; grp < Cfg::GROUPS; grp++) {
                    if (!ready && !mbar_wait(&bfull_bar[bst], bph, ab)) { ok = false; break; }
                    const uint32_t b_lo = b_desc0 + bst * (PC::B_BYTES >> 4);
                    uint64_t *eb = &bempty_bar[bst];
                    if (++bst == BST) { bst = 0; bph ^= 1u; }
                    ready = mbar_test_wait(&bfull_bar[bst], bph);
                    tc_fence_after();
                    if (elect_one()) {
                        constexpr uint32_t a_hi = desc_hi(PITCH * 128, 0);
#pragma unroll
                        for (int ti = 0; ti < TPS; ti++) {
                            const int tap = grp * TPS + ti, r = tap / 3, s_ = tap % 3;
#pragma unroll
                            for (int m = 0; m < SUB; m++)
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    const uint32_t al = a_buf_lo + (((r * PITCH + s_ + 8 * m) * 128 + k * 32) >> 4);
                                    const uint32_t bl = b_lo + ((ti * PC::B_TAP_BYTES + k * 32) >> 4);
                                    mma_tf32_pair(d0 + m * NT, al, a_hi, bl, kBHi, idesc, (uint32_t)(!(k == 0 && tap == 0 && item == 0)));
                                }
                        }
                        tc_commit_pair(eb);
                        if (grp == Cfg::GROUPS - 1) {
                            tc_commit_pair(&aempty_bar[abuf]);
                            if (item == items_per_tile - 1) tc_commit_pair(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (both CTAs, own accumulator rows)
        const int e = warp - 2;
        const int m = e >> 2;
        const int q = warp & 3;
        const uint32_t sbuf = epi_base + e * kEpiWarpBytes;
        const uint32_t row_off = lane * 128;
        const int sw = lane & 7;
        uint32_t tcount = 0;
        bool ok = true;
        for (int tile = pair; tile < p.n_tiles && ok; tile += n_pairs, tcount++) {
            const TileCoord t = tile_of(tile);
            const int hw = t.h0 + 4 * q, ww = t.w0 + 8 * m;
            const int ph_ = hw + (lane >> 3), pw = ww + (lane & 7);
            const bool inside = ph_ < p.H && pw < p.W;
            const long long mask_idx = (((long long)t.img * p.H + ph_) * p.W + pw) * p.mask_words + (t.n0 >> 5);
            const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
            if (!mbar_wait(&tfull_bar[acc], aph, ab)) { ok = false; break; }
            tc_fence_after();
            constexpr int kChunks = NT / 32;
#pragma unroll 1
            for (int ch = 0; ch < kChunks; ch++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + m * NT + ch * 32, v);
                uint32_t mword = 0xFFFFFFFFu;
                if (p.mask_in && inside) mword = __ldg(p.mask_in + mask_idx + ch);
                tmem_wait_ld();
                if (ch == kChunks - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
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
                if (p.mask_in) {
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = (mword >> j) & 1u ? f[j] : 0.f;
                }
                if (p.mask_out) {
                    uint32_t w = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) w |= (f[j] > 0.f ? 1u : 0u) << j;
                    if (inside) p.mask_out[mask_idx + ch] = w;
                }
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + row_off + ((j ^ sw) << 4)), "f"(f[4 * j]), "f"(f[4 * j + 1]),
                                 "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&map_out, sbuf, t.n0 + ch * 32, ww, hw, t.img);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // neither CTA leaves (or frees tensor memory) while its peer may still signal / read it
    if (threadIdx.x == 0 && abort_flag && p.status) atomicOr(p.status, GOM_STATUS_TIMEOUT);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------ helpers
__global__ void k_pack_weights(GomConvPackArgs a) {
    const int K = a.c_out, C = a.c_in;
    const int taps = a.kernel_size == 1 ? 1 : 9;
    const long long total = (long long)taps * K * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        float w;
        if (!a.transpose) {          // packed[tap][k][c] = W[k][c][tap]
            const int c = (int)(e % C), k = (int)((e / C) % K), tap = (int)(e / ((long long)C * K));
            w = a.weight[((long long)k * C + c) * taps + tap];
        } else {                     // packed[tap][c][k] = W[k][c][8 - tap]
            const int k = (int)(e % K), c = (int)((e / K) % C), tap = (int)(e / ((long long)C * K));
            w = a.weight[((long long)k * C + c) * taps + (taps - 1 - tap)];
        }
        uint32_t hi, lo;
        split_tf32(w, hi, lo);
        a.packed[e] = __uint_as_float(hi);
        if (a.split) a.packed[total + e] = __uint_as_float(lo);
    }
}

// After a K-split convolution: out <- act(out + bias), ReLU bit mask written / applied.  One thread per pixel and 32 channels.
struct FinishDev { long long n_words; int words_per_pixel, relu; float *out; const float *bias; const uint32_t *mask_in; uint32_t *mask_out; };
__global__ void __launch_bounds__(256) k_conv_finish(FinishDev a) {
    gom_pdl_trigger();
    gom_pdl_wait();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_words) return;
    const int wi = (int)(i % a.words_per_pixel);
    float4 *o = reinterpret_cast<float4 *>(a.out + i * 32);
    const uint32_t mw = a.mask_in ? __ldg(a.mask_in + i) : 0xFFFFFFFFu;
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float4 v = o[j];
        if (a.bias) { const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + wi * 32) + j); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        v.x = (mw >> (4 * j)) & 1u ? v.x : 0.f; v.y = (mw >> (4 * j + 1)) & 1u ? v.y : 0.f;
        v.z = (mw >> (4 * j + 2)) & 1u ? v.z : 0.f; v.w = (mw >> (4 * j + 3)) & 1u ? v.w : 0.f;
        w |= (v.x > 0.f ? 1u : 0u) << (4 * j) | (v.y > 0.f ? 1u : 0u) << (4 * j + 1) | (v.z > 0.f ? 1u : 0u) << (4 * j + 2) | (v.w > 0.f ? 1u : 0u) << (4 * j + 3);
        o[j] = v;
    }
    if (a.mask_out) a.mask_out[i] = w;
}

// lo = x - trunc_tf32(x): the part of x the tensor core drops when it reads the fp32 word x as a TF32 operand.
// With col_sum: x is a row-major matrix of n_cols columns (n_cols divides 1024 = the floats one block covers per step, and the
// grid stride is a multiple of it, so a thread meets the same four columns every time); the column sums of x (the bias
// gradient of the Linear layer whose output gradient x is) are accumulated in registers, reduced over the block's row groups
// in shared memory and added to col_sum with one atomic per column and block.
__global__ void __launch_bounds__(256) k_tf32_split(GomTf32SplitArgs a) {
    __shared__ float4 s_sum[256];
    const long long n4 = a.n / 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(a.x) + i);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
        if (a.hi) reinterpret_cast<float4 *>(a.hi)[i] = h;
        reinterpret_cast<float4 *>(a.lo)[i] = l;
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    if (a.col_sum) {
        const int tpr = a.n_cols / 4;                       // threads per row; 256 / tpr row groups in this block
        s_sum[threadIdx.x] = acc;
        __syncthreads();
        if ((int)threadIdx.x < tpr) {
            for (int r = threadIdx.x + tpr; r < 256; r += tpr) {
                const float4 o = s_sum[r];
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            float *c = a.col_sum + 4 * threadIdx.x;
            atomicAdd(c, acc.x); atomicAdd(c + 1, acc.y); atomicAdd(c + 2, acc.z); atomicAdd(c + 3, acc.w);
        }
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

// NHWC activation [N,H,W,C] as a 4-D tensor (C, W, H, N) with a (32, box_w, box_h, 1) box, 128-byte swizzle, zero fill
int make_act_map(CUtensorMap *m, const float *base, int N, int H, int W, int C, int box_w, int box_h, bool tf32_round) {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = g_encode(m, tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gom_set_error("gom_conv3x3: cuTensorMapEncodeTiled (activation) failed: %d", (int)r); return GOM_ERR_CUDA; }
    return GOM_OK;
}
// packed weight [taps][c_out][c_in] as a 3-D tensor (c_in, c_out, taps) with a (32, nt, taps_per_box) box
int make_weight_map(CUtensorMap *m, const float *base, int taps, int c_out, int c_in, int nt, int taps_per_box) {
    const cuuint64_t dims[3] = {(cuuint64_t)c_in, (cuuint64_t)c_out, (cuuint64_t)taps};
    const cuuint64_t strides[2] = {(cuuint64_t)c_in * 4, (cuuint64_t)c_out * c_in * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)nt, (cuuint32_t)taps_per_box};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gom_set_error("gom_conv3x3: cuTensorMapEncodeTiled (weights) failed: %d", (int)r); return GOM_ERR_CUDA; }
    return GOM_OK;
}

template <int NT, int SUB, int TPS, int BST, int TAPS = 9>
int launch_conv(const GomConv3x3Args *p, ConvDev &d, cudaStream_t stream) {
    using Cfg = ConvCfg<NT, SUB, TPS, BST, TAPS>;
    constexpr int PITCH = Cfg::PITCH;
    static bool configured = false;
    if (!configured) {
        GOM_CUDA(cudaFuncSetAttribute(k_conv3x3<NT, SUB, TPS, BST, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        configured = true;
    }
    d.tiles_w = gom_div_up(p->width, Cfg::TILE_W);
    d.tiles_h = gom_div_up(p->height, kTileH);
    d.n_tiles_n = p->c_out / NT;
    const long long n_tiles = (long long)p->n_images * d.tiles_h * d.tiles_w * d.n_tiles_n * d.k_splits;
    GOM_REQUIRE(n_tiles < (1ll << 30), "too many tiles");
    d.n_tiles = (int)n_tiles;
    CUtensorMap ma, malo, mb, mo;
    const bool round = p->tma_round && p->precision == 0;
    if (int rc = make_act_map(&ma, p->x, p->n_images, p->height, p->width, p->c_in, PITCH, Cfg::HALO_H, round)) return rc;
    if (int rc = make_act_map(&malo, p->precision == 1 ? p->x_lo : p->x, p->n_images, p->height, p->width, p->c_in, PITCH, Cfg::HALO_H, false)) return rc;
    if (int rc = make_weight_map(&mb, p->w_packed, (p->precision == 1 ? 2 : 1) * TAPS, p->c_out, p->c_in, NT, TPS)) return rc;
    if (int rc = make_act_map(&mo, p->out, p->n_images, p->height, p->width, p->c_out, 8, 4, false)) return rc;
    const int grid = d.n_tiles < g_sms ? d.n_tiles : g_sms;
    const size_t out_elems = (size_t)p->n_images * p->height * p->width * p->c_out;
    if (d.k_splits > 1) GOM_CUDA(cudaMemsetAsync(p->out, 0, out_elems * sizeof(float), stream));
    GOM_CUDA(gom_launch_pdl(k_conv3x3<NT, SUB, TPS, BST, TAPS>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, stream, ma, malo, mb, mo, d));
    GOM_LAUNCH_CHECK();
    if (d.k_splits > 1) {
        FinishDev f{(long long)(out_elems / 32), p->c_out / 32, p->relu, p->out, p->bias, p->mask_in, p->mask_out};
        GOM_CUDA(gom_launch_pdl(k_conv_finish, dim3(gom_div_up(f.n_words, 256)), dim3(256), 0, stream, f));
        GOM_LAUNCH_CHECK();
    }
    return GOM_OK;
}


template <int NT, int SUB, int TPS, int BST>
int launch_conv_pair(const GomConv3x3Args *p, ConvDev &d, cudaStream_t stream) {
    using Cfg = ConvCfg<NT, SUB, TPS, BST, 9>;
    using PC = PairCfg<NT, SUB, TPS, BST>;
    static bool configured = false;
    if (!configured) {
        GOM_CUDA(cudaFuncSetAttribute(k_conv3x3_pair<NT, SUB, TPS, BST>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC::SMEM));
        configured = true;
    }
    d.tiles_w = gom_div_up(p->width, 2 * Cfg::TILE_W);           // PAIR tiles
    d.tiles_h = gom_div_up(p->height, kTileH);
    d.n_tiles_n = p->c_out / NT;
    d.k_splits = 1;
    d.c_blocks = p->c_in / 32;
    const long long n_tiles = (long long)p->n_images * d.tiles_h * d.tiles_w * d.n_tiles_n;
    GOM_REQUIRE(n_tiles < (1ll << 30), "too many tiles");
    d.n_tiles = (int)n_tiles;
    CUtensorMap ma, malo, mb, mo;
    const bool round = p->tma_round && p->precision == 0;
    if (int rc = make_act_map(&ma, p->x, p->n_images, p->height, p->width, p->c_in, Cfg::PITCH, Cfg::HALO_H, round)) return rc;
    if (int rc = make_act_map(&malo, p->precision == 1 ? p->x_lo : p->x, p->n_images, p->height, p->width, p->c_in, Cfg::PITCH, Cfg::HALO_H, false)) return rc;
    if (int rc = make_weight_map(&mb, p->w_packed, (p->precision == 1 ? 2 : 1) * 9, p->c_out, p->c_in, NT / 2, TPS)) return rc;
    if (int rc = make_act_map(&mo, p->out, p->n_images, p->height, p->width, p->c_out, 8, 4, false)) return rc;
    const int max_pairs = g_sms / 2;
    const int grid = 2 * (d.n_tiles < max_pairs ? d.n_tiles : max_pairs);
    GOM_CUDA(gom_launch_pdl(k_conv3x3_pair<NT, SUB, TPS, BST>, dim3(grid), dim3(Cfg::THREADS), PC::SMEM, stream, ma, malo, mb, mo, d));
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

}  // namespace

extern "C" int gom_conv3x3_pack_weights(const GomConvPackArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->c_out > 0 && p->c_in > 0, "sizes");
    GOM_REQUIRE(p->weight && p->packed, "null pointer");
    GOM_REQUIRE(p->kernel_size == 3 || p->kernel_size == 1, "kernel_size must be 3 or 1");
    const long long total = (p->kernel_size == 1 ? 1ll : 9ll) * p->c_out * p->c_in;
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
    GOM_REQUIRE(!p->col_sum || (p->n_cols >= 4 && p->n_cols % 4 == 0 && 1024 % p->n_cols == 0 && p->n % p->n_cols == 0),
                "col_sum: n_cols must be a multiple of 4 that divides 1024 and n");
    int blocks = gom_div_up(p->n / 4, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (p->col_sum && blocks > 148 * 2) blocks = 148 * 2;       // one atomic per column and block: keep the blocks few
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
    GOM_REQUIRE(p->kernel_size == 3 || p->kernel_size == 1, "kernel_size must be 3 or 1");
    GOM_REQUIRE(p->kernel_size == 3 || p->c_out % 64 == 0, "1x1: c_out must be a multiple of 64");
    GOM_REQUIRE(((uintptr_t)p->x % 16) == 0 && ((uintptr_t)p->out % 16) == 0 && ((uintptr_t)p->w_packed % 16) == 0 &&
                ((uintptr_t)p->x_lo % 16) == 0 && ((uintptr_t)p->bias % 16) == 0, "16-byte alignment");
    if (int rc = conv_setup()) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;

    ConvDev d{};
    d.H = p->height; d.W = p->width;
    d.c_blocks = p->c_in / 32;
    d.k_splits = 1;
    d.n_pass = p->precision == 1 ? 3 : 1;
    d.relu = p->relu;
    d.mask_words = p->c_out / 32;
    d.bias = p->bias;
    d.mask_in = p->mask_in;
    d.mask_out = p->mask_out;
    d.status = p->status;

    if (p->kernel_size == 1) {                  // plain GEMM over pixel rows (the MLP layers): no halo, no tile-shape search
        gom_prof_begin(GOM_PROF_GEMM_TC, stream);
        const int rc1 = p->c_out % 128 == 0 ? launch_conv<128, 2, 1, 4, 1>(p, d, stream) : launch_conv<64, 2, 1, 4, 1>(p, d, stream);
        if (rc1) return rc1;
        gom_prof_end(GOM_PROF_GEMM_TC, stream);
        return GOM_OK;
    }
    const int slot = p->relu ? GOM_PROF_CONV3X3_FWD : GOM_PROF_CONV3X3_DGRAD;
    gom_prof_begin(slot, stream);
    int rc;
    // The halo pitch is exactly the tile width + 2 columns, so the 8-row groups of a tap's view start at arbitrary multiples of
    // 128 B inside the 1024-byte swizzle pattern.  The tensor core handles that because the SWIZZLE_128B XOR is a function of
    // the shared-memory ADDRESS bits (measured on B200: results identical to a 1024-byte-aligned pitch of 24 pixels, and wrong
    // if the start phase is additionally written into the descriptor's base-offset field: profiles/r5_conv_probe_*.json).
    //
    // Tile shape: the largest one that still gives the 148 SMs about two waves of tiles; cost model = waves x MMA cycles of one
    // tile (128 x NT x 8 TF32 MMA: NT / 2 cycles, but at least 48 — below NT = 128 the operand reads from shared memory bind;
    // a single sub-tile does not share its weight tile, measured ~25 % slower per MMA).
    // A layer with too few tiles even of the smallest shape (the 32 x 32 and 64 x 64 layers at one frame per step) additionally
    // splits the channel blocks of a tile over KS CTAs (partial sums meet in global memory through TMA reduce-add stores; a
    // small finishing kernel applies bias / ReLU / masks): KS x more tiles of 1 / KS the length.
    struct Shape { int nt, sub, ks; };
    const int total_cb = p->c_in / 32;
    Shape best = {p->c_out % 128 == 0 ? 128 : p->c_out % 64 == 0 ? 64 : 32, 2, 1};
    // GOM_CONV_KSPLIT=0: never split K.  Every unsplit shape (single CTA or pair, any NT / SUB) accumulates an output element over
    // the same sequence of MMAs, so results are bit-identical whatever the number of images in the launch; a K-split adds its
    // partial sums in another order, which moves results at the 1e-5 level of the tensor core's fp32 accumulation — per-image
    // results then depend on the batch size through the shape choice (tests/host_harness/dist_grad_check.py pins this switch).
    const char *ks_env = getenv("GOM_CONV_KSPLIT");
    const int max_ks = (ks_env && atoi(ks_env) == 0) ? 1 : 8;
    if (best.nt >= 64) {
        double best_cost = 1e30;
        for (int nt = 128; nt >= 64; nt /= 2)
            for (int sub = 2; sub >= 1; sub--)
                for (int ks = 1; ks <= max_ks; ks *= 2) {
                    if (p->c_out % nt || total_cb % ks || total_cb / ks < 2) continue;
                    const long long tiles = (long long)p->n_images * gom_div_up(p->height, kTileH) * gom_div_up(p->width, 8 * sub) * (p->c_out / nt) * ks;
                    const long long waves = (tiles + g_sms - 1) / g_sms;
                    // measured on B200 (tools/conv_overhead.py, tools/conv_shape_sweep.py): a 128 x NT x 8 TF32 MMA takes NT / 2
                    // cycles (64 at NT = 128; 48, not 32, at NT = 64: operand reads from shared memory bind), ~20 % more in the
                    // stream of a real tile (barrier round trips), another ~25 % when a single sub-tile does not share its weight
                    // tile; a launch costs ~7.5 us (14 000 cycles) before / after its MMAs, a K-split ~10 us more (memset,
                    // finishing pass, two launches) plus three passes over the output
                    const double mma = (nt >= 128 ? 64.0 : 48.0) * 1.2 * (sub == 1 ? 1.25 : 1.0);
                    const double out_bytes = (double)p->n_images * p->height * p->width * p->c_out * 4.0;
                    const double finish = ks > 1 ? 19000.0 + 3.0 * out_bytes / 3000.0 : 0.0;
                    const double cost = (double)waves * ((double)(total_cb / ks) * 9 * sub * 4 * mma) + 14000.0 + finish;
                    if (cost < best_cost * 0.97) { best_cost = cost; best = {nt, sub, ks}; }   // prefer the earlier (larger, unsplit) shape on near ties
                }
    }
    bool forced = false;
    if (const char *force = getenv("GOM_CONV_SHAPE")) {
        forced = true;            // tests: "NT,SUB[,KS]" pins the tile shape (ignored if it does not divide)
        int nt = 0, sub = 0, ks = 1;
        const int got = sscanf(force, "%d,%d,%d", &nt, &sub, &ks);
        if (got >= 2 && (nt == 64 || nt == 128) && (sub == 1 || sub == 2) && p->c_out % nt == 0) {
            if (got < 3 || ks < 1 || total_cb % ks) ks = 1;
            best = {nt, sub, ks};
        } else if (got >= 2 && nt == 256 && sub == 1 && p->c_out % 256 == 0) {      // CTA pairs only (see below)
            best = {256, 1, 1};
        }
    }
    d.k_splits = best.ks;
    d.c_blocks = total_cb / best.ks;
    // CTA pairs (cta_group::2, k_conv3x3_pair) for every shape without a K-split, when the image is at least one pair tile wide;
    // GOM_CONV_PAIR=0 keeps the single-CTA kernel, 2 = pairs without the 256-channel tiles (A/B measurements: profiles/)
    const char *pair_env = getenv("GOM_CONV_PAIR");
    const int pair_mode = pair_env ? atoi(pair_env) : 1;
    const bool pair_ok = pair_mode > 0 && best.ks == 1 && best.nt >= 64 && p->width >= 16 * best.sub;
    if (best.nt == 256 && !pair_ok) best = {128, 2, 1};
    // 256 output channels per pair tile (one sub-tile per CTA, the same number of work units): a CTA then reads 8 KB of operands
    // per 128-cycle MMA instead of 6 KB per 64-cycle MMA (measured: -0.6 % of the step; GOM_CONV_PAIR=2 keeps NT = 128)
    if (pair_mode == 1 && !forced && pair_ok && best.nt == 128 && best.sub == 2 && p->c_out % 256 == 0) best = {256, 1, 1};
    if (pair_ok && best.nt == 256) rc = launch_conv_pair<256, 1, 1, 5>(p, d, stream);
    else if (pair_ok && best.nt == 128 && best.sub == 2) rc = launch_conv_pair<128, 2, 1, 5>(p, d, stream);
    else if (pair_ok && best.nt == 128) rc = launch_conv_pair<128, 1, 3, 3>(p, d, stream);
    else if (pair_ok && best.sub == 2) rc = launch_conv_pair<64, 2, 3, 4>(p, d, stream);
    else if (pair_ok) rc = launch_conv_pair<64, 1, 3, 4>(p, d, stream);
    else if (best.nt == 128 && best.sub == 2) rc = launch_conv<128, 2, 1, 5>(p, d, stream);
    else if (best.nt == 128) rc = launch_conv<128, 1, 3, 3>(p, d, stream);
    else if (best.nt == 64 && best.sub == 2) rc = launch_conv<64, 2, 3, 4>(p, d, stream);
    else if (best.nt == 64) rc = launch_conv<64, 1, 3, 4>(p, d, stream);
    else rc = launch_conv<32, 2, 3, 6>(p, d, stream);
    if (rc) return rc;
    gom_prof_end(slot, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_conv3x3_args(void) { return sizeof(GomConv3x3Args); }
extern "C" size_t gom_sizeof_conv_pack_args(void) { return sizeof(GomConvPackArgs); }
extern "C" size_t gom_sizeof_tf32_split_args(void) { return sizeof(GomTf32SplitArgs); }

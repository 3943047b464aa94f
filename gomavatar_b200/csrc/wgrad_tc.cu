// wgrad_tc.cu — weight gradient of a Linear layer over a tall batch of rows on the tcgen05 tensor cores
// (gom_linear_wgrad, include/gom_b200.h):   gw[m, n] += sum_r g[r, m] * x[r, n],   r over ~1e5 rows.
// Reference: the `weight.grad` torch autograd forms for every nn.Linear of models/modules/non_rigid_module.py:75-147
// (width 128, depth 6, one [B * V, 128] activation per layer).
//
// This is a GEMM whose contraction runs over the ROWS of two row-major matrices, i.e. both operands are "MN-major" for
// the tensor core: a k-block is 32 rows, and a 32-row x 32-column box fetched by TMA (one row of 32 floats = one 128-byte line)
// with the swizzle mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks XORed with the row index, period 4 rows) is
// exactly the one canonical shared-memory layout the tensor core accepts for MN-major 32-bit operands (UMMA layout type
// SWIZZLE_128B_BASE32B: LBO = distance between 32-column boxes, SBO = 512 B between groups of 4 k-rows, 8 k-rows = one MMA;
// with the ordinary 16-byte-chunk SWIZZLE_128B the MMA silently produces zeros — measured).  No transposition anywhere.
//   * split-K over CTAs: every CTA owns a contiguous range of k-blocks, accumulates the whole m x n product (m = 128 lanes x
//     n <= 256 fp32 columns of tensor memory) and adds it to the output with TMA reduce-add stores (the output is zeroed first);
//   * 3xTF32 (fp32-GEMM accuracy, what nn.Linear's backward computes in the reference): the tensor core reads the fp32 words
//     of g and x as TF32 (truncation), and the dropped low parts g_lo, x_lo (gom_tf32_split, already formed for the input
//     gradient / the forward of the same layer) arrive as two more operand tiles: g x + g_lo x + g x_lo;
//   * the kernel is bound by HBM (4 operand matrices streamed once, ~250 MB per 128 x 128 layer at 120 k rows), the MMAs of a
//     k-block (12 x 128 x n x 8) take a quarter of its TMA time.
// Warp roles (192 threads, one CTA per SM): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 epilogue.
#include <cuda.h>

#include "gom_common.cuh"
#include "gom_tcgen05.cuh"

namespace {

using namespace gomtc;

constexpr int kRowsPerBlock = 32;                   // rows (k) per pipeline stage
constexpr int kBoxBytes = kRowsPerBlock * 128;      // one 32-row x 32-column operand box
constexpr int kM = 128;                             // output rows = columns of g (one UMMA M)
constexpr int kMaxN = 256;
constexpr int kThreads = 192;
constexpr int kEpiWarpBytes = 4096;

struct WgradDev {
    int n_kblocks, kb_per_cta, n;
    uint32_t *status;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_tf32_d(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// MN-major SWIZZLE_128B_BASE32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor; atom = 4 k-rows of 128 B = 32
// elements of the M / N dimension, Swizzle<2,5,2> on the byte address): start address >> 4, LBO = bytes between 32-element
// blocks of the M / N dimension, SBO = bytes between groups of 4 k-rows, version 1 (bit 46), layout type 1 (bits 61-63).
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
// instruction descriptor: fp32 accumulate, TF32 x TF32, A and B MN-major (bits 15, 16), N, M = 128
__host__ __device__ constexpr uint32_t instr_desc_mn(int n) { return instr_desc_n(n) | (1u << 15) | (1u << 16); }

template <int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
k_linear_wgrad(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_g_lo,
               const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x_lo,
               const __grid_constant__ CUtensorMap map_out, const WgradDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int abort_flag;

    const int n_blocks_b = p.n / 32;                                // 32-column boxes of x per stage
    const uint32_t a_bytes = (kM / 32) * kBoxBytes, b_bytes = n_blocks_b * kBoxBytes;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;        // g, g_lo, x, x_lo
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + STAGES * stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&done_bar, 1);
        abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kMaxN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    volatile int *ab = &abort_flag;

    const int kb0 = blockIdx.x * p.kb_per_cta;
    const int kb1 = min(p.n_kblocks, kb0 + p.kb_per_cta);
    const int n_kb = kb1 - kb0;                                     // > 0 by construction of the grid

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------- TMA producer
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        }
        uint32_t st = 0, ph = 0;
        for (int i = 0; i < n_kb; i++) {
            if (!mbar_wait(&empty_bar[st], ph ^ 1u, ab)) break;
            if (elect_one()) {
                const int row = (kb0 + i) * kRowsPerBlock;
                const uint32_t base = smem_base + st * stage_bytes;
                mbar_expect_tx(&full_bar[st], stage_bytes);
                for (int j = 0; j < kM / 32; j++) {
                    tma_load_2d(base + j * kBoxBytes, &map_g, j * 32, row, &full_bar[st]);
                    tma_load_2d(base + a_bytes + j * kBoxBytes, &map_g_lo, j * 32, row, &full_bar[st]);
                }
                for (int j = 0; j < n_blocks_b; j++) {
                    tma_load_2d(base + 2 * a_bytes + j * kBoxBytes, &map_x, j * 32, row, &full_bar[st]);
                    tma_load_2d(base + 2 * a_bytes + b_bytes + j * kBoxBytes, &map_x_lo, j * 32, row, &full_bar[st]);
                }
            }
            __syncwarp();
            if (++st == STAGES) { st = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // --------------------------------------------------------------------------------------------- MMA issuer
        const uint32_t idesc = instr_desc_mn(p.n);
        uint32_t st = 0, ph = 0;
        for (int i = 0; i < n_kb; i++) {
            if (!mbar_wait(&full_bar[st], ph, ab)) break;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t base = smem_base + st * stage_bytes;
#pragma unroll
                for (int ks = 0; ks < kRowsPerBlock / 8; ks++) {           // 8 rows (one swizzle atom of every box) per MMA
                    const uint64_t a_hi = make_mn_desc(base + ks * 1024, kBoxBytes, 512);
                    const uint64_t a_lo = make_mn_desc(base + a_bytes + ks * 1024, kBoxBytes, 512);
                    const uint64_t b_hi = make_mn_desc(base + 2 * a_bytes + ks * 1024, kBoxBytes, 512);
                    const uint64_t b_lo = make_mn_desc(base + 2 * a_bytes + b_bytes + ks * 1024, kBoxBytes, 512);
                    mma_tf32_d(tmem, a_hi, b_hi, idesc, (uint32_t)(i != 0 || ks != 0));
                    mma_tf32_d(tmem, a_lo, b_hi, idesc, 1u);
                    mma_tf32_d(tmem, a_hi, b_lo, idesc, 1u);
                }
                tc_commit(&empty_bar[st]);
                if (i == n_kb - 1) tc_commit(&done_bar);
            }
            __syncwarp();
            if (++st == STAGES) { st = 0; ph ^= 1u; }
        }
    } else {
        // ----------------------------------------------------- epilogue: tensor memory -> shared memory -> TMA reduce-add
        const int q = warp & 3;                                     // TMEM lane quarter of this warp = output rows 32 q .. 32 q + 31
        const uint32_t sbuf = epi_base + (warp - 2) * kEpiWarpBytes;
        const int sw = lane & 7;
        if (mbar_wait(&done_bar, 0u, ab)) {
            tc_fence_after();
            for (int ch = 0; ch < p.n / 32; ch++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ch * 32, v);
                tmem_wait_ld();
                if (lane == 0) bulk_wait_read0();                   // the staging buffer of the previous chunk has been read
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + lane * 128 + ((j ^ sw) << 4)), "r"(v[4 * j]),
                                 "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_reduce_add_2d(&map_out, sbuf, ch * 32, q * 32);
                    bulk_commit();
                }
            }
            if (lane == 0) bulk_wait_all();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && abort_flag && p.status) atomicOr(p.status, GOM_STATUS_TIMEOUT);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kMaxN) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_sms = 0;

int wgrad_setup(void) {
    if (g_encode && g_sms) return GOM_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GOM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        gom_set_error("gom_linear_wgrad: cuTensorMapEncodeTiled is not available from this driver");
        return GOM_ERR_UNSUPPORTED;
    }
    int dev = 0, sms = 0;
    GOM_CUDA(cudaGetDevice(&dev));
    GOM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_encode = (EncodeTiledFn)fn;
    g_sms = sms;
    return GOM_OK;
}

// row-major [rows, cols] fp32 matrix, 32-column x box_rows boxes, 128-byte swizzle, zero fill beyond the last row
int make_matrix_map(CUtensorMap *m, const float *base, long long rows, int cols, int box_rows, CUtensorMapSwizzle swizzle) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gom_set_error("gom_linear_wgrad: cuTensorMapEncodeTiled failed: %d", (int)r); return GOM_ERR_CUDA; }
    return GOM_OK;
}

template <int STAGES> int launch_wgrad(const CUtensorMap &mg, const CUtensorMap &mgl, const CUtensorMap &mx, const CUtensorMap &mxl,
                                       const CUtensorMap &mo, const WgradDev &d, int grid, cudaStream_t stream) {
    const int stage_bytes = 2 * (kM / 32) * kBoxBytes + 2 * (d.n / 32) * kBoxBytes;
    const int smem = STAGES * stage_bytes + 4 * kEpiWarpBytes + 1024;
    static int configured = 0;
    if (configured < smem) {
        GOM_CUDA(cudaFuncSetAttribute(k_linear_wgrad<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    k_linear_wgrad<STAGES><<<grid, kThreads, smem, stream>>>(mg, mgl, mx, mxl, mo, d);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

}  // namespace

extern "C" int gom_linear_wgrad(const GomLinearWgradArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->rows > 0, "rows");
    GOM_REQUIRE(p->m == kM, "m (columns of g) must be 128");
    GOM_REQUIRE(p->n > 0 && p->n % 32 == 0 && p->n <= kMaxN, "n (columns of x) must be a multiple of 32, at most 256");
    GOM_REQUIRE(p->g && p->g_lo && p->x && p->x_lo && p->out, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->g % 16) == 0 && ((uintptr_t)p->g_lo % 16) == 0 && ((uintptr_t)p->x % 16) == 0 &&
                ((uintptr_t)p->x_lo % 16) == 0 && ((uintptr_t)p->out % 16) == 0, "16-byte alignment");
    if (int rc = wgrad_setup()) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    CUtensorMap mg, mgl, mx, mxl, mo;
    if (int rc = make_matrix_map(&mg, p->g, p->rows, p->m, kRowsPerBlock, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_matrix_map(&mgl, p->g_lo, p->rows, p->m, kRowsPerBlock, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_matrix_map(&mx, p->x, p->rows, p->n, kRowsPerBlock, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_matrix_map(&mxl, p->x_lo, p->rows, p->n, kRowsPerBlock, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_matrix_map(&mo, p->out, p->m, p->n, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    WgradDev d{};
    d.n = p->n;
    d.n_kblocks = gom_div_up(p->rows, kRowsPerBlock);
    // at least 4 k-blocks per CTA (below that the reduce-add of the 128 x n partial product costs more than its MMAs)
    int grid = g_sms;
    if (d.n_kblocks < 4 * grid) grid = gom_div_up(d.n_kblocks, 4);
    d.kb_per_cta = gom_div_up(d.n_kblocks, grid);
    grid = gom_div_up(d.n_kblocks, d.kb_per_cta);                    // no CTA without work
    d.status = p->status;
    gom_prof_begin(GOM_PROF_WGRAD_TC, stream);
    if (p->zero_first) GOM_CUDA(cudaMemsetAsync(p->out, 0, sizeof(float) * (size_t)p->m * p->n, stream));
    // stage = 32 rows of g, g_lo (2 x 16 KB) and of x, x_lo (2 x n / 32 x 4 KB): 64 KB at n = 128 (3 stages), 96 KB at n = 256 (2)
    const int rc = p->n <= 128 ? launch_wgrad<3>(mg, mgl, mx, mxl, mo, d, grid, stream) : launch_wgrad<2>(mg, mgl, mx, mxl, mo, d, grid, stream);
    if (rc) return rc;
    gom_prof_end(GOM_PROF_WGRAD_TC, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_linear_wgrad_args(void) { return sizeof(GomLinearWgradArgs); }

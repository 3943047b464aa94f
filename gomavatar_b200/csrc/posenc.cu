// posenc.cu — input of the non-rigid deformation MLP (reference models/modules/non_rigid_module.py:15-72,128-140) in one launch
// each way: per (frame, vertex) row  [ pose vector (C) | Hann-windowed positional encoding of the canonical vertex (6 L) | 0 pad ]
// written straight in the padded row-major form the tensor-core GEMM of csrc/conv3x3_tc.cu reads, plus the encoding alone
// (re-read by the skip layer).  torch needs ~12 elementwise / stack / cat / pad kernels forward and ~20 backward for the same.
//   enc[k * 6 + s * 3 + d] = w_k * (s == 0 ? sin : cos)(2^k x_d),  w_k = (1 - cos(pi clamp(alpha - k, 0, 1))) / 2
#include <math.h>

#include "gom_common.cuh"

namespace {

constexpr int kRows = 64;            // rows per block
constexpr int kThreads = 256;
constexpr int kMaxL = 10;

__device__ __forceinline__ float hann(float alpha, int k) {
    return (1.0f - cosf(3.14159265358979323846f * fminf(fmaxf(alpha - (float)k, 0.f), 1.f))) * 0.5f;
}

__global__ void __launch_bounds__(kThreads) k_nonrigid_input_fwd(GomNonRigidInputArgs a) {
    __shared__ float s_enc[kRows][6 * kMaxL + 1];
    const int E = 6 * a.multires;
    const long long R = (long long)a.n_frames * a.n_verts, r0 = (long long)blockIdx.x * kRows;
    if (threadIdx.x < kRows) {
        const long long r = r0 + threadIdx.x;
        if (r < R) {
            const int b = (int)(r / a.n_verts), v = (int)(r - (long long)b * a.n_verts);
            const float *x = a.xyz + (long long)(a.xyz_frames == 1 ? 0 : b) * 3 * a.n_verts;
            const float p[3] = {x[v], x[a.n_verts + v], x[2 * a.n_verts + v]};
            float f = 1.f;
            for (int k = 0; k < a.multires; k++, f *= 2.f) {
                const float w = hann(a.alpha, k);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    float s, c;
                    sincosf(p[d] * f, &s, &c);
                    s_enc[threadIdx.x][k * 6 + d] = s * w;
                    s_enc[threadIdx.x][k * 6 + 3 + d] = c * w;
                }
            }
        } else {
            for (int j = 0; j < E; j++) s_enc[threadIdx.x][j] = 0.f;
        }
    }
    __syncthreads();
    const int C = a.cond, cols = a.cols;
    for (long long i = threadIdx.x; i < (long long)kRows * cols; i += kThreads) {     // coalesced rows of the padded matrix
        const int lr = (int)(i / cols), col = (int)(i - (long long)lr * cols);
        const long long r = r0 + lr;
        if (r >= a.rows_padded) break;
        float val = 0.f;
        if (r < R) {
            if (col < C) val = a.posevec[(r / a.n_verts) * C + col];
            else if (col < C + E) val = s_enc[lr][col - C];
        }
        a.h0[r * cols + col] = val;
    }
    for (int i = threadIdx.x; i < kRows * E; i += kThreads) {
        const int lr = i / E, j = i - lr * E;
        const long long r = r0 + lr;
        if (r < a.rows_padded) a.enc[r * E + j] = s_enc[lr][j];
    }
}

__global__ void __launch_bounds__(kThreads) k_nonrigid_input_bwd(GomNonRigidInputArgs a) {
    __shared__ float s_g[kRows][6 * kMaxL + 1];
    const int E = 6 * a.multires, C = a.cond, cols = a.cols;
    const long long R = (long long)a.n_frames * a.n_verts, r0 = (long long)blockIdx.x * kRows;
    for (int i = threadIdx.x; i < kRows * E; i += kThreads) {
        const int lr = i / E, j = i - lr * E;
        const long long r = r0 + lr;
        float g = 0.f;
        if (r < R) {
            if (a.g_h0) g += a.g_h0[r * cols + C + j];
            if (a.g_enc) g += a.g_enc[r * E + j];
        }
        s_g[lr][j] = g;
    }
    __syncthreads();
    if (threadIdx.x >= kRows) return;
    const long long r = r0 + threadIdx.x;
    if (r >= R) return;
    const int b = (int)(r / a.n_verts), v = (int)(r - (long long)b * a.n_verts);
    const int bx = a.xyz_frames == 1 ? 0 : b;
    const float *x = a.xyz + (long long)bx * 3 * a.n_verts;
    const float p[3] = {x[v], x[a.n_verts + v], x[2 * a.n_verts + v]};
    float gx[3] = {0.f, 0.f, 0.f};
    float f = 1.f;
    for (int k = 0; k < a.multires; k++, f *= 2.f) {
        const float wf = hann(a.alpha, k) * f;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float s, c;
            sincosf(p[d] * f, &s, &c);
            gx[d] += wf * (s_g[threadIdx.x][k * 6 + d] * c - s_g[threadIdx.x][k * 6 + 3 + d] * s);
        }
    }
    float *o = a.g_xyz + (long long)bx * 3 * a.n_verts;
    if (a.xyz_frames == 1 && a.n_frames > 1) {       // one canonical vertex set shared by all frames: the frames' gradients add up
        atomicAdd(o + v, gx[0]); atomicAdd(o + a.n_verts + v, gx[1]); atomicAdd(o + 2 * a.n_verts + v, gx[2]);
    } else {
        o[v] = gx[0]; o[a.n_verts + v] = gx[1]; o[2 * a.n_verts + v] = gx[2];
    }
}

int check(const GomNonRigidInputArgs *p) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_verts > 0 && (p->xyz_frames == 1 || p->xyz_frames == p->n_frames), "sizes");
    GOM_REQUIRE(p->multires >= 1 && p->multires <= kMaxL && p->cond >= 0, "multires / cond");
    GOM_REQUIRE(p->cols >= p->cond + 6 * p->multires, "cols must hold the pose vector and the encoding");
    GOM_REQUIRE(p->rows_padded >= (long long)p->n_frames * p->n_verts, "rows_padded");
    GOM_REQUIRE(p->xyz, "null pointer");
    return GOM_OK;
}

}  // namespace

extern "C" int gom_nonrigid_input_forward(const GomNonRigidInputArgs *p, gom_stream_t stream) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE(p->posevec && p->h0 && p->enc, "null pointer");
    k_nonrigid_input_fwd<<<gom_div_up(p->rows_padded, kRows), kThreads, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" int gom_nonrigid_input_backward(const GomNonRigidInputArgs *p, gom_stream_t stream_) {
    if (int rc = check(p)) return rc;
    GOM_REQUIRE((p->g_h0 || p->g_enc) && p->g_xyz, "null gradient pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (p->xyz_frames == 1 && p->n_frames > 1) GOM_CUDA(cudaMemsetAsync(p->g_xyz, 0, sizeof(float) * 3 * (size_t)p->n_verts, stream));
    k_nonrigid_input_bwd<<<gom_div_up((long long)p->n_frames * p->n_verts, kRows), kThreads, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

extern "C" size_t gom_sizeof_nonrigid_input_args(void) { return sizeof(GomNonRigidInputArgs); }

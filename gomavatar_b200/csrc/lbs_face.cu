// lbs_face.cu — skeleton chain, linear-blend skinning and Gaussians-on-mesh transform, forward + backward (sm_100a).
//
// Replaces reference utils/body_util.py:591-644 (get_global_RTs, apply_lbs) and models/model.py:27-41,225-234
// (Steiner frame, so3 exp, covariance) — ~90 small torch launches and a B x J x 3 x V temporary per frame — with six
// kernels that read the model's own SoA buffers ([3,V], [J+1,V], [3,F]) once.  A batch of B frames shares one read of
// the skinning weights; weight tiles are staged through shared memory with TMA bulk copies.
#include "gom_common.cuh"
#include "gom_face.cuh"
#include "gom_joints.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxJ = GOM_MAX_JOINTS;

// =================================================================================================== skeleton chain
struct JointDev {
    int B, J;
    const int32_t *parents;
    const float *cnl, *dst_Rs, *dst_Ts;
    float *out_Rs, *out_Ts, *chain_G, *cnl_inv;
    const float *dRs, *dTs;
    float *d_dst_Rs, *d_dst_Ts;
};

__device__ __forceinline__ void load_local(const float *Rs, const float *Ts, float L[12]) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
        L[4 * r + 0] = Rs[3 * r + 0]; L[4 * r + 1] = Rs[3 * r + 1]; L[4 * r + 2] = Rs[3 * r + 2];
        L[4 * r + 3] = Ts[r];
    }
}

__global__ void __launch_bounds__(kMaxJ) k_joint_fwd(JointDev a) {
    __shared__ float sL[kMaxJ][12];
    __shared__ int spar[kMaxJ];
    const int b = blockIdx.x, j = threadIdx.x;
    const long long bj = (long long)b * a.J + j;
    float cinv[16];
    if (j < a.J) {
        float L[12];
        load_local(a.dst_Rs + 9 * bj, a.dst_Ts + 3 * bj, L);
#pragma unroll
        for (int k = 0; k < 12; k++) sL[j][k] = L[k];
        spar[j] = a.parents[j];
        float m[16];
#pragma unroll
        for (int k = 0; k < 16; k++) m[k] = a.cnl[16 * bj + k];
        gomjoint::inverse4x4(m, cinv);
#pragma unroll
        for (int k = 0; k < 16; k++) a.cnl_inv[16 * bj + k] = cinv[k];
    }
    __syncthreads();
    if (j >= a.J) return;
    // ancestors root..j, multiplied left to right exactly like the reference's G_i = G_parent(i) · L_i
    unsigned char path[kMaxJ];
    int n = 0;
    for (int k = j; k >= 0 && n < kMaxJ; k = spar[k]) path[n++] = (unsigned char)k;
    float G[12];
#pragma unroll
    for (int k = 0; k < 12; k++) G[k] = sL[path[n - 1]][k];
    for (int i = n - 2; i >= 0; i--) {
        float Lk[12], C[12];
#pragma unroll
        for (int k = 0; k < 12; k++) Lk[k] = sL[path[i]][k];
        gomjoint::affine_mul(G, Lk, C);
#pragma unroll
        for (int k = 0; k < 12; k++) G[k] = C[k];
    }
#pragma unroll
    for (int k = 0; k < 12; k++) a.chain_G[12 * bj + k] = G[k];
    // F = G · inv(cnl) (general 4x4 on the right); G's bottom row is [0 0 0 1]
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float f[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            f[c] = G[4 * r] * cinv[c] + G[4 * r + 1] * cinv[4 + c] + G[4 * r + 2] * cinv[8 + c] + G[4 * r + 3] * cinv[12 + c];
        a.out_Rs[9 * bj + 3 * r + 0] = f[0]; a.out_Rs[9 * bj + 3 * r + 1] = f[1]; a.out_Rs[9 * bj + 3 * r + 2] = f[2];
        a.out_Ts[3 * bj + r] = f[3];
    }
}

__global__ void __launch_bounds__(kMaxJ) k_joint_bwd(JointDev a) {
    __shared__ float sdG[kMaxJ][12];
    const int b = blockIdx.x, j = threadIdx.x;
    const long long bj = (long long)b * a.J + j;
    if (j < a.J) {     // dG = dF · inv(cnl)^T   (top three rows)
        const float *ci = a.cnl_inv + 16 * bj;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const float d0 = a.dRs[9 * bj + 3 * r], d1 = a.dRs[9 * bj + 3 * r + 1], d2 = a.dRs[9 * bj + 3 * r + 2];
            const float d3 = a.dTs[3 * bj + r];
#pragma unroll
            for (int k = 0; k < 4; k++)
                sdG[j][4 * r + k] = d0 * ci[4 * k] + d1 * ci[4 * k + 1] + d2 * ci[4 * k + 2] + d3 * ci[4 * k + 3];
        }
    }
    __syncthreads();
    if (j >= 12) return;                       // 12 lanes of warp 0 walk the chain leaf-to-root
    const int r = j >> 2, k = j & 3;           // this lane owns entry (r,k) of a 3x4 block
    for (int i = a.J - 1; i >= 1; i--) {
        const int p = a.parents[i];
        const long long bi = (long long)b * a.J + i, bp = (long long)b * a.J + p;
        float Li[12];
        load_local(a.dst_Rs + 9 * bi, a.dst_Ts + 3 * bi, Li);
        const float *dGi = sdG[i];
        // dG_p[r][k] += sum_c dG_i[r][c] L_i[k][c]   (L_i row 3 = [0 0 0 1])
        float add = (k < 3) ? dGi[4 * r] * Li[4 * k] + dGi[4 * r + 1] * Li[4 * k + 1] + dGi[4 * r + 2] * Li[4 * k + 2] +
                                  dGi[4 * r + 3] * Li[4 * k + 3]
                            : dGi[4 * r + 3];
        // dL_i[r][k] = sum_q G_p[q][r] dG_i[q][k]     (rows r < 3 of L_i are variables)
        const float *Gp = a.chain_G + 12 * bp;
        const float dl = Gp[r] * dGi[k] + Gp[4 + r] * dGi[4 + k] + Gp[8 + r] * dGi[8 + k];
        __syncwarp(0xfffu);
        sdG[p][4 * r + k] += add;
        if (k < 3) a.d_dst_Rs[9 * bi + 3 * r + k] = dl;
        else a.d_dst_Ts[3 * bi + r] = dl;
        __syncwarp(0xfffu);
    }
    const long long b0 = (long long)b * a.J;
    const float dl0 = sdG[0][4 * r + k];       // root: L_0 = G_0
    if (k < 3) a.d_dst_Rs[9 * b0 + 3 * r + k] = dl0;
    else a.d_dst_Ts[3 * b0 + r] = dl0;
}

// ============================================================================================ linear-blend skinning
constexpr int kTile = 256;        // vertices per block
constexpr int kWRow = kTile + 8;  // smem row pitch of a staged weight row (room for the 16-B alignment lead)
constexpr int kFB = 8;            // frames per block

struct LbsDev {
    int B, J, V, use_tma;
    const float *xyz; long long xyz_stride;
    const float *w, *Rs, *Ts;
    float *out;
    const float *dout;
    float *dxyz; long long dxyz_stride;
    float *dRs, *dTs;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage rows 0..J-1 of w[:, v0 : v0+cnt] into sW (pitch kWRow).  TMA path: one cp.async.bulk per row from the
// 16-byte-aligned address at or below the row start (the returned per-row lead is then added when indexing); the
// buffer has J+1 rows, so rounding a row up to 16 B never leaves it.  Falls back to coalesced loads.
__device__ __forceinline__ void stage_weights(const LbsDev &a, float *sW, unsigned long long *bar, int v0, int cnt) {
    const int tid = threadIdx.x;
    if (a.use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t total = 0;
            for (int j = 0; j < a.J; j++) {
                const long long e = (long long)j * a.V + v0;
                total += (uint32_t)((((int)(e & 3) + cnt + 3) & ~3) * 4);
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total) : "memory");
            for (int j = 0; j < a.J; j++) {
                const long long e = (long long)j * a.V + v0;
                const int lead = (int)(e & 3);
                const uint32_t bytes = (uint32_t)(((lead + cnt + 3) & ~3) * 4);
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(sW + j * kWRow)),
                    "l"(a.w + (e - lead)), "r"(bytes), "r"(smem_u32(bar))
                    : "memory");
            }
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar))
                : "memory");
        }
    } else {
        for (int j = 0; j < a.J; j++) {
            const long long e = (long long)j * a.V + v0;
            if (tid < cnt) sW[j * kWRow + (int)(e & 3) + tid] = a.w[e + tid];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) k_lbs_fwd(LbsDev a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sW = reinterpret_cast<float *>(smem_raw);                       // [J][kWRow]
    float *sA = sW + a.J * kWRow;                                          // [kFB][J][12]
    __shared__ unsigned long long bar;
    const int tid = threadIdx.x;
    const int v0 = blockIdx.x * kTile, cnt = min(kTile, a.V - v0);
    const int b0 = blockIdx.y * kFB, nb = min(kFB, a.B - b0);
    for (int i = tid; i < nb * a.J * 12; i += kThreads) {
        const int fb = i / (a.J * 12), rem = i - fb * a.J * 12, j = rem / 12, k = rem - j * 12;
        const long long bj = (long long)(b0 + fb) * a.J + j;
        const int r = k >> 2, c = k & 3;
        sA[i] = (c < 3) ? a.Rs[9 * bj + 3 * r + c] : a.Ts[3 * bj + r];
    }
    stage_weights(a, sW, &bar, v0, cnt);       // ends with a block-wide sync on both paths (barrier wait / syncthreads)
    __syncthreads();
    if (tid >= cnt) return;
    const int v = v0 + tid;
    float acc[kFB][3];
    float px[kFB], py[kFB], pz[kFB];
#pragma unroll
    for (int fb = 0; fb < kFB; fb++) {
        acc[fb][0] = acc[fb][1] = acc[fb][2] = 0.f;
        if (fb < nb) {
            const float *x = a.xyz + (long long)(b0 + fb) * a.xyz_stride;
            px[fb] = x[v]; py[fb] = x[a.V + v]; pz[fb] = x[2LL * a.V + v];
        }
    }
    for (int j = 0; j < a.J; j++) {
        const float w = sW[j * kWRow + (int)(((long long)j * a.V + v0) & 3) + tid];
        if (w == 0.f) continue;
#pragma unroll
        for (int fb = 0; fb < kFB; fb++) {
            if (fb < nb) {
                const float4 *A = reinterpret_cast<const float4 *>(sA + (fb * a.J + j) * 12);
                const float4 r0 = A[0], r1 = A[1], r2 = A[2];
                const float tx = r0.x * px[fb] + r0.y * py[fb] + r0.z * pz[fb] + r0.w;
                const float ty = r1.x * px[fb] + r1.y * py[fb] + r1.z * pz[fb] + r1.w;
                const float tz = r2.x * px[fb] + r2.y * py[fb] + r2.z * pz[fb] + r2.w;
                acc[fb][0] += w * tx; acc[fb][1] += w * ty; acc[fb][2] += w * tz;
            }
        }
    }
#pragma unroll
    for (int fb = 0; fb < kFB; fb++)
        if (fb < nb) {
            float *o = a.out + (long long)(b0 + fb) * 3 * a.V;
            o[v] = acc[fb][0]; o[a.V + v] = acc[fb][1]; o[2LL * a.V + v] = acc[fb][2];
        }
}

__global__ void __launch_bounds__(kThreads) k_lbs_bwd(LbsDev a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sW = reinterpret_cast<float *>(smem_raw);                       // [J][kWRow]
    float *sA = sW + a.J * kWRow;                                          // [kFB][J][12]
    float *sG = sA + kFB * a.J * 12;                                       // [kFB][J][12] pose-gradient accumulators
    __shared__ unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31;
    const int v0 = blockIdx.x * kTile, cnt = min(kTile, a.V - v0);
    const int b0 = blockIdx.y * kFB, nb = min(kFB, a.B - b0);
    const bool pose = a.dRs != nullptr;
    for (int i = tid; i < nb * a.J * 12; i += kThreads) {
        const int fb = i / (a.J * 12), rem = i - fb * a.J * 12, j = rem / 12, k = rem - j * 12;
        const long long bj = (long long)(b0 + fb) * a.J + j;
        const int r = k >> 2, c = k & 3;
        sA[i] = (c < 3) ? a.Rs[9 * bj + 3 * r + c] : a.Ts[3 * bj + r];
        sG[i] = 0.f;
    }
    stage_weights(a, sW, &bar, v0, cnt);
    __syncthreads();
    const bool active = tid < cnt;
    const int v = v0 + (active ? tid : 0);
    const bool shared_xyz = a.dxyz_stride == 0;
    float dsum[3] = {0.f, 0.f, 0.f};
    for (int fb = 0; fb < nb; fb++) {
        const int b = b0 + fb;
        const float *x = a.xyz + (long long)b * a.xyz_stride;
        const float *g = a.dout + (long long)b * 3 * a.V;
        float p[3] = {0.f, 0.f, 0.f}, gr[3] = {0.f, 0.f, 0.f};
        if (active) {
            p[0] = x[v]; p[1] = x[a.V + v]; p[2] = x[2LL * a.V + v];
            gr[0] = g[v]; gr[1] = g[a.V + v]; gr[2] = g[2LL * a.V + v];
        }
        float dv[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < a.J; j++) {
            const float w = active ? sW[j * kWRow + (int)(((long long)j * a.V + v0) & 3) + tid] : 0.f;
            const bool nz = w != 0.f;
            if (nz) {     // dv += w R^T g
                const float *A = sA + (fb * a.J + j) * 12;
                dv[0] += w * (A[0] * gr[0] + A[4] * gr[1] + A[8] * gr[2]);
                dv[1] += w * (A[1] * gr[0] + A[5] * gr[1] + A[9] * gr[2]);
                dv[2] += w * (A[2] * gr[0] + A[6] * gr[1] + A[10] * gr[2]);
            }
            if (pose && __ballot_sync(0xffffffffu, nz)) {   // dR += w g p^T, dT += w g  (warp-reduced first)
                float *G = sG + (fb * a.J + j) * 12;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const float wg = nz ? w * gr[r] : 0.f;
                    const float s0 = warp_sum(wg * p[0]), s1 = warp_sum(wg * p[1]), s2 = warp_sum(wg * p[2]);
                    const float s3 = warp_sum(wg);
                    if (lane == 0) {
                        atomicAdd(G + 4 * r, s0); atomicAdd(G + 4 * r + 1, s1); atomicAdd(G + 4 * r + 2, s2);
                        atomicAdd(G + 4 * r + 3, s3);
                    }
                }
            }
        }
        if (active) {
            if (shared_xyz) { dsum[0] += dv[0]; dsum[1] += dv[1]; dsum[2] += dv[2]; }
            else {
                float *o = a.dxyz + (long long)b * a.dxyz_stride;
                o[v] = dv[0]; o[a.V + v] = dv[1]; o[2LL * a.V + v] = dv[2];
            }
        }
    }
    if (active && shared_xyz) {    // frames of other blocks (blockIdx.y) add into the same [3,V] buffer
        atomicAdd(a.dxyz + v, dsum[0]); atomicAdd(a.dxyz + a.V + v, dsum[1]); atomicAdd(a.dxyz + 2LL * a.V + v, dsum[2]);
    }
    if (pose) {
        __syncthreads();
        for (int i = tid; i < nb * a.J * 12; i += kThreads) {
            const float s = sG[i];
            if (s == 0.f) continue;
            const int fb = i / (a.J * 12), rem = i - fb * a.J * 12, j = rem / 12, k = rem - j * 12;
            const long long bj = (long long)(b0 + fb) * a.J + j;
            const int r = k >> 2, c = k & 3;
            if (c < 3) atomicAdd(a.dRs + 9 * bj + 3 * r + c, s);
            else atomicAdd(a.dTs + 3 * bj + r, s);
        }
    }
}

// ============================================================================================== Gaussians on mesh
struct FaceDev {
    int B, F, V, faces_int64;
    float sigma;
    const float *verts; const void *faces; const float *so3, *scale;
    float *means3D, *cov3D;
    const float *dmeans, *dcov;
    float *dverts, *dso3, *dscale;
};

__device__ __forceinline__ void load_face(const FaceDev &a, int f, int idx[3]) {
    if (a.faces_int64) {
        const long long *p = reinterpret_cast<const long long *>(a.faces) + 3LL * f;
        idx[0] = (int)p[0]; idx[1] = (int)p[1]; idx[2] = (int)p[2];
    } else {
        const int *p = reinterpret_cast<const int *>(a.faces) + 3LL * f;
        idx[0] = p[0]; idx[1] = p[1]; idx[2] = p[2];
    }
}

__device__ __forceinline__ void load_tri(const FaceDev &a, int b, const int idx[3], float v[3][3]) {
    const float *x = a.verts + (long long)b * 3 * a.V;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        v[k][0] = x[idx[k]]; v[k][1] = x[a.V + idx[k]]; v[k][2] = x[2LL * a.V + idx[k]];
    }
}

__global__ void __launch_bounds__(kThreads) k_face_fwd(FaceDev a) {
    const int f = blockIdx.x * kThreads + threadIdx.x, b = blockIdx.y;
    if (f >= a.F) return;
    int idx[3];
    load_face(a, f, idx);
    float v[3][3];
    load_tri(a, b, idx, v);
    const float w[3] = {a.so3[f], a.so3[a.F + f], a.so3[2LL * a.F + f]};
    const float s[3] = {a.scale[f], a.scale[a.F + f], a.scale[2LL * a.F + f]};
    float R[9], L[9], M[9], mean[3], cov6[6];
    gomface::Frame fr;
    gomface::so3_exp(w, R);
    gomface::local_factor(R, s, L);
    gomface::steiner_frame(v[0], v[1], v[2], a.sigma, mean, fr);
    gomface::world_cov(fr.A, L, M, cov6);
    const long long o = (long long)b * a.F + f;
    a.means3D[3 * o] = mean[0]; a.means3D[3 * o + 1] = mean[1]; a.means3D[3 * o + 2] = mean[2];
    float2 *c = reinterpret_cast<float2 *>(a.cov3D + 6 * o);
    c[0] = make_float2(cov6[0], cov6[1]); c[1] = make_float2(cov6[2], cov6[3]); c[2] = make_float2(cov6[4], cov6[5]);
}

__global__ void __launch_bounds__(kThreads) k_face_bwd(FaceDev a) {
    const int f = blockIdx.x * kThreads + threadIdx.x, b = blockIdx.y;
    if (f >= a.F) return;
    int idx[3];
    load_face(a, f, idx);
    float v[3][3];
    load_tri(a, b, idx, v);
    const float w[3] = {a.so3[f], a.so3[a.F + f], a.so3[2LL * a.F + f]};
    const float s[3] = {a.scale[f], a.scale[a.F + f], a.scale[2LL * a.F + f]};
    float R[9], L[9], M[9], mean[3], cov6[6];
    gomface::Frame fr;
    gomface::so3_exp(w, R);
    gomface::local_factor(R, s, L);
    gomface::steiner_frame(v[0], v[1], v[2], a.sigma, mean, fr);
    gomface::world_cov(fr.A, L, M, cov6);
    const long long o = (long long)b * a.F + f;
    const float dmean[3] = {a.dmeans[3 * o], a.dmeans[3 * o + 1], a.dmeans[3 * o + 2]};
    const float2 *gc = reinterpret_cast<const float2 *>(a.dcov + 6 * o);
    const float2 g01 = gc[0], g23 = gc[1], g45 = gc[2];
    const float g[6] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y};
    float dv[3][3], dL[9], dR[9], ds[3], dw[3];
    gomface::face_bwd(fr, L, M, a.sigma, dmean, g, dv[0], dv[1], dv[2], dL);
    gomface::local_factor_bwd(R, s, dL, dR, ds);
    gomface::so3_exp_bwd(w, dR, dw);
    float *gx = a.dverts + (long long)b * 3 * a.V;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        atomicAdd(gx + idx[k], dv[k][0]); atomicAdd(gx + a.V + idx[k], dv[k][1]); atomicAdd(gx + 2LL * a.V + idx[k], dv[k][2]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        atomicAdd(a.dso3 + (long long)k * a.F + f, dw[k]);
        atomicAdd(a.dscale + (long long)k * a.F + f, ds[k]);
    }
}

size_t lbs_smem_bytes(int J, bool bwd) { return sizeof(float) * ((size_t)J * kWRow + (size_t)(bwd ? 2 : 1) * kFB * J * 12); }

}  // namespace

// ===================================================================================================== C ABI
extern "C" int gom_joint_transforms_forward(const GomJointFwdArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_joints > 0 && p->n_joints <= kMaxJ, "n_frames / n_joints (<= 64)");
    GOM_REQUIRE(p->parents && p->cnl_gtfms && p->dst_Rs && p->dst_Ts && p->global_Rs && p->global_Ts && p->chain_G &&
                    p->cnl_inv, "null pointer");
    JointDev a{};
    a.B = p->n_frames; a.J = p->n_joints; a.parents = p->parents; a.cnl = p->cnl_gtfms; a.dst_Rs = p->dst_Rs;
    a.dst_Ts = p->dst_Ts; a.out_Rs = p->global_Rs; a.out_Ts = p->global_Ts; a.chain_G = p->chain_G; a.cnl_inv = p->cnl_inv;
    gom_prof_begin(GOM_PROF_JOINT_FWD, (cudaStream_t)stream);
    k_joint_fwd<<<a.B, kMaxJ, 0, (cudaStream_t)stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_JOINT_FWD, (cudaStream_t)stream);
    return GOM_OK;
}

extern "C" int gom_joint_transforms_backward(const GomJointBwdArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_joints > 0 && p->n_joints <= kMaxJ, "n_frames / n_joints (<= 64)");
    GOM_REQUIRE(p->parents && p->dst_Rs && p->dst_Ts && p->chain_G && p->cnl_inv && p->dL_dglobal_Rs && p->dL_dglobal_Ts &&
                    p->dL_ddst_Rs && p->dL_ddst_Ts, "null pointer");
    JointDev a{};
    a.B = p->n_frames; a.J = p->n_joints; a.parents = p->parents; a.dst_Rs = p->dst_Rs; a.dst_Ts = p->dst_Ts;
    a.chain_G = const_cast<float *>(p->chain_G); a.cnl_inv = const_cast<float *>(p->cnl_inv);
    a.dRs = p->dL_dglobal_Rs; a.dTs = p->dL_dglobal_Ts; a.d_dst_Rs = p->dL_ddst_Rs; a.d_dst_Ts = p->dL_ddst_Ts;
    gom_prof_begin(GOM_PROF_JOINT_BWD, (cudaStream_t)stream);
    k_joint_bwd<<<a.B, kMaxJ, 0, (cudaStream_t)stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_JOINT_BWD, (cudaStream_t)stream);
    return GOM_OK;
}

static int lbs_common(LbsDev &a, bool bwd, cudaStream_t stream) {
    a.use_tma = (((uintptr_t)a.w) % 16 == 0) ? 1 : 0;
    const size_t smem = lbs_smem_bytes(a.J, bwd);
    static bool attr_fwd = false, attr_bwd = false;
    if (smem > 48 * 1024) {
        bool &done = bwd ? attr_bwd : attr_fwd;
        if (!done) {
            GOM_CUDA(cudaFuncSetAttribute(bwd ? k_lbs_bwd : k_lbs_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            done = true;
        }
    }
    dim3 grid(gom_div_up(a.V, kTile), gom_div_up(a.B, kFB));
    gom_prof_begin(bwd ? GOM_PROF_LBS_BWD : GOM_PROF_LBS_FWD, stream);
    if (bwd) k_lbs_bwd<<<grid, kThreads, smem, stream>>>(a);
    else k_lbs_fwd<<<grid, kThreads, smem, stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(bwd ? GOM_PROF_LBS_BWD : GOM_PROF_LBS_FWD, stream);
    return GOM_OK;
}

extern "C" int gom_lbs_forward(const GomLbsFwdArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_verts > 0 && p->n_joints > 0 && p->n_joints <= kMaxJ, "sizes");
    GOM_REQUIRE(p->n_frames <= 65535 * kFB, "n_frames");
    GOM_REQUIRE(p->xyz && p->lbs_weights && p->global_Rs && p->global_Ts && p->out, "null pointer");
    LbsDev a{};
    a.B = p->n_frames; a.J = p->n_joints; a.V = p->n_verts; a.xyz = p->xyz; a.xyz_stride = p->xyz_stride;
    a.w = p->lbs_weights; a.Rs = p->global_Rs; a.Ts = p->global_Ts; a.out = p->out;
    return lbs_common(a, false, (cudaStream_t)stream);
}

extern "C" int gom_lbs_backward(const GomLbsBwdArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->n_verts > 0 && p->n_joints > 0 && p->n_joints <= kMaxJ, "sizes");
    GOM_REQUIRE(p->xyz && p->lbs_weights && p->global_Rs && p->global_Ts && p->dL_dout && p->dL_dxyz, "null pointer");
    GOM_REQUIRE((p->dL_dglobal_Rs == nullptr) == (p->dL_dglobal_Ts == nullptr), "dL_dglobal_Rs/Ts must be both set or both null");
    cudaStream_t stream = (cudaStream_t)stream_;
    LbsDev a{};
    a.B = p->n_frames; a.J = p->n_joints; a.V = p->n_verts; a.xyz = p->xyz; a.xyz_stride = p->xyz_stride;
    a.w = p->lbs_weights; a.Rs = p->global_Rs; a.Ts = p->global_Ts; a.dout = p->dL_dout;
    a.dxyz = p->dL_dxyz; a.dxyz_stride = p->dL_dxyz_stride; a.dRs = p->dL_dglobal_Rs; a.dTs = p->dL_dglobal_Ts;
    if (a.dxyz_stride == 0) GOM_CUDA(cudaMemsetAsync(a.dxyz, 0, sizeof(float) * 3 * (size_t)a.V, stream));
    if (a.dRs) {
        GOM_CUDA(cudaMemsetAsync(a.dRs, 0, sizeof(float) * 9 * (size_t)a.B * a.J, stream));
        GOM_CUDA(cudaMemsetAsync(a.dTs, 0, sizeof(float) * 3 * (size_t)a.B * a.J, stream));
    }
    return lbs_common(a, true, stream);
}

static int face_check(int B, int F, int V) { return B > 0 && B <= 65535 && F > 0 && V > 0; }

extern "C" int gom_face_gaussians_forward(const GomFaceFwdArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(face_check(p->n_frames, p->n_faces, p->n_verts), "sizes");
    GOM_REQUIRE(p->verts && p->faces && p->so3 && p->scale && p->means3D && p->cov3D, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->cov3D % 8) == 0, "cov3D must be 8-byte aligned");
    FaceDev a{};
    a.B = p->n_frames; a.F = p->n_faces; a.V = p->n_verts; a.faces_int64 = p->faces_int64; a.sigma = p->sigma;
    a.verts = p->verts; a.faces = p->faces; a.so3 = p->so3; a.scale = p->scale; a.means3D = p->means3D; a.cov3D = p->cov3D;
    dim3 grid(gom_div_up(a.F, kThreads), a.B);
    gom_prof_begin(GOM_PROF_FACE_FWD, (cudaStream_t)stream);
    k_face_fwd<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_FACE_FWD, (cudaStream_t)stream);
    return GOM_OK;
}

extern "C" int gom_face_gaussians_backward(const GomFaceBwdArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(face_check(p->n_frames, p->n_faces, p->n_verts), "sizes");
    GOM_REQUIRE(p->verts && p->faces && p->so3 && p->scale && p->dL_dmeans3D && p->dL_dcov3D && p->dL_dverts && p->dL_dso3 &&
                    p->dL_dscale, "null pointer");
    GOM_REQUIRE(((uintptr_t)p->dL_dcov3D % 8) == 0, "dL_dcov3D must be 8-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    FaceDev a{};
    a.B = p->n_frames; a.F = p->n_faces; a.V = p->n_verts; a.faces_int64 = p->faces_int64; a.sigma = p->sigma;
    a.verts = p->verts; a.faces = p->faces; a.so3 = p->so3; a.scale = p->scale;
    a.dmeans = p->dL_dmeans3D; a.dcov = p->dL_dcov3D; a.dverts = p->dL_dverts; a.dso3 = p->dL_dso3; a.dscale = p->dL_dscale;
    GOM_CUDA(cudaMemsetAsync(a.dverts, 0, sizeof(float) * 3 * (size_t)a.B * a.V, stream));
    GOM_CUDA(cudaMemsetAsync(a.dso3, 0, sizeof(float) * 3 * (size_t)a.F, stream));
    GOM_CUDA(cudaMemsetAsync(a.dscale, 0, sizeof(float) * 3 * (size_t)a.F, stream));
    dim3 grid(gom_div_up(a.F, kThreads), a.B);
    gom_prof_begin(GOM_PROF_FACE_BWD, (cudaStream_t)stream);
    k_face_bwd<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_FACE_BWD, (cudaStream_t)stream);
    return GOM_OK;
}

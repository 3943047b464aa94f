// gom_core.cu — error reporting, ABI introspection and the camera-setup kernel of libgom_b200.so
#include <stdarg.h>
#include <string.h>

#include "gom_common.cuh"

static thread_local char g_err[512] = "";

void gom_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *gom_last_error(void) { return g_err; }
extern "C" int gom_abi_version(void) { return GOM_ABI_VERSION; }
extern "C" size_t gom_sizeof_camera_args(void) { return sizeof(GomCameraArgs); }
extern "C" size_t gom_sizeof_raster_fwd_args(void) { return sizeof(GomRasterFwdArgs); }
extern "C" size_t gom_sizeof_raster_bwd_args(void) { return sizeof(GomRasterBwdArgs); }
extern "C" size_t gom_sizeof_joint_fwd_args(void) { return sizeof(GomJointFwdArgs); }
extern "C" size_t gom_sizeof_joint_bwd_args(void) { return sizeof(GomJointBwdArgs); }
extern "C" size_t gom_sizeof_lbs_fwd_args(void) { return sizeof(GomLbsFwdArgs); }
extern "C" size_t gom_sizeof_lbs_bwd_args(void) { return sizeof(GomLbsBwdArgs); }
extern "C" size_t gom_sizeof_face_fwd_args(void) { return sizeof(GomFaceFwdArgs); }
extern "C" size_t gom_sizeof_face_bwd_args(void) { return sizeof(GomFaceBwdArgs); }

// Host math of reference models/modules/renderer/gaussian.py:30-47,60-61 moved on device: the four .item() syncs and
// the host-built K_ndc + H2D copy disappear.  Scalars are formed in fp64 from the fp32 K entries and rounded to fp32
// exactly like the reference's `torch.tensor([...python floats...]).float()`; tanfov = tan(focal2fov/2) = W/(2 fx).
__global__ void k_camera(GomCameraArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_frames) return;
    const float *K = a.K + 9 * b, *E = a.E + 16 * b;
    const double fx = K[0], fy = K[4], px = K[2], py = K[5];
    const double w = a.width, h = a.height, znear = 0.001, zfar = 100.0;
    float Kn[4][4] = {{(float)(2 * fx / w), 0.f, (float)((2 * px - w) / w), 0.f},
                      {0.f, (float)(2 * fy / h), (float)((2 * py - h) / h), 0.f},
                      {0.f, 0.f, (float)(zfar / (zfar - znear)), (float)(-zfar * znear / (zfar - znear))},
                      {0.f, 0.f, 1.f, 0.f}};
    float *view = a.viewmatrix + 16 * b, *proj = a.projmatrix + 16 * b;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            view[4 * i + j] = E[4 * j + i];                 // E^T
            float s = 0.f;                                  // (E^T K_ndc^T)[i][j] = sum_k E[k][i] K_ndc[j][k]
            for (int k = 0; k < 4; k++) s += E[4 * k + i] * Kn[j][k];
            proj[4 * i + j] = s;
        }
    a.tanfov[2 * b + 0] = (float)(w / (2.0 * fx));
    a.tanfov[2 * b + 1] = (float)(h / (2.0 * fy));
    if (a.campos) {   // camera centre -R^T t (E is rigid in the reference's datasets)
        for (int i = 0; i < 3; i++)
            a.campos[3 * b + i] = -(E[0 * 4 + i] * E[3] + E[1 * 4 + i] * E[7] + E[2 * 4 + i] * E[11]);
    }
}

extern "C" int gom_camera_from_KE(const GomCameraArgs *p, gom_stream_t stream) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n_frames > 0 && p->height > 0 && p->width > 0, "sizes");
    GOM_REQUIRE(p->K && p->E && p->viewmatrix && p->projmatrix && p->tanfov, "null pointer");
    k_camera<<<gom_div_up(p->n_frames, 64), 64, 0, (cudaStream_t)stream>>>(*p);
    GOM_LAUNCH_CHECK();
    return GOM_OK;
}

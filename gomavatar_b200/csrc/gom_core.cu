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

// ------------------------------------------------------------------------------------ launch counter / kernel timers
#include <atomic>
#include <vector>
static std::atomic<long long> g_launches{0};
void gom_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }

#include <stdlib.h>
bool gom_pdl_enabled(void) {
    // measured on B200 (bench.py, whole step in one CUDA graph): 7.73 -> 7.99 ms at 8 frames, 1.636 -> 1.665 ms at one frame with
    // the attribute on the LPIPS chain — the early-resident CTAs cost more than the hidden prologues — so it is opt-in
    static const bool on = [] { const char *e = getenv("GOM_PDL"); return e && e[0] == '1'; }();
    return on;
}
extern "C" long long gom_launch_count(void) { return g_launches.load(); }

struct ProfSlot { std::vector<cudaEvent_t> ev; size_t used = 0; };
static bool g_prof_on = false;
static ProfSlot g_prof[GOM_PROF_NSLOTS];
static const char *kProfNames[GOM_PROF_NSLOTS] = {"preprocess", "scan_tiles", "emit", "blend_fwd", "blend_bwd",
    "preprocess_bwd", "joint_fwd", "joint_bwd", "lbs_fwd", "lbs_bwd", "face_fwd", "face_bwd", "photo_fwd", "photo_bwd",
    "camera", "lpips_input", "bias_relu", "relu_bwd", "lpips_tap_fwd", "lpips_tap_bwd", "eval_metrics", "conv_first_fwd", "conv_first_bwd", "adam", "mesh_bin", "mesh_raster_fwd", "mesh_raster_bwd",
    "shadow_compact", "shadow_mlp_fwd", "shadow_mlp_bwd_data", "shadow_mlp_bwd_weights", "mesh_regularizers", "conv3x3_fwd", "conv3x3_dgrad", "worklist", "tile_sort", "gemm_tc", "wgrad_tc"};

static cudaEvent_t prof_next(int slot) {
    ProfSlot &s = g_prof[slot];
    if (s.used == s.ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        s.ev.push_back(e);
    }
    return s.ev[s.used++];
}
void gom_prof_begin(int slot, cudaStream_t stream) { if (g_prof_on) cudaEventRecord(prof_next(slot), stream); }
void gom_prof_end(int slot, cudaStream_t stream) { if (g_prof_on) cudaEventRecord(prof_next(slot), stream); }

extern "C" void gom_profile_enable(int on) {
    g_prof_on = on != 0;
    if (on) for (auto &s : g_prof) s.used = 0;
}
extern "C" int gom_profile_num_slots(void) { return GOM_PROF_NSLOTS; }
extern "C" const char *gom_profile_slot_name(int slot) { return (slot >= 0 && slot < GOM_PROF_NSLOTS) ? kProfNames[slot] : ""; }
// Sum of elapsed ms and number of launches recorded for a slot since gom_profile_enable(1).  Synchronises the events.
extern "C" int gom_profile_read(int slot, double *total_ms, int *count) {
    if (slot < 0 || slot >= GOM_PROF_NSLOTS || !total_ms || !count) return GOM_ERR_INVALID;
    ProfSlot &s = g_prof[slot];
    double tot = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < s.used; i += 2) {
        float ms = 0.f;
        if (cudaEventSynchronize(s.ev[i + 1]) != cudaSuccess) return GOM_ERR_CUDA;
        if (cudaEventElapsedTime(&ms, s.ev[i], s.ev[i + 1]) != cudaSuccess) return GOM_ERR_CUDA;
        tot += ms;
        n++;
    }
    *total_ms = tot;
    *count = n;
    return GOM_OK;
}
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
extern "C" size_t gom_sizeof_photo_args(void) { return sizeof(GomPhotoArgs); }

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

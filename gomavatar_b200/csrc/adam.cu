// adam.cu — one-launch Adam over the flat parameter / gradient arena (sm_100a).
//
// Replaces `optimizer.step()` of reference train.py:339 (torch.optim.Adam over the 6-8 parameter groups of
// models/model.py:305-324; per-group learning rates, exponentially decayed by train.py:166-175) for the parameters
// that live in gomavatar_b200.dist.FlatArena.  torch's fused Adam costs ~50 us per tensor list on B200 (4 launches,
// profiles/r1s_launches_step.md) for 1.3 MB of state; here the whole arena is one grid-stride pass, the per-group
// learning rate is looked up from a segment table passed by value, and the 1/world_size of the gradient all-reduce
// is folded in (grad_scale).  Same arithmetic as torch.optim.Adam (amsgrad off, weight decay 0):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#include "gom_common.cuh"

namespace {
__global__ void __launch_bounds__(256) k_adam(GomAdamArgs a) {
    __shared__ float s_step_size[GOM_ADAM_MAX_SEGMENTS], s_bc2_sqrt[GOM_ADAM_MAX_SEGMENTS];
    if (threadIdx.x < a.n_segments) {                       // per-segment scalars (double: 1 - beta^t loses digits in fp32)
        const int s = threadIdx.x;
        const long long t = a.dev_steps ? a.dev_steps[s] + 1 : a.seg_step[s];
        const long long iter = a.dev_steps ? a.dev_steps[GOM_ADAM_MAX_SEGMENTS] : a.iter;
        double lr = a.seg_lr[s];
        if (a.lr_decay_steps > 0.f) lr *= pow((double)a.lr_decay_rate, (double)iter / (double)a.lr_decay_steps);
        const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
        s_step_size[s] = (float)(lr / bc1);
        s_bc2_sqrt[s] = (float)sqrt(bc2);
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (long long)gridDim.x * 256) {
        int s = 0;
        while (s + 1 < a.n_segments && i >= a.seg_end[s]) s++;
        if (!a.seg_active[s]) continue;
        const float g = a.grad[i] * a.grad_scale;
        const float m = a.beta1 * a.exp_avg[i] + (1.f - a.beta1) * g;
        const float v = a.beta2 * a.exp_avg_sq[i] + (1.f - a.beta2) * g * g;
        a.exp_avg[i] = m;
        a.exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / s_bc2_sqrt[s] + a.eps;
        a.param[i] -= s_step_size[s] * (m / denom);
    }
}
// device-resident counters: runs after k_adam on the same stream
__global__ void k_adam_tick(GomAdamArgs a) {
    const int s = threadIdx.x;
    if (s < a.n_segments && a.seg_active[s]) a.dev_steps[s] += 1;
    if (s == GOM_ADAM_MAX_SEGMENTS) a.dev_steps[s] += 1;
}
}  // namespace

extern "C" int gom_adam_step(const GomAdamArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n > 0 && p->param && p->grad && p->exp_avg && p->exp_avg_sq, "null pointer / empty arena");
    GOM_REQUIRE(p->n_segments >= 1 && p->n_segments <= GOM_ADAM_MAX_SEGMENTS, "n_segments");
    GOM_REQUIRE(p->seg_end[p->n_segments - 1] >= p->n, "the last segment must end at or after n");
    if (!p->dev_steps)
        for (int s = 0; s < p->n_segments; s++) GOM_REQUIRE(!p->seg_active[s] || p->seg_step[s] >= 1, "seg_step must be >= 1 for active segments");
    cudaStream_t stream = (cudaStream_t)stream_;
    long long blocks = (p->n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gom_prof_begin(GOM_PROF_ADAM, stream);
    k_adam<<<(unsigned)blocks, 256, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    if (p->dev_steps) {
        k_adam_tick<<<1, 32, 0, stream>>>(*p);
        GOM_LAUNCH_CHECK();
    }
    gom_prof_end(GOM_PROF_ADAM, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_adam_args(void) { return sizeof(GomAdamArgs); }

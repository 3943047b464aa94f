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
    const float bc2_sqrt = sqrtf(a.bias_correction2);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (long long)gridDim.x * 256) {
        int s = 0;
        while (s + 1 < a.n_segments && i >= a.seg_end[s]) s++;
        const float g = a.grad[i] * a.grad_scale;
        const float m = a.beta1 * a.exp_avg[i] + (1.f - a.beta1) * g;
        const float v = a.beta2 * a.exp_avg_sq[i] + (1.f - a.beta2) * g * g;
        a.exp_avg[i] = m;
        a.exp_avg_sq[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + a.eps;
        a.param[i] -= (a.seg_lr[s] / a.bias_correction1) * (m / denom);
    }
}
}  // namespace

extern "C" int gom_adam_step(const GomAdamArgs *p, gom_stream_t stream_) {
    GOM_REQUIRE(p != nullptr, "args");
    GOM_REQUIRE(p->n > 0 && p->param && p->grad && p->exp_avg && p->exp_avg_sq, "null pointer / empty arena");
    GOM_REQUIRE(p->n_segments >= 1 && p->n_segments <= GOM_ADAM_MAX_SEGMENTS, "n_segments");
    GOM_REQUIRE(p->seg_end[p->n_segments - 1] >= p->n, "the last segment must end at or after n");
    GOM_REQUIRE(p->bias_correction1 > 0.f && p->bias_correction2 > 0.f, "bias corrections must be positive (step >= 1)");
    cudaStream_t stream = (cudaStream_t)stream_;
    long long blocks = (p->n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gom_prof_begin(GOM_PROF_ADAM, stream);
    k_adam<<<(unsigned)blocks, 256, 0, stream>>>(*p);
    GOM_LAUNCH_CHECK();
    gom_prof_end(GOM_PROF_ADAM, stream);
    return GOM_OK;
}

extern "C" size_t gom_sizeof_adam_args(void) { return sizeof(GomAdamArgs); }

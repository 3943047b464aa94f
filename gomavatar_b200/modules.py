"""The three small MLPs around the hot path, in plain torch, so that ``gomavatar_b200.model.Model`` is complete behind
the reference's ``Model(model_cfg, canonical_info)`` constructor without importing the reference (SURVEY.md §2 rows
6-8; they are dense cuBLAS work outside the north-star path):

* ``ShadowModule``          reference models/modules/shadow_module.py:67-117 — per-pixel normal -> positional encoding
                            (input + sin/cos at 2^0..2^(multires-1)) -> MLP -> sigmoid
* ``NonRigidModule``        reference models/modules/non_rigid_module.py:75-147 — pose-conditioned vertex offsets,
                            positional encoding faded in by a Hann window over training iterations (:15-72)
* ``PoseRefinementModule``  reference models/modules/pose_refinement_module.py:10-48 — pose vector -> per-joint axis-angle
                            corrections -> rotation matrices (root = identity)

State-dict keys (``block_mlps.N.weight / bias``) and initialisation match the reference (Xavier-uniform with the ReLU
gain, last layer U(-1e-5, 1e-5) with zero bias: utils/network_util.py:403-461), so reference checkpoints load unchanged;
tests/golden/golden_modules.npz pins the forward passes against the reference's own modules.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


_TC_MLP = True          # False: plain torch (cuBLAS fp32) everywhere — the A/B reference of tests/test_modules_gpu.py


def _init_linear(m, gain):
    nn.init.xavier_uniform_(m.weight, gain)
    nn.init.zeros_(m.bias)


def _mlp(in_dim, cond_dim, width, depth, out_dim, skips, init_last=1e-5):
    """Linear/ReLU stack of the reference: layer i in `skips` (counted in Linear layers, from 1) re-reads the encoding."""
    layers, cat_at = [nn.Linear(in_dim + cond_dim, width), nn.ReLU()], []
    for i in range(1, depth):
        if i in skips:
            cat_at.append(len(layers))
            layers += [nn.Linear(width + in_dim, width), nn.ReLU()]
        else:
            layers += [nn.Linear(width, width), nn.ReLU()]
    layers.append(nn.Linear(width, out_dim))
    mods = nn.ModuleList(layers)
    relu_gain = nn.init.calculate_gain("relu")
    for m in layers[:-1]:
        if isinstance(m, nn.Linear):
            _init_linear(m, relu_gain)
    nn.init.uniform_(layers[-1].weight, -init_last, init_last)
    nn.init.zeros_(layers[-1].bias)
    return mods, cat_at


class _Linear(torch.autograd.Function):
    """``F.linear`` whose bias gradient is a GEMV with a vector of ones instead of ``grad.sum(0)``: for the tall matrices of
    the non-rigid MLP ([B*V, 128], 120 k rows) torch's column reduction runs at 0.7 TB/s (83 us per layer on B200, 0.8 ms
    per step); the GEMV is bandwidth-bound (~10 us).  Same values up to summation order."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        gx = g @ weight if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            # [out, R] @ [R, in] with R ~ 1e5: one GEMM has only (out/64)(in/64) = 4 output tiles to spread over 148 SMs;
            # split the rows into S independent slabs (batched GEMM) and add the S partial products
            R = g.shape[0]
            S = 64
            while S > 1 and R % S:
                S //= 2
            if S >= 8 and R // S >= 512:
                gw = torch.bmm(g.view(S, R // S, -1).transpose(1, 2), x.reshape(S, R // S, -1)).sum(0)
            else:
                gw = g.t() @ x
        gb = (torch.ones(1, g.shape[0], dtype=g.dtype, device=g.device) @ g)[0] if ctx.needs_input_grad[2] else None
        return gx, gw, gb


class _NarrowLinear(torch.autograd.Function):
    """A Linear layer with 1 .. 4 outputs over a tall batch of rows (the non-rigid MLP's 128 -> 3 output layer) as one pass over
    the rows each way (csrc/narrow_linear.cu) instead of three SIMT cuBLAS GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from ._lib import GomNarrowLinearArgs, call, ptr
        xc, w = x.detach().contiguous().float(), weight.detach().contiguous().float().clone()      # clone: 16-byte aligned (arena views are not)
        b = None if bias is None else bias.detach().contiguous().float()
        R, C = xc.shape
        y = torch.empty(R, w.shape[0], dtype=torch.float32, device=xc.device)
        call("gom_narrow_linear_forward", GomNarrowLinearArgs(rows=R, c_in=C, n_out=w.shape[0], x=ptr(xc), weight=ptr(w), bias=ptr(b), y=ptr(y)))
        ctx.save_for_backward(xc, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        from ._lib import GomNarrowLinearArgs, call, ptr
        xc, w = ctx.saved_tensors
        R, C = xc.shape
        gc = g.contiguous().float()
        gx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w)
        gb = torch.empty(w.shape[0], dtype=torch.float32, device=xc.device) if ctx.has_bias else None
        call("gom_narrow_linear_backward", GomNarrowLinearArgs(rows=R, c_in=C, n_out=w.shape[0], x=ptr(xc), weight=ptr(w), g_y=ptr(gc), g_x=ptr(gx),
                                                               g_weight=ptr(gw), g_bias=ptr(gb)))
        return gx, gw, gb


def _linear(m, h):
    if (h.is_cuda and _TC_MLP and h.dim() == 2 and h.shape[0] >= 4096 and m.out_features <= 4 and h.shape[1] % 4 == 0
            and h.shape[1] <= 256 and (h.data_ptr() % 16) == 0):
        return _NarrowLinear.apply(h, m.weight, m.bias)
    if h.is_cuda and h.dim() >= 2 and h.numel() // h.shape[-1] >= 4096 and torch.is_grad_enabled():
        return _Linear.apply(h.reshape(-1, h.shape[-1]), m.weight, m.bias).reshape(*h.shape[:-1], m.out_features)
    return m(h)


def _pad_cols(t, mult):
    """[R, C] -> contiguous [R16, Cm]: rows padded to a multiple of 16, columns to a multiple of ``mult``, with zeros"""
    R, C = t.shape
    R16, Cm = (R + 15) // 16 * 16, (C + mult - 1) // mult * mult
    if R16 == R and Cm == C and t.is_contiguous():
        return t
    out = t.new_zeros(R16, Cm)
    out[:R, :C] = t
    return out


class _TcMlpStack(torch.autograd.Function):
    """The hidden Linear + ReLU layers of an MLP (reference models/modules/non_rigid_module.py:75-147: width 128, depth 6,
    the encoding re-read by the layer in ``skips``) on the tcgen05 kernel of csrc/conv3x3_tc.cu (kernel_size 1: a GEMM over
    rows, bias + ReLU fused, the ReLU bit mask written by the forward and consumed by the dgrad of the layer above) with 3xTF32
    products = fp32-GEMM accuracy, what ``nn.Linear`` computes in the reference (torch never enables TF32 for matmul by
    default).  Weight gradients (contractions over the ~1e5 rows) run on csrc/wgrad_tc.cu (tcgen05 with MN-major operands,
    split-K over the SMs) when the layer is 128 wide, the bias gradients are column sums formed by the TF32 split pass that
    reads the gradient anyway; other widths keep the split-K batched GEMM of ``_Linear``.

    forward(x0 [R, in0], enc [R, E], cat_layers, *weights_and_biases) -> last hidden activation [R, width]."""

    @staticmethod
    def forward(ctx, x0, enc, cat_layers, *params):
        from . import conv as C
        R = x0.shape[0]
        ws, bs = params[0::2], params[1::2]
        h = _pad_cols(x0.detach().float(), 32)
        encd = enc.detach().float()
        saved_in, saved_lo, masks = [], [], []
        for i, (w, b) in enumerate(zip(ws, bs)):
            if i in cat_layers:
                h = _pad_cols(torch.cat([h[:R, :ws[i - 1].shape[0]], encd], dim=1), 32)
            wp = w.detach().new_zeros(w.shape[0], h.shape[1])
            wp[:, :w.shape[1]] = w.detach()
            mask = C.new_mask(1, h.shape[0] // 16, 16, w.shape[0], h.device).view(h.shape[0], -1)
            h_lo = C.tf32_low_part(h)                      # kept: second operand of the weight gradient as well
            y = C.linear(h, C.pack_weights(wp, split=True), bias=b.detach().clone(), relu=True, mask_out=mask, precision="fp32", x_lo=h_lo)     # clone: 16-byte aligned (arena views are not)
            saved_in.append(h)
            saved_lo.append(h_lo)
            masks.append(mask)
            h = y
        ctx.cat_layers, ctx.R, ctx.n = tuple(cat_layers), R, len(ws)
        ctx.dims = (x0.shape[1], enc.shape[1])
        ctx.save_for_backward(*saved_in, *masks, *ws, *saved_lo)
        ctx.last = h
        return h[:R]

    @staticmethod
    def backward(ctx, g):
        from . import conv as C
        from ._lib import GomReluBwdArgs, call, ptr
        n, R = ctx.n, ctx.R
        t = ctx.saved_tensors
        saved_in, masks, ws, saved_lo = t[:n], t[n:2 * n], t[2 * n:3 * n], t[3 * n:]
        in0, E = ctx.dims
        R16 = saved_in[0].shape[0]
        if R16 == R:
            gp = g.contiguous().clone()                   # masked in place below
        else:
            gp = g.new_zeros(R16, g.shape[1])
            gp[:R] = g
        call("gom_relu_backward", GomReluBwdArgs(n=gp.numel(), act=ptr(ctx.last), grad=ptr(gp)))      # ReLU of the last hidden layer
        g_enc = None
        grads = [None] * (2 * n)
        ones = None                                          # GEMV partner of the fallback bias gradient, made on first use
        for i in reversed(range(n)):
            w, hin = ws[i], saved_in[i]
            # weight / bias gradients: contractions over the rows.  128-wide layers: tensor cores (csrc/wgrad_tc.cu), the bias
            # gradient = column sums formed by the TF32 split of gp; other widths: split-K batched GEMM, GEMV with ones
            gp_lo = None
            if gp.shape[1] == 128 and hin.shape[1] <= 256 and 1024 % gp.shape[1] == 0:
                gb = torch.zeros(gp.shape[1], dtype=gp.dtype, device=gp.device)
                gp_lo = C.tf32_low_part(gp, col_sum=gb)
                gw = C.linear_wgrad(gp, gp_lo, hin, saved_lo[i])
            else:
                S = 64
                while S > 1 and R16 % S:
                    S //= 2
                if S >= 8 and R16 // S >= 512:
                    gw = torch.bmm(gp.view(S, R16 // S, -1).transpose(1, 2), hin.view(S, R16 // S, -1)).sum(0)
                else:
                    gw = gp.t() @ hin
                if ones is None:
                    ones = torch.ones(1, R16, dtype=gp.dtype, device=gp.device)
                gb = (ones @ gp)[0]
            grads[2 * i] = gw[:, :w.shape[1]]
            grads[2 * i + 1] = gb
            if i == 0 and not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
                break
            # input gradient with the ReLU backward of the layer below fused (bit mask; the re-read encoding columns and the
            # first layer's input have no ReLU: all-ones mask words)
            wp = w.new_zeros(w.shape[0], hin.shape[1])
            wp[:, :w.shape[1]] = w
            c_out = hin.shape[1]
            mask_in = None
            if i > 0 and masks[i - 1].shape[1] == c_out // 32:
                mask_in = masks[i - 1]                    # the whole input is the previous layer's activation
            elif i > 0:
                mask_in = torch.full((R16, c_out // 32), -1, dtype=torch.int32, device=gp.device)
                mask_in[:, :masks[i - 1].shape[1]] = masks[i - 1]
            if c_out % 64:                               # the kernel writes 64-column tiles: pad the transposed weight's rows
                wp = torch.cat([wp, wp.new_zeros(wp.shape[0], 64 - c_out % 64)], dim=1)
                if mask_in is not None:
                    mask_in = torch.cat([mask_in, mask_in.new_full((R16, 1), -1)], dim=1).contiguous()
            g_in = C.linear(gp, C.pack_weights(wp, transpose=True, split=True), mask_in=mask_in, precision="fp32", x_lo=gp_lo)
            if i in ctx.cat_layers:
                wprev = ws[i - 1].shape[0]
                g_enc = g_in[:R, wprev:wprev + E] if g_enc is None else g_enc + g_in[:R, wprev:wprev + E]
                gp = g_in[:, :wprev].contiguous()
            elif i > 0:
                gp = g_in if g_in.shape[1] == ws[i - 1].shape[0] else g_in[:, :ws[i - 1].shape[0]].contiguous()
            else:
                gp = g_in
        g_x0 = gp[:R, :in0] if ctx.needs_input_grad[0] else None
        ctx.last = None
        return (g_x0, g_enc if ctx.needs_input_grad[1] else None, None) + tuple(grads)


class _NonRigidInput(torch.autograd.Function):
    """(canonical vertices [1|B,3,V], pose vectors [B,C]) -> (h0 [R16, cols] = [pose | windowed positional encoding | 0], enc
    [R16, 6L]) in the padded layout ``_TcMlpStack`` reads, one launch each way (csrc/posenc.cu) instead of torch's sin / cos /
    stack / window / cat / pad chain.  Reference models/modules/non_rigid_module.py:15-72,128-140."""

    @staticmethod
    def forward(ctx, xyz, posevec, alpha, multires, cols, rows_padded):
        from ._lib import GomNonRigidInputArgs, call, ptr
        B, C = posevec.shape
        Bx, _, V = xyz.shape
        x, pv = xyz.detach().contiguous().float(), posevec.detach().contiguous().float()
        h0 = torch.empty(rows_padded, cols, dtype=torch.float32, device=x.device)
        enc = torch.empty(rows_padded, 6 * multires, dtype=torch.float32, device=x.device)
        ctx.args = dict(n_frames=B, n_verts=V, xyz_frames=Bx, cond=C, multires=multires, cols=cols, rows_padded=rows_padded, alpha=float(alpha))
        call("gom_nonrigid_input_forward", GomNonRigidInputArgs(xyz=ptr(x), posevec=ptr(pv), h0=ptr(h0), enc=ptr(enc), **ctx.args))
        ctx.save_for_backward(x)
        return h0, enc

    @staticmethod
    def backward(ctx, g_h0, g_enc):
        from ._lib import GomNonRigidInputArgs, call, ptr
        (x,) = ctx.saved_tensors
        gh = None if g_h0 is None else g_h0.contiguous().float()
        ge = None if g_enc is None else g_enc.contiguous().float()
        gx = torch.empty_like(x)
        call("gom_nonrigid_input_backward", GomNonRigidInputArgs(xyz=ptr(x), g_h0=ptr(gh), g_enc=ptr(ge), g_xyz=ptr(gx), **ctx.args))
        return gx, None, None, None, None, None


def _run(mods, cat_at, h, enc):
    # hidden stack on the tensor cores when it is tall enough to matter and shaped like the reference's (width % 64 == 0)
    lin = [m for m in mods if isinstance(m, nn.Linear)]
    if (h.is_cuda and h.dim() >= 2 and h.numel() // h.shape[-1] >= 4096 and len(lin) >= 2
            and all(m.out_features % 64 == 0 for m in lin[:-1]) and _TC_MLP):
        lead = h.shape[:-1]
        cat_layers = tuple(sorted(i // 2 for i in cat_at))            # module index -> Linear index
        params = [p for m in lin[:-1] for p in (m.weight, m.bias)]
        hid = _TcMlpStack.apply(h.reshape(-1, h.shape[-1]), enc.reshape(-1, enc.shape[-1]), cat_layers, *params)
        return _linear(lin[-1], hid).reshape(*lead, lin[-1].out_features)
    for i, m in enumerate(mods):
        if i in cat_at:
            h = torch.cat([h, enc], dim=-1)
        h = _linear(m, h) if isinstance(m, nn.Linear) else m(h)
    return h


def posenc(x, multires, include_input=True, window=None):
    """[..., 3] -> [..., 3 (+3) * 2 * multires]: (x,) then for each frequency 2^k: sin(2^k x), cos(2^k x), each block
    optionally scaled by window[k]."""
    freqs = 2.0 ** torch.arange(multires, dtype=x.dtype, device=x.device)
    xf = x[..., None, :] * freqs[:, None]                                   # [..., F, 3]
    sc = torch.stack([torch.sin(xf), torch.cos(xf)], dim=-2)                # [..., F, 2, 3]
    if window is not None:
        sc = sc * window.to(x.dtype).to(x.device)[:, None, None]
    sc = sc.reshape(*x.shape[:-1], multires * 6)
    return torch.cat([x, sc], dim=-1) if include_input else sc


def hann_window(multires, i_iter, kick_in_iter, full_band_iter, device=None):
    """reference non_rigid_module.py:33-43: w_k = (1 - cos(pi clamp(alpha - k, 0, 1))) / 2, alpha = m t / N.  Built on
    `device` from Python scalars only (no host-to-device tensor copy: the step stays CUDA-graph capturable)."""
    t = max(float(i_iter) - float(kick_in_iter), 0.0)
    alpha = multires * t / (float(full_band_iter) - float(kick_in_iter))
    k = torch.arange(multires, dtype=torch.float32, device=device)
    return (1.0 - torch.cos(math.pi * torch.clamp(alpha - k, min=0.0, max=1.0))) / 2.0


class ShadowModule(nn.Module):
    def __init__(self, module_cfg=None, **kwargs):
        super().__init__()
        self.multires = int(_get(module_cfg, "multires", 6))
        width, depth = int(_get(module_cfg, "mlp_width", 128)), int(_get(module_cfg, "mlp_depth", 3))
        skips = list(_get(module_cfg, "skips", [4]))
        self.block_mlps, self.layers_to_cat_inputs = _mlp(3 + 6 * self.multires, 0, width, depth, 1, skips,
                                                          float(_get(module_cfg, "init_scale", 1e-5)))

    def forward(self, normals, **kwargs):
        enc = posenc(normals, self.multires, include_input=True)
        return torch.sigmoid(_run(self.block_mlps, self.layers_to_cat_inputs, enc, enc))


class NonRigidModule(nn.Module):
    def __init__(self, module_cfg=None, **kwargs):
        super().__init__()
        self.multires = int(_get(module_cfg, "multires", 6))
        self.kick_in_iter = float(_get(module_cfg, "kick_in_iter", 0))
        self.full_band_iter = float(_get(module_cfg, "full_band_iter", 50000))
        self.update_rot, self.update_scale = bool(_get(module_cfg, "update_rot", False)), bool(_get(module_cfg, "update_scale", False))
        if self.update_rot or self.update_scale:
            raise NotImplementedError("update_rot / update_scale are off in every reference config (exps/*.yaml)")
        width, depth = int(_get(module_cfg, "mlp_width", 128)), int(_get(module_cfg, "mlp_depth", 6))
        self.block_mlps, self.layers_to_cat_inputs = _mlp(6 * self.multires, int(_get(module_cfg, "condition_code_size", 69)),
                                                          width, depth, 3, list(_get(module_cfg, "skips", [4])),
                                                          float(_get(module_cfg, "init_scale", 1e-5)))

    def forward(self, xyzs_skeleton, dst_posevec, i_iter, R=None, S=None):
        """xyzs_skeleton [B,3,V], dst_posevec [B,69] -> (xyzs_skeleton + offset [B,3,V], R, S)"""
        lin = [m for m in self.block_mlps if isinstance(m, nn.Linear)]
        Bp, V = dst_posevec.shape[0], xyzs_skeleton.shape[2]
        if (xyzs_skeleton.is_cuda and _TC_MLP and Bp * V >= 4096 and xyzs_skeleton.shape[0] in (1, Bp) and self.multires <= 10
                and not dst_posevec.requires_grad and len(lin) >= 2 and all(m.out_features % 64 == 0 for m in lin[:-1])):
            # fused input rows (csrc/posenc.cu) -> tensor-core hidden stack -> last layer
            C, E = dst_posevec.shape[1], 6 * self.multires
            cols, rows = (C + E + 31) // 32 * 32, (Bp * V + 15) // 16 * 16
            t = max(float(i_iter) - self.kick_in_iter, 0.0)
            alpha = self.multires * t / (self.full_band_iter - self.kick_in_iter)
            h0, enc = _NonRigidInput.apply(xyzs_skeleton, dst_posevec, alpha, self.multires, cols, rows)
            cat_layers = tuple(sorted(i // 2 for i in self.layers_to_cat_inputs))
            params = [p for m in lin[:-1] for p in (m.weight, m.bias)]
            hid = _TcMlpStack.apply(h0, enc, cat_layers, *params)
            offset = _linear(lin[-1], hid[:Bp * V]).reshape(Bp, V, 3)
            return xyzs_skeleton + offset.permute(0, 2, 1), R, S
        xyzs = xyzs_skeleton.permute(0, 2, 1)
        B, N, _ = xyzs.shape
        if B != dst_posevec.shape[0]:
            xyzs = xyzs.expand(dst_posevec.shape[0], -1, -1)
            B = dst_posevec.shape[0]
        enc = posenc(xyzs, self.multires, include_input=False,
                     window=hann_window(self.multires, i_iter, self.kick_in_iter, self.full_band_iter, device=xyzs.device))
        h = torch.cat([dst_posevec[:, None, :].expand(B, N, -1), enc], dim=-1)
        offset = _run(self.block_mlps, self.layers_to_cat_inputs, h, enc)
        return xyzs_skeleton + offset.permute(0, 2, 1), R, S


class PoseRefinementModule(nn.Module):
    def __init__(self, module_cfg=None, **kwargs):
        super().__init__()
        emb, width, depth = int(_get(module_cfg, "embedding_size", 69)), int(_get(module_cfg, "mlp_width", 256)), int(_get(module_cfg, "mlp_depth", 4))
        self.refine_root = bool(_get(module_cfg, "refine_root", False))
        total = int(_get(module_cfg, "total_bones", 24))
        self.total_bones = total if self.refine_root else total - 1
        layers = [nn.Linear(emb, width), nn.ReLU()]
        for _ in range(depth - 1):
            layers += [nn.Linear(width, width), nn.ReLU()]
        layers.append(nn.Linear(width, 3 * self.total_bones))
        self.block_mlps = nn.Sequential(*layers)
        g = nn.init.calculate_gain("relu")
        for m in layers[:-1]:
            if isinstance(m, nn.Linear):
                _init_linear(m, g)
        nn.init.uniform_(layers[-1].weight, -1e-5, 1e-5)
        nn.init.zeros_(layers[-1].bias)

    def forward(self, dst_posevec, **kwargs):
        from .model import rodrigues_grouped
        rvec = self.block_mlps(dst_posevec).view(-1, 3)
        return rodrigues_grouped(rvec, self.total_bones, prepend_identity=True)     # identity in front: pose_refinement_module.py:44-46

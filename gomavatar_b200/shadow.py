"""``ShadowModule`` on the tcgen05 kernels of csrc/shadow_mlp.cu — drop-in for reference
``models/modules/shadow_module.py::ShadowModule`` (constructor ``ShadowModule(module_cfg)``, ``forward(normals [B,N,3])
-> [B,N,1]`` in (0,1), state-dict keys ``block_mlps.{0,2,..}.{weight,bias}``), SURVEY.md §8 f-2.

Forward: ``gom_shadow_mlp_forward`` — foreground compaction (the mesh renderer writes exact zeros on the background,
reference mesh.py:103-112, so those pixels share ONE value), weights split into TF32 hi/lo, and a persistent kernel that
keeps the activations in tensor memory (3xTF32 products, fp32 accumulation: fp32-GEMM accuracy, see the kernel header).

Backward: ``gom_shadow_mlp_backward`` — the forward saved its operands as TF32 hi/lo images; one tcgen05 kernel runs the
chain backwards to dL/dnormal, a second one accumulates the weight/bias gradients as split-K GEMMs in tensor memory, a
fixed-order reduction makes the result deterministic.  Static shapes, no host sync: CUDA-graph capturable.  The
background pixels are ONE extra row with normal 0 carrying the summed gradient of all background pixels (the reference
evaluates the MLP on them too, so they contribute weight gradients and receive a normal gradient): a handful of tiny torch
ops (``background_row``).  There is no CPU path: inputs must be CUDA tensors and the library must be built.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomShadowMlpArgs, call, ptr
from .modules import ShadowModule as _TorchShadowModule
from .modules import posenc

MAX_TRAIN_DEPTH = 3          # tensor-memory columns of the weight-gradient accumulators (csrc/shadow_mlp.cu)


def _posenc_backward(x, g_enc, multires):
    """d posenc(x) / dx applied to g_enc [R, 3 + 6*multires] -> [R,3]."""
    g = g_enc[:, :3].clone()
    for k in range(multires):
        f = float(2 ** k)
        xf = x * f
        g = g + f * (torch.cos(xf) * g_enc[:, 3 + 6 * k: 6 + 6 * k] - torch.sin(xf) * g_enc[:, 6 + 6 * k: 9 + 6 * k])
    return g


def background_row(weights, biases, w_out, b_out, multires):
    """The MLP at normal = 0 and its gradients for a unit upstream gradient (device-agnostic torch, one row).

    weights/biases: the Linear+ReLU layers; w_out [W], b_out [1].  Returns (y0 scalar tensor, g_x0 [1,3] = d out / d normal,
    [dW_l], [db_l], dw_out [W], db_out scalar) — everything is linear in the upstream gradient, the caller scales it."""
    dev = w_out.device
    x0 = torch.zeros(1, 3, device=dev)
    enc0 = posenc(x0, multires, include_input=True)
    h, hid = enc0, []
    for W, b in zip(weights, biases):
        h = torch.relu(h @ W.t() + b)
        hid.append(h)
    y0 = torch.sigmoid(h @ w_out[:, None] + b_out).reshape(())
    dz = y0 * (1.0 - y0)
    dw_out, db_out = hid[-1][0] * dz, dz
    dzl = (w_out * dz) * (hid[-1][0] > 0)                                         # [W]
    dWs, dbs = [None] * len(weights), [None] * len(weights)
    for l in range(len(weights) - 1, 0, -1):
        dWs[l], dbs[l] = torch.outer(dzl, hid[l - 1][0]), dzl
        dzl = (weights[l].t() @ dzl) * (hid[l - 1][0] > 0)
    dWs[0], dbs[0] = torch.outer(dzl, enc0[0]), dzl
    g_x0 = _posenc_backward(x0, (dzl @ weights[0])[None], multires)
    return y0, g_x0, dWs, dbs, dw_out, db_out


class _ShadowMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, need_grad, normals, w_out, b_out, *wb):
        """normals [N,3] contiguous fp32 CUDA; wb = (W_0, b_0, W_1, b_1, ...) of the Linear+ReLU layers."""
        depth = len(wb) // 2
        if need_grad and depth > MAX_TRAIN_DEPTH:
            raise NotImplementedError(f"FusedShadowModule: the backward pass supports mlp_depth <= {MAX_TRAIN_DEPTH}")
        weights, biases = list(wb[0::2]), list(wb[1::2])
        N = normals.shape[0]
        W_hid = torch.stack([w.detach() for w in weights[1:]]).contiguous() if depth > 1 else None
        b_hid = torch.stack([b.detach() for b in biases[1:]]).contiguous() if depth > 1 else None
        W_in, b_in = weights[0].detach().contiguous(), biases[0].detach().contiguous()
        wo, bo = w_out.detach().reshape(-1).contiguous(), b_out.detach().reshape(-1).contiguous()
        out = torch.empty(N, dtype=torch.float32, device=normals.device)
        while True:
            ws = module._workspace(N, depth, normals.device, need_grad)
            a = GomShadowMlpArgs(n_pixels=N, capacity=ws["capacity"], multires=module.multires, width=module.width, depth=depth,
                                 save_hidden=int(need_grad), normals=ptr(normals), W_in=ptr(W_in), b_in=ptr(b_in),
                                 W_hid=ptr(W_hid), b_hid=ptr(b_hid), W_out=ptr(wo), b_out=ptr(bo),
                                 block_count=ptr(ws["block_count"]), fg_index=ptr(ws["fg_index"]), n_fg=ptr(ws["n_fg"]),
                                 w_images=ptr(ws["w_images"]), bg_value=ptr(ws["bg_value"]), out=ptr(out),
                                 act_img=ptr(ws["act_img"]) if need_grad else None, status=ptr(ws["status"]))
            call("gom_shadow_mlp_forward", a)
            if not module.strict or torch.cuda.is_current_stream_capturing():
                break                                   # status stays on the device: FusedShadowModule.check_status()
            status = int(ws["status"].item())
            if status & _lib.STATUS_TIMEOUT:
                raise _lib.GomError("gom_shadow_mlp_forward: tcgen05 pipeline wait timed out (status TIMEOUT)")
            if status & _lib.STATUS_OVERFLOW:           # more foreground than rows kept for backward: regrow, rerun
                n_fg = int(ws["n_fg"].item())
                module.capacity = (int(n_fg * 1.25) + 127) // 128 * 128
                continue
            break
        if need_grad:
            ws["generation"] = ws.get("generation", 0) + 1          # the saved operand images live in the shared workspace
            ctx.generation = ws["generation"]
            ctx.module, ctx.depth, ctx.ws = module, depth, ws
            ctx.fwd_args = (W_in, b_in, W_hid, b_hid, wo, bo)
            ctx.save_for_backward(normals, out)
        return out

    @staticmethod
    def backward(ctx, g_out):
        module, depth, ws = ctx.module, ctx.depth, ctx.ws
        if ws.get("generation") != ctx.generation:
            raise _lib.GomError("FusedShadowModule: another forward pass with gradients ran on this module before this backward; "
                                "its saved activations were overwritten (use one module instance per concurrent graph)")
        normals, out = ctx.saved_tensors
        W_in, b_in, W_hid, b_hid, wo, bo = ctx.fwd_args
        dev = normals.device
        N = normals.shape[0]
        g_out = g_out.contiguous().float()
        # the background row (normal 0) is handled by csrc/shadow_bg.cu: `prepare` writes g_normals for every pixel (the row's input
        # gradient times g_out on the background, 0 elsewhere) and sums g_out over the background, `apply` adds that sum times the
        # row's parameter gradients after the tcgen05 backward has written the foreground's
        g_normals = torch.empty(N, 3, dtype=torch.float32, device=dev)
        bg_scratch = torch.empty(int(_lib.lib().gom_shadow_mlp_bg_scratch_floats()), dtype=torch.float32, device=dev)
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        gW_in, gb_in, gw_out, gb_out = e(module.width, W_in.shape[1]), e(module.width), e(module.width), e(1)
        gW_hid = e(depth - 1, module.width, module.width) if depth > 1 else None
        gb_hid = e(depth - 1, module.width) if depth > 1 else None
        a = GomShadowMlpArgs(n_pixels=N, capacity=ws["capacity"], multires=module.multires, width=module.width, depth=depth,
                             save_hidden=1, normals=ptr(normals), W_in=ptr(W_in), b_in=ptr(b_in), W_hid=ptr(W_hid), b_hid=ptr(b_hid),
                             W_out=ptr(wo), b_out=ptr(bo), block_count=ptr(ws["block_count"]), fg_index=ptr(ws["fg_index"]),
                             n_fg=ptr(ws["n_fg"]), w_images=ptr(ws["w_images"]), bg_value=ptr(ws["bg_value"]), out=ptr(out),
                             act_img=ptr(ws["act_img"]), status=ptr(ws["status"]), g_out=ptr(g_out), dz_img=ptr(ws["dz_img"]),
                             g_normals=ptr(g_normals), dzo_sums=ptr(ws["dzo_sums"]), partials=ptr(ws["partials"]),
                             g_W_in=ptr(gW_in), g_b_in=ptr(gb_in), g_W_hid=ptr(gW_hid), g_b_hid=ptr(gb_hid),
                             g_w_out=ptr(gw_out), g_b_out=ptr(gb_out), bg_scratch=ptr(bg_scratch))
        call("gom_shadow_mlp_background_prepare", a)
        call("gom_shadow_mlp_backward", a)
        call("gom_shadow_mlp_background_apply", a)
        grads = [gW_in, gb_in]
        for l in range(1, depth):
            grads += [gW_hid[l - 1], gb_hid[l - 1]]
        return (None, None, g_normals, gw_out.reshape(1, -1), gb_out.reshape(1), *grads)


class FusedShadowModule(_TorchShadowModule):
    """Same parameters / state dict as ``modules.ShadowModule`` (and therefore as the reference's); forward on the
    tcgen05 kernels.  ``capacity``: foreground rows kept for the backward pass (default: a quarter of the pixels, regrown
    automatically when ``strict``); ``strict=False`` never reads the device status (for CUDA-graph capture).  Training
    needs ``mlp_depth <= 3`` (the reference's configs use 3); inference supports up to 8."""

    def __init__(self, module_cfg=None, capacity=None, strict=True, **kwargs):
        super().__init__(module_cfg, **kwargs)
        linears = [m for m in self.block_mlps if isinstance(m, torch.nn.Linear)]
        self.width = linears[0].out_features
        if self.layers_to_cat_inputs or self.width != 128 or any(m.out_features != self.width for m in linears[:-1]) \
                or linears[-1].out_features != 1 or not (1 <= len(linears) - 1 <= 8) or self.multires > 10:
            raise NotImplementedError("FusedShadowModule supports the reference's shipped configuration family: width 128, "
                                      "no skip connection inside the depth, 1..8 hidden layers, multires <= 10")
        self.capacity, self.strict = capacity, strict
        self._ws = None

    def _linears(self):
        return [m for m in self.block_mlps if isinstance(m, torch.nn.Linear)]

    def _workspace(self, n_pixels, depth, device, need_backward):
        cap = self.capacity if self.capacity else max(128, (n_pixels // 4 + 127) // 128 * 128)
        cap = min((cap + 127) // 128 * 128, (n_pixels + 127) // 128 * 128)
        key = (n_pixels, depth, str(device), cap)
        L = _lib.lib()
        if self._ws is None or self._ws["key"] != key:
            img_bytes = int(L.gom_shadow_mlp_weight_image_bytes(depth))
            e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=device)
            self._ws = dict(key=key, capacity=cap, block_count=e((n_pixels + 1023) // 1024 + 1, dtype=torch.int32),
                            fg_index=e(n_pixels, dtype=torch.int32), n_fg=torch.zeros(1, dtype=torch.int32, device=device),
                            w_images=e(img_bytes // 4), bg_value=e(1), status=torch.zeros(1, dtype=torch.int32, device=device),
                            act_img=None)
        if need_backward and self._ws["act_img"] is None:
            tiles = cap // 128
            e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=device)
            self._ws.update(
                act_img=e(tiles * int(L.gom_shadow_mlp_tile_words(depth, 0)), dtype=torch.int32),
                # zeros once: rows 1..15 of the dz_out images are never written and must stay 0
                dz_img=torch.zeros(tiles * int(L.gom_shadow_mlp_tile_words(depth, 1)), dtype=torch.int32, device=device),
                dzo_sums=e(tiles * 4), partials=e(int(L.gom_shadow_mlp_num_ctas()) * int(L.gom_shadow_mlp_partial_floats())))
        return self._ws

    def check_status(self):
        """Read the device status of the last forward (a host sync): raises on overflow / timeout."""
        if self._ws is None:
            return
        status = int(self._ws["status"].item())
        if status & _lib.STATUS_TIMEOUT:
            raise _lib.GomError("shadow MLP: tcgen05 pipeline wait timed out")
        if status & _lib.STATUS_OVERFLOW:
            raise _lib.GomError(f"shadow MLP: {int(self._ws['n_fg'].item())} foreground pixels exceed capacity "
                                f"{self._ws['capacity']}; raise `capacity`")

    def forward(self, normals, **kwargs):
        if normals.device.type != "cuda":
            raise _lib.GomError("FusedShadowModule: inputs must live on a CUDA device (no CPU path exists)")
        shape = normals.shape[:-1]
        flat = normals.reshape(-1, 3).contiguous().float()
        lin = self._linears()
        wb = []
        for m in lin[:-1]:
            wb += [m.weight, m.bias]
        # grad mode is off inside Function.forward and ctx.needs_input_grad ignores torch.no_grad(): decide here whether
        # the kernel has to keep the hidden activations for a backward pass
        need_grad = torch.is_grad_enabled() and (flat.requires_grad or any(p.requires_grad for p in self.parameters()))
        out = _ShadowMlp.apply(self, need_grad, flat, lin[-1].weight, lin[-1].bias, *wb)
        return out.reshape(*shape, 1)

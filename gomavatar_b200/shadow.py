"""``ShadowModule`` on the tcgen05 kernel of csrc/shadow_mlp.cu — drop-in for reference
``models/modules/shadow_module.py::ShadowModule`` (constructor ``ShadowModule(module_cfg)``, ``forward(normals [B,N,3])
-> [B,N,1]`` in (0,1), state-dict keys ``block_mlps.{0,2,..}.{weight,bias}``), SURVEY.md §8 f-2.

Forward: ``gom_shadow_mlp_forward`` — foreground compaction (the mesh renderer writes exact zeros on the background,
reference mesh.py:103-112, so those pixels share ONE value), weights split into TF32 hi/lo, and a persistent kernel that
keeps the activations in tensor memory (3xTF32 products, fp32 accumulation: fp32-GEMM accuracy, see the kernel header).

Backward: the kernel saved the post-ReLU activations of the foreground rows feature-major ``[depth,128,capacity]``; the
gradient GEMMs (dW = dZ^T H, dH = dZ W, K = foreground rows) are plain fp32 cuBLAS calls on those fixed-capacity buffers
— static shapes, no host sync, CUDA-graph capturable; padded rows carry zero gradient.  The background pixels are one
extra row with normal 0 and the summed gradient of all background pixels (the reference evaluates the MLP on them too,
so they contribute weight gradients).  There is no CPU path: inputs must be CUDA tensors and the library must be built.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GomShadowMlpArgs, call, ptr
from .modules import ShadowModule as _TorchShadowModule
from .modules import posenc


def _posenc_backward(x, g_enc, multires):
    """d posenc(x) / dx applied to g_enc [R, 3 + 6*multires] -> [R,3]."""
    g = g_enc[:, :3].clone()
    for k in range(multires):
        f = float(2 ** k)
        xf = x * f
        g = g + f * (torch.cos(xf) * g_enc[:, 3 + 6 * k: 6 + 6 * k] - torch.sin(xf) * g_enc[:, 6 + 6 * k: 9 + 6 * k])
    return g


def _mlp_backward(enc, hidden_t, weights, w_out, dz_out):
    """Backward of Linear/ReLU x depth + Linear(width,1) given the saved activations.

    enc [R,E] layer-0 input; hidden_t: list of [W,R] post-ReLU activations (feature-major); weights: list of [W,in];
    w_out [W]; dz_out [R] gradient of the pre-sigmoid output.  Returns (g_enc [R,E], [dW_l], [db_l], dw_out [W], db_out)."""
    depth = len(hidden_t)
    dw_out = hidden_t[-1] @ dz_out
    db_out = dz_out.sum()
    dzt = (w_out[:, None] * dz_out[None, :]) * (hidden_t[-1] > 0)                # [W,R]
    dWs, dbs = [None] * depth, [None] * depth
    for l in range(depth - 1, 0, -1):
        dWs[l] = dzt @ hidden_t[l - 1].t()                                       # [W,W]
        dbs[l] = dzt.sum(dim=1)
        dzt = (weights[l].t() @ dzt) * (hidden_t[l - 1] > 0)
    dWs[0] = dzt @ enc                                                           # [W,E]
    dbs[0] = dzt.sum(dim=1)
    g_enc = dzt.t() @ weights[0]                                                 # [R,E]
    return g_enc, dWs, dbs, dw_out, db_out


def shadow_backward(normals, out, fg_index, n_fg, hidden, weights, biases, w_out, b_out, g_out, multires, cap):
    """Gradients of sum(out * g_out) from what gom_shadow_mlp_forward left behind (device-agnostic torch, static shapes).

    normals [N,3]; out [N] sigmoid outputs; fg_index [>=cap] foreground pixel ids; n_fg [1] their count (device tensor);
    hidden [depth,W,cap] post-ReLU activations of foreground row r in column r.  Returns (g_normals [N,3], g_w_out [1,W],
    g_b_out [1], [g_W0, g_b0, g_W1, g_b1, ...])."""
    depth = len(weights)
    dev = normals.device
    g_out = g_out.contiguous().float()
    rows = torch.arange(cap, device=dev)
    valid = rows < n_fg.clamp(max=cap)                                            # device-side, no sync
    idx = torch.where(valid, fg_index[:cap].long(), torch.zeros_like(rows))
    y = out[idx]
    gy = torch.where(valid, g_out[idx], torch.zeros_like(y))
    dz_out = gy * y * (1.0 - y)
    x = torch.where(valid[:, None], normals[idx], torch.zeros(1, 3, device=dev))
    enc = posenc(x, multires, include_input=True)
    # rows >= n_fg of `hidden` are stale but finite (the buffer starts as zeros) and their dz is exactly 0
    hidden_t = [hidden[l] for l in range(depth)]
    g_enc, dWs, dbs, dw_out, db_out = _mlp_backward(enc, hidden_t, weights, w_out, dz_out)
    # the background: one row with normal 0 carrying the summed gradient of every background pixel (the reference
    # evaluates the MLP there too); the MLP backward is linear in dz, so it is run for dz = y0 (1 - y0) and scaled
    x0 = torch.zeros(1, 3, device=dev)
    enc0 = posenc(x0, multires, include_input=True)
    h, hid0 = enc0, []
    for l in range(depth):
        h = torch.relu(h @ weights[l].t() + biases[l])
        hid0.append(h.t().contiguous())
    y0 = torch.sigmoid(h @ w_out[:, None] + b_out)
    g_bg = g_out.sum() - gy.sum()
    g_enc0, dWs0, dbs0, dw_out0, db_out0 = _mlp_backward(enc0, hid0, weights, w_out, (y0 * (1.0 - y0)).reshape(1))
    g_x0 = _posenc_backward(x0, g_enc0, multires)                                 # [1,3]: d out / d normal at normal = 0
    is_bg = (normals == 0).all(dim=1, keepdim=True)
    g_normals = torch.where(is_bg, g_out[:, None] * g_x0, torch.zeros_like(normals))
    g_normals.index_add_(0, idx, _posenc_backward(x, g_enc, multires))
    g_wb = []
    for l in range(depth):
        g_wb += [dWs[l] + g_bg * dWs0[l], dbs[l] + g_bg * dbs0[l]]
    return g_normals, (dw_out + g_bg * dw_out0).reshape(1, -1), (db_out + g_bg * db_out0).reshape(1), g_wb


class _ShadowMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, normals, w_out, b_out, *wb):
        """normals [N,3] contiguous fp32 CUDA; wb = (W_0, b_0, W_1, b_1, ...) of the Linear+ReLU layers."""
        depth = len(wb) // 2
        weights, biases = list(wb[0::2]), list(wb[1::2])
        N = normals.shape[0]
        need_grad = any(ctx.needs_input_grad)            # (grad mode is off inside Function.forward; this accounts for it)
        ws = module._workspace(N, depth, normals.device, need_grad)
        W_hid = torch.stack([w.detach() for w in weights[1:]]).contiguous() if depth > 1 else None
        b_hid = torch.stack([b.detach() for b in biases[1:]]).contiguous() if depth > 1 else None
        W_in, b_in = weights[0].detach().contiguous(), biases[0].detach().contiguous()
        wo, bo = w_out.detach().reshape(-1).contiguous(), b_out.detach().reshape(-1).contiguous()
        out = torch.empty(N, dtype=torch.float32, device=normals.device)
        while True:
            a = GomShadowMlpArgs(n_pixels=N, capacity=ws["capacity"], multires=module.multires, width=module.width, depth=depth,
                                 save_hidden=int(need_grad), normals=ptr(normals), W_in=ptr(W_in), b_in=ptr(b_in),
                                 W_hid=ptr(W_hid), b_hid=ptr(b_hid), W_out=ptr(wo), b_out=ptr(bo),
                                 block_count=ptr(ws["block_count"]), fg_index=ptr(ws["fg_index"]), n_fg=ptr(ws["n_fg"]),
                                 w_images=ptr(ws["w_images"]), bg_value=ptr(ws["bg_value"]), out=ptr(out),
                                 hidden=ptr(ws["hidden"]) if need_grad else None, status=ptr(ws["status"]))
            call("gom_shadow_mlp_forward", a)
            if not module.strict or torch.cuda.is_current_stream_capturing():
                break                                   # status stays on the device: FusedShadowModule.check_status()
            status = int(ws["status"].item())
            if status & _lib.STATUS_TIMEOUT:
                raise _lib.GomError("gom_shadow_mlp_forward: tcgen05 pipeline wait timed out (status TIMEOUT)")
            if status & _lib.STATUS_OVERFLOW:           # more foreground than rows kept for backward: regrow, rerun
                n_fg = int(ws["n_fg"].item())
                module.capacity = min(N, (int(n_fg * 1.25) + 127) // 128 * 128)
                ws = module._workspace(N, depth, normals.device, need_grad)
                continue
            break
        if need_grad:
            ctx.module, ctx.depth, ctx.capacity = module, depth, ws["capacity"]
            ctx.save_for_backward(normals, out, ws["fg_index"], ws["n_fg"], ws["hidden"], wo, *weights)
        return out

    @staticmethod
    def backward(ctx, g_out):
        normals, out, fg_index, n_fg, hidden, wo = ctx.saved_tensors[:6]
        weights = [w.detach() for w in ctx.saved_tensors[6:]]
        m = ctx.module
        g_normals, g_wo, g_bo, g_wb = shadow_backward(normals, out, fg_index, n_fg, hidden, weights, m._biases_detached(), wo,
                                                      m._b_out_detached(), g_out, m.multires, ctx.capacity)
        return (None, g_normals, g_wo, g_bo, *g_wb)


class FusedShadowModule(_TorchShadowModule):
    """Same parameters / state dict as ``modules.ShadowModule`` (and therefore as the reference's); forward on the
    tcgen05 kernel.  ``capacity``: foreground rows kept for the backward pass (default: a quarter of the pixels, regrown
    automatically when ``strict``); ``strict=False`` never reads the device status (for CUDA-graph capture)."""

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

    def _biases_detached(self):
        return [m.bias.detach() for m in self._linears()[:-1]]

    def _b_out_detached(self):
        return self._linears()[-1].bias.detach()

    def _workspace(self, n_pixels, depth, device, need_hidden):
        cap = self.capacity if self.capacity else max(128, (n_pixels // 4 + 127) // 128 * 128)
        cap = min(cap, (n_pixels + 127) // 128 * 128)
        key = (n_pixels, depth, str(device), cap)
        if self._ws is None or self._ws["key"] != key:
            img_bytes = int(_lib.lib().gom_shadow_mlp_weight_image_bytes(depth))
            e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=device)
            self._ws = dict(key=key, capacity=cap, block_count=e((n_pixels + 1023) // 1024 + 1, dtype=torch.int32),
                            fg_index=torch.zeros(max(n_pixels, cap), dtype=torch.int32, device=device),
                            n_fg=torch.zeros(1, dtype=torch.int32, device=device), w_images=e(img_bytes // 4),
                            bg_value=e(1), status=torch.zeros(1, dtype=torch.int32, device=device), hidden=None)
        if need_hidden and self._ws["hidden"] is None:
            # zeros once: rows the kernel never writes must stay finite for the padded backward GEMMs
            self._ws["hidden"] = torch.zeros(depth, self.width, cap, dtype=torch.float32, device=device)
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
        out = _ShadowMlp.apply(self, flat, lin[-1].weight, lin[-1].bias, *wb)
        return out.reshape(*shape, 1)

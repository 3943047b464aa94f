"""Oracle (test infrastructure, not product code): float64 restatement of the reference's pseudo-shading MLP.

Follows reference models/modules/shadow_module.py:
  * Embedder / get_embedder (:14-62): [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(m-1) x), cos(2^(m-1) x)], m = multires,
    frequencies 2 ** linspace(0, m-1, m);
  * ShadowModule.forward (:107-117): Linear+ReLU x mlp_depth (no skip inside the depth for the shipped configs:
    skips = [4] > mlp_depth = 3, exps/zju-mocap_377.yaml:91-98), Linear(width, 1), sigmoid.
Pinned by tests/golden/golden_modules.npz (the reference's own module run by oracle/make_golden.py::modules_golden):
tests/test_oracle_golden.py::test_shadow_oracle_matches_reference_module.
"""
from __future__ import annotations

import numpy as np


def posenc(x, multires):
    x = np.asarray(x, dtype=np.float64)
    out = [x]
    for k in range(multires):
        out += [np.sin(x * 2.0 ** k), np.cos(x * 2.0 ** k)]
    return np.concatenate(out, axis=-1)


def shadow_forward(normals, weights, biases, multires=6):
    """normals [...,3]; weights/biases: the Linear layers in order (last one has 1 output) -> sigmoid output [...,1]."""
    h = posenc(normals, multires)
    n = len(weights)
    for i, (W, b) in enumerate(zip(weights, biases)):
        h = h @ np.asarray(W, np.float64).T + np.asarray(b, np.float64)
        if i < n - 1:
            h = np.maximum(h, 0.0)
    return 1.0 / (1.0 + np.exp(-h))


def shadow_forward_backward(normals, weights, biases, g_out, multires=6):
    """float64 torch autograd over the same restatement: returns (out, d normals, [dW], [db]) for loss = sum(out * g_out)."""
    import torch
    x = torch.tensor(np.asarray(normals), dtype=torch.float64, requires_grad=True)
    Ws = [torch.tensor(np.asarray(W), dtype=torch.float64, requires_grad=True) for W in weights]
    bs = [torch.tensor(np.asarray(b), dtype=torch.float64, requires_grad=True) for b in biases]
    enc = [x]
    for k in range(multires):
        enc += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
    h = torch.cat(enc, dim=-1)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        h = h @ W.T + b
        if i < len(Ws) - 1:
            h = torch.relu(h)
    out = torch.sigmoid(h)
    (out * torch.tensor(np.asarray(g_out), dtype=torch.float64).reshape(out.shape)).sum().backward()
    return (out.detach().numpy(), x.grad.numpy(), [W.grad.numpy() for W in Ws], [b.grad.numpy() for b in bs])

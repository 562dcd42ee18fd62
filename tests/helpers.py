"""Shared helpers for the parity tests (seeded synthetic RBMs, tolerance checks)."""
import numpy as np
import torch

RTOL = 1e-10  # north_star: energies, gradients and normalisations within 1e-10 relative of the reference


def make_binary_params(n, nh, seed, bias_scale=0.1, weight_scale=None):
    rng = np.random.default_rng(seed)
    ws = 1.0 / np.sqrt(n) if weight_scale is None else weight_scale
    W = rng.standard_normal((nh, n)) * ws
    b = rng.standard_normal(n) * bias_scale
    c = rng.standard_normal(nh) * bias_scale
    return W, b, c


def make_purification_params(n, nh, na, seed, bias_scale=0.1):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((nh, n)) / np.sqrt(n)
    U = rng.standard_normal((na, n)) / np.sqrt(n)
    b = rng.standard_normal(n) * bias_scale
    c = rng.standard_normal(nh) * bias_scale
    d = rng.standard_normal(na) * bias_scale
    return W, U, b, c, d


def load_binary(rbm, W, b, c):
    with torch.no_grad():
        rbm.weights.copy_(torch.as_tensor(W, dtype=torch.double))
        rbm.visible_bias.copy_(torch.as_tensor(b, dtype=torch.double))
        rbm.hidden_bias.copy_(torch.as_tensor(c, dtype=torch.double))
    return rbm


def load_purification(rbm, W, U, b, c, d):
    with torch.no_grad():
        rbm.weights_W.copy_(torch.as_tensor(W, dtype=torch.double))
        rbm.weights_U.copy_(torch.as_tensor(U, dtype=torch.double))
        rbm.visible_bias.copy_(torch.as_tensor(b, dtype=torch.double))
        rbm.hidden_bias.copy_(torch.as_tensor(c, dtype=torch.double))
        rbm.aux_bias.copy_(torch.as_tensor(d, dtype=torch.double))
    return rbm


def random_states(rows, n, seed):
    return (np.random.default_rng(seed).random((rows, n)) < 0.5).astype(np.float64)


def assert_close(actual, expected, rtol=RTOL, what=""):
    a = actual.detach().cpu().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual)
    e = np.asarray(expected)
    assert a.shape == e.shape, f"{what}: shape {a.shape} != {e.shape}"
    scale = max(float(np.max(np.abs(e))) if e.size else 0.0, 1e-300)
    err = float(np.max(np.abs(a - e))) / scale if e.size else 0.0
    assert err <= rtol, f"{what}: max relative error {err:.3e} > {rtol:.1e}"

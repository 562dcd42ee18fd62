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


ATOL_FRACTION = 1e-13  # absolute floor of the elementwise gate, as a fraction of the largest reference entry


def assert_close(actual, expected, rtol=RTOL, what="", elementwise=True, atol_fraction=ATOL_FRACTION, terms=None):
    """Two gates. (1) max-norm: max|a - e| <= rtol * max|e|. (2) elementwise: |a_i - e_i| <= rtol * |e_i| + floor for
    EVERY entry, so that a small entry cannot hide behind a large one; the floor is ``atol_fraction * max|e|`` (1e-13
    of the largest entry: rounding noise of the reference's own float64 sums).
    ``terms`` (same shape as ``expected``) replaces both gates by a written cancellation bound for entries that are
    differences of larger quantities: |a_i - e_i| <= rtol * terms_i, where terms_i is the sum of the ABSOLUTE values of
    the quantities entry i is built from (>= |e_i|) — the relative error of every ingredient is within rtol."""
    a = actual.detach().cpu().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual)
    e = np.asarray(expected)
    assert a.shape == e.shape, f"{what}: shape {a.shape} != {e.shape}"
    if not e.size:
        return
    if terms is not None:
        t = np.broadcast_to(np.asarray(terms, dtype=np.float64), e.shape)
        assert np.all(t >= np.abs(e) * (1 - 1e-12)), f"{what}: `terms` must bound |expected|"
        bad = np.abs(a - e) > rtol * t
        assert not np.any(bad), (f"{what}: {int(bad.sum())} of {e.size} entries exceed rtol {rtol:.0e} x their terms; worst ratio "
                                 f"{float(np.max(np.abs(a - e) / np.maximum(t, 1e-300))):.3e}")
        return
    scale = max(float(np.max(np.abs(e))), 1e-300)
    err = float(np.max(np.abs(a - e))) / scale
    assert err <= rtol, f"{what}: max relative error {err:.3e} > {rtol:.1e}"
    if elementwise:
        floor = atol_fraction * scale
        bad = np.abs(a - e) > rtol * np.abs(e) + floor
        if np.any(bad):
            i = np.unravel_index(np.argmax(np.where(bad, np.abs(a - e) / np.maximum(np.abs(e), 1e-300), 0.0)), e.shape)
            raise AssertionError(f"{what}: {int(bad.sum())} of {e.size} entries fail the elementwise gate (rtol {rtol:.0e}, floor "
                                 f"{atol_fraction:.0e} x max|e|); worst at {i}: got {a[i]!r}, want {e[i]!r}")

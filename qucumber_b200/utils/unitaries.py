"""Basis rotations for the CUDA drop-in (reference: qucumber/utils/unitaries.py).

``rotate_psi_inner_prod`` / ``rotate_rho_probs`` call the fused kernels qcb_cwf_rotated /
qcb_dm_rotated: the expansion of every sample over its non-Z sites, the unitary products, psi / rho~
and the sums all happen on the GPU (the reference builds the expansion with NumPy on the host).
``rotate_psi`` / ``rotate_rho`` rotate a full-space state by the Kronecker product of one-site
unitaries with one tensordot per site on the device.
"""
import numpy as np
import torch

from .. import _lib
from . import cplx


def create_dict(**kwargs):
    """Dictionary of one-site unitaries stored as (2, 2, 2) = [re/im][row][col]; X, Y, Z rotate
    from the computational basis into that Pauli basis (unitaries.py:22-61)."""
    r = 1.0 / np.sqrt(2.0)
    out = {
        "X": torch.tensor([[[r, r], [r, -r]], [[0.0, 0.0], [0.0, 0.0]]], dtype=torch.double),
        "Y": torch.tensor([[[r, 0.0], [r, 0.0]], [[0.0, -r], [0.0, r]]], dtype=torch.double),
        "Z": torch.tensor([[[1.0, 0.0], [0.0, 1.0]], [[0.0, 0.0], [0.0, 0.0]]], dtype=torch.double),
    }
    for name, mat in kwargs.items():
        t = mat.clone().detach() if isinstance(mat, torch.Tensor) else torch.tensor(mat)
        out[name] = t.to(dtype=torch.double)
    return out


_encode_cache = {}


def encode_basis(basis, unitary_dict):
    """basis (sequence of names) -> (uint8 codes on the host, (K, 2, 2, 2) table, names).
    Code 0 is the identity "Z"; every other name in the dictionary gets its own code. The training loop asks for the same
    few bases in every batch: the result is cached per (names of the dictionary, basis)."""
    key = (tuple(unitary_dict), tuple(str(b) for b in basis))
    hit = _encode_cache.get(key)
    if hit is not None:
        return hit
    if len(_encode_cache) > 4096:
        _encode_cache.clear()
    _encode_cache[key] = out = _encode_basis(basis, unitary_dict)
    return out


def _encode_basis(basis, unitary_dict):
    names = ["Z"] + sorted(k for k in unitary_dict if k != "Z")
    if len(names) > 255:
        raise ValueError("too many unitaries")
    lookup = {n: i for i, n in enumerate(names)}
    try:
        codes = np.array([lookup[str(b)] for b in basis], dtype=np.uint8)
    except KeyError as exc:
        raise KeyError(f"basis element {exc} is not in the unitary dictionary") from None
    return codes, names


_table_cache = {}


def unitary_table(unitary_dict, names, device):
    """(K, 2, 2, 2) device table of the dictionary's unitaries in ``names`` order. The training loop asks for the
    same table for every basis of every batch, so it is cached per (dictionary contents, device): the key holds
    each tensor's identity and version counter, i.e. an in-place edit or a replaced entry builds a new table."""
    key = (str(device), tuple((n, id(unitary_dict[n]), unitary_dict[n]._version) for n in names))
    hit = _table_cache.get(key)
    if hit is None:
        if len(_table_cache) > 64:
            _table_cache.clear()
        hit = torch.stack([unitary_dict[n].to(dtype=torch.double) for n in names]).to(device).contiguous()
        _table_cache[key] = (hit, [unitary_dict[n] for n in names])  # keep the sources alive: ids stay unique
        return hit
    return hit[0]


def _kron_apply(us, x):
    """Apply U_0 (x) U_1 (x) ... to the state index of x (2, 2^n, ...) — one contraction per site
    (what the reference's _kron_mult butterfly computes, unitaries.py:64-83)."""
    n = len(us)
    if x.shape[1] != (1 << n):
        raise ValueError(f"state has {x.shape[1]} rows, expected 2^{n}")
    _lib.require_cuda(x, "rotate")
    lib = _lib.load()
    out = x.to(dtype=torch.double).contiguous().clone()  # (2, 2^n, *tail): re and im planes, rotated in place
    eye = torch.tensor([[[1.0, 0.0], [0.0, 1.0]], [[0.0, 0.0], [0.0, 0.0]]], dtype=torch.double)
    mats = [u.detach().to(device="cpu", dtype=torch.double) for u in us]
    codes = np.array([0 if torch.equal(m, eye) else s + 1 for s, m in enumerate(mats)], dtype=np.uint8)
    table = torch.stack([eye] + mats).to(out.device).contiguous()
    batch = out[0].numel() >> n
    _lib.check(lib.qcb_kron_rotate(_lib.ptr(out[0]), _lib.ptr(out[1]), n, batch, codes.ctypes.data, _lib.ptr(table), n + 1,
                                   _lib.stream_ptr(out.device)), "qcb_kron_rotate")
    return out


def rotate_psi(nn_state, basis, space, unitaries=None, psi=None):
    """Rotate the full-space wavefunction into ``basis`` (unitaries.py:86-115)."""
    psi = nn_state.psi(space) if psi is None else psi.to(dtype=torch.double, device=nn_state.device)
    unitaries = unitaries if unitaries else nn_state.unitary_dict
    return _kron_apply([unitaries[b] for b in basis], psi)


def rotate_rho(nn_state, basis, space, unitaries=None, rho=None):
    """U rho U^dagger over the full space (unitaries.py:118-150)."""
    rho = nn_state.rho(space, space) if rho is None else rho.to(dtype=torch.double, device=nn_state.device)
    unitaries = unitaries if unitaries else nn_state.unitary_dict
    us = [unitaries[b] for b in basis]
    rho_r = _kron_apply(us, rho)
    return _kron_apply(us, cplx.conjugate(rho_r))


def _convert_basis_element_to_index(states):
    powers = (2 ** (torch.arange(states.shape[-1], 0, -1) - 1)).to(states)
    return torch.matmul(states, powers)


def _expand(nn_state, basis, states, unitaries):
    """Host/torch restatement of the expansion used only for the table-lookup variants (psi= / rho=
    given): returns Ut (complex, (2^S, B)) and v (2^S, B, N)."""
    unitaries = unitaries if unitaries else nn_state.unitary_dict
    basis = np.array(list(basis))
    sites = np.where(basis != "Z")[0]
    S = sites.size
    dev = states.device
    if S == 0:
        return torch.ones(1, states.shape[0], dtype=torch.complex128, device=dev), states.unsqueeze(0)
    idx = torch.arange(1 << S, device=dev)
    bits = ((idx[:, None] >> torch.arange(S - 1, -1, -1, device=dev)[None, :]) & 1)  # (2^S, S)
    v = states.unsqueeze(0).repeat(1 << S, 1, 1)
    v[..., torch.as_tensor(sites, device=dev)] = bits.to(states).unsqueeze(1)
    sig = states[:, torch.as_tensor(sites, device=dev)].round().long()  # (B, S)
    Ut = torch.ones(1 << S, states.shape[0], dtype=torch.complex128, device=dev)
    for t, site in enumerate(sites):
        u = unitaries[basis[site]].to(dev)
        uc = torch.complex(u[0], u[1])
        Ut = Ut * uc[sig[None, :, t], bits[:, None, t]]
    return Ut, v


def rotate_psi_inner_prod(nn_state, basis, states, unitaries=None, psi=None, include_extras=False):
    """Amplitude of each sample after rotating the wavefunction into ``basis``
    (unitaries.py:189-233). With ``psi=None`` this is one fused CUDA call; ``include_extras``
    additionally returns the per-term products and the expanded states like the reference."""
    states = states.to(device=nn_state.device, dtype=torch.double)
    if psi is None and not include_extras:
        return nn_state._rotated_amplitudes(basis, states, unitaries)
    Ut, v = _expand(nn_state, basis, states, unitaries)
    if psi is None:
        p = nn_state.psi(v).detach()
    else:
        idx = _convert_basis_element_to_index(v).long()
        p = psi.to(device=nn_state.device, dtype=torch.double)[:, idx]
    Upsi_v = cplx.make_complex(Ut.real * p[0] - Ut.imag * p[1], Ut.real * p[1] + Ut.imag * p[0])
    Upsi = torch.sum(Upsi_v, dim=1)
    return (Upsi, Upsi_v, v) if include_extras else Upsi


def rotate_rho_probs(nn_state, basis, states, unitaries=None, rho=None, include_extras=False):
    """Born-rule probability of each sample after rotating the density matrix into ``basis``
    (unitaries.py:236-283)."""
    states = states.to(device=nn_state.device, dtype=torch.double)
    if rho is None and not include_extras:
        return nn_state._rotated_probs(basis, states, unitaries)
    Ut, v = _expand(nn_state, basis, states, unitaries)
    Ut2 = Ut.unsqueeze(1) * torch.conj(Ut).unsqueeze(0)  # (i, j, b)
    if rho is None:
        r = nn_state.rho(v).detach()
    else:
        idx = _convert_basis_element_to_index(v).long()
        r = rho.to(device=nn_state.device, dtype=torch.double)[:, idx.unsqueeze(0), idx.unsqueeze(1)]
    UrhoU_v = cplx.make_complex(Ut2.real * r[0] - Ut2.imag * r[1], Ut2.real * r[1] + Ut2.imag * r[0])
    UrhoU = torch.sum(cplx.real(UrhoU_v), dim=(0, 1))
    return (UrhoU, UrhoU_v, v) if include_extras else UrhoU


def _call_cwf(rbm_am, rbm_ph, basis, states, unitary_dict, g_am=None, g_ph=None):
    """qcb_cwf_rotated wrapper: returns Upsi (2, B); accumulates into g_am / g_ph when given."""
    lib = _lib.load()
    codes, names = encode_basis(basis, unitary_dict)
    dev = states.device
    table = unitary_table(unitary_dict, names, dev)
    B, N = states.shape
    states = states.contiguous()
    Upsi = torch.empty(2, B, dtype=torch.double, device=dev)
    ws = _lib.workspace(dev, lib.qcb_workspace_bytes(_lib.OP_CWF_ROTATED, N, rbm_am.num_hidden, 0, B))
    wa, ba, ca = rbm_am._params()
    wp, bp, cp = rbm_ph._params()
    rc = lib.qcb_cwf_rotated(_lib.ptr(wa), _lib.ptr(ba), _lib.ptr(ca), _lib.ptr(wp), _lib.ptr(bp), _lib.ptr(cp), N,
                             rbm_am.num_hidden, _lib.ptr(states), B, codes.ctypes.data, _lib.ptr(table), len(names),
                             _lib.ptr(Upsi), _lib.ptr(g_am), _lib.ptr(g_ph), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "qcb_cwf_rotated")
    return Upsi


def _call_dm(rbm_am, rbm_ph, basis, states, unitary_dict, g_am=None, g_ph=None):
    """qcb_dm_rotated wrapper: returns UrhoU (B,); accumulates into g_am / g_ph when given."""
    lib = _lib.load()
    codes, names = encode_basis(basis, unitary_dict)
    dev = states.device
    table = unitary_table(unitary_dict, names, dev)
    B, N = states.shape
    states = states.contiguous()
    out = torch.empty(B, dtype=torch.double, device=dev)
    ws = _lib.workspace(dev, lib.qcb_workspace_bytes(_lib.OP_DM_ROTATED, N, rbm_am.num_hidden, rbm_am.num_aux, B))
    ta, tp = rbm_am._tensors(), rbm_ph._tensors()
    rc = lib.qcb_dm_rotated(*[_lib.ptr(t) for t in ta], *[_lib.ptr(t) for t in tp], N, rbm_am.num_hidden, rbm_am.num_aux,
                            _lib.ptr(states), B, codes.ctypes.data, _lib.ptr(table), len(names), _lib.ptr(out),
                            _lib.ptr(g_am), _lib.ptr(g_ph), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "qcb_dm_rotated")
    return out


def _call_groups(kind, rbm_am, rbm_ph, unique_bases, counts, grouped, unitary_dict, g_am, g_ph):
    """qcb_cwf_rotated_groups / qcb_dm_rotated_groups: the rows of ``grouped`` (already sorted by basis: ``counts[i]`` rows
    of ``unique_bases[i]`` each) in ONE call that accumulates into g_am / g_ph — the per-basis loop runs in C."""
    lib = _lib.load()
    dev = grouped.device
    B, N = grouped.shape
    grouped = grouped.contiguous()
    counts = np.asarray(counts, dtype=np.int64)
    offsets = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    codes = None
    names = None
    rows = []
    for basis in unique_bases:
        c, names = encode_basis(basis, unitary_dict)
        rows.append(c)
    codes = np.ascontiguousarray(np.stack(rows).astype(np.uint8))
    table = unitary_table(unitary_dict, names, dev)
    bmax = int(counts.max()) if len(counts) else 0
    scratch = torch.empty(2 * max(bmax, 1), dtype=torch.double, device=dev)
    if kind == "cwf":
        ws = _lib.workspace(dev, lib.qcb_workspace_bytes(_lib.OP_CWF_ROTATED, N, rbm_am.num_hidden, 0, bmax))
        wa, ba, ca = rbm_am._params()
        wp, bp, cp = rbm_ph._params()
        rc = lib.qcb_cwf_rotated_groups(_lib.ptr(wa), _lib.ptr(ba), _lib.ptr(ca), _lib.ptr(wp), _lib.ptr(bp), _lib.ptr(cp), N,
                                        rbm_am.num_hidden, _lib.ptr(grouped), offsets.ctypes.data, codes.ctypes.data, len(counts),
                                        _lib.ptr(table), len(names), _lib.ptr(scratch), _lib.ptr(g_am), _lib.ptr(g_ph), _lib.ptr(ws),
                                        ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "qcb_cwf_rotated_groups")
    else:
        ws = _lib.workspace(dev, lib.qcb_workspace_bytes(_lib.OP_DM_ROTATED, N, rbm_am.num_hidden, rbm_am.num_aux, bmax))
        ta, tp = rbm_am._tensors(), rbm_ph._tensors()
        rc = lib.qcb_dm_rotated_groups(*[_lib.ptr(t) for t in ta], *[_lib.ptr(t) for t in tp], N, rbm_am.num_hidden, rbm_am.num_aux,
                                       _lib.ptr(grouped), offsets.ctypes.data, codes.ctypes.data, len(counts), _lib.ptr(table),
                                       len(names), _lib.ptr(scratch), _lib.ptr(g_am), _lib.ptr(g_ph), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(dev))
        _lib.check(rc, "qcb_dm_rotated_groups")

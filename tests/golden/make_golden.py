#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE itself (PIQuIL/QuCumber v1.3.3 imported
read-only from /root/reference) on seeded inputs. Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin oracle/rbm_oracle.py to the reference (tests/test_oracle_golden.py) and travel to
the GPU box, where /root/reference does not exist. matplotlib is absent here, and only the
reference's LivePlotting callback needs it, so three stub modules are injected before import.
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.ticker"].MultipleLocator = type("MultipleLocator", (), {"__init__": lambda self, *a, **k: None})
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import qucumber  # noqa: F401

    return qucumber


def cnp(t):
    """qucumber complex tensor (2, ...) -> numpy complex128."""
    t = t.detach().cpu().numpy()
    return t[0] + 1j * t[1]


def replay_uniforms(seed, k, chains, widths):
    """The uniforms torch.bernoulli consumes under `seed` in a k-sweep Gibbs run:
    one torch.rand block per conditional, in draw order."""
    torch.manual_seed(seed)
    rows = []
    for _ in range(k):
        rows.append(torch.cat([torch.rand(chains, w, dtype=torch.double) for w in widths], dim=1))
    return torch.stack(rows).numpy()


def main():
    qucumber = import_reference()
    from qucumber.nn_states import ComplexWaveFunction, DensityMatrix, PositiveWaveFunction
    from qucumber.rbm import BinaryRBM, PurificationRBM
    from qucumber.utils import training_statistics as ts
    from qucumber.utils import unitaries as un
    import qucumber.utils.data as data

    ex = os.path.join(REF, "examples")
    td = os.path.join(REF, "tests", "data")

    # ------------------------------------------------------------------ BinaryRBM / PositiveWaveFunction
    init = np.load(os.path.join(td, "test_training_init_pos_params.npz"))
    train, target = data.load_data(
        os.path.join(ex, "Tutorial1_TrainPosRealWaveFunction", "tfim1d_data.txt"),
        os.path.join(ex, "Tutorial1_TrainPosRealWaveFunction", "tfim1d_psi.txt"),
    )
    nn = PositiveWaveFunction(10, 10, gpu=False)
    with torch.no_grad():
        nn.rbm_am.weights.copy_(torch.tensor(init["weights"], dtype=torch.double))
        nn.rbm_am.visible_bias.copy_(torch.tensor(init["visible_bias"], dtype=torch.double))
        nn.rbm_am.hidden_bias.copy_(torch.tensor(init["hidden_bias"], dtype=torch.double))
    space = nn.generate_hilbert_space()
    rng = np.random.default_rng(11)
    states = torch.tensor((rng.random((64, 10)) < 0.5).astype(np.float64))
    hstates = torch.tensor((rng.random((64, 10)) < 0.5).astype(np.float64))
    k, chains = 5, 64
    u = replay_uniforms(4321, k, chains, [10, 10])
    torch.manual_seed(4321)
    vk = nn.rbm_am.gibbs_steps(k, states)
    out = {
        "weights": init["weights"], "visible_bias": init["visible_bias"], "hidden_bias": init["hidden_bias"],
        "train_packed": np.packbits(train.numpy().astype(np.uint8), axis=1), "target": cnp(target),
        "states": states.numpy(), "hstates": hstates.numpy(),
        "energy": nn.rbm_am.effective_energy(states).numpy(),
        "prob_h": nn.rbm_am.prob_h_given_v(states).numpy(),
        "prob_v": nn.rbm_am.prob_v_given_h(hstates).numpy(),
        "grad_reduce_train": nn.rbm_am.effective_energy_gradient(train).numpy(),
        "grad_batch": nn.rbm_am.effective_energy_gradient(states[:16], reduce=False).numpy(),
        "gibbs_uniforms": u, "gibbs_vk": vk.numpy(), "gibbs_k": k,
        "Z": nn.normalization(space).item(),
        "fidelity": ts.fidelity(nn, target, space),
        "KL": ts.KL(nn, target, space),
        "NLL": ts.NLL(nn, train, space),
        "psi_space": cnp(nn.psi(space)),
        "exact_grad_100": nn.compute_exact_gradients(train[:100], space)[0].numpy(),
    }
    np.savez_compressed(os.path.join(OUT, "binary_rbm_tfim10.npz"), **out)

    # N = nh = 64 (benchmark shape) and an odd shape
    for (n, nh, tag) in [(64, 64, "n64"), (7, 12, "n7")]:
        torch.manual_seed(99 + n)
        rbm = BinaryRBM(n, nh, gpu=False)
        with torch.no_grad():
            rbm.visible_bias.copy_(0.1 * torch.randn(n, dtype=torch.double))
            rbm.hidden_bias.copy_(0.1 * torch.randn(nh, dtype=torch.double))
        v = (torch.rand(48, n) < 0.5).double()
        k, chains = 3, 48
        u = replay_uniforms(777, k, chains, [nh, n])
        torch.manual_seed(777)
        vk = rbm.gibbs_steps(k, v)
        out = {
            "weights": rbm.weights.numpy(), "visible_bias": rbm.visible_bias.numpy(), "hidden_bias": rbm.hidden_bias.numpy(),
            "states": v.numpy(), "energy": rbm.effective_energy(v).numpy(),
            "grad_reduce": rbm.effective_energy_gradient(v).numpy(),
            "grad_batch": rbm.effective_energy_gradient(v[:4], reduce=False).numpy(),
            "gibbs_uniforms": u, "gibbs_vk": vk.numpy(), "gibbs_k": k,
        }
        if n <= 20:
            sp = PositiveWaveFunction(n, nh, gpu=False).generate_hilbert_space()
            out["Z"] = rbm.partition(sp).item()
        np.savez_compressed(os.path.join(OUT, f"binary_rbm_{tag}.npz"), **out)

    # ------------------------------------------------------------------ PurificationRBM
    torch.manual_seed(2024)
    n, nh, na = 5, 4, 3
    prbm = PurificationRBM(n, nh, na, gpu=False)
    with torch.no_grad():
        prbm.visible_bias.copy_(0.2 * torch.randn(n, dtype=torch.double))
        prbm.hidden_bias.copy_(0.2 * torch.randn(nh, dtype=torch.double))
        prbm.aux_bias.copy_(0.2 * torch.randn(na, dtype=torch.double))
    v = (torch.rand(40, n) < 0.5).double()
    vp = (torch.rand(40, n) < 0.5).double()
    h = (torch.rand(40, nh) < 0.5).double()
    a = (torch.rand(40, na) < 0.5).double()
    k, chains = 4, 40
    u = replay_uniforms(31337, k, chains, [nh, na, n])
    torch.manual_seed(31337)
    vk = prbm.gibbs_steps(k, v)
    sp = DensityMatrix(n, nh, na, gpu=False).generate_hilbert_space()
    out = {
        "weights_W": prbm.weights_W.numpy(), "weights_U": prbm.weights_U.numpy(), "visible_bias": prbm.visible_bias.numpy(),
        "hidden_bias": prbm.hidden_bias.numpy(), "aux_bias": prbm.aux_bias.numpy(),
        "v": v.numpy(), "vp": vp.numpy(), "h": h.numpy(), "a": a.numpy(),
        "energy": prbm.effective_energy(v).numpy(), "energy_va": prbm.effective_energy(v, a).numpy(),
        "prob_h": prbm.prob_h_given_v(v).numpy(), "prob_a": prbm.prob_a_given_v(v).numpy(),
        "prob_v": prbm.prob_v_given_ha(h, a).numpy(),
        "grad_reduce": prbm.effective_energy_gradient(v).numpy(),
        "grad_batch": prbm.effective_energy_gradient(v[:6], reduce=False).numpy(),
        "gibbs_uniforms": u, "gibbs_vk": vk.numpy(), "gibbs_k": k,
        "Z": prbm.partition(sp).item(),
        "mixing": prbm.mixing_term(v).numpy(),
        "gamma_plus": prbm.gamma(v[:6], vp[:5], eta=+1, expand=True).numpy(),
        "gamma_minus": prbm.gamma(v[:6], vp[:5], eta=-1, expand=True).numpy(),
        "gamma_minus_noexp": prbm.gamma(v, vp, eta=-1, expand=False).numpy(),
        "gamma_grad_plus": cnp(prbm.gamma_grad(v[:4], vp[:3], eta=+1, expand=True)),
        "gamma_grad_minus_noexp": cnp(prbm.gamma_grad(v[:4], vp[:4], eta=-1, expand=False)),
    }
    np.savez_compressed(os.path.join(OUT, "purification_rbm.npz"), **out)

    # ------------------------------------------------------------------ ComplexWaveFunction
    torch.manual_seed(555)
    n, nh = 5, 4
    cwf = ComplexWaveFunction(n, nh, gpu=False)
    for r in (cwf.rbm_am, cwf.rbm_ph):
        with torch.no_grad():
            r.visible_bias.copy_(0.3 * torch.randn(n, dtype=torch.double))
            r.hidden_bias.copy_(0.3 * torch.randn(nh, dtype=torch.double))
    samples = (torch.rand(12, n) < 0.5).double()
    basis = np.array(list("XZYZX"))
    Upsi, Upsi_v, vexp = un.rotate_psi_inner_prod(cwf, basis, samples, include_extras=True)
    rg = cwf.rotated_gradient(basis, samples)
    bases_mixed = np.array([list("XZYZX"), list("ZZZZZ"), list("ZYZZZ"), list("XXXXX")] * 3)
    gmix = cwf.gradient(samples, bases_mixed)
    sp = cwf.generate_hilbert_space()
    out = {
        "am_weights": cwf.rbm_am.weights.numpy(), "am_visible_bias": cwf.rbm_am.visible_bias.numpy(),
        "am_hidden_bias": cwf.rbm_am.hidden_bias.numpy(), "ph_weights": cwf.rbm_ph.weights.numpy(),
        "ph_visible_bias": cwf.rbm_ph.visible_bias.numpy(), "ph_hidden_bias": cwf.rbm_ph.hidden_bias.numpy(),
        "samples": samples.numpy(), "basis": "".join(basis), "Upsi": cnp(Upsi), "Upsi_v": cnp(Upsi_v), "v_expanded": vexp.numpy(),
        "rot_grad_am": rg[0].numpy(), "rot_grad_ph": rg[1].numpy(),
        "bases_mixed": np.array(["".join(b) for b in bases_mixed]), "grad_mixed_am": gmix[0].numpy(), "grad_mixed_ph": gmix[1].numpy(),
        "psi_space": cnp(cwf.psi(sp)), "Z": cwf.normalization(sp).item(),
        "psi_rotated_full": cnp(un.rotate_psi(cwf, basis, sp)),
    }
    np.savez_compressed(os.path.join(OUT, "complex_wavefunction.npz"), **out)

    # the reference's own gradient-test fixture (tests/data/test_grad_data.pkl), 2-qubit complex case + Tutorial 2 metrics
    with open(os.path.join(td, "test_grad_data.pkl"), "rb") as f:
        tg = pickle.load(f)
    qucumber.set_random_seed(1234, cpu=True, gpu=False, quiet=True)
    q2 = tg["2qubits"]
    s2 = torch.tensor(q2["train_samples"], dtype=torch.double)
    b2 = q2["train_bases"]
    c2 = ComplexWaveFunction(2, 2, gpu=False)
    sp2 = c2.generate_hilbert_space()
    tpsi = torch.tensor(q2["target_psi"], dtype=torch.double).t()
    all_bases = np.array([list(b) for b in np.array(q2["bases"]).reshape(-1)]) if np.array(q2["bases"]).ndim == 1 else np.array(q2["bases"])
    tdict = {}
    for i, bb in enumerate(all_bases):
        tdict["".join(bb)] = tpsi[:, 4 * i: 4 * (i + 1)]
    bases_list = list(tdict.keys())
    g2 = c2.compute_exact_gradients(s2, sp2, bases_batch=b2)
    out = {
        "am_weights": c2.rbm_am.weights.numpy(), "am_visible_bias": c2.rbm_am.visible_bias.numpy(),
        "am_hidden_bias": c2.rbm_am.hidden_bias.numpy(), "ph_weights": c2.rbm_ph.weights.numpy(),
        "ph_visible_bias": c2.rbm_ph.visible_bias.numpy(), "ph_hidden_bias": c2.rbm_ph.hidden_bias.numpy(),
        "train_samples": s2.numpy(), "train_bases": np.array(["".join(b) for b in b2]),
        "bases": np.array(bases_list), "target_by_basis": np.stack([cnp(tdict[b]) for b in bases_list]),
        "KL": ts.KL(c2, tdict, sp2, bases=bases_list), "NLL": ts.NLL(c2, s2, sp2, sample_bases=b2),
        "exact_grad_am": g2[0].numpy(), "exact_grad_ph": g2[1].numpy(),
        "fidelity_ZZ": ts.fidelity(c2, tdict[bases_list[0]] if "ZZ" not in tdict else tdict["ZZ"], sp2),
    }
    np.savez_compressed(os.path.join(OUT, "complex_2qubits.npz"), **out)

    # ------------------------------------------------------------------ DensityMatrix
    torch.manual_seed(808)
    n, nh, na = 4, 3, 5
    dm = DensityMatrix(n, nh, na, gpu=False)
    for r in (dm.rbm_am, dm.rbm_ph):
        with torch.no_grad():
            r.visible_bias.copy_(0.3 * torch.randn(n, dtype=torch.double))
            r.hidden_bias.copy_(0.3 * torch.randn(nh, dtype=torch.double))
    with torch.no_grad():
        dm.rbm_am.aux_bias.copy_(0.3 * torch.randn(na, dtype=torch.double))
    samples = (torch.rand(10, n) < 0.5).double()
    basis = np.array(list("XZYY"))
    UrhoU, UrhoU_v, vexp = un.rotate_rho_probs(dm, basis, samples, include_extras=True)
    rg = dm.rotated_gradient(basis, samples)
    bases_mixed = np.array([list("XZYY"), list("ZZZZ"), list("ZZXZ"), list("YYYY"), list("ZXZZ")] * 2)
    gmix = dm.gradient(samples, bases_mixed)
    sp = dm.generate_hilbert_space()
    out = {}
    for tag, r in (("am", dm.rbm_am), ("ph", dm.rbm_ph)):
        out.update({f"{tag}_weights_W": r.weights_W.numpy(), f"{tag}_weights_U": r.weights_U.numpy(),
                    f"{tag}_visible_bias": r.visible_bias.numpy(), f"{tag}_hidden_bias": r.hidden_bias.numpy(),
                    f"{tag}_aux_bias": r.aux_bias.numpy()})
    out.update({
        "samples": samples.numpy(), "basis": "".join(basis), "UrhoU": UrhoU.numpy(), "UrhoU_v": cnp(UrhoU_v),
        "rot_grad_am": rg[0].numpy(), "rot_grad_ph": rg[1].numpy(),
        "bases_mixed": np.array(["".join(b) for b in bases_mixed]), "grad_mixed_am": gmix[0].numpy(), "grad_mixed_ph": gmix[1].numpy(),
        "rho_space": cnp(dm.rho(sp, sp)), "pi_space": cnp(dm.pi(sp, sp)), "Z": dm.normalization(sp).item(),
        "pi_grad_am": cnp(dm.pi_grad(sp[:5], sp[3:7], phase=False, expand=True)),
        "pi_grad_ph": cnp(dm.pi_grad(sp[:5], sp[3:7], phase=True, expand=True)),
    })
    np.savez_compressed(os.path.join(OUT, "density_matrix.npz"), **out)

    # Tutorial-3 / test_grad_data density-matrix case: metrics on the reference's own fixture
    qucumber.set_random_seed(1234, cpu=True, gpu=False, quiet=True)
    qd = tg["density_matrix"]
    sd = torch.tensor(qd["train_samples"], dtype=torch.double)
    bd = qd["train_bases"]
    d2 = DensityMatrix(2, 2, 2, gpu=False)
    spd = d2.generate_hilbert_space()
    trho = torch.tensor(qd["density_matrix"], dtype=torch.double)
    basesd = ["".join(b) for b in np.array(qd["bases"])]
    gd = d2.compute_exact_gradients(sd, spd, bases_batch=bd)
    out = {}
    for tag, r in (("am", d2.rbm_am), ("ph", d2.rbm_ph)):
        out.update({f"{tag}_weights_W": r.weights_W.numpy(), f"{tag}_weights_U": r.weights_U.numpy(),
                    f"{tag}_visible_bias": r.visible_bias.numpy(), f"{tag}_hidden_bias": r.hidden_bias.numpy(),
                    f"{tag}_aux_bias": r.aux_bias.numpy()})
    out.update({
        "train_samples": sd.numpy(), "train_bases": np.array(["".join(b) for b in bd]), "bases": np.array(basesd),
        "target_rho": cnp(trho), "KL": ts.KL(d2, trho, spd, bases=basesd), "NLL": ts.NLL(d2, sd, spd, sample_bases=bd),
        "fidelity": ts.fidelity(d2, trho, spd), "exact_grad_am": gd[0].numpy(), "exact_grad_ph": gd[1].numpy(),
    })
    np.savez_compressed(os.path.join(OUT, "density_2qubits.npz"), **out)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

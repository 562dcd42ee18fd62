"""Pins the CPU oracle (oracle/rbm_oracle.py) to the reference: every golden vector below was
produced by the reference implementation itself (tests/golden/make_golden.py, run in the build
container with /root/reference imported read-only). CPU only."""
import os

import numpy as np
import pytest

from oracle import rbm_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    assert float(np.max(np.abs(a - b))) / scale <= tol, float(np.max(np.abs(a - b))) / scale


def split_bases(arr):
    return np.array([list(s) for s in arr])


def test_binary_tfim10():
    g = load("binary_rbm_tfim10.npz")
    rbm = orc.BinaryRBMOracle(g["weights"], g["visible_bias"], g["hidden_bias"])
    train = np.unpackbits(g["train_packed"], axis=1)[:, :10].astype(np.float64)
    close(rbm.effective_energy(g["states"]), g["energy"])
    close(rbm.prob_h_given_v(g["states"]), g["prob_h"])
    close(rbm.prob_v_given_h(g["hstates"]), g["prob_v"])
    close(rbm.effective_energy_gradient(train), g["grad_reduce_train"])
    close(rbm.effective_energy_gradient(g["states"][:16], reduce=False), g["grad_batch"])
    # seed replay: the reference's gibbs_steps under torch.manual_seed == thresholding the same uniforms
    assert np.array_equal(rbm.gibbs_steps_fed(int(g["gibbs_k"]), g["states"], g["gibbs_uniforms"]), g["gibbs_vk"])
    Z = rbm.partition()
    close(Z, g["Z"])
    space = orc.hilbert_rows(10)
    close(rbm.partition(space), g["Z"])
    psi = orc.psi_positive(rbm, space)
    close(psi, g["psi_space"])
    close(orc.fidelity_pure(psi, Z, g["target"]), g["fidelity"], 1e-11)
    close(orc.KL_positive(rbm, g["target"]), g["KL"], 1e-11)
    close(orc.NLL_positive(rbm, train), g["NLL"], 1e-11)
    close(orc.exact_gradient_positive(rbm, train[:100]), g["exact_grad_100"], 1e-11)
    # known answers quoted in SURVEY.md 8(c)
    assert abs(g["Z"] - 1.320537595512757e12) <= 1e-3 and abs(g["fidelity"] - 5.510884086324108e-02) <= 1e-14
    # independent identity: tracing the visible units instead of the hidden ones
    assert abs(rbm.log_partition_hidden_side() - np.log(Z)) <= 1e-7


@pytest.mark.parametrize("tag", ["n64", "n7"])
def test_binary_shapes(tag):
    g = load(f"binary_rbm_{tag}.npz")
    rbm = orc.BinaryRBMOracle(g["weights"], g["visible_bias"], g["hidden_bias"])
    close(rbm.effective_energy(g["states"]), g["energy"])
    close(rbm.effective_energy_gradient(g["states"]), g["grad_reduce"])
    close(rbm.effective_energy_gradient(g["states"][:4], reduce=False), g["grad_batch"])
    assert np.array_equal(rbm.gibbs_steps_fed(int(g["gibbs_k"]), g["states"], g["gibbs_uniforms"]), g["gibbs_vk"])
    if "Z" in g.files:
        close(rbm.partition(), g["Z"])


def test_hilbert_order():
    rows = orc.hilbert_rows(4)
    assert rows.shape == (16, 4) and rows[1].tolist() == [0, 0, 0, 1] and rows[8].tolist() == [1, 0, 0, 0]
    assert np.array_equal(orc.hilbert_rows(5, 7, 9), orc.hilbert_rows(5)[7:16])
    assert np.array_equal(orc.basis_index(orc.hilbert_rows(6)), np.arange(64))


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 distribution)."""
    def run(ctr, key):
        return [int(x) for x in orc.philox4x32_10(*[np.uint64(c) for c in ctr], key[0], key[1])]

    assert run((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = orc.philox_uniforms(12345, np.arange(4), 3, 8, 6)
    assert u.shape == (4, 6) and np.all((u >= 0) & (u < 1))
    assert np.all(u * 2.0 ** 53 == np.floor(u * 2.0 ** 53))
    # layer convention of the Gibbs kernel: 53-bit grid, eight units per call, all units distinct streams
    lu = orc.layer_uniforms(77, np.arange(3), 5, orc.visible_q0(40), 70)
    assert lu.shape == (3, 70) and np.all((lu >= 0) & (lu < 1)) and np.all(lu * 2.0 ** 53 == np.floor(lu * 2.0 ** 53))
    assert len(np.unique(lu)) == lu.size
    hi = np.stack(orc.philox4x32_10(np.uint64(1), np.uint64(0), np.uint64(5), np.uint64(orc.visible_q0(40) + 4 * 1 + 2), 77, 0), -1)
    # unit 32*1 + 16*1 + 4*2 + 3 = 59 of chain 1: call q0 + 4*1 + 2, word 3, low half
    assert int(lu[1, 59] * 2.0 ** 53) >> 37 == int(hi[3]) & 0xFFFF
    assert orc.visible_q0(64) == 8 and orc.visible_q0(10) == 4 and orc.visible_q0(33) == 8


def test_purification():
    g = load("purification_rbm.npz")
    rbm = orc.PurificationRBMOracle(g["weights_W"], g["weights_U"], g["visible_bias"], g["hidden_bias"], g["aux_bias"])
    close(rbm.effective_energy(g["v"]), g["energy"])
    close(rbm.effective_energy(g["v"], g["a"]), g["energy_va"])
    close(rbm.prob_h_given_v(g["v"]), g["prob_h"])
    close(rbm.prob_a_given_v(g["v"]), g["prob_a"])
    close(rbm.prob_v_given_ha(g["h"], g["a"]), g["prob_v"])
    close(rbm.effective_energy_gradient(g["v"]), g["grad_reduce"])
    close(rbm.effective_energy_gradient(g["v"][:6], reduce=False), g["grad_batch"])
    assert np.array_equal(rbm.gibbs_steps_fed(int(g["gibbs_k"]), g["v"], g["gibbs_uniforms"]), g["gibbs_vk"])
    close(rbm.partition(), g["Z"])
    close(rbm.mixing_term(g["v"]), g["mixing"])
    close(rbm.gamma(g["v"][:6], g["vp"][:5], +1, True), g["gamma_plus"])
    close(rbm.gamma(g["v"][:6], g["vp"][:5], -1, True), g["gamma_minus"])
    close(rbm.gamma(g["v"], g["vp"], -1, False), g["gamma_minus_noexp"])
    close(rbm.gamma_grad(g["v"][:4], g["vp"][:3], +1, True), g["gamma_grad_plus"].real)
    close(rbm.gamma_grad(g["v"][:4], g["vp"][:4], -1, False), g["gamma_grad_minus_noexp"].real)
    assert np.all(g["gamma_grad_plus"].imag == 0)


def _complex_pair(g):
    am = orc.BinaryRBMOracle(g["am_weights"], g["am_visible_bias"], g["am_hidden_bias"])
    ph = orc.BinaryRBMOracle(g["ph_weights"], g["ph_visible_bias"], g["ph_hidden_bias"])
    return am, ph


def test_complex_wavefunction():
    g = load("complex_wavefunction.npz")
    am, ph = _complex_pair(g)
    basis = list(str(g["basis"]))
    Upsi, Upsi_v, v = orc.rotate_psi_inner_prod(lambda x: orc.psi_complex(am, ph, x), basis, g["samples"], None, True)
    assert np.array_equal(v, g["v_expanded"])
    close(Upsi_v, g["Upsi_v"])
    close(Upsi, g["Upsi"])
    ga, gp = orc.complex_rotated_gradient(am, ph, basis, g["samples"])
    close(ga, g["rot_grad_am"], 1e-11)
    close(gp, g["rot_grad_ph"], 1e-11)
    gm = orc.complex_gradient(am, ph, g["samples"], split_bases(g["bases_mixed"]))
    close(gm[0], g["grad_mixed_am"], 1e-11)
    close(gm[1], g["grad_mixed_ph"], 1e-11)
    space = orc.hilbert_rows(am.n)
    psi = orc.psi_complex(am, ph, space)
    close(psi, g["psi_space"])
    close(am.partition(), g["Z"])
    close(orc.kron_rotate(basis, psi), g["psi_rotated_full"])
    # rotate_psi_inner_prod over the full space == rotate_psi (reference tests/test_models_misc.py:318-345)
    close(orc.rotate_psi_inner_prod(lambda x: orc.psi_complex(am, ph, x), basis, space), g["psi_rotated_full"], 1e-11)


def test_complex_2qubits_reference_fixture():
    g = load("complex_2qubits.npz")
    am, ph = _complex_pair(g)
    bases = [str(b) for b in g["bases"]]
    tdict = {b: g["target_by_basis"][i] for i, b in enumerate(bases)}
    close(orc.KL_complex(am, ph, tdict, bases), g["KL"], 1e-11)
    tb = split_bases(g["train_bases"])
    close(orc.NLL_complex(am, ph, g["train_samples"], tb), g["NLL"], 1e-11)
    ge = orc.exact_gradient_with_bases("complex", am, ph, g["train_samples"], tb)
    close(ge[0], g["exact_grad_am"], 1e-10)
    close(ge[1], g["exact_grad_ph"], 1e-10)


def _dm_pair(g):
    am = orc.PurificationRBMOracle(g["am_weights_W"], g["am_weights_U"], g["am_visible_bias"], g["am_hidden_bias"], g["am_aux_bias"])
    ph = orc.PurificationRBMOracle(g["ph_weights_W"], g["ph_weights_U"], g["ph_visible_bias"], g["ph_hidden_bias"], g["ph_aux_bias"])
    return am, ph


def test_density_matrix():
    g = load("density_matrix.npz")
    am, ph = _dm_pair(g)
    basis = list(str(g["basis"]))
    space = orc.hilbert_rows(am.n)
    close(orc.dm_pi(am, ph, space, space), g["pi_space"])
    rho = orc.dm_rho(am, ph, space, space)
    close(rho, g["rho_space"])
    assert np.allclose(rho, rho.conj().T, atol=0, rtol=1e-13)  # Hermitian (reference tests/test_models_misc.py:148-156)
    close(am.partition(), g["Z"])
    close(np.trace(rho).real / am.partition(), 1.0)
    UrhoU, UrhoU_v, _ = orc.rotate_rho_probs(lambda x: orc.dm_rho(am, ph, x), basis, g["samples"], None, True)
    close(UrhoU_v, g["UrhoU_v"])
    close(UrhoU, g["UrhoU"])
    close(orc.dm_pi_grad(am, ph, space[:5], space[3:7], phase=False), g["pi_grad_am"])
    close(orc.dm_pi_grad(am, ph, space[:5], space[3:7], phase=True), g["pi_grad_ph"])
    ga, gp = orc.dm_rotated_gradient(am, ph, basis, g["samples"])
    close(ga, g["rot_grad_am"], 1e-11)
    close(gp, g["rot_grad_ph"], 1e-11)
    gm = orc.dm_gradient(am, ph, g["samples"], split_bases(g["bases_mixed"]))
    close(gm[0], g["grad_mixed_am"], 1e-11)
    close(gm[1], g["grad_mixed_ph"], 1e-11)


def test_density_2qubits_reference_fixture():
    g = load("density_2qubits.npz")
    am, ph = _dm_pair(g)
    bases = [str(b) for b in g["bases"]]
    close(orc.KL_dm(am, ph, g["target_rho"], bases), g["KL"], 1e-11)
    tb = split_bases(g["train_bases"])
    close(orc.NLL_dm(am, ph, g["train_samples"], tb), g["NLL"], 1e-11)
    close(orc.fidelity_dm(am, ph, g["target_rho"]), g["fidelity"], 1e-10)
    ge = orc.exact_gradient_with_bases("dm", am, ph, g["train_samples"], tb)
    close(ge[0], g["exact_grad_am"], 1e-10)
    close(ge[1], g["exact_grad_ph"], 1e-10)


def test_observables_reference_fixture():
    """SigmaX / SigmaY / SigmaZ / SWAP of the reference (observables/pauli.py, entanglement.py) and the Tutorial-4 TFIM energy on positive,
    complex and density-matrix states (tests/golden/make_golden_observables.py)."""
    g = load("observables.npz")
    t = load("binary_rbm_tfim10.npz")
    am = orc.BinaryRBMOracle(t["weights"], t["visible_bias"], t["hidden_bias"])
    w = orc.wavefunction_weight(lambda v: orc.psi_positive(am, v))
    v = g["pos_states"]
    close(orc.pauli_magnetisation("X", w, v), g["pos_sigma_x"])
    close(orc.pauli_magnetisation("X", w, v, absolute=True), g["pos_abs_sigma_x"])
    assert np.all(g["pos_sigma_y"] == 0) and np.all(orc.pauli_magnetisation("Y", w, v) == 0)
    close(orc.pauli_magnetisation("Z", w, v), g["pos_sigma_z"])
    close(orc.pauli_magnetisation("Z", w, v, absolute=True), g["pos_abs_sigma_z"])
    close(orc.tfim_chain_energy(am, v, 1.0), g["pos_tfim_energy_h1"])
    close(orc.tfim_chain_energy(am, v, 0.5), g["pos_tfim_energy_h05"])
    close(orc.swap_estimator(w, v, g["pos_swap_A"]), g["pos_swap"])

    c = load("complex_wavefunction.npz")
    cam, cph = _complex_pair(c)
    w = orc.wavefunction_weight(lambda v: orc.psi_complex(cam, cph, v))
    v = g["cwf_states"]
    close(orc.pauli_magnetisation("X", w, v), g["cwf_sigma_x"])
    close(orc.pauli_magnetisation("Y", w, v), g["cwf_sigma_y"])
    close(orc.pauli_magnetisation("Y", w, v, absolute=True), g["cwf_abs_sigma_y"])
    close(orc.pauli_magnetisation("Z", w, v), g["cwf_sigma_z"])
    close(orc.swap_estimator(w, v, g["cwf_swap_A"]), g["cwf_swap"])

    d = load("density_matrix.npz")
    w = orc.dm_weight(*_dm_pair(d))
    v = g["dm_states"]
    close(orc.pauli_magnetisation("X", w, v), g["dm_sigma_x"])
    close(orc.pauli_magnetisation("Y", w, v), g["dm_sigma_y"], 1e-11)
    close(orc.pauli_magnetisation("Z", w, v), g["dm_sigma_z"])
    close(orc.swap_estimator(w, v, g["dm_swap_A"]), g["dm_swap"], 1e-11)

#!/usr/bin/env python
"""Generates tests/golden/observables.npz by running the REFERENCE's observables (qucumber/observables/pauli.py
SigmaX / SigmaY / SigmaZ, qucumber/observables/entanglement.py SWAP and examples/Tutorial4_DataGeneration_CalculateObservables/quantum_ising_chain.py
TFIMChainEnergy, imported read-only from /root/reference) on the parameters already stored in
binary_rbm_tfim10.npz, complex_wavefunction.npz and density_matrix.npz. Run in the build container only:
    python tests/golden/make_golden_observables.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, import_reference  # noqa: E402


def main():
    import_reference()
    from qucumber.nn_states import ComplexWaveFunction, DensityMatrix, PositiveWaveFunction
    from qucumber.observables import SWAP, SigmaX, SigmaY, SigmaZ

    sys.path.insert(0, os.path.join(REF, "examples", "Tutorial4_DataGeneration_CalculateObservables"))
    from quantum_ising_chain import TFIMChainEnergy

    out = {}
    g = np.load(os.path.join(HERE, "binary_rbm_tfim10.npz"))
    nn = PositiveWaveFunction(10, 10, gpu=False)
    with torch.no_grad():
        nn.rbm_am.weights.copy_(torch.tensor(g["weights"]))
        nn.rbm_am.visible_bias.copy_(torch.tensor(g["visible_bias"]))
        nn.rbm_am.hidden_bias.copy_(torch.tensor(g["hidden_bias"]))
    states = torch.tensor(g["states"])
    out["pos_states"] = states.numpy()
    for name, obs in (("sigma_x", SigmaX()), ("sigma_y", SigmaY()), ("sigma_z", SigmaZ()), ("abs_sigma_x", SigmaX(absolute=True)),
                      ("abs_sigma_z", SigmaZ(absolute=True)), ("tfim_energy_h1", TFIMChainEnergy(1.0)),
                      ("tfim_energy_h05", TFIMChainEnergy(0.5))):
        out["pos_" + name] = obs.apply(nn, states.clone()).numpy()
    out["pos_swap_A"] = np.array([0, 1, 2, 3, 4])
    out["pos_swap"] = SWAP([0, 1, 2, 3, 4]).apply(nn, states.clone()).numpy()

    c = np.load(os.path.join(HERE, "complex_wavefunction.npz"))
    n, nh = c["am_weights"].shape[1], c["am_weights"].shape[0]
    cwf = ComplexWaveFunction(n, nh, gpu=False)
    with torch.no_grad():
        for tag, r in (("am", cwf.rbm_am), ("ph", cwf.rbm_ph)):
            r.weights.copy_(torch.tensor(c[tag + "_weights"]))
            r.visible_bias.copy_(torch.tensor(c[tag + "_visible_bias"]))
            r.hidden_bias.copy_(torch.tensor(c[tag + "_hidden_bias"]))
    space = cwf.generate_hilbert_space()
    out["cwf_states"] = space.numpy()
    for name, obs in (("sigma_x", SigmaX()), ("sigma_y", SigmaY()), ("sigma_z", SigmaZ()), ("abs_sigma_y", SigmaY(absolute=True))):
        out["cwf_" + name] = obs.apply(cwf, space.clone()).numpy()
    out["cwf_swap_A"] = np.array([1, 3])
    out["cwf_swap"] = SWAP([1, 3]).apply(cwf, space.clone()).numpy()
    d = np.load(os.path.join(HERE, "density_matrix.npz"))
    na, n = d["am_weights_U"].shape
    dm = DensityMatrix(n, d["am_weights_W"].shape[0], na, gpu=False)
    with torch.no_grad():
        for tag, r in (("am", dm.rbm_am), ("ph", dm.rbm_ph)):
            r.weights_W.copy_(torch.tensor(d[tag + "_weights_W"]))
            r.weights_U.copy_(torch.tensor(d[tag + "_weights_U"]))
            r.visible_bias.copy_(torch.tensor(d[tag + "_visible_bias"]))
            r.hidden_bias.copy_(torch.tensor(d[tag + "_hidden_bias"]))
            r.aux_bias.copy_(torch.tensor(d[tag + "_aux_bias"]))
    dspace = dm.generate_hilbert_space()
    out["dm_states"] = dspace.numpy()
    for name, obs in (("sigma_x", SigmaX()), ("sigma_y", SigmaY()), ("sigma_z", SigmaZ())):
        out["dm_" + name] = obs.apply(dm, dspace.clone()).numpy()
    out["dm_swap_A"] = np.array([0, 2])
    out["dm_swap"] = SWAP([0, 2]).apply(dm, dspace.clone()).numpy()
    np.savez_compressed(os.path.join(HERE, "observables.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/training_2qubits.npz: the inputs of the reference's own training regression
(tests/test_training.py:143-279, `slow` cases for ComplexWaveFunction and DensityMatrix) read through the REFERENCE's
loaders (utils/data.py: float32 parse, then float64), plus its fixed initial parameters for the complex case
(tests/data/test_training_init_complex_params.npz). The statistical targets of that test (fidelity / KL after 5 epochs)
are written next to them. Run in the build container only:  python tests/golden/make_golden_training.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, import_reference  # noqa: E402


def main():
    import_reference()
    import qucumber.utils.data as data

    ex = os.path.join(REF, "examples")
    t2 = os.path.join(ex, "Tutorial2_TrainComplexWaveFunction")
    samples, psi, tr_bases, bases = data.load_data(os.path.join(t2, "qubits_train.txt"), os.path.join(t2, "qubits_psi.txt"),
                                                   os.path.join(t2, "qubits_train_bases.txt"), os.path.join(t2, "qubits_bases.txt"))
    init = np.load(os.path.join(REF, "tests", "data", "test_training_init_complex_params.npz"))
    out = {
        "cwf_samples": samples.numpy().astype(np.uint8), "cwf_target_psi": psi.numpy(),
        "cwf_train_bases": np.array(["".join(b) for b in tr_bases]), "cwf_bases": np.array(["".join(b) for b in bases]),
        "cwf_fid_target": 0.38, "cwf_kl_target": 0.33, "cwf_pos_batch": 50, "cwf_neg_batch": 10,  # test_training.py:188-189
    }
    for k in init.files:
        out["cwf_init_" + k] = init[k]
    t3 = os.path.join(ex, "Tutorial3_TrainDensityMatrix")
    dsamples, rho, dtr_bases, dbases = data.load_data_DM(
        os.path.join(t3, "N2_W_state_100_samples_data.txt"), os.path.join(t3, "N2_W_state_target_real.txt"),
        os.path.join(t3, "N2_W_state_target_imag.txt"), os.path.join(t3, "N2_W_state_100_samples_bases.txt"),
        os.path.join(t3, "N2_IC_bases.txt"))
    out.update({
        "dm_samples": dsamples.numpy().astype(np.uint8), "dm_target_rho": rho.numpy(),
        "dm_train_bases": np.array(["".join(b) for b in dtr_bases]), "dm_bases": np.array(["".join(b) for b in dbases]),
        "dm_fid_target": 0.45, "dm_kl_target": 0.42, "dm_pos_batch": 100, "dm_neg_batch": 10,  # test_training.py:209-210
    })
    np.savez_compressed(os.path.join(HERE, "training_2qubits.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

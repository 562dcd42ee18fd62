#!/usr/bin/env python
"""Generates tests/golden/ref_checkpoint_*.pt: checkpoints written by the REFERENCE's NeuralState.save
(nn_states/neural_state.py:199-226, imported read-only from /root/reference) with the states' energies on a few rows as
metadata, to check that stock-QuCumber checkpoints load into the drop-in classes and that both write the same layout.
Run in the build container only:  python tests/golden/make_golden_checkpoints.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    import_reference()
    from qucumber.nn_states import ComplexWaveFunction, DensityMatrix, PositiveWaveFunction

    torch.manual_seed(2024)
    for tag, nn in (("positive", PositiveWaveFunction(6, 5, gpu=False)), ("complex", ComplexWaveFunction(5, 4, gpu=False)),
                    ("density", DensityMatrix(4, 3, 2, gpu=False))):
        for net in nn.networks:
            rbm = getattr(nn, net)
            with torch.no_grad():
                for prm in rbm.parameters():
                    prm.add_(0.2 * torch.randn_like(prm))
        rows = (torch.rand(9, nn.num_visible) < 0.5).double()
        meta = {"rows": rows, "energy_am": nn.rbm_am.effective_energy(rows), "epochs_done": 7}
        if hasattr(nn, "rbm_ph"):
            meta["energy_ph"] = nn.rbm_ph.effective_energy(rows)
        nn.save(os.path.join(HERE, f"ref_checkpoint_{tag}.pt"), metadata=meta)
        print(tag, sorted(torch.load(os.path.join(HERE, f"ref_checkpoint_{tag}.pt"), weights_only=False).keys()))


if __name__ == "__main__":
    main()

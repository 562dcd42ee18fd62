"""pytest plugin (``-p qcb_inject``) for running the REFERENCE's own test-suite (oracle/_ref/tests) with the
CUDA RBMs of qucumber_b200 injected at the seam INTEGRATION.md section 1 describes.

The reference's state classes build their RBMs by name (qucumber/nn_states/positive_wavefunction.py:46,
complex_wavefunction.py:66-67, density_matrix.py:65-66); this plugin rebinds exactly those names to
``qucumber_b200.BinaryRBM`` / ``PurificationRBM`` before the tests are collected. Everything above the seam —
NeuralStateBase.fit / gradient / sample, utils.unitaries, utils.training_statistics, observables, callbacks —
stays the unmodified reference code.

QCB_INJECT_FORCE_GPU=1 (default): most reference tests hard-code ``gpu=False``; the injected classes ignore
the flag and always live on the CUDA device (there is no CPU path to fall back to), so those tests exercise
the kernels as well.
"""
import os

import torch

import qucumber  # the vendored reference (PYTHONPATH is set by tests/test_reference_suite_gpu.py)
import qucumber.nn_states.complex_wavefunction as _cwf
import qucumber.nn_states.density_matrix as _dm
import qucumber.nn_states.positive_wavefunction as _pwf
import qucumber.rbm as _rbm
import qucumber_b200 as qb

FORCE = os.environ.get("QCB_INJECT_FORCE_GPU", "1") == "1"

assert "oracle/_ref" in os.path.abspath(qucumber.__file__).replace(os.sep, "/"), qucumber.__file__
assert torch.cuda.is_available(), "the injected RBMs have no CPU path"


class BinaryRBM(qb.BinaryRBM):
    def __init__(self, num_visible, num_hidden=None, zero_weights=False, gpu=True):
        super().__init__(num_visible, num_hidden, zero_weights=zero_weights, gpu=True if FORCE else gpu)


class PurificationRBM(qb.PurificationRBM):
    def __init__(self, num_visible, num_hidden=None, num_aux=None, zero_weights=False, gpu=False):
        super().__init__(num_visible, num_hidden, num_aux, zero_weights=zero_weights, gpu=True if FORCE else gpu)


_rbm.BinaryRBM = BinaryRBM
_rbm.PurificationRBM = PurificationRBM
_pwf.BinaryRBM = BinaryRBM
_cwf.BinaryRBM = BinaryRBM
_dm.PurificationRBM = PurificationRBM


def pytest_report_header(config):
    from qucumber_b200 import _lib

    return [f"qcb_inject: reference {qucumber.__file__} with CUDA RBMs from {_lib.LIB_PATH} (force_gpu={FORCE})"]


def pytest_sessionfinish(session, exitstatus):
    from qucumber_b200 import _lib

    n = _lib.load().qcb_launch_count()
    print(f"\nqcb_inject: {n} kernel launches from libqucumber_b200.so during this session")
    assert n > 0 or session.testscollected == 0, "no CUDA kernel of the drop-in was launched: the injection did not take"

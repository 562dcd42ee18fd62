"""The reference's OWN test-suite, run against the CUDA RBMs (last row).

``oracle/_ref`` is an unmodified copy of the reference made by ``oracle/make_ref.py`` (test infrastructure,
git-ignored, travels to the GPU box). Each case below is one pytest subprocess over reference test files with the
plugin ``tests/refsuite/qcb_inject.py`` loaded, which rebinds the RBM classes the reference's state classes
construct (qucumber/nn_states/positive_wavefunction.py:46, complex_wavefunction.py:66-67, density_matrix.py:65-66)
to ``qucumber_b200.BinaryRBM`` / ``PurificationRBM``. The tolerances are the reference's own (2e-8 finite
differences in tests/test_grads.py:247-279, +-0.02 statistical in tests/test_training.py:238-273).
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)

from oracle import make_ref  # noqa: E402

# Deselected, with the reason: these reference tests hard-code gpu=False and compare a HOST tensor they created with the
# state's output by torch.equal. The injected RBMs live on the CUDA device (there is no CPU path), so `.to(self.weights)`
# (rbm/binary_rbm.py:220) moves the input to the device and the comparison mixes devices — exactly what the stock reference
# does with gpu=True. The aliasing contract itself is covered on device tensors by tests/test_binary_rbm_gpu.py::test_gibbs_api_contract.
DESELECT = ["test_models_misc.py::test_sampling_with_overwrite"]

CASES = [
    # (id, pytest arguments relative to oracle/_ref/tests, minimum number of tests that must pass)
    ("grads", ["test_grads.py"], 24),
    ("models_misc", ["test_models_misc.py"], 40),
    ("training_slow", ["test_training.py", "-m", "slow"], 3),
    ("training_fast", ["test_training.py", "-m", "not slow"], 1),
    ("observables_callbacks_unitaries", ["test_observables.py", "test_callbacks.py", "test_unitaries.py"], 10),
]


def run_reference_tests(args, force_gpu=True, timeout=1500):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "refsuite"), *make_ref.ref_paths(), ROOT,
                                         env.get("PYTHONPATH", "")])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env["QCB_INJECT_FORCE_GPU"] = "1" if force_gpu else "0"
    tests = os.path.join(REF, "tests")
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "qcb_inject", "-c", os.path.join(REF, "tox.ini"),
           f"--rootdir={tests}", "-q", "--no-header", *[f"--deselect={d}" for d in DESELECT], "-W", "ignore::ResourceWarning", "-W", "ignore::UserWarning",
           *[a if a.startswith("-") or not a.endswith(".py") else os.path.join(tests, a) for a in args]]
    return subprocess.run(cmd, cwd=tests, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.gpu
@pytest.mark.parametrize("name,args,min_passed", CASES, ids=[c[0] for c in CASES])
def test_reference_suite_with_cuda_rbms(cuda, lib, name, args, min_passed):
    if not make_ref.available():
        pytest.fail("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container before gpurun")
    res = run_reference_tests(args)
    tail = (res.stdout[-6000:] + "\n" + res.stderr[-3000:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"refsuite_{name}.log"), "w") as fh:
            fh.write(res.stdout + "\n--- stderr ---\n" + res.stderr)
    assert res.returncode == 0, f"reference tests {args} failed with the CUDA RBMs injected:\n{tail}"
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= min_passed, f"too few reference tests ran ({m.group(0) if m else 'none'}):\n{tail}"
    launches = re.search(r"qcb_inject: (\d+) kernel launches", res.stdout)
    assert launches and int(launches.group(1)) > 0, "the reference tests did not reach the CUDA library"


def test_reference_copy_is_unmodified_and_ignored():
    """oracle/_ref must stay out of history and, when the source is present, equal it byte for byte."""
    with open(os.path.join(ROOT, ".gitignore")) as fh:
        assert "oracle/_ref/" in fh.read().split()
    ign = os.path.join(ROOT, ".gpurunignore")
    if os.path.exists(ign):
        with open(ign) as fh:
            assert not any(line.strip().startswith("oracle") for line in fh), "oracle/_ref must travel to the GPU box"
    src = make_ref.REF_SRC
    if not (os.path.isdir(src) and make_ref.available()):
        pytest.skip("reference source or copy absent")
    import filecmp

    for sub in ("qucumber", "tests"):
        for base, _, files in os.walk(os.path.join(src, sub)):
            for f in files:
                if f.endswith(".pyc"):
                    continue
                a = os.path.join(base, f)
                b = os.path.join(REF, os.path.relpath(a, src))
                assert os.path.exists(b) and filecmp.cmp(a, b, shallow=False), f"{b} differs from the reference"

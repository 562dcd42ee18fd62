"""TEST INFRASTRUCTURE — builds ``oracle/_ref/``: an unmodified copy of the reference (PIQuIL/QuCumber,
pure Python) so that it can travel to the GPU box, where ``/root/reference`` does not exist.

  python oracle/make_ref.py            # copy /root/reference/{qucumber,tests,examples} -> oracle/_ref/
                                       # + a 3-file matplotlib stub (matplotlib is not installable here)

``oracle/_ref/`` is git-ignored (never committed: no reference source enters the history) but NOT
gpurun-ignored. Nothing in the product (``qucumber_b200/``) imports it; users are

* ``tests/test_reference_suite_gpu.py`` — runs the reference's own test-suite with the CUDA RBMs injected,
* ``bench.py --impl reference`` — times the reference's own ``BinaryRBM.gibbs_steps`` on the host cores
  (``cpu_baseline.kind = "reference"``) and, on the GPU box, its ``gpu=True`` torch-CUDA op chain
  (``torch_cuda_baseline``),
* ``tests/test_oracle_golden.py`` — cross-checks the numpy restatement against the live reference.

The stub exists because ``qucumber/callbacks/liveplotting.py:20-21`` imports matplotlib at package import;
only LivePlotting would ever call it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("QCB_REFERENCE_SRC", "/root/reference")
DEST = os.path.join(HERE, "_ref")

STUB = {
    "matplotlib/__init__.py": '"""Stub: matplotlib is absent in this image; only LivePlotting needs the real one."""\n',
    "matplotlib/pyplot.py": (
        'def subplots(*args, **kwargs):\n'
        '    raise RuntimeError("matplotlib stub: plotting is not available in this environment")\n'
    ),
    "matplotlib/ticker.py": (
        'class MultipleLocator:\n'
        '    def __init__(self, *args, **kwargs):\n'
        '        pass\n'
    ),
}


def have_matplotlib():
    try:
        import importlib.util

        return importlib.util.find_spec("matplotlib") is not None
    except Exception:
        return False


def ref_paths():
    """sys.path entries that make ``import qucumber`` resolve to the vendored reference."""
    paths = [DEST]
    if not have_matplotlib():
        paths.insert(0, os.path.join(DEST, "_stub"))
    return paths


def available():
    return os.path.isfile(os.path.join(DEST, "qucumber", "__init__.py"))


def _make_writable(root):
    for base, dirs, files in os.walk(root):
        for name in dirs + files:
            path = os.path.join(base, name)
            os.chmod(path, os.stat(path).st_mode | 0o200)
    os.chmod(root, os.stat(root).st_mode | 0o200)


def make(force=False):
    """Copy the reference. Returns DEST, or None when neither the source nor a previous copy exists."""
    if not os.path.isdir(os.path.join(REF_SRC, "qucumber")):
        return DEST if available() else None
    stamp = os.path.join(DEST, ".stamp")
    if available() and os.path.exists(stamp) and not force:
        return DEST
    if os.path.isdir(DEST):
        _make_writable(DEST)
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb", ".pytest_cache")
    for sub in ("qucumber", "tests", "examples"):
        shutil.copytree(os.path.join(REF_SRC, sub), os.path.join(DEST, sub), ignore=ignore)
    shutil.copy(os.path.join(REF_SRC, "tox.ini"), os.path.join(DEST, "tox.ini"))
    shutil.copy(os.path.join(REF_SRC, "LICENSE"), os.path.join(DEST, "LICENSE"))
    for rel, text in STUB.items():
        path = os.path.join(DEST, "_stub", rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    _make_writable(DEST)  # the source tree is read-only; the copy must be removable and re-creatable
    with open(stamp, "w") as fh:
        fh.write("copied from %s by oracle/make_ref.py; unmodified\n" % REF_SRC)
    return DEST


def import_reference():
    """Put the vendored reference on sys.path (ahead of anything else called qucumber) and import it."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    for p in reversed(ref_paths()):
        if p not in sys.path:
            sys.path.insert(0, p)
    import qucumber  # noqa: F401

    if not os.path.abspath(qucumber.__file__).startswith(DEST):
        raise RuntimeError(f"`import qucumber` resolved to {qucumber.__file__}, not to oracle/_ref")
    return qucumber


if __name__ == "__main__":
    out = make(force="--force" in sys.argv)
    print(out if out else "no reference source and no previous copy: oracle/_ref not built")

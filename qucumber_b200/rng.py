"""Seeding for the in-kernel Philox streams.

Every sampling call takes a fresh 64-bit Philox key from torch's default CPU generator, so
``torch.manual_seed`` / ``qucumber_b200.set_random_seed`` make a whole run reproducible (the
reference's contract, qucumber/__init__.py:22-37, tests/test_models_misc.py:44-76) while
successive calls still draw independent numbers. Host-only; nothing here touches the device.
"""
import torch


def set_random_seed(seed, cpu=True, gpu=False, quiet=False):
    """Mirror of qucumber.set_random_seed: seeds torch's generators (the Philox keys derive from the CPU one)."""
    if gpu and torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if cpu or not gpu:
        torch.manual_seed(seed)


def next_seed():
    """A fresh 62-bit key drawn from torch's default CPU generator."""
    return int(torch.randint(0, 1 << 62, (1,), dtype=torch.int64).item())

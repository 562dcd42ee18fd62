"""Text-file loaders with the reference's conventions (qucumber/utils/data.py:22-133): numbers are
parsed as float32 and then widened to float64, bases are arrays of one-character strings."""
import numpy as np
import torch

from . import cplx


def _f32_as_double(path):
    return torch.tensor(np.loadtxt(path, dtype="float32"), dtype=torch.double)


def load_data(tr_samples_path, tr_psi_path=None, tr_bases_path=None, bases_path=None):
    """[samples, (target psi as (2, D)), (per-sample bases), (list of bases)] — entries appear only
    for the paths that were given."""
    out = [_f32_as_double(tr_samples_path)]
    if tr_psi_path is not None:
        psi = np.loadtxt(tr_psi_path, dtype="float32")
        target = torch.zeros(2, len(psi), dtype=torch.double)
        target[0] = torch.tensor(psi[:, 0], dtype=torch.double)
        target[1] = torch.tensor(psi[:, 1], dtype=torch.double)
        out.append(target)
    if tr_bases_path is not None:
        out.append(np.loadtxt(tr_bases_path, dtype=str))
    if bases_path is not None:
        out.append(np.loadtxt(bases_path, dtype=str, ndmin=1))
    return out


def load_data_DM(tr_samples_path, tr_mtx_real_path=None, tr_mtx_imag_path=None, tr_bases_path=None, bases_path=None):
    """Same for density matrices: the target is given as separate real and imaginary text files."""
    out = [_f32_as_double(tr_samples_path)]
    if (tr_mtx_real_path is None) != (tr_mtx_imag_path is None):
        raise ValueError("Must provide a real and imaginary part of target matrix!")
    if tr_mtx_real_path is not None:
        out.append(cplx.make_complex(_f32_as_double(tr_mtx_real_path), _f32_as_double(tr_mtx_imag_path)))
    if tr_bases_path is not None:
        out.append(np.loadtxt(tr_bases_path, dtype=str))
    if bases_path is not None:
        out.append(np.loadtxt(bases_path, dtype=str, ndmin=1))
    return out


def extract_refbasis_samples(train_samples, train_bases):
    """Rows of ``train_samples`` whose measurement basis is all-Z."""
    mask = torch.as_tensor(np.all(np.asarray(train_bases) == "Z", axis=1), device=train_samples.device)
    return train_samples[mask]

"""Local estimators on top of the CUDA RBMs (reference: qucumber/observables/pauli.py,
observable.py:223-249). QuCumber's own observable classes keep working with the drop-in states (they
call ``nn_state.sample`` / ``psi`` / ``importance_sampling_*``); the classes here compute the same
estimators from ONE fused kernel call per RBM: qcb_rbm_flip_logratio gives
log p(v^(i)) - log p(v) for all N single-spin flips of every sample, instead of N full energy
evaluations of flipped copies of the batch."""
import numpy as np
import torch

from . import _lib


def flip_logratio(rbm, samples):
    """(rows, N) tensor of E(v) - E(v with unit i flipped) for a BinaryRBM / PurificationRBM."""
    lib = _lib.load()
    args, keep = rbm._param_args()
    like = keep[0]
    rows = samples.to(like).reshape(-1, rbm.num_visible).contiguous()
    out = torch.empty_like(rows)
    _lib.check(lib.qcb_rbm_flip_logratio(*args, _lib.ptr(rows), rows.shape[0], _lib.ptr(out), _lib.stream_ptr(rows.device)),
               "qcb_rbm_flip_logratio")
    return out


def _is_wavefunction(nn_state):
    return hasattr(nn_state, "phase")


def _importance_sampled_pauli(nn_state, samples, axis, max_rows=1 << 22):
    """(1/N) sum_i Re[c_i rho(v^(i), v) / rho(v, v)] (pauli.py:62-83, 113-135) for states without a psi(v) factorisation
    (DensityMatrix). The reference loops over the sites; here all N single-flip copies of a block of samples form ONE batch
    (N, rows, N), so each of the drop-in's CUDA paths inside rho() runs once per block instead of once per site."""
    from .utils import cplx

    n = samples.shape[-1]
    flat = samples.reshape(-1, n)
    out = torch.empty(flat.shape[0], dtype=torch.double, device=flat.device)
    flips = torch.eye(n, dtype=torch.double, device=flat.device).unsqueeze(1)  # (N, 1, N): copy i flips unit i
    block = max(1, max_rows // n)
    for first in range(0, flat.shape[0], block):  # bounded scratch: n * block rows at a time
        v = flat[first:first + block]
        denom = nn_state.importance_sampling_denominator(v)
        tiled = v.unsqueeze(0).expand(n, -1, -1)
        flipped = (tiled - flips).abs_()  # |v - e_i|: unit i flipped, the others unchanged
        numer = nn_state.importance_sampling_numerator(flipped.reshape(-1, n), tiled.reshape(-1, n))
        numer = numer.reshape(2, n, v.shape[0])
        if axis == "Y":  # coefficient i (2 v_i - 1) per site
            coeff = v.t().mul(2.0).sub(1.0)
            numer = cplx.make_complex(-cplx.imag(numer) * coeff, cplx.real(numer) * coeff)
        out[first:first + block] = cplx.real(cplx.elementwise_division(numer.sum(1), denom)).div_(n)
    return out.reshape(samples.shape[:-1])


class _Estimator:
    def statistics_from_samples(self, nn_state, samples):
        """mean / variance / std_error / num_samples of the estimator over the given samples."""
        obs = self.apply(nn_state, samples)
        variance, mean = torch.stack(torch.var_mean(obs)).tolist()  # one device-to-host read
        return {"mean": mean, "variance": variance, "std_error": float(np.sqrt(variance / len(obs))), "num_samples": len(obs)}

    def statistics(self, nn_state, num_samples, num_chains=0, burn_in=1000, steps=1):
        """Draw samples with block Gibbs (burn-in, then `steps` sweeps between recorded states, chains kept
        on the device) and accumulate the statistics batch by batch like observable.py:135-221."""
        num_chains = num_chains if num_chains > 0 else num_samples
        chains = nn_state.sample(k=burn_in, num_samples=min(num_chains, num_samples))
        total, s1, s2 = 0, 0.0, 0.0
        while total < num_samples:
            if total > 0:
                chains = nn_state.sample(k=steps, initial_state=chains, overwrite=True)
            take = min(chains.shape[0], num_samples - total)
            obs = self.apply(nn_state, chains[:take])
            s1 += float(obs.sum())
            s2 += float((obs * obs).sum())
            total += take
        mean = s1 / total
        variance = max(s2 - total * mean * mean, 0.0) / max(total - 1, 1)
        return {"mean": mean, "variance": variance, "std_error": float(np.sqrt(variance / total)), "num_samples": total}


class SigmaZ(_Estimator):
    """Average magnetisation along Z (pauli.py:138-171): mean of the +-1 spins of each sample."""

    def __init__(self, absolute=False):
        self.name, self.symbol, self.absolute = "SigmaZ", "Z", absolute

    def apply(self, nn_state, samples):
        res = samples.to(device=nn_state.device, dtype=torch.double).mean(1).mul_(2.0).sub_(1.0)  # mean of the +-1 spins
        return res.abs_() if self.absolute else res


class SigmaX(_Estimator):
    """Average magnetisation along X (pauli.py:36-83): (1/N) sum_i Re[psi(v^(i)) / psi(v)]."""

    def __init__(self, absolute=False):
        self.name, self.symbol, self.absolute = "SigmaX", "X", absolute

    def apply(self, nn_state, samples):
        samples = samples.to(device=nn_state.device, dtype=torch.double)
        if _is_wavefunction(nn_state):
            ratio = flip_logratio(nn_state.rbm_am, samples).mul_(0.5).exp_()  # fresh tensor: in place
            if getattr(nn_state, "rbm_ph", None) is not None:
                ratio.mul_(flip_logratio(nn_state.rbm_ph, samples).mul_(0.5).cos_())
            res = ratio.mean(1)
        else:
            res = _importance_sampled_pauli(nn_state, samples, "X")
        return res.abs() if self.absolute else res


class SigmaY(_Estimator):
    """Average magnetisation along Y (pauli.py:86-135): (1/N) sum_i Re[i s_i psi(v^(i)) / psi(v)], s_i = 2 v_i - 1;
    identically zero for a positive wavefunction."""

    def __init__(self, absolute=False):
        self.name, self.symbol, self.absolute = "SigmaY", "Y", absolute

    def apply(self, nn_state, samples):
        samples = samples.to(device=nn_state.device, dtype=torch.double)
        if not _is_wavefunction(nn_state):
            res = _importance_sampled_pauli(nn_state, samples, "Y")
        elif getattr(nn_state, "rbm_ph", None) is None:
            return torch.zeros(samples.shape[:-1], dtype=torch.double, device=samples.device)
        else:
            ratio = (0.5 * flip_logratio(nn_state.rbm_am, samples)).exp()
            res = -(samples.mul(2.0).sub(1.0) * ratio * (0.5 * flip_logratio(nn_state.rbm_ph, samples)).sin()).mean(1)
        return res.abs() if self.absolute else res


class TFIMChainEnergy(_Estimator):
    """Energy per site of the 1-D transverse-field Ising chain, -J sum s_i s_{i+1} - h sum sigma^x_i
    (examples/Tutorial4_DataGeneration_CalculateObservables/quantum_ising_chain.py:24-78), open
    boundaries unless ``periodic_bcs``; positive wavefunctions."""

    def __init__(self, h, periodic_bcs=False):
        self.name, self.symbol, self.h, self.periodic_bcs = "TFIMChainEnergy", "E", h, periodic_bcs

    def apply(self, nn_state, samples):
        samples = samples.to(device=nn_state.device, dtype=torch.double)
        spins = samples.mul(2.0).sub(1.0)
        interaction = (spins[:, :-1] * spins[:, 1:]).sum(1)
        if self.periodic_bcs:
            interaction = interaction + spins[:, 0] * spins[:, -1]
        transverse = flip_logratio(nn_state.rbm_am, samples).mul_(0.5).exp_().sum(1)
        return -(interaction + self.h * transverse) / samples.shape[-1]


class SWAP(_Estimator):
    """Swap_A between two replicas (observables/entanglement.py:40-80): S_2 = -ln <SWAP_A>. The second replica is the
    batch rolled by one row, like the reference; the importance weights psi(s') / psi(s) (or rho(s', s) / rho(s, s))
    come from the drop-in state's CUDA energy / density-matrix paths."""

    def __init__(self, A):
        self.name, self.symbol, self.A = "SWAP", "S", A

    def apply(self, nn_state, samples):
        from .utils import cplx

        s1 = samples.to(device=nn_state.device, dtype=torch.double)
        s2 = torch.roll(s1, 1, 0)
        s1_, s2_ = s1.clone(), s2.clone()
        s1_[:, self.A] = s2[:, self.A]
        s2_[:, self.A] = s1[:, self.A]
        w1 = nn_state.importance_sampling_weight(s1_, s1)
        w2 = nn_state.importance_sampling_weight(s2_, s2)
        return cplx.real(cplx.elementwise_mult(w1, w2))

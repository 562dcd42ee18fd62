"""Complex numbers in QuCumber's storage convention: a real tensor whose leading dimension of size
2 holds [real, imaginary] (reference: qucumber/utils/cplx.py). These helpers exist for API
compatibility on the host side of the drop-in; they convert to torch's native complex dtype, do
the operation, and convert back. None of them is on the CUDA hot path — rotated amplitudes and
gradients are computed by fused kernels (qcb_cwf_rotated / qcb_dm_rotated)."""
import numpy as np
import torch

I = torch.tensor([0.0, 1.0], dtype=torch.double)  # noqa: E741  the imaginary unit as a complex scalar


def _c(x):
    return torch.complex(x[0].contiguous(), x[1].contiguous())


def _r(z):
    return torch.stack([z.real, z.imag], dim=0)


def make_complex(x, y=None):
    """[x, y] = x + iy; a complex numpy array is split into its parts."""
    if isinstance(x, np.ndarray):
        return torch.stack([torch.tensor(x.real), torch.tensor(x.imag)], dim=0).contiguous()
    if y is None:
        y = torch.zeros_like(x)
    return torch.stack([x, y], dim=0)


def numpy(x):
    return x[0].detach().cpu().numpy() + 1j * x[1].detach().cpu().numpy()


def real(x):
    return x[0, ...]


def imag(x):
    return x[1, ...]


def scalar_mult(x, y, out=None):
    """Elementwise product of complex tensors (either may be a complex scalar of shape (2,))."""
    y = y.to(x)
    xr, xi, yr, yi = x[0], x[1], y[0], y[1]
    res = torch.stack([xr * yr - xi * yi, xr * yi + xi * yr], dim=0)
    if out is not None:
        out.copy_(res)
        return out
    return res


def elementwise_mult(x, y):
    return scalar_mult(x, y)


def matmul(x, y):
    return _r(torch.matmul(_c(x), _c(y.to(x))))


def inner_prod(x, y):
    """<x|y> = sum conj(x) y for complex vectors."""
    y = y.to(x)
    if x.dim() == 2 and y.dim() == 2:
        return _r(torch.sum(torch.conj(_c(x)) * _c(y)))
    if x.dim() == 1 and y.dim() == 1:
        return scalar_mult(conjugate_scalar(x), y)
    raise ValueError("Unsupported input shapes!")


def conjugate_scalar(x):
    return torch.stack([x[0], -x[1]], dim=0)


def outer_prod(x, y):
    if x.dim() != 2 or y.dim() != 2:
        raise ValueError("An input is not of the right dimension.")
    return _r(torch.outer(_c(x), _c(y.to(x))))


def einsum(equation, a, b, real_part=True, imag_part=True):
    """Complex einsum on the trailing dimensions (the leading [re, im] axis is implicit)."""
    if not (real_part or imag_part):
        raise ValueError("at least one of real_part / imag_part must be requested")
    z = torch.einsum(equation, _c(a), _c(b.to(a)))
    if real_part and imag_part:
        return _r(z)
    return z.real if real_part else z.imag


def conjugate(x):
    """Hermitian conjugate: transpose (for matrices) and negate the imaginary part."""
    if x.dim() < 3:
        return conj(x)
    return torch.stack([x[0].transpose(0, 1), -x[1].transpose(0, 1)], dim=0)


def conj(x):
    return torch.stack([x[0], -x[1]], dim=0)


def elementwise_division(x, y):
    if x.shape != y.shape:
        raise ValueError("x and y must have the same shape!")
    return _r(_c(x) / _c(y.to(x)))


def absolute_value(x):
    return torch.sqrt(x[0] ** 2 + x[1] ** 2)


def kronecker_prod(x, y):
    if not (x.dim() == y.dim() == 3):
        raise ValueError("Inputs must be complex matrices!")
    return _r(torch.kron(_c(x), _c(y.to(x))))


def sigmoid(x, y):
    """1 / (1 + exp(-(x + iy)))."""
    z = torch.complex(x, y)
    return _r(1.0 / (1.0 + torch.exp(-z)))


def scalar_divide(x, y):
    return _r(_c(x) / _c(y.to(x)))


def inverse(z):
    """Elementwise 1/z."""
    return conj(z) / (z[0] ** 2 + z[1] ** 2)


def norm_sqr(x):
    """<x|x> (real)."""
    return real(inner_prod(x, x))


def norm(x):
    return norm_sqr(x).sqrt()

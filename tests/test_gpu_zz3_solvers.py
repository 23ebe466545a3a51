"""faer_b200.solvers on the GPU through the C ABI: the reference's `test_all_solvers` identities (solvers.rs:2919-2977)
and the accessor contracts, shared with the CPU host-logic run (tests/solvers_cases.py), then the same decompositions on
device-resident tensors."""
import numpy as np
import pytest

from solvers_cases import approx, run_all

pytestmark = pytest.mark.gpu


def test_all_solvers_host_arrays(fb, cuda_dev):
    run_all(fb.solvers)


def test_solvers_on_device_tensors(fb, cuda_dev):
    import torch
    sv = fb.solvers
    rng = np.random.default_rng(9)
    n, k = 300, 5
    A = rng.standard_normal((n, n)); B = rng.standard_normal((n, k))
    cond = np.linalg.cond(A)
    dA = torch.from_numpy(A).to(cuda_dev)            # row-major on the device: any strides are accepted
    dB = torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda_dev).t()
    for dec in (sv.partial_piv_lu(dA), sv.qr(dA)):
        X = dec.solve(dB)
        assert X.is_cuda and X.stride(0) == 1       # results are column-major device tensors
        assert approx(A @ X.cpu().numpy(), B, n, cond)
        Xt = dec.solve_transpose(dB)
        assert approx(A.T @ Xt.cpu().numpy(), B, n, cond)
        assert approx(dec.reconstruct().cpu().numpy(), A, n, np.abs(A).max() * n)
    S = A @ A.T
    dS = torch.from_numpy(S).to(cuda_dev)
    llt = sv.llt(dS)
    Lh = llt.L().cpu().numpy()
    assert np.all(np.triu(Lh, 1) == 0) and approx(Lh @ Lh.T, S, n, np.abs(S).max())
    assert approx(S @ llt.solve(dB).cpu().numpy(), B, n, np.linalg.cond(S))
    tall = rng.standard_normal((1000, 60)); rhs = rng.standard_normal((1000, 3))
    d = sv.qr(torch.from_numpy(tall).to(cuda_dev))
    x = d.solve_lstsq(torch.from_numpy(rhs).to(cuda_dev))
    assert tuple(x.shape) == (60, 3)
    assert approx(x.cpu().numpy(), np.linalg.lstsq(tall, rhs, rcond=None)[0], 1000, np.linalg.cond(tall))
    assert torch.equal(dA.cpu(), torch.from_numpy(A))  # inputs are never modified

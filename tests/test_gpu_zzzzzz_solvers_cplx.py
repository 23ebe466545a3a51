"""faer_b200.solvers on c64 through the C ABI: the reference's `test_all_solvers` (solvers.rs:2919-2977) runs on c64 — the eight
solve / rsolve identities with conjugate / adjoint meaning what they say — for PartialPivLu, Qr and Llt, plus the accessor
contracts (shared cases: tests/solvers_cases.py, `cplx=True`; the same cases run on the CPU against the oracle-backed stand-in),
then on device-resident complex tensors."""
import numpy as np
import pytest

from solvers_cases import approx, run_all

pytestmark = pytest.mark.gpu


def test_all_solvers_c64_host_arrays(fb, cuda_dev):
    run_all(fb.solvers, cplx=True)


def test_solvers_c64_on_device_tensors(fb, cuda_dev):
    import torch
    sv = fb.solvers
    rng = np.random.default_rng(19)
    n, k = 200, 4
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    cond = np.linalg.cond(A)
    dA = torch.from_numpy(A).to(cuda_dev)
    dB = torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda_dev).t()
    for dec in (sv.partial_piv_lu(dA), sv.qr(dA)):
        assert approx(A @ dec.solve(dB).cpu().numpy(), B, n, cond)
        assert approx(A.conj() @ dec.solve_conjugate(dB).cpu().numpy(), B, n, cond)
        assert approx(A.T @ dec.solve_transpose(dB).cpu().numpy(), B, n, cond)
        assert approx(A.conj().T @ dec.solve_adjoint(dB).cpu().numpy(), B, n, cond)
        assert approx(dec.reconstruct().cpu().numpy(), A, n, np.abs(A).max() * n)
    S = A @ A.conj().T
    llt = sv.llt(torch.from_numpy(S).to(cuda_dev))
    Lh = llt.L().cpu().numpy()
    assert np.all(np.triu(Lh, 1) == 0) and approx(Lh @ Lh.conj().T, S, n, np.abs(S).max())
    assert approx(S @ llt.solve(dB).cpu().numpy(), B, n, np.linalg.cond(S))
    assert approx(S.T @ llt.solve_transpose(dB).cpu().numpy(), B, n, np.linalg.cond(S))
    assert approx(llt.reconstruct().cpu().numpy(), S, n, np.abs(S).max())
    assert torch.equal(dA.cpu(), torch.from_numpy(A))  # inputs are never modified

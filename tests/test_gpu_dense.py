"""The hand-written dense inverse (csrc/hank_dense.cu) against LAPACK through NumPy, and the Newton solve with it
against the cuSOLVER A/B path.  Tolerances: the inverse is compared through its residual, scaled by the condition
number (both LAPACK's and the Gauss-Jordan inverse are backward-stable to ~n*eps*cond)."""
import os

import numpy as np
import pytest

from common import model_inputs, make_block

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def blk():
    b = make_block(model_inputs(50, 3), 4)
    yield b
    b.close()


@pytest.mark.parametrize("n", [1, 2, 5, 16, 17, 31, 32, 100, 299, 1196, 1700, 2500, 3301])   # block widths 16 (n <= ~1750), 8, 4
def test_inverse_matches_lapack(blk, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    if n > 2:
        A[0, 0] = 0.0            # forces a row interchange in the first column
        A[n // 2] *= 1e-6        # badly scaled row: pivoting has to avoid it
    X = blk.dense_inverse(A)
    Xl = np.linalg.inv(A)
    cond = np.linalg.cond(A)
    res = np.abs(X @ A - np.eye(n)).max()
    assert res <= 64 * n * np.finfo(float).eps * cond, (n, res, cond)
    assert np.abs(X - Xl).max() <= 64 * n * np.finfo(float).eps * cond * np.abs(Xl).max(), (n, cond)


def test_inverse_needs_pivoting(blk):
    """A permutation-like matrix: every pivot comes from a row interchange, several of them out of the block."""
    n = 97
    rng = np.random.default_rng(0)
    perm = rng.permutation(n)
    A = np.zeros((n, n)); A[np.arange(n), perm] = rng.uniform(1.0, 2.0, n)
    A += 1e-3 * rng.standard_normal((n, n))
    X = blk.dense_inverse(A)
    assert np.abs(X @ A - np.eye(n)).max() < 1e-10


def test_inverse_reports_singular(blk):
    from hankb200 import HankError
    A = np.eye(40); A[:, 7] = 0.0
    with pytest.raises(HankError) as ei:
        blk.dense_inverse(A)
    assert "singular" in ei.value.msg


def test_inverse_of_golden_jbar(blk):
    """The real preconditioner: J̅ of the 100x3, T=30 golden case (cond ~1e4) and the 500x7, T=300 one if present."""
    g = np.load(os.path.join(GOLD, "ks_100x3_T30.npz"))
    J = np.array(g["Jbar"])
    X = blk.dense_inverse(J)
    Xl = np.linalg.inv(J)
    assert np.abs(X - Xl).max() <= 1e-10 * np.abs(Xl).max()

"""Oracle (TEST INFRASTRUCTURE): the reference's numeric helpers, restated with plain numpy.

Follows reference ``num_utils.py``: BLAS-1 wrappers :20-42, ``ssyrk``/``ssymm`` :53-66,
``norm2`` :69-71, ``p_norm`` :74-82, ``normalize`` :85-87, ``roll2`` :136-140,
``gram_matrix`` :143-147, ``tv_norm`` :150-162.  Everything is float32 like the reference; sums
are taken by numpy's float32 pairwise reductions instead of BLAS sdot/sasum (differences are at
float32 round-off and are covered by the tolerances in tests/test_oracle_golden.py, which pins
these functions against outputs of the reference's own module).
"""

import numpy as np

EPS = np.finfo(np.float32).eps        # num_utils.py:14


def sdot(x, y):
    return np.float32(np.dot(x.ravel(), y.ravel()))


def sasum(x):
    return np.float32(np.abs(x).sum(dtype=np.float32))


def norm2(arr):
    """Half the squared L2 norm (num_utils.py:69-71)."""
    return sdot(arr, arr) / 2


def normalize_(arr):
    """In place: scale so that mean|arr| == 1 (num_utils.py:85-87)."""
    arr *= np.float32(1 / (sasum(arr) / arr.size + EPS))
    return arr


def gram_lower(feat):
    """tril(F F^T) / F.size for F = feat reshaped [C, H*W]; the strict upper triangle is exactly 0
    because the reference calls SSYRK, which writes one triangle only (num_utils.py:53-56,143-147).
    """
    f = feat.reshape(feat.shape[0], -1)
    return np.tril(f @ f.T).astype(np.float32) * np.float32(1 / f.size)


def symm_times(lower, mat):
    """sym(lower) @ mat where only the lower triangle of ``lower`` is looked at (SSYMM,
    num_utils.py:60-66)."""
    full = np.tril(lower) + np.tril(lower, -1).T
    return (full @ mat).astype(np.float32)


def p_norm(arr, p=2):
    """sum |arr|^p and its gradient (num_utils.py:74-82)."""
    if p == 1:
        return sasum(arr), np.sign(arr)
    if p == 2:
        return sdot(arr, arr), 2 * arr
    mag = np.abs(arr)
    mag_p1 = mag ** (p - 1)
    return sdot(mag_p1, mag), p * np.sign(arr) * mag_p1


def roll2_(arr, xy):
    """In place circular shift: xy[0] along the LAST axis, xy[1] along the second-to-last
    (num_utils.py:136-140 -- ``np.roll(arr, xy, axis=(-1, -2))``)."""
    if xy is not None and np.any(np.asarray(xy) != 0):
        arr[...] = np.roll(arr, (int(xy[0]), int(xy[1])), axis=(-1, -2))
    return arr


def tv_norm(x, beta=2):
    """Periodic total-variation norm sum((dx^2+dy^2+EPS)^(beta/2)) and gradient
    (num_utils.py:150-162); forward differences x[i]-x[i+1] with wrap-around."""
    dx = x - np.roll(x, -1, axis=2)
    dy = x - np.roll(x, -1, axis=1)
    g2 = dx ** 2 + dy ** 2 + EPS
    loss = np.sum(g2 ** (beta / 2))
    dg = (beta / 2) * g2 ** (beta / 2 - 1)
    ddx = 2 * dx * dg
    ddy = 2 * dy * dg
    grad = ddx + ddy - np.roll(ddx, 1, axis=2) - np.roll(ddy, 1, axis=1)
    return loss, grad

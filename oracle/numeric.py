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


# ---------------------------------------------------------------------------------------------
# resize (num_utils.py:90-108): per-channel float resampling through PIL 'F' images.  Restated from
# Pillow's published algorithm (libImaging/Resample.c: precompute_coeffs + the 32-bit-per-channel
# horizontal / vertical passes) so that the device kernel has something other than PIL itself to be
# checked against; pinned bit for bit against PIL and against the reference's own num_utils.resize
# (tests/golden/resize.npz, tests/test_oracle_golden.py).
# ---------------------------------------------------------------------------------------------
def _resample_filter(kind):
    import math

    def sinc(t):
        if t == 0.0:
            return 1.0
        t = t * math.pi
        return math.sin(t) / t

    def lanczos(x):
        return sinc(x) * sinc(x / 3) if -3.0 <= x < 3.0 else 0.0

    def bilinear(x):
        x = abs(x)
        return 1.0 - x if x < 1.0 else 0.0
    return {'lanczos': (lanczos, 3.0), 'bilinear': (bilinear, 1.0)}[kind]


def resample_coeffs(in_size, out_size, kind='lanczos'):
    """Pillow's precompute_coeffs for the full source range: returns (bounds int32[out, 2] =
    (first source index, count), weights float64[out, ksize]), weights normalised per output."""
    import math
    filt, support0 = _resample_filter(kind)
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = filt((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_rows(a, out_w, kind):
    """One pass along the last axis: float64 accumulation in source order, float32 store."""
    bounds, kk = resample_coeffs(a.shape[-1], out_w, kind)
    out = np.zeros(a.shape[:-1] + (out_w,), np.float32)
    for xx, (xmin, count) in enumerate(bounds):
        acc = np.zeros(a.shape[:-1], np.float64)
        for x in range(count):
            acc += a[..., x + xmin].astype(np.float64) * kk[xx, x]
        out[..., xx] = acc
    return out


def resize(a, hw, method='lanczos'):
    """Resamples [C,H,W] (or [H,W]) float32 to hw like ``num_utils.resize``: horizontal pass first
    (float32 intermediate), then vertical; a pass whose size does not change is skipped."""
    a = np.float32(a)
    h, w = hw
    t = _resample_rows(a, w, method) if w != a.shape[-1] else a
    if h != a.shape[-2]:
        t = np.swapaxes(_resample_rows(np.ascontiguousarray(np.swapaxes(t, -1, -2)), h, method), -1, -2)
    return np.ascontiguousarray(t, dtype=np.float32)


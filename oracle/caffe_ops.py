"""Oracle (TEST INFRASTRUCTURE): the Caffe CPU layer arithmetic the reference's hot path runs.

Restated from BVLC Caffe's published CPU algorithms (external, unpinned -- see oracle/__init__.py):
``ConvolutionLayer`` = im2col + SGEMM, ``ReLULayer`` in place, ``PoolingLayer`` MAX/AVE in ceil
mode.  The reference reaches them through ``net.forward`` / ``net.backward``
(style_transfer.py:425, 566, 608-610) on the six bundled VGG prototxts, which use only
3x3/pad-1/stride-1 convolutions, in-place ReLU and 2x2/stride-2 pooling
(vgg19.prototxt:16-60).  All arrays are float32, single image, CHW.
"""

import numpy as np
from scipy.linalg import blas

FLT_MAX = np.finfo(np.float32).max


# --------------------------------------------------------------------------------------------
# Convolution 3x3, pad 1, stride 1  (Caffe: cross-correlation, weights OIHW, im2col + SGEMM)
# --------------------------------------------------------------------------------------------

def im2col_3x3(x):
    """x f32[C,H,W] -> col f32[C*9, H*W]; row index = (c*3 + ky)*3 + kx, zero padding of 1."""
    c, h, w = x.shape
    xp = np.zeros((c, h + 2, w + 2), np.float32)
    xp[:, 1:-1, 1:-1] = x
    col = np.empty((c, 3, 3, h, w), np.float32)
    for ky in range(3):
        for kx in range(3):
            col[:, ky, kx] = xp[:, ky:ky + h, kx:kx + w]
    return col.reshape(c * 9, h * w)


def col2im_3x3(col, c, h, w):
    """Adjoint of im2col_3x3: scatter-add col f32[C*9, H*W] back into f32[C,H,W]."""
    col = col.reshape(c, 3, 3, h, w)
    xp = np.zeros((c, h + 2, w + 2), np.float32)
    for ky in range(3):
        for kx in range(3):
            xp[:, ky:ky + h, kx:kx + w] += col[:, ky, kx]
    return np.ascontiguousarray(xp[:, 1:-1, 1:-1])


def _sgemm(a, b, trans_a=False, trans_b=False):
    """C-contiguous float32 a @ b through BLAS SGEMM (as Caffe's caffe_cpu_gemm<float> does)."""
    # scipy BLAS is column-major: compute (b^T a^T)^T by swapping operands on transposed views.
    return blas.sgemm(1.0, b.T, a.T, trans_a=trans_b, trans_b=trans_a).T


def conv3x3_forward(x, weight, bias):
    """top = W (*) x + b.  x f32[Cin,H,W], weight f32[Cout,Cin,3,3], bias f32[Cout]."""
    cout = weight.shape[0]
    _, h, w = x.shape
    col = im2col_3x3(x)
    top = _sgemm(np.ascontiguousarray(weight.reshape(cout, -1)), col)
    top += bias.reshape(cout, 1)
    return np.ascontiguousarray(top.reshape(cout, h, w), dtype=np.float32)


def conv3x3_backward_data(top_diff, weight):
    """bottom_diff = col2im(W^T top_diff) (overwrites, never accumulates)."""
    cout, cin = weight.shape[:2]
    _, h, w = top_diff.shape
    wmat = np.ascontiguousarray(weight.reshape(cout, cin * 9))
    col = _sgemm(wmat, np.ascontiguousarray(top_diff.reshape(cout, h * w)), trans_a=True)
    return col2im_3x3(np.ascontiguousarray(col), cin, h, w)


def conv3x3_backward_weight(top_diff, x):
    """dW = top_diff col(x)^T, db = sum(top_diff).  Computed (and discarded) by the reference's
    Caffe path on every backward; only used here by the CPU-baseline timing leg."""
    cout, h, w = top_diff.shape
    col = im2col_3x3(x)
    dw = _sgemm(np.ascontiguousarray(top_diff.reshape(cout, h * w)), col, trans_b=True)
    db = top_diff.reshape(cout, -1).sum(axis=1)
    return dw.reshape(cout, x.shape[0], 3, 3), db


# --------------------------------------------------------------------------------------------
# ReLU (in place)
# --------------------------------------------------------------------------------------------

def relu_forward_(blob):
    np.maximum(blob, 0, out=blob)
    return blob


def relu_backward_(diff, data):
    """In place: diff *= (data > 0), data being the (post-ReLU) shared top/bottom blob."""
    diff *= (data > 0)
    return diff


# --------------------------------------------------------------------------------------------
# Pooling 2x2, stride 2, no padding, ceil mode
# --------------------------------------------------------------------------------------------

def pooled_size(n):
    """Caffe: ceil((n + 2*pad - kernel) / stride) + 1 with pad=0, kernel=stride=2."""
    return int(np.ceil((n - 2) / 2)) + 1 if n >= 2 else 1


def _windows(x, fill):
    """f32[C,H,W] -> f32[C,Ho,Wo,4] of the 2x2 windows in Caffe scan order (h outer, w inner)."""
    c, h, w = x.shape
    ho, wo = pooled_size(h), pooled_size(w)
    xp = np.full((c, 2 * ho, 2 * wo), fill, np.float32)
    xp[:, :h, :w] = x
    return xp.reshape(c, ho, 2, wo, 2).transpose(0, 1, 3, 2, 4).reshape(c, ho, wo, 4)


def maxpool_forward(x):
    """Returns (top, argmax) where argmax in {0,1,2,3} = first maximum in scan order (strict >
    against an accumulator initialised to -FLT_MAX)."""
    win = _windows(x, -FLT_MAX)
    arg = np.argmax(win, axis=-1)          # numpy returns the FIRST maximal index
    top = np.take_along_axis(win, arg[..., None], axis=-1)[..., 0]
    return np.ascontiguousarray(top), arg.astype(np.int8)


def maxpool_backward(top_diff, arg, in_shape):
    """bottom_diff zero-filled, then += top_diff at the saved argmax."""
    c, h, w = in_shape
    ho, wo = top_diff.shape[1:]
    win = np.zeros((c, ho, wo, 4), np.float32)
    np.put_along_axis(win, arg[..., None].astype(np.int64), top_diff[..., None], axis=-1)
    xp = win.reshape(c, ho, wo, 2, 2).transpose(0, 1, 3, 2, 4).reshape(c, 2 * ho, 2 * wo)
    return np.ascontiguousarray(xp[:, :h, :w])


def _ave_counts(h, w):
    ho, wo = pooled_size(h), pooled_size(w)
    ch = np.minimum(2 * np.arange(ho) + 2, h) - 2 * np.arange(ho)
    cw = np.minimum(2 * np.arange(wo) + 2, w) - 2 * np.arange(wo)
    return (ch[:, None] * cw[None, :]).astype(np.float32)


def avepool_forward(x):
    """Mean over the window clipped to the (unpadded) input."""
    win = _windows(x, 0.0)
    return np.ascontiguousarray(win.sum(axis=-1) / _ave_counts(*x.shape[1:]))


def avepool_backward(top_diff, in_shape):
    c, h, w = in_shape
    ho, wo = top_diff.shape[1:]
    share = top_diff / _ave_counts(h, w)
    xp = np.broadcast_to(share[:, :, None, :, None], (c, ho, 2, wo, 2)).reshape(c, 2 * ho, 2 * wo)
    return np.ascontiguousarray(xp[:, :h, :w])

"""Device-side mirrors of the reference's optimizers (optimizers.py), same method names.

State lives in HBM as float32 CUDA tensors; one fused kernel performs the Adam/EWMA update
(st_adam_step), the L-BFGS two-loop recursion runs as a chain of dot/axpy kernels whose scalars
stay on the device (st_lbfgs_inv_hv).  ``roll`` is virtual: all state stays in the un-rolled
frame, which is equivalent because every operation here is element-wise or a global dot product.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class EWMA:
    """Restates ``average.EWMA`` bookkeeping (third-party, optimizers.py:22-24); the array update
    itself happens inside st_adam_step."""

    def __init__(self, like, beta, correct_bias=True):
        self.beta = float(beta)
        self.beta_accum = 1.0 if correct_bias else 0.0
        self.value = torch.zeros_like(like)

    def tick(self):
        self.beta_accum *= self.beta
        return np.float32(1 - self.beta_accum)


class AdamOptimizer:
    """Adam with iterate averaging (optimizers.py:11-61)."""

    def __init__(self, params, step_size=1, b1=0.9, b2=0.999, bp1=0, decay=0, power=1,
                 biased_g1=False):
        self.params = params
        self.step_size = step_size
        self.decay, self.power = decay, power
        self.i = 1
        self.xy = np.zeros(2, dtype=np.int32)
        self.g1 = EWMA(params, b1, correct_bias=not biased_g1)
        self.g2 = EWMA(params, b2)
        self.p1 = EWMA(params, bp1)
        # the averaged iterate is returned alternately in two buffers, so that a consumer on another
        # stream (statistics / picture of step i) may still be reading while step i + 1 is enqueued
        self._avg = [torch.empty_like(params), torch.empty_like(params)]
        self._avg_i = 0
        self.avg = self._avg[0]

    def update(self, opfunc):
        """Returns (averaged iterate, loss); ``opfunc(params) -> (loss, grad)`` on the device."""
        step_size = self.step_size / self.i ** self.power
        self.i += self.decay
        loss, grad = opfunc(self.params)
        self._avg_i ^= 1
        self.avg = self._avg[self._avg_i]
        _lib.call('st_adam_step', _ptr(self.params), _ptr(grad), _ptr(self.g1.value),
                  _ptr(self.g2.value), _ptr(self.p1.value), _ptr(self.avg), self.params.numel(),
                  step_size, self.g1.beta, self.g2.beta, self.p1.beta, self.g1.tick(),
                  self.g2.tick(), self.p1.tick(), _stream())
        return self.avg, loss

    def roll(self, xy):
        """Virtual: the state stays in the un-rolled frame (see module docstring)."""
        self.xy += np.asarray(xy, dtype=np.int32)

    def set_params(self, last_iterate, resize=None):
        """Cross-scale restart (optimizers.py:53-61).  ``resize(tensor, hw, method)`` resamples a
        state array; ``beta_accum`` is deliberately NOT reset, as in the reference."""
        self.i = 1
        self.params = last_iterate
        hw = tuple(self.params.shape[-2:])
        if tuple(self.g1.value.shape[-2:]) != hw:
            if resize is None:
                raise ValueError('set_params with a new shape needs a resize function')
            self.g1.value = resize(self.g1.value, hw, 'lanczos')
            self.g2.value = torch.clamp_min(resize(self.g2.value, hw, 'bilinear'), 0)
            self.p1.value = resize(self.p1.value, hw, 'lanczos')
        self._avg = [torch.empty_like(self.params), torch.empty_like(self.params)]
        self.avg = self._avg[self._avg_i]


class LBFGSOptimizer:
    """L-BFGS with fixed size steps, no line search (optimizers.py:64-138).  The curvature pairs,
    their s.y products and the pair count live on the device (``st_lbfgs_step`` /
    ``st_lbfgs_commit``): an update enqueues kernels only -- no ``.item()``, no host decision --
    so the host runs ahead of the GPU exactly as it does with Adam."""

    def __init__(self, params, initial_step=0.1, n_corr=10):
        self.params = params
        self.initial_step = initial_step
        self.n_corr = n_corr
        self.xy = np.zeros(2, dtype=np.int32)
        self.loss, self.grad = None, None
        self._alloc()

    def _alloc(self):
        n, dev = self.params.numel(), self.params.device
        self._ring_s = torch.empty((self.n_corr + 1, n), dtype=torch.float32, device=dev)
        self._ring_y = torch.empty((self.n_corr + 1, n), dtype=torch.float32, device=dev)
        self._state = torch.zeros(64, dtype=torch.float64, device=dev)     # empty memory
        self._scratch = torch.empty(n, dtype=torch.float32, device=dev)

    def update(self, opfunc):
        if self.loss is None:                                                    # (:76-77)
            self.loss, self.grad = opfunc(self.params)
        n = self.params.numel()
        # s = -H g, scaled (:80-84), params += s (:85)
        _lib.call('st_lbfgs_step', _ptr(self.grad), n, self.n_corr, _ptr(self._ring_s),
                  _ptr(self._ring_y), _ptr(self._state), _ptr(self._scratch), _ptr(self.params),
                  self.initial_step, _stream())
        loss, grad = opfunc(self.params)                                         # (:88)
        # y = grad - self.grad; the pair is kept iff s.y > 1e-10 (:89-103)
        _lib.call('st_lbfgs_commit', _ptr(grad), _ptr(self.grad), n, self.n_corr,
                  _ptr(self._ring_s), _ptr(self._ring_y), _ptr(self._state), _stream())
        self.loss, self.grad = loss, grad
        return self.params, loss

    @property
    def sk(self):
        """The stored s vectors, oldest first (host view for tests / inspection: synchronises)."""
        st = self._state.cpu().numpy()
        count, head = int(st[0]), int(st[1])
        slots = [(head - 1 - k) % (self.n_corr + 1) for k in range(count)]
        return [self._ring_s[i] for i in reversed(slots)]

    def roll(self, xy):
        self.xy += np.asarray(xy, dtype=np.int32)

    def set_params(self, last_iterate, resize=None):
        """Cross-scale restart (optimizers.py:134-138): new parameters, memory cleared."""
        self.params = last_iterate
        self.loss, self.grad = None, None
        if self._scratch.numel() != self.params.numel():
            self._alloc()
        else:
            self._state.zero_()

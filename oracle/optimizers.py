"""Oracle (TEST INFRASTRUCTURE): the reference's optimizers, restated.

Follows reference ``optimizers.py``: Adam with iterate averaging :11-61, fixed-step L-BFGS
:64-138.  ``Ewma`` restates the third-party ``average.EWMA`` (requirements.txt:3 ``average>=1.0``,
not installed anywhere here) from its published behaviour as used at optimizers.py:22-24, 35-42:
value <- beta*value + (1-beta)*x, optional bias correction by 1 - beta^t.

``set_params`` (cross-scale restart, optimizers.py:53-61 / :134-138) resamples Adam's state with
``oracle.numeric.resize`` (== the reference's ``num_utils.resize`` == PIL, bit for bit); pinned
against the reference's own module by tests/golden/set_params.npz.
"""

import numpy as np

from .numeric import EPS, resize, roll2_, sdot


class Ewma:
    def __init__(self, like, beta, correct_bias=True):
        self.beta = float(beta)
        self.beta_accum = 1.0
        self.correct_bias = correct_bias
        self.value = np.zeros_like(like)

    def update(self, x):
        self.beta_accum *= self.beta
        self.value *= np.float32(self.beta)
        self.value += np.float32(1 - self.beta) * x

    def get(self):
        if self.correct_bias:
            return self.value / np.float32(1 - self.beta_accum)
        return self.value.copy()


class Adam:
    """optimizers.py:11-51."""

    def __init__(self, params, step_size=1, b1=0.9, b2=0.999, bp1=0, decay=0, power=1,
                 biased_g1=False):
        self.params = params
        self.step_size, self.decay, self.power = step_size, decay, power
        self.i = 1
        self.xy = np.zeros(2, dtype=np.int32)
        self.g1 = Ewma(params, b1, correct_bias=not biased_g1)
        self.g2 = Ewma(params, b2)
        self.p1 = Ewma(params, bp1)

    def update(self, opfunc):
        step_size = self.step_size / self.i ** self.power          # :29
        self.i += self.decay                                       # :30
        loss, grad = opfunc(self.params)                           # :32
        self.g1.update(grad)                                       # :35
        self.g2.update(grad ** 2)                                  # :36
        step = self.g1.get() / (np.sqrt(self.g2.get()) + EPS)      # :37
        self.params += np.float32(-step_size) * step               # :38 (saxpy, in place)
        self.p1.update(self.params)                                # :41
        return roll2_(self.p1.get(), -self.xy), loss               # :42

    def roll(self, xy):
        xy = np.asarray(xy)
        if (xy == 0).all():
            return
        self.xy += xy
        for ew in (self.g1, self.g2, self.p1):
            roll2_(ew.value, xy)

    def set_params(self, last_iterate):
        """optimizers.py:53-61: step counter back to 1, state resampled to the new size (g1 / p1
        Lanczos, g2 bilinear and clamped at 0); the EWMAs' ``beta_accum`` is NOT reset."""
        self.i = 1
        self.params = last_iterate
        hw = self.params.shape[-2:]
        self.g1.value = resize(self.g1.value, hw)
        self.g2.value = np.maximum(0, resize(self.g2.value, hw, 'bilinear'))
        self.p1.value = resize(self.p1.value, hw)


class Lbfgs:
    """optimizers.py:64-138."""

    def __init__(self, params, initial_step=0.1, n_corr=10):
        self.params = params
        self.initial_step, self.n_corr = initial_step, n_corr
        self.xy = np.zeros(2, dtype=np.int32)
        self.loss = self.grad = None
        self.sk, self.yk, self.syk = [], [], []

    def update(self, opfunc):
        if self.loss is None:                                      # :76-77
            self.loss, self.grad = opfunc(self.params)
        s = -self.inv_hv(self.grad)                                # :80
        if not self.sk:                                            # :81-84
            s *= self.initial_step / np.mean(abs(s))
        elif len(self.sk) < self.n_corr:
            s *= len(self.sk) / self.n_corr
        self.params += s                                           # :85
        loss, grad = opfunc(self.params)                           # :88
        self._store(s, grad - self.grad)                           # :89-90
        self.loss, self.grad = loss, grad
        return self.params, loss

    def _store(self, s, y):
        sy = sdot(s, y)                                            # :97-103
        if sy > 1e-10:
            self.sk.append(s)
            self.yk.append(y)
            self.syk.append(sy)
        if len(self.sk) > self.n_corr:
            self.sk, self.yk, self.syk = self.sk[1:], self.yk[1:], self.syk[1:]

    def inv_hv(self, p):
        """Two-loop recursion (optimizers.py:105-121)."""
        p = p.copy()
        alphas = []
        for s, y, sy in zip(reversed(self.sk), reversed(self.yk), reversed(self.syk)):
            alphas.append(sdot(s, p) / sy)
            p += np.float32(-alphas[-1]) * y
        if self.sk:
            p *= self.syk[-1] / sdot(self.yk[-1], self.yk[-1])
        for s, y, sy, alpha in zip(self.sk, self.yk, self.syk, reversed(alphas)):
            beta = sdot(y, p) / sy
            p += np.float32(alpha - beta) * s
        return p

    def roll(self, xy):
        xy = np.asarray(xy)
        if (xy == 0).all():
            return
        self.xy += xy
        if self.grad is not None:
            roll2_(self.grad, xy)
        for s, y in zip(self.sk, self.yk):
            roll2_(s, xy)
            roll2_(y, xy)

    def set_params(self, last_iterate):
        """optimizers.py:134-138: new parameters, memory cleared."""
        self.params = last_iterate
        self.loss, self.grad = None, None
        self.sk, self.yk, self.syk = [], [], []

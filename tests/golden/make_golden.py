#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the REFERENCE's own modules (in the build container
only -- /root/reference does not exist on the GPU box; the committed .npz files travel instead).

    python tests/golden/make_golden.py [/root/reference]

``num_utils.py`` needs ``pywt`` only for the out-of-scope SWT regulariser, and ``optimizers.py``
needs the third-party ``average.EWMA``; neither package is installed, so import stubs are
registered first.  The ``average`` stub restates the published EWMA behaviour (see
oracle/optimizers.py); everything else executed below is the reference's unmodified code.
"""

import os
import sys
import types

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    sys.modules['pywt'] = types.ModuleType('pywt')
    average = types.ModuleType('average')

    class EWMA:
        def __init__(self, shape=(), dtype=np.float64, beta=0.9, correct_bias=True):
            self.beta = beta
            self.beta_accum = 1 if correct_bias else 0
            self.value = np.zeros(shape, dtype)

        @classmethod
        def like(cls, arr, beta=0.9, correct_bias=True):
            return cls(arr.shape, arr.dtype, beta, correct_bias)

        def get(self):
            return self.value / (1 - self.beta_accum)

        def update(self, datum):
            self.beta_accum *= self.beta
            self.value *= self.beta
            self.value += (1 - self.beta) * datum
            return self.get()

    average.EWMA = EWMA
    sys.modules['average'] = average


def opfunc_factory(target, coupling):
    """A smooth non-separable test objective: 0.5|A(x-t)|^2-like with a periodic stencil."""
    def opfunc(x):
        d = x - target
        lap = d + coupling * (np.roll(d, 1, -1) + np.roll(d, 1, -2))
        loss = 0.5 * float(np.sum(lap * lap))
        grad = lap + coupling * (np.roll(lap, -1, -1) + np.roll(lap, -1, -2))
        return loss, np.float32(grad)
    return opfunc


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import num_utils as nu
    import optimizers as ro

    rs = np.random.RandomState(20240917)
    out = {}

    # ---- num_utils -------------------------------------------------------------------------
    feat = rs.randn(12, 9, 7).astype(np.float32)
    feat[feat < -0.3] = 0                      # post-ReLU-like sparsity
    style = nu.gram_matrix(rs.rand(12, 9, 7).astype(np.float32))
    gram = nu.gram_matrix(feat)
    out['nu_feat'], out['nu_style_gram'], out['nu_gram'] = feat, style, gram
    gdiff = gram - style
    out['nu_ssymm'] = nu.ssymm(gdiff, feat.reshape(12, -1))
    out['nu_norm2_gdiff'] = np.float32(nu.norm2(gdiff))
    out['nu_normalize'] = nu.normalize(out['nu_ssymm'].copy())
    x = (rs.rand(3, 10, 13).astype(np.float32) - 0.5) * 2
    out['nu_x'] = x
    for p in (1, 2, 6, 3.5):
        loss, grad = nu.p_norm(x.copy(), p)
        out['nu_pnorm_loss_%s' % p], out['nu_pnorm_grad_%s' % p] = np.float32(loss), np.float32(grad)
    for beta in (2, 1.5, 1):
        loss, grad = nu.tv_norm(x.copy(), beta)
        out['nu_tv_loss_%s' % beta], out['nu_tv_grad_%s' % beta] = np.float32(loss), np.float32(grad)
    out['nu_roll2'] = nu.roll2(x.copy(), np.array([3, -4], dtype=np.int32))
    out['nu_eps'] = np.float32(nu.EPS)
    np.savez(os.path.join(OUT, 'num_utils.npz'), **out)

    # ---- optimizers ------------------------------------------------------------------------
    out = {}
    shape = (3, 8, 12)
    target = rs.randn(*shape).astype(np.float32) * 20
    x0 = rs.randn(*shape).astype(np.float32) * 50
    rolls = rs.randint(-4, 5, size=(16, 2)).astype(np.int32)
    out['target'], out['x0'], out['rolls'] = target, x0, rolls
    rolled_target = target.copy()

    def rolled_opfunc(cum):
        # the objective lives in the rolled frame, like the reference's rolled image/features
        return opfunc_factory(np.roll(target, tuple(cum), axis=(-1, -2)), 0.25)

    for name, biased in (('adam', False), ('adam_biased', True)):
        params = x0.copy()
        opt = ro.AdamOptimizer(params, step_size=15, bp1=1 - 1 / 20, decay=0.05, power=0.5,
                               biased_g1=biased)
        traj, losses = [], []
        for it in range(8):
            xy = rolls[it]
            nu.roll2(params, xy)
            opt.roll(xy)
            avg, loss = opt.update(rolled_opfunc(xy))
            nu.roll2(params, -xy)
            opt.roll(-xy)
            traj.append(avg.copy())
            losses.append(loss)
        out[name + '_avg'], out[name + '_loss'] = np.stack(traj), np.float64(losses)
        out[name + '_params'] = params.copy()

    params = x0.copy()
    opt = ro.LBFGSOptimizer(params)
    traj, losses = [], []
    for it in range(16):
        xy = rolls[it]
        nu.roll2(params, xy)
        opt.roll(xy)
        _, loss = opt.update(rolled_opfunc(xy))
        nu.roll2(params, -xy)
        opt.roll(-xy)
        traj.append(params.copy())
        losses.append(loss)
    out['lbfgs_params'], out['lbfgs_loss'] = np.stack(traj), np.float64(losses)
    out['lbfgs_mem'] = np.int32(len(opt.sk))
    np.savez(os.path.join(OUT, 'optimizers.npz'), **out)

    # ---- num_utils.resize (its own RNG and file: the fixtures above stay bit-identical) -----------
    from PIL import Image
    rr = np.random.RandomState(77)
    out = {}
    cases = [((3, 37, 53), (52, 75)), ((3, 64, 48), (91, 68)), ((2, 50, 70), (36, 49)),
             ((1, 45, 45), (45, 64))]
    for i, (shape, hw) in enumerate(cases):
        x = (rr.rand(*shape) * 300 - 120).astype(np.float32)
        out['in_%d' % i], out['hw_%d' % i] = x, np.int32(hw)
        out['lanczos_%d' % i] = nu.resize(x, hw)
        out['bilinear_%d' % i] = nu.resize(x, hw, method=Image.BILINEAR)
    out['n_cases'] = np.int32(len(cases))
    np.savez(os.path.join(OUT, 'resize.npz'), **out)
    # ---- AdamOptimizer.set_params across a scale change (optimizers.py:53-61; its own file) --------
    rs2 = np.random.RandomState(5150)
    out = {}
    shape0, shape1 = (3, 12, 17), (3, 17, 24)
    tgt0 = rs2.randn(*shape0).astype(np.float32) * 20
    tgt1 = rs2.randn(*shape1).astype(np.float32) * 20
    x0 = rs2.randn(*shape0).astype(np.float32) * 50
    out['target0'], out['target1'], out['x0'] = tgt0, tgt1, x0
    params = x0.copy()
    opt = ro.AdamOptimizer(params, step_size=15, bp1=1 - 1 / 20, decay=0.05, power=0.5)
    for it in range(5):
        avg, _ = opt.update(opfunc_factory(tgt0, 0.25))
    out['avg_scale0'] = avg.copy()
    new_params = nu.resize(avg, shape1[-2:])                  # style_transfer.py:877-881
    out['params_scale1'] = new_params.copy()
    opt.set_params(new_params)
    out['g1_after'], out['g2_after'], out['p1_after'] = (opt.g1.value.copy(), opt.g2.value.copy(),
                                                         opt.p1.value.copy())
    out['i_after'] = np.float64(opt.i)
    out['beta_accum_after'] = np.float64([opt.g1.beta_accum, opt.g2.beta_accum, opt.p1.beta_accum])
    traj = []
    for it in range(4):
        avg, _ = opt.update(opfunc_factory(tgt1, 0.25))
        traj.append(avg.copy())
    out['avg_scale1'] = np.stack(traj)
    out['params_final'] = opt.params.copy()
    np.savez(os.path.join(OUT, 'set_params.npz'), **out)
    print('wrote', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()

"""Pins oracle/numeric.py and oracle/optimizers.py against outputs of the REFERENCE's own
num_utils.py / optimizers.py (tests/golden/*.npz, written by tests/golden/make_golden.py)."""

import os

import numpy as np
import pytest

from oracle import numeric as on
from oracle import optimizers as oo


@pytest.fixture(scope='module')
def nu(golden_dir):
    return np.load(os.path.join(golden_dir, 'num_utils.npz'))


@pytest.fixture(scope='module')
def og(golden_dir):
    return np.load(os.path.join(golden_dir, 'optimizers.npz'))


def close(a, b, rtol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= rtol * scale, (np.abs(a - b).max(), scale)


def test_eps(nu):
    assert on.EPS == nu['nu_eps']


def test_gram_is_lower_triangular_and_matches(nu):
    g = on.gram_lower(nu['nu_feat'])
    assert np.all(np.triu(g, 1) == 0)
    assert np.all(np.triu(nu['nu_gram'], 1) == 0)       # the reference's SSYRK quirk itself
    close(g, nu['nu_gram'])


def test_ssymm_norm2_normalize(nu):
    gdiff = nu['nu_gram'] - nu['nu_style_gram']
    s = on.symm_times(gdiff, nu['nu_feat'].reshape(12, -1))
    close(s, nu['nu_ssymm'])
    close(on.norm2(gdiff), nu['nu_norm2_gdiff'])
    close(on.normalize_(s.copy()), nu['nu_normalize'])


@pytest.mark.parametrize('p', [1, 2, 6, 3.5])
def test_p_norm(nu, p):
    loss, grad = on.p_norm(nu['nu_x'].copy(), p)
    close(loss, nu['nu_pnorm_loss_%s' % p])
    close(grad, nu['nu_pnorm_grad_%s' % p])


@pytest.mark.parametrize('beta', [2, 1.5, 1])
def test_tv_norm(nu, beta):
    loss, grad = on.tv_norm(nu['nu_x'].copy(), beta)
    close(loss, nu['nu_tv_loss_%s' % beta])
    close(grad, nu['nu_tv_grad_%s' % beta], rtol=1e-4)


def test_roll2_axis_convention(nu):
    x = nu['nu_x'].copy()
    rolled = on.roll2_(x, np.array([3, -4]))
    assert np.array_equal(rolled, nu['nu_roll2'])
    # xy[0] moves along the WIDTH axis, xy[1] along the HEIGHT axis
    assert rolled[0, (0 - 4) % 10, 3] == nu['nu_x'][0, 0, 0]


def _opfunc(target, cum, coupling=0.25):
    tgt = np.roll(target, tuple(cum), axis=(-1, -2))

    def opfunc(x):
        d = x - tgt
        lap = d + coupling * (np.roll(d, 1, -1) + np.roll(d, 1, -2))
        loss = 0.5 * float(np.sum(lap * lap))
        grad = lap + coupling * (np.roll(lap, -1, -1) + np.roll(lap, -1, -2))
        return loss, np.float32(grad)
    return opfunc


@pytest.mark.parametrize('name,biased', [('adam', False), ('adam_biased', True)])
def test_adam_trajectory(og, name, biased):
    params = og['x0'].copy()
    opt = oo.Adam(params, step_size=15, bp1=1 - 1 / 20, decay=0.05, power=0.5, biased_g1=biased)
    for it in range(8):
        xy = og['rolls'][it]
        on.roll2_(params, xy)
        opt.roll(xy)
        avg, loss = opt.update(_opfunc(og['target'], xy))
        on.roll2_(params, -xy)
        opt.roll(-xy)
        close(avg, og[name + '_avg'][it], rtol=1e-5)
        close(loss, og[name + '_loss'][it], rtol=1e-5)
    close(params, og[name + '_params'], rtol=1e-5)


def test_lbfgs_trajectory(og):
    params = og['x0'].copy()
    opt = oo.Lbfgs(params)
    for it in range(16):
        xy = og['rolls'][it]
        on.roll2_(params, xy)
        opt.roll(xy)
        _, loss = opt.update(_opfunc(og['target'], xy))
        on.roll2_(params, -xy)
        opt.roll(-xy)
        # fixed-step L-BFGS amplifies float32 round-off along the trajectory; compare loosely late
        tol = 1e-4 if it < 8 else 2e-2
        close(params, og['lbfgs_params'][it], rtol=tol)
    assert len(opt.sk) == int(og['lbfgs_mem'])


def test_resize_matches_reference_golden(golden_dir):
    """oracle.numeric.resize (a restatement of Pillow's float resampling) against outputs of the
    reference's own num_utils.resize (:90-108): bit for bit, Lanczos and bilinear, up and down."""
    g = np.load(os.path.join(golden_dir, 'resize.npz'))
    for i in range(int(g['n_cases'])):
        hw = tuple(int(v) for v in g['hw_%d' % i])
        for method in ('lanczos', 'bilinear'):
            got = on.resize(g['in_%d' % i], hw, method)
            assert got.dtype == np.float32 and got.shape == g['%s_%d' % (method, i)].shape
            assert np.array_equal(got, g['%s_%d' % (method, i)]), (i, method)


def test_resize_matches_pil_directly():
    """The same against PIL itself on a fresh case (PIL is present on both machines)."""
    from PIL import Image
    rs = np.random.RandomState(5)
    a = (rs.rand(41, 29) * 255 - 100).astype(np.float32)
    for hw in ((58, 41), (29, 20), (41, 50)):
        for method, pil in (('lanczos', Image.LANCZOS), ('bilinear', Image.BILINEAR)):
            want = np.asarray(Image.fromarray(a).resize((hw[1], hw[0]), pil), dtype=np.float32)
            assert np.array_equal(on.resize(a, hw, method), want), (hw, method)


def test_adam_set_params_across_a_scale_change(golden_dir):
    """``AdamOptimizer.set_params`` (optimizers.py:53-61) pinned against the reference's own module
    (tests/golden/set_params.npz): five steps at 12x17, the averaged iterate resampled to 17x24 and
    handed back, state resampled (g1 / p1 Lanczos, g2 bilinear + clamp), i = 1, beta_accum kept, four
    more steps."""
    g = np.load(os.path.join(golden_dir, 'set_params.npz'))
    params = g['x0'].copy()
    opt = oo.Adam(params, step_size=15, bp1=1 - 1 / 20, decay=0.05, power=0.5)
    for it in range(5):
        avg, _ = opt.update(_opfunc(g['target0'], (0, 0)))
    close(avg, g['avg_scale0'], rtol=1e-5)
    new_params = on.resize(avg, g['target1'].shape[-2:])
    close(new_params, g['params_scale1'], rtol=1e-5)
    opt.set_params(new_params)
    assert opt.i == float(g['i_after']) == 1
    close(opt.g1.value, g['g1_after'], rtol=1e-5)
    close(opt.g2.value, g['g2_after'], rtol=1e-5)
    close(opt.p1.value, g['p1_after'], rtol=1e-5)
    assert (opt.g2.value >= 0).all()
    close(np.float64([opt.g1.beta_accum, opt.g2.beta_accum, opt.p1.beta_accum]),
          g['beta_accum_after'], rtol=1e-12)
    for it in range(4):
        avg, _ = opt.update(_opfunc(g['target1'], (0, 0)))
        close(avg, g['avg_scale1'][it], rtol=1e-5)
    close(opt.params, g['params_final'], rtol=1e-5)

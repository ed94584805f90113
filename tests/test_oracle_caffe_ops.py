"""Anchors oracle/caffe_ops.py + caffe_net.py (restated, un-pinnable Caffe) on torch's CPU
functional ops and on autograd: the segmented backward with gradient injection must equal the
gradient of the surrogate  sum_l <inj_l, Z_l>  (Z_l = conv output of loss layer l, pre-ReLU)."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import caffe_ops as ops
from oracle.caffe_net import OracleNet, he_normal_weights, model_layers, vgg_layers

torch.set_num_threads(4)


def rnd(rs, *shape):
    return rs.randn(*shape).astype(np.float32)


@pytest.mark.parametrize('hw', [(5, 7), (16, 16), (9, 1), (1, 1)])
def test_conv_forward_backward_vs_torch(hw):
    rs = np.random.RandomState(1)
    x, w, b = rnd(rs, 6, *hw), rnd(rs, 10, 6, 3, 3), rnd(rs, 10)
    top = ops.conv3x3_forward(x, w, b)
    ref = F.conv2d(torch.from_numpy(x)[None], torch.from_numpy(w), torch.from_numpy(b), padding=1)
    np.testing.assert_allclose(top, ref[0].numpy(), rtol=1e-4, atol=1e-4)
    dy = rnd(rs, 10, *hw)
    xt = torch.from_numpy(x)[None].requires_grad_()
    wt = torch.from_numpy(w).requires_grad_()
    F.conv2d(xt, wt, padding=1).backward(torch.from_numpy(dy)[None])
    np.testing.assert_allclose(ops.conv3x3_backward_data(dy, w), xt.grad[0].numpy(),
                               rtol=1e-4, atol=1e-4)
    dw, db = ops.conv3x3_backward_weight(dy, x)
    np.testing.assert_allclose(dw, wt.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(db, dy.reshape(10, -1).sum(1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('n,expect', [(1, 1), (2, 1), (3, 2), (224, 112), (181, 91), (7, 4)])
def test_pooled_size_ceil_mode(n, expect):
    assert ops.pooled_size(n) == expect


@pytest.mark.parametrize('hw', [(6, 8), (7, 9), (1, 5), (3, 3)])
def test_pools_vs_torch(hw):
    rs = np.random.RandomState(2)
    x = np.maximum(rnd(rs, 4, *hw), 0)          # post-ReLU: plenty of exact ties at 0
    xt = torch.from_numpy(x)[None].requires_grad_()
    dy = rnd(rs, 4, ops.pooled_size(hw[0]), ops.pooled_size(hw[1]))
    top, arg = ops.maxpool_forward(x)
    ref = F.max_pool2d(xt, 2, 2, ceil_mode=True)
    np.testing.assert_array_equal(top, ref[0].detach().numpy())
    ref.backward(torch.from_numpy(dy)[None])
    np.testing.assert_array_equal(ops.maxpool_backward(dy, arg, x.shape), xt.grad[0].numpy())
    xt.grad = None
    ref = F.avg_pool2d(xt, 2, 2, ceil_mode=True, count_include_pad=False)
    np.testing.assert_allclose(ops.avepool_forward(x), ref[0].detach().numpy(), rtol=1e-6)
    ref.backward(torch.from_numpy(dy)[None])
    np.testing.assert_allclose(ops.avepool_backward(dy, x.shape), xt.grad[0].numpy(), rtol=1e-6)


def test_maxpool_first_max_wins():
    x = np.zeros((1, 2, 4), np.float32)
    x[0, :, 2:] = 3.0                            # window 1: all four equal -> index 0
    _, arg = ops.maxpool_forward(x)
    assert arg.tolist() == [[[0, 0]]]
    x[0, 1, 1] = 1.0                             # window 0: unique max at (1,1) -> index 3
    _, arg = ops.maxpool_forward(x)
    assert arg.tolist() == [[[3, 0]]]


def torch_forward(layers, params, x, stop_after):
    """Same graph with torch ops; returns {blob: pre-ReLU conv output / pool output}."""
    blobs, pre = {'data': x}, {}
    for kind, name, bottom, top, extra in layers:
        if kind == 'conv':
            w, b = params[name]
            pre[top] = F.conv2d(blobs[bottom], torch.from_numpy(w).double(),
                                torch.from_numpy(b).double(), padding=1)
            blobs[top] = pre[top]
            if name == stop_after:
                break
        elif kind == 'relu':
            blobs[top] = F.relu(blobs[top])
        elif extra == 'max':
            blobs[top] = pre[top] = F.max_pool2d(blobs[bottom], 2, 2, ceil_mode=True)
        else:
            blobs[top] = pre[top] = F.avg_pool2d(blobs[bottom], 2, 2, ceil_mode=True,
                                                 count_include_pad=False)
    return pre


@pytest.mark.parametrize('pool,big,hw', [('max', False, (24, 20)), ('ave', False, (22, 26)),
                                         ('max', True, (12, 12))])
def test_segmented_backward_equals_autograd_of_surrogate(pool, big, hw):
    layers = vgg_layers('vgg16', pool, big)
    params = he_normal_weights(layers, seed=3)
    rs = np.random.RandomState(4)
    img = rnd(rs, 3, *hw) * 50
    loss_layers = ['conv4_2', 'conv3_1', 'pool1', 'conv1_1']      # deepest first
    net = OracleNet(layers, params)
    net.set_input(img)
    net.forward(end=loss_layers[0])
    ops.relu_forward_(net.data[loss_layers[0]])
    inj = {l: rnd(rs, *net.data[l].shape) for l in loss_layers}
    for l in loss_layers:
        net.diff[l][...] = 0
    for i, l in enumerate(loss_layers):
        net.diff[l] += inj[l]
        if i + 1 == len(loss_layers):
            net.backward(start=l)
        else:
            net.backward(start=l, end=loss_layers[i + 1])
    got = net.diff['data']

    x = torch.from_numpy(img).double()[None].requires_grad_()
    pre = torch_forward(layers, params, x, stop_after=loss_layers[0])
    surrogate = sum((torch.from_numpy(inj[l]).double()[None] * pre[l]).sum() for l in loss_layers)
    surrogate.backward()
    want = x.grad[0].numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 2e-4, err


def test_forward_end_stops_before_relu():
    layers = model_layers('vgg19.prototxt')
    params = he_normal_weights(layers, seed=5)
    net = OracleNet(layers, params)
    net.set_input(rnd(np.random.RandomState(6), 3, 8, 8) * 30)
    net.forward(end='conv1_2')
    assert (net.data['conv1_2'] < 0).any()       # conv1_2 not yet rectified
    assert (net.data['conv1_1'] >= 0).all()      # conv1_1 was (in place)
    assert 'pool1' not in net.data


def test_model_layers_variants():
    assert [l[1] for l in model_layers('vgg16.prototxt') if l[0] == 'conv'][-1] == 'conv5_3'
    assert [l[1] for l in model_layers('vgg19.prototxt') if l[0] == 'conv'][-1] == 'conv5_4'
    assert all(l[4] == 'ave' for l in model_layers('vgg19_avgpool.prototxt') if l[0] == 'pool')
    big = {l[1]: l for l in model_layers('/x/y/vgg19_big.prototxt')}
    assert big['conv2_1'][2] == 'conv1_2' and big['conv3_1'][2] == 'pool2'

"""Two-GPU test of the exchange step: skipped on a machine with fewer than two GPUs (the driver's
single-GPU test box); runs tools/multi_gpu_check.py under torch.distributed.run, one rank per GPU."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_step_is_bit_identical_to_one_rank():
    """Image after 3 Adam steps on 2 ranks (round-robin tiles, one ncclAllGather from the C ABI,
    regulariser + optimizer replicated) == the 1-rank image, bit for bit; losses agree to 1e-9."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    env = dict(os.environ, CHECK_SIZE='512', CHECK_TILE='128')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(29600 + os.getpid() % 300),
           os.path.join(ROOT, 'tools', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0


def test_host_image_path_equals_resident_image():
    """One GPU: ``stage_host_image`` (two-phase upload overlapping the first half of the tiles)
    gives the bits of the resident-image evaluation, with and without a roll."""
    import numpy as np
    import torch
    from style_transfer_b200 import netdesc, weights
    from style_transfer_b200.engine import TileEngine
    from style_transfer_b200.transfer import StyleTransfer, default_args
    size, tile = 256, 64
    args = default_args(size=size, min_size=size, tile_size=tile, model='vgg16.prototxt',
                        content_layers=['conv3_2'], style_layers=['conv1_1', 'conv2_1'])
    net = netdesc.from_model(args.model)
    params = weights.he_normal(net)
    rs = np.random.RandomState(3)
    content, style = (rs.randint(0, 256, (size, size, 3)).astype(np.uint8) for _ in range(2))
    outs = []
    for staged in (False, True):
        eng = TileEngine(net, params, mean=args.mean, precision='fp16')
        st = StyleTransfer(eng, args)
        np.random.seed(0)
        st.init_first_scale(size, size)
        st.prepare([eng.pil_to_image(content)], [eng.pil_to_image(style)])
        for i in range(3):
            if staged:
                host = torch.empty(eng.img.shape, dtype=torch.float32).pin_memory()
                host.copy_(eng.img)
                torch.cuda.synchronize()
                eng.img.zero_()
                eng.stage_host_image(host)
            avg, loss = st.step()
        torch.cuda.synchronize()
        outs.append((avg.clone(), float(loss)))
    assert torch.equal(outs[0][0], outs[1][0])
    assert abs(outs[0][1] - outs[1][1]) <= 1e-12 * abs(outs[0][1])

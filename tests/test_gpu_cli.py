"""GPU test of the command-line surface: a two-scale run through style_transfer.py's entry point."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cli_two_scales(tmp_path, monkeypatch):
    from PIL import Image
    from style_transfer_b200.cli import main
    monkeypatch.chdir(tmp_path)
    rs = np.random.RandomState(0)
    Image.fromarray(rs.randint(0, 256, (80, 120, 3)).astype(np.uint8)).save('content.png')
    Image.fromarray(rs.randint(0, 256, (90, 90, 3)).astype(np.uint8)).save('style.png')
    rc = main(['-ci', 'content.png', '-si', 'style.png', '-oi', 'out.png', '-s', '96', '--min-size',
               '64', '-i', '4', '3', '--tile-size', '64', '--weights', 'random', '--model',
               'vgg16.prototxt', '--content-layers', 'conv3_2', '--style-layers', 'conv1_1',
               'conv2_1', '--save-every', '3'])
    assert rc == 0
    out = Image.open('out.png')
    assert out.size == (96, 64)                      # 120x80 fitted into 96 (resize_to_fit)
    arr = np.asarray(out)
    assert arr.std() > 1                             # not a constant image
    rows = open(sorted(p for p in tmp_path.iterdir() if p.name.endswith('_log.csv'))[0]).read().splitlines()
    assert rows[0] == 'iteration,scale,step,time,content_h,content_w,update_size,loss,tv_norm'
    assert len(rows) == 1 + 4 + 3                    # two scales: 68 -> 96
    assert len([p for p in tmp_path.iterdir() if '_out_' in p.name]) == 2

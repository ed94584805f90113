"""CPU checks of bench.py's contract pieces that need no GPU: both arms name the same workload, the
roofline traffic figure comes from the committed ncu artefact by key, the argument defaults."""

import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench(monkeypatch):
    monkeypatch.setattr(sys, 'argv', ['bench.py'])
    sys.path.insert(0, ROOT)
    import importlib
    import bench
    return importlib.reload(bench)


def test_both_arms_report_the_same_config(monkeypatch):
    bench = load_bench(monkeypatch)
    a = bench.parse_args()
    assert (a.gpus, a.impl, a.precision, a.size, a.tile_size, a.optimizer) == (1, 'engine', 'fp16', 2048, 512, 'adam')
    assert a.warmup >= 3
    # the reference arm prints workload_config(a, 1), the engine arm workload_config(a, world): equal
    # for every world size (layout keys -- tiles per GPU, precision -- live in `layout`)
    for world in (1, 2, 4, 8):
        assert bench.workload_config(a, world) == bench.workload_config(a, 1)
        lay = bench.layout_config(a, world)
        assert lay['tiles_per_gpu'] == 16 // world and lay['precision'] == 'fp16'
    cfg = bench.workload_config(a, 1)
    assert 'workload' in cfg and 'l2' in cfg and 'model' not in cfg
    assert cfg['tiles'] == 16 and 'cfg3' in cfg['workload']
    assert bench.METRIC.startswith('style-transfer iterations/sec')


def test_roofline_traffic_comes_from_the_committed_profile(monkeypatch):
    bench = load_bench(monkeypatch)
    a = SimpleNamespace(size=2048, tile_size=512, precision='fp16')
    traffic, source = bench.measured_traffic('conv_tc2_kernel', a, 1)
    table = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    key = 'conv_tc2_kernel|size=2048|tile=512|precision=fp16|gpus=1'
    assert traffic == table[key]['dram_bytes_per_launch'] and 'ncu' in source
    # algorithmic bytes per 3x3 launch are ~280 MB: measured traffic within 10 % above
    assert 280e6 <= traffic <= 1.1 * 280e6
    # no capture for another configuration -> null, never a literal
    assert bench.measured_traffic('conv_tc2_kernel', SimpleNamespace(size=1024, tile_size=512, precision='fp16'), 1) == (None, None)
    assert bench.measured_traffic('conv_tc2_kernel', a, 8) == (None, None)

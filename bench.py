#!/usr/bin/env python
"""Benchmark of the hot path: optimiser iterations per second of one scale of style transfer.

Workload (BASELINE.json metric, SURVEY section 8d "cfg3"): a 2048x2048 image cut into 4x4 tiles of
512x512, VGG-19 (random-init He-normal weights), 5 style layers + conv4_2 content layer, TV and
p-norm regularisers at their default weights, Adam with iterate averaging -- i.e. one pass of the
loop body of the reference's ``StyleTransfer.transfer`` (style_transfer.py:771-828) per step,
steady state, preprocessing excluded.  With N GPUs the 16 tiles are dealt round-robin to the ranks
(tile i -> rank i mod N, TileWorkerPool.request :284-298), the gradient tiles are all-gathered and
every rank runs the regularisers + optimizer redundantly: total work is fixed ("strong" scaling).

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA engine
  python bench.py --impl reference [...]                      the reference's CPU algorithm (the
        oracle port: Caffe's im2col+SGEMM forward/backward incl. dW, reference Gram/loss code,
        reference optimizer) on the host cores, one tile-evaluation per step, scaled to 16 tiles

Rank 0 prints ONE JSON line.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1; the CPU legs (reference arm, cpu_baseline) are meant to use
# every host core, and OpenBLAS reads the variable when numpy loads -- so fix it before that import.
_HOST_CORES = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
if '--impl' in sys.argv and 'reference' in sys.argv or int(os.environ.get('WORLD_SIZE', '1')) == 1:
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(_HOST_CORES)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'style-transfer iterations/sec (VGG-19, 2048px/512-tile)'
STYLE_LAYERS = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']
CONTENT_LAYERS = ['conv4_2']

# Algorithmic work per 512x512 VGG-19 tile evaluation to conv5_1 (SURVEY 8d / DESIGN.md):
CONV_GFLOP_PER_TILE = 378.70          # forward 189.35 + backward-data 189.35
GRAM_GFLOP_PER_TILE = 18.25           # F^T F and dG x F on the five style layers


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=50)
    p.add_argument('--warmup', type=int, default=5)
    p.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    p.add_argument('--precision', default='fp16', choices=['bf16', 'fp16', 'tc32', 'fp32'])
    p.add_argument('--size', type=int, default=2048)
    p.add_argument('--tile-size', type=int, default=512)
    p.add_argument('--optimizer', default='adam', choices=['adam', 'lbfgs'])
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-e2e', action='store_true')
    p.add_argument('--no-extra', action='store_true', help='skip the tc32 and cfg4 records')
    return p.parse_args()


def workload_config(a, world):
    ntiles = ((a.size - 1) // a.tile_size + 1) ** 2
    return {
        'workload': 'cfg3: %dx%d image, %d tiles of <=%dpx, vgg19.prototxt, 5 style + 1 content '
                    'layer, tv+p-norm regularisers, %s' % (a.size, a.size, ntiles, a.tile_size,
                                                           a.optimizer),
        'size': a.size, 'tile_size': a.tile_size, 'tiles': ntiles, 'optimizer': a.optimizer,
        # both arms name the same workload: its working set per step is far larger than any cache
        'l2': 'no explicit flush: one step streams %d MB of image + optimizer state and >= 152 MB of '
              'activations per tile, both larger than the 126 MB L2' % (6 * 3 * a.size * a.size * 4 >> 20),
    }


def layout_config(a, world):
    ntiles = ((a.size - 1) // a.tile_size + 1) ** 2
    return {'tiles_per_gpu': -(-ntiles // world),
            'parallelism': 'tiles round-robin over %d GPU(s), one NCCL all-gather per evaluation' % world,
            'precision': a.precision}


def synthetic_rgb(seed, size):
    return np.random.RandomState(seed).randint(0, 256, (size, size, 3)).astype(np.uint8)


# =======================================================================================================
# CPU leg: the oracle port of the reference's algorithm (test infrastructure, timed as a baseline)
# =======================================================================================================
class CpuReference:
    """One tile-evaluation of the reference's path (forward, losses, backward incl. the dW Caffe
    computes and discards) + the full-image regularisers/Adam, on the host cores."""

    def __init__(self, a, crop=None):
        from oracle import numeric as on
        from oracle.caffe_net import he_normal_weights, model_layers
        from oracle.optimizers import Adam
        from oracle.tile_operator import OracleModel
        from oracle.transfer import to_params
        self.a, self.on = a, on
        model = 'vgg19.prototxt'
        params = he_normal_weights(model_layers(model))
        self.ora = ora = OracleModel(model, params, compute_weight_grads=True)
        # the sampled unit: one tile, or (to bound the run time) a crop x crop corner of one
        t = min(a.tile_size, a.size)
        self.crop = t = min(t, crop) if crop else t
        self.tile_px = min(a.tile_size, a.size)
        content = to_params(synthetic_rgb(1, a.size)[:t, :t])
        style = to_params(synthetic_rgb(2, a.size)[:t, :t])
        np.random.seed(0)
        self.full = to_params(np.random.uniform(0, 255, size=(a.size, a.size, 3)))
        self.tile = np.ascontiguousarray(self.full[:, :t, :t])
        # targets of the sampled tile (values do not influence the timing)
        ora.img = style
        ora.styles = [{l: on.gram_lower(f) for l, f in ora.features_once(STYLE_LAYERS).items()}]
        ora.img = content
        ora.contents = [ora.features_once(CONTENT_LAYERS)]
        ora.publish()
        self.layers = ora.ordered_layers(CONTENT_LAYERS, STYLE_LAYERS)
        self.lw = {l: 1.0 for l in ora.layers()}
        self.sw = {l: 1.0 / len(STYLE_LAYERS) for l in STYLE_LAYERS}
        self.adam = Adam(self.full.copy(), step_size=15.0, bp1=1 - 1 / 20.0, decay=0.05, power=0.5)
        self.ntiles = ((a.size - 1) // a.tile_size + 1) ** 2
        # threads the BLAS behind numpy / scipy will really use
        self.cores = _HOST_CORES
        try:
            from threadpoolctl import threadpool_info, threadpool_limits
            threadpool_limits(limits=_HOST_CORES)
            blas = [p['num_threads'] for p in threadpool_info() if p.get('user_api') == 'blas']
            if blas:
                self.cores = max(blas)
        except Exception:
            pass

    def tile_eval(self):
        t0 = time.perf_counter()
        self.ora.sc_grad_tile(self.tile, np.array([0, 0]), self.layers, CONTENT_LAYERS,
                              STYLE_LAYERS, [], self.lw, {'conv4_2': 0.05}, self.sw, {})
        return time.perf_counter() - t0

    def full_iteration(self):
        """ONE complete iteration: every tile of the image, one after the other, + the tail."""
        ts = self.tile_px
        t0 = time.perf_counter()
        for ty in range(0, self.a.size, ts):
            for tx in range(0, self.a.size, ts):
                tile = np.ascontiguousarray(self.full[:, ty:ty + ts, tx:tx + ts])
                self.ora.sc_grad_tile(tile, np.array([0, 0]), self.layers, CONTENT_LAYERS,
                                      STYLE_LAYERS, [], self.lw, {'conv4_2': 0.05}, self.sw, {})
        self.full_image_tail()
        return time.perf_counter() - t0

    def full_image_tail(self):
        """Regularisers (style_transfer.py:710-727) + Adam update (optimizers.py:26-42)."""
        on = self.on
        mean = np.float32((103.939, 116.779, 123.68)).reshape(3, 1, 1)

        def opfunc(img):
            tv_loss, tv_grad = on.tv_norm(img / np.float32(127.5), beta=2.0)
            p_loss, p_grad = on.p_norm((img + mean - np.float32(127.5)) / np.float32(127.5), p=6.0)
            return 5 * tv_loss + 2 * p_loss, np.float32(5) * tv_grad + np.float32(2) * np.float32(p_grad)
        t0 = time.perf_counter()
        self.adam.update(opfunc)
        return time.perf_counter() - t0

    def iteration_seconds(self, t_tile, t_tail):
        """Seconds per iteration from the time of one sampled unit (work is linear in pixels)."""
        return self.ntiles * (self.tile_px / self.crop) ** 2 * t_tile + t_tail

    def sample_text(self, n):
        unit = ('%dx%d tile-evaluation' % (self.crop, self.crop) if self.crop == self.tile_px else
                '%dx%d corner of one %dx%d tile-evaluation' % (self.crop, self.crop, self.tile_px,
                                                              self.tile_px))
        return ('%d x one %s (VGG-19 forward + losses + backward incl. dW, as Caffe does) timed and '
                'scaled to the %d tiles of an iteration (x%d in pixels), plus one full-image '
                'regulariser+Adam pass; numpy/OpenBLAS oracle port' %
                (n, unit, self.ntiles, round(self.ntiles * (self.tile_px / self.crop) ** 2)))


class OneDnnConvs:
    """Context manager for the "best-case CPU" stand-in of SURVEY section 8d: the oracle's three
    convolution ops (Caffe's im2col + SGEMM restated in numpy / OpenBLAS) are replaced by torch's CPU
    convolutions (oneDNN, all host cores) for the duration of a timing, so that the reported speed-up
    is not inflated by a naive im2col.  Everything else (ReLU, pooling, Gram / loss code, the dW that
    Caffe computes and discards) stays the oracle's.  Timing infrastructure only."""

    def __enter__(self):
        import torch
        import torch.nn.functional as F
        from oracle import caffe_ops as ops
        self.ops, self.saved = ops, (ops.conv3x3_forward, ops.conv3x3_backward_data,
                                     ops.conv3x3_backward_weight)
        torch.set_num_threads(_HOST_CORES)

        def fwd(x, weight, bias):
            y = F.conv2d(torch.from_numpy(x)[None], torch.from_numpy(weight), torch.from_numpy(bias),
                         padding=1)
            return np.ascontiguousarray(y[0].numpy())

        def bwd_data(top_diff, weight):
            cin = weight.shape[1]
            g = torch.nn.grad.conv2d_input((1, cin) + top_diff.shape[1:], torch.from_numpy(weight),
                                           torch.from_numpy(np.ascontiguousarray(top_diff))[None],
                                           padding=1)
            return np.ascontiguousarray(g[0].numpy())

        def bwd_weight(top_diff, x):
            cout, cin = top_diff.shape[0], x.shape[0]
            td = torch.from_numpy(np.ascontiguousarray(top_diff))[None]
            dw = torch.nn.grad.conv2d_weight(torch.from_numpy(np.ascontiguousarray(x))[None],
                                             (cout, cin, 3, 3), td, padding=1)
            return dw.numpy(), top_diff.reshape(cout, -1).sum(axis=1)

        ops.conv3x3_forward, ops.conv3x3_backward_data, ops.conv3x3_backward_weight = fwd, bwd_data, bwd_weight
        return self

    def __exit__(self, *exc):
        (self.ops.conv3x3_forward, self.ops.conv3x3_backward_data,
         self.ops.conv3x3_backward_weight) = self.saved
        return False


def run_reference(a):
    """The reference arm: the reference's CPU algorithm (oracle port, all host cores) on the engine
    arm's workload.  A whole 16-tile iteration costs ~50 s here, so every STEP is a bounded sample --
    one 512x512 tile evaluation (1/16 of an iteration's tiles) -- and ``ms_per_step`` is the measured
    duration of that sample; ``value`` scales the mean sample to an iteration (tiles are evaluated
    one after the other and independently by the reference, style_transfer.py:623-643) and adds one
    measured full-image regulariser + Adam pass.  When it fits (~75 s) ONE complete iteration -- all
    16 distinct tiles + the tail -- is timed as well and reported beside the scaled figure."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ref = CpuReference(a)
    t_probe = ref.tile_eval()
    crop = ref.crop
    while crop > 64 and (a.steps + a.warmup) * t_probe * (crop / ref.tile_px) ** 2 > 180.0:
        crop //= 2
    if crop != ref.crop:
        ref = CpuReference(a, crop=crop)
    t_tail = ref.full_image_tail()
    for _ in range(a.warmup):
        ref.tile_eval()
    times = [ref.tile_eval() for _ in range(a.steps)]
    t_tile = float(np.mean(times))
    sec = ref.iteration_seconds(t_tile, t_tail)
    value = 1.0 / sec
    full = None
    if ref.crop == ref.tile_px and ref.ntiles * t_tile + t_tail <= 75.0:
        t_full = ref.full_iteration()
        full = {'seconds': t_full, 'value': 1.0 / t_full,
                'what': 'ONE complete iteration measured: all %d distinct tiles of the image one '
                        'after the other + the regulariser / Adam pass' % ref.ntiles}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'iterations/s',
        'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': t_tile * 1e3,
        'ms_per_iteration': sec * 1e3, 'step_is': 'one tile evaluation = 1/%d of an iteration' %
        round(ref.ntiles * (ref.tile_px / ref.crop) ** 2),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(a, 1),
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': ref.cores, 'kind': 'port',
                         'sample': 'each step = ' + ref.sample_text(1)},
        'full_iteration': full,
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# =======================================================================================================
# clocks
# =======================================================================================================
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few
    milliseconds from a thread (the timed region is ~0.1-1 s), nvidia-smi -lms as a fallback."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20),
               ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(',')]
            try:
                mask = sum(bit for (name, bit), v in zip(self.REASONS, r[3:7])
                           if v.lower().startswith('active'))
                self.max_mhz = float(r[1])
                self.rows.append((time.perf_counter(), float(r[0]), mask))
            except (ValueError, IndexError):
                continue

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.nvml is None and self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no NVML / nvidia-smi']}
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        reasons = sorted({name for _, _, mask in rows for name, bit in self.REASONS if mask & bit})
        sm = [r[1] for r in rows]
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': getattr(self, 'max_mhz', None), 'reasons': reasons,
                'samples': len(sm), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# =======================================================================================================
# engine arm
# =======================================================================================================
PROFILE_TRAFFIC = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')


def measured_traffic(kernel, a, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu
    `--set full` capture of THIS workload (profiles/ncu_traffic.json, written by
    tools/ncu_traffic.py from the .ncu-rep), or None when no capture matches the configuration."""
    try:
        with open(PROFILE_TRAFFIC) as f:
            table = json.load(f)
    except (OSError, ValueError):
        return None, None
    key = '%s|size=%d|tile=%d|precision=%s|gpus=%d' % (kernel, a.size, a.tile_size, a.precision, world)
    rec = table.get(key)
    if not rec:
        return None, None
    return rec.get('dram_bytes_per_launch'), rec.get('source')


class Workload:
    """One configured run of the loop body: engine + StyleTransfer, images prepared."""

    def __init__(self, rank, world, local, size, tile, model, optimizer, precision, **flags):
        import torch
        from style_transfer_b200 import netdesc, weights
        from style_transfer_b200.engine import TileEngine
        from style_transfer_b200.transfer import StyleTransfer, default_args
        self.torch = torch
        args = default_args(size=size, min_size=size, tile_size=tile, optimizer=optimizer,
                            model=model, **flags)
        net = netdesc.from_model(args.model)
        self.eng = TileEngine(net, weights.he_normal(net), mean=args.mean, device=local,
                              precision=precision, rank=rank, world=world)
        self.eng.init_comm()
        self.st = StyleTransfer(self.eng, args)
        content = self.eng.pil_to_image(synthetic_rgb(1, size))
        style = self.eng.pil_to_image(synthetic_rgb(2, size))
        np.random.seed(args.seed)
        self.st.init_first_scale(size, size)
        self.st.prepare([content], [style])
        self.old = self.eng.img.clone()
        self.stats = torch.zeros(2, dtype=torch.float64, device=self.eng.img.device)
        self.rank = rank
        # the output step runs where the reference runs it -- on the master (rank 0) only -- and on
        # its own stream: it reads the averaged iterate of step i while step i + 1 is being computed
        self.out_stream = torch.cuda.Stream(device=self.eng.img.device)
        self.out_done = []
        self.picture = self.loss = None
        self.after_output = None          # e2e: called on the output stream after the output step

    def step(self):
        """One pass of the loop body (style_transfer.py:777-821): roll, objective, optimizer step,
        roll back; then, on rank 0, the update-size / TV statistics and the uint8 picture of the
        averaged iterate (device-resident, on the output stream)."""
        torch = self.torch
        main = torch.cuda.current_stream()
        if len(self.out_done) >= 2:
            # the buffer this step's average goes to was read by the output step two steps ago
            main.wait_event(self.out_done.pop(0))
        avg, loss = self.st.step()
        self.loss = loss
        if self.rank == 0:
            ready = torch.cuda.Event()
            ready.record(main)
            self.out_stream.wait_event(ready)
            with torch.cuda.stream(self.out_stream):
                self.picture = self.st.output_step(avg, self.old, self.stats)
                if self.after_output is not None:
                    self.after_output(loss)
                done = torch.cuda.Event()
                done.record(self.out_stream)
            self.out_done.append(done)
            if avg.data_ptr() == self.eng.img.data_ptr():
                # L-BFGS returns the parameters themselves (no averaged copy): the next step updates
                # them in place, so the output step must have read them first
                main.wait_event(done)
        return avg, loss

    def drain(self):
        self.out_stream.synchronize()

    def close(self):
        self.st = self.eng = self.old = self.picture = None
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def run_engine(a):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from style_transfer_b200 import _lib

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the engine has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except OSError:
        pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, join=None):
        """(device ms, host enqueue ms) of `steps` calls: CUDA events on the launching stream,
        barrier + synchronize on both sides, max over ranks.  `join`: a stream whose work (the output
        step of the last iterations) must be inside the timed region: the launching stream waits for
        it before the closing event."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        h1 = time.perf_counter()
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1), (h1 - h0) * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0].item()), float(ms[1].item())

    def host_enqueue_ms(fn, steps=6):
        """Host time to ENQUEUE one step (no synchronisation inside): measured over a few steps on
        an empty stream, so that the launch queue never fills up and blocks the host."""
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        return (h1 - h0) * 1e3 / steps

    def measure(w, steps, warmup):
        for _ in range(warmup):
            w.step()
        ms, _ = timed(w.step, steps, join=w.out_stream)
        return steps / (ms * 1e-3), ms / steps, host_enqueue_ms(w.step)

    def conv_roofline(w, steps, precision, sm_mhz, sm_max):
        """Per-launch CUDA events (st_timing_*) around every kernel group of `steps` steps."""
        lib.st_timing_reset()
        lib.st_timing_enable(1)
        ms_timing, _ = timed(w.step, steps, join=w.out_stream)
        lib.st_timing_enable(0)
        cats = ['conv_tc', 'conv_edge', 'pool', 'gram', 'style_grad', 'loss', 'image']
        breakdown = {}
        for i, name in enumerate(cats):
            t, wk, k = C.c_double(), C.c_double(), C.c_uint64()
            _lib.call('st_timing_read', i, C.byref(t), C.byref(wk), C.byref(k))
            breakdown[name] = {'ms_per_step': t.value / steps, 'work_per_step': wk.value / steps,
                               'launch_groups_per_step': k.value / steps}
        lib.st_timing_reset()
        dom = 'conv_tc' if breakdown['conv_tc']['ms_per_step'] > 0 else 'conv_edge'
        d = breakdown[dom]
        if d['ms_per_step'] <= 0:
            return None, breakdown
        tensor = precision in ('bf16', 'fp16', 'tc32')
        if tensor:
            # burst figure when the SM clock sat at its maximum during the timed region (a short
            # run on a cool GPU), the sustained one when the power cap had pulled it down
            at_max = sm_mhz is not None and sm_max and sm_mhz >= 0.97 * sm_max
            kind = 'burst' if at_max else 'sustained'
            peak = peaks.get('bf16_tflops' if at_max else 'bf16_tflops_sustained',
                             1650.0 if at_max else 1400.0)
            src = ('MEASURED_PEAKS.json bf16_tflops%s' % ('' if at_max else '_sustained')) if peaks \
                else 'fallback (B200_PROFILING.md)'
        else:
            kind, peak = 'nominal', 0.5 * 148 * 128 * 2 * 1.965e-3 * 2
            src = 'nominal fp32 FFMA peak (no tensor cores in fp32 mode)'
        ach = d['work_per_step'] / (d['ms_per_step'] * 1e-3) / 1e12
        kernel = 'conv_tc2_kernel' if dom == 'conv_tc' else 'conv3x3_kernel'
        traffic, tsrc = measured_traffic(kernel, a, world) if precision == a.precision else (None, None)
        roof = {'bound': 'tensor', 'kernel': kernel, 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': ach / peak, 'peak_kind': kind, 'peak_source': src,
                'sm_mhz_during_run': sm_mhz, 'traffic': traffic, 'traffic_source': tsrc,
                'flops_per_launch': d['work_per_step'] / max(d['launch_groups_per_step'], 1),
                'avg_launch_ms': d['ms_per_step'] / max(d['launch_groups_per_step'], 1),
                'share_of_step': d['ms_per_step'] / (ms_timing / steps),
                'work': 'ALGORITHMIC flops: 2*9*Cin*Cout*H*W per conv launch, forward + backward-data'
                        + ('; the split-operand mode executes 3 MMAs per product, i.e. 3x these flops'
                           ' on the tensor pipe' if precision == 'tc32' else '')}
        for key in ('bf16_tflops', 'bf16_tflops_sustained'):
            if tensor and peaks.get(key):
                roof['frac_of_' + key] = ach / peaks[key]
        return roof, breakdown

    # ---- the headline workload: cfg3 in the fast tensor-core mode -----------------------------------
    w = Workload(rank, world, local, a.size, a.tile_size, 'vgg19.prototxt', a.optimizer, a.precision)
    eng, st = w.eng, w.st
    n = eng.img.numel()
    for _ in range(a.warmup):
        w.step()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.st_launch_count()
    t0 = time.perf_counter()
    ms, host_ms = timed(w.step, a.steps, join=w.out_stream)
    t1 = time.perf_counter()
    launches = lib.st_launch_count() - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    lt = torch.tensor([launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    launches = int(lt.item())
    value = a.steps / (ms * 1e-3)
    host_ms = host_enqueue_ms(w.step) * a.steps
    sm_mhz = clocks and clocks.get('sm_mhz')
    sm_max = clocks and clocks.get('sm_max_mhz')
    if world > 1:                      # every rank needs the same peak choice
        cl = torch.tensor([sm_mhz or 0.0, sm_max or 0.0], dtype=torch.float64, device=dev)
        dist.broadcast(cl, src=0)
        sm_mhz, sm_max = float(cl[0].item()) or None, float(cl[1].item()) or None

    # ---- end to end: host buffers in, host buffers out, every step ---------------------------------
    e2e = None
    if not a.no_e2e:
        H, W = eng.img.shape[-2:]
        rows = H // world if H % world == 0 else None
        host_params = torch.empty(eng.img.shape, dtype=torch.float32).pin_memory()
        host_params.copy_(eng.img)
        host_pic = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        host_scalars = torch.empty(3, dtype=torch.float64).pin_memory()
        scal = torch.zeros(3, dtype=torch.float64, device=dev)
        pending = []

        def fetch_results(loss):
            # (runs on the output stream, after the output step) D2H: picture, loss, statistics
            host_pic.copy_(w.picture, non_blocking=True)
            scal[0:1].copy_(loss)
            scal[1:3].copy_(w.stats)
            host_scalars.copy_(scal, non_blocking=True)
        if rank == 0:
            w.after_output = fetch_results

        def e2e_step():
            # before it enqueues step k the host consumes the results of step k - 2 (complete in
            # pinned memory once that step's output stream is done): the device always has the next
            # step queued, every step's results are read, at most two steps are in flight
            if len(pending) >= 2:
                pending.pop(0).synchronize()
            # H2D: this step's image (every rank its 1/world slab of rows through its own PCIe link,
            # NVLink all-gather; one GPU: two halves, the second streaming in behind the first tiles)
            eng.stage_host_image(host_params)
            w.step()
            if rank == 0:
                pending.append(w.out_done[-1])
            else:
                ev = torch.cuda.Event()
                ev.record()
                pending.append(ev)
        for _ in range(3):
            e2e_step()
        ms_e2e, _ = timed(e2e_step, a.steps, join=w.out_stream)
        w.drain()
        w.after_output = None
        slab = n * 4 // world if rows is not None else n * 4
        e2e = {'value': a.steps / (ms_e2e * 1e-3), 'unit': 'iterations/s',
               'h2d_bytes_per_step': n * 4, 'd2h_bytes_per_step': H * W * 3 + 24,
               'h2d_bytes_per_step_per_gpu': slab,
               'd2h_bytes_per_step_per_gpu': 'rank 0: %d (picture + loss + statistics), others: 0' % (H * W * 3 + 24),
               'ms_per_step': ms_e2e / a.steps, 'fraction_of_value': a.steps / (ms_e2e * 1e-3) / value,
               'boundary': 'pinned host f32[3,H,W] image in (TileEngine.stage_host_image), one pass of '
                           'the loop body (StyleTransfer.step + statistics + picture), uint8 RGB '
                           'picture + loss + update-size / TV statistics back to pinned host memory, '
                           'every step; the host reads step k - 2\'s results before it enqueues step k (at most '
                           'two steps in flight; the timed region ends with a device-wide synchronisation); '
                           'N > 1: every rank uploads its 1/N slab of rows through its own PCIe link, '
                           'the slabs are all-gathered over NCCL, results leave from rank 0 (the master)'}

    if os.environ.get('ST_NCU_RANGE') == '1':
        # for `ncu --profile-from-start off`: exactly one step of the headline workload is profiled
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        w.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    roofline, breakdown = conv_roofline(w, min(a.steps, 3), a.precision, sm_mhz, sm_max)

    # Gram kernels against the HBM roofline: algorithmic bytes = every style layer's 16-bit feature
    # map read once (64.9 MB per 512x512 VGG-19 tile) over gram_tc_kernel + finalize time
    roofline_gram = None
    if breakdown['gram']['ms_per_step'] > 0 and a.tile_size == 512 and a.precision in ('fp16', 'bf16'):
        tiles_here = -(-(((a.size - 1) // a.tile_size + 1) ** 2) // world)
        gbytes = 64.9e6 * tiles_here
        ach = gbytes / (breakdown['gram']['ms_per_step'] * 1e-3) / 1e9
        hbm = peaks.get('hbm_gbs', 6650.0)
        roofline_gram = {'bound': 'hbm', 'kernel': 'gram_tc_kernel + gram_tc_finish_kernel + delta_pack_kernel',
                         'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                         'traffic': None, 'algorithmic_bytes_per_step': gbytes,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 (recipe)'}
    w.close()

    # ---- further records: the reference-precision tensor-core mode and BASELINE config 4 -------------
    records = []
    if not a.no_extra:
        k, wu = min(a.steps, 6), 3
        try:
            w2 = Workload(rank, world, local, a.size, a.tile_size, 'vgg19.prototxt', a.optimizer, 'tc32')
            v2, ms2, h2 = measure(w2, k, wu)
            roof2, bd2 = conv_roofline(w2, 2, 'tc32', sm_mhz, sm_max)
            w2.close()
            records.append({
                'record': 'cfg3 in the tc32 mode (fp32 storage, split fp16 hi+lo tensor-core operands, '
                          'chained fp32 accumulation): the reference-precision tensor-core path',
                'value': v2, 'unit': 'iterations/s', 'ms_per_step': ms2, 'steps': k, 'warmup': wu,
                'dtype': 'f32 storage / split f16 operands / f32 accumulate', 'roofline': roof2,
                'breakdown_ms': {kk: vv['ms_per_step'] for kk, vv in bd2.items()},
                'tolerance': 'tests/test_gpu_parity_configs.py: loss 1e-4, gradient relative L2 <= 3e-3 '
                             'vs the CPU oracle at 256^2 .. 2048^2 (measured 3e-4 .. 1.3e-3, the fp32 '
                             'SIMT mode measures 1e-4 .. 1.0e-3)'})
        except Exception as e:                                   # a reporting extra
            records.append({'record': 'cfg3 tc32', 'unavailable': repr(e)[:300]})
        try:
            w4 = Workload(rank, world, local, 4096, 1024, 'vgg19_avgpool.prototxt', 'lbfgs', a.precision,
                          tv_weight=5.0)
            v4, ms4, h4 = measure(w4, k, wu)
            w4.close()
            records.append({
                'record': 'cfg4: 4096x4096 image, 16 tiles of 1024px, vgg19_avgpool.prototxt, 5 style + 1 '
                          'content layer, L-BFGS (device-resident memory), tv-weight 5',
                'value': v4, 'unit': 'iterations/s', 'ms_per_step': ms4, 'host_enqueue_ms_per_step': h4,
                'steps': k, 'warmup': wu, 'precision': a.precision})
        except Exception as e:
            records.append({'record': 'cfg4', 'unavailable': repr(e)[:300]})

    # ---- CPU baseline (rank 0, N = 1) ---------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        ref = CpuReference(a)
        t_tail = ref.full_image_tail()
        ref.tile_eval()
        t_tile = float(np.mean([ref.tile_eval() for _ in range(2)]))
        cpu = {'value': 1.0 / ref.iteration_seconds(t_tile, t_tail), 'unit': 'iterations/s',
               'cores': ref.cores, 'kind': 'port', 'sample': ref.sample_text(2),
               'tile_eval_s': t_tile, 'full_image_tail_s': t_tail}
        # "best-case CPU" beside it: the same evaluation with oneDNN convolutions (torch CPU)
        try:
            with OneDnnConvs():
                ref.tile_eval()
                t_best = float(np.mean([ref.tile_eval() for _ in range(2)]))
            cpu['best_case'] = {'value': 1.0 / ref.iteration_seconds(t_best, t_tail),
                                'unit': 'iterations/s', 'tile_eval_s': t_best,
                                'what': 'same sample with the convolutions (forward, backward-data, '
                                        'dW) on torch CPU / oneDNN instead of im2col + OpenBLAS SGEMM'}
        except Exception as e:                       # a reporting extra: never fail the bench line
            cpu['best_case'] = {'unavailable': repr(e)[:200]}

    if rank == 0:
        cfg = workload_config(a, world)
        layout = layout_config(a, world)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'iterations/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': {'bf16': 'bf16', 'fp16': 'f16 forward / bf16 backward', 'fp32': 'f32',
                      'tc32': 'f32 storage, split f16 hi+lo tensor-core operands, f32 accumulate'}[a.precision],
            'data': 'synthetic', 'config': cfg, 'layout': layout,
            'step_is': 'one pass of the loop body (style_transfer.py:777-821): roll, 16-tile objective, '
                       'regularisers, optimizer step, update-size / TV statistics, uint8 picture',
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'host_enqueue_ms_per_step': host_ms / a.steps, 'roofline': roofline,
            'roofline_gram': roofline_gram, 'records': records,
            'cpu_baseline': cpu, 'breakdown': breakdown,
            'tile_eval_ms': breakdown and sum(v['ms_per_step'] for k, v in breakdown.items()
                                              if k != 'image') / layout['tiles_per_gpu'],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_engine(a)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""bench.py - RGB-D fused frames/s at 768x384 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): two-stream VGG16-FCN (rgb 3-ch + depth 1-ch, num_units
64, 12 classes) inference + confusion-matrix Bayes fusion + confusion-matrix accumulation of
score(), batch 16 per GPU, synthetic inputs, random-init weights of the named architecture.
A "step" = one such batch through `BayesFusion` (experts -> fusion -> score accumulation).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo's CUDA path
    python bench.py --impl reference [...]                        # CPU restatement of the
                                                                   # reference on the host cores
For N > 1 launch through torchrun (one rank per GPU); images are sharded over ranks with no
data-path collective ("weak" scaling: 16 frames per GPU per step), the only collective being
the int64 confusion-matrix all-reduce at the end of score().

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value     device-resident throughput: inputs already in HBM when the timed region starts;
            the timed K steps follow W warm-up steps and an untimed soak of the same step
            (`soak_s`, >= 2 s) so that clocks and power have settled
  e2e       the same metric through the public API `net.score(host arrays)`: pinned host
            buffers, H2D of rgb/depth/labels every step, D2H of the confusion matrix;
            `e2e.dataset_call` = ONE net.score() over K x 16 frames with batchsize 16 (the
            reference's API shape, base_model.py:294-313)
  roofline  tensor-core stack (all tcgen05 conv launches): algorithmic FLOPs / CUDA-event time
            measured live in the timed region, against both MEASURED_PEAKS.json figures
            (`frac` = sustained, `frac_burst` = burst)
  fusion_hbm  HBM fraction of every per-pixel fusion / score kernel (algorithmic bytes /
            CUDA-event time / measured copy bandwidth), batch-16 sizes
  dirichlet_mc_T20  BASELINE configs[2]: Dirichlet fusion of T = 20 MC-dropout samples per
            modality, frames/s (device-resident)
  fit       BASELINE configs[4]: data-parallel fit() of the depth stream, frames/s
  sharded_check  N > 1: the same global batch through score() with images sharded over the
            ranks equals rank 0's single-rank confusion matrix
  cpu_baseline  the oracle (CPU restatement of the reference; TensorFlow is not installable)
            timed on this box's host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, C, NU, BATCH = 768, 384, 12, 64, 16
METRIC = 'rgbd_fused_frames_per_s_768x384'
WORKLOAD = ('two-stream VGG16-FCN (rgb+depth, num_units=64, 12 classes) + confusion-matrix '
            'Bayes fusion + score() accumulation, 768x384, batch 16 per GPU')
# algorithmic tensor-core FLOPs per fused frame: 181.23 + 180.55 GFLOP
GFLOP_PER_FRAME = 361.78


def _data_description():
    return ({'rgb': np.float32, 'depth': np.float32, 'labels': np.int32},
            {'rgb': (None, None, 3), 'depth': (None, None, 1), 'labels': (None, None)}, C)


def _confusion_matrices(rng):
    """Measure-set confusion matrices of sacred run 868 if the golden fixture is present,
    otherwise Dirichlet(1)-sampled rows + 1."""
    path = os.path.join(ROOT, 'tests', 'golden', 'exp868.npz')
    if os.path.exists(path):
        g = np.load(path)
        return {'rgb': g['cm_measure_rgb'], 'depth': g['cm_measure_depth']}
    return {m: np.floor(rng.dirichlet(np.ones(C), size=C) * 1e6) + 1 for m in ('rgb', 'depth')}


def _peaks():
    """(sustained bf16 TFLOP/s, burst bf16 TFLOP/s, HBM GB/s, 'measured' | 'fallback')."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return (p.get('bf16_tflops_sustained', 1400.0), p.get('bf16_tflops', 1590.0),
                p.get('hbm_gbs', 6650.0), 'measured')
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler(object):
    """SM clock + throttle reasons sampled during the timed region: NVML polled every 10 ms from
    a thread (a 20-step timed region lasts ~0.1 s); nvidia-smi -lms when NVML is unavailable, one
    nvidia-smi query right after the region when NVML delivered nothing inside it."""

    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40),
               ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []            # (time, sm_mhz, sm_max_mhz, [reasons])
        self.smi_rows = []
        self.proc = None
        self._stop = False
        self.source = None
        self._index = index
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if visible:
                ids = [v for v in visible.split(',') if v.strip() != '']
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self.source = 'nvml'
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.source = None
        # nvidia-smi -lms only when NVML is unavailable (a second poller would add driver queries
        # to the timed region for nothing); if NVML is there but delivers no sample in the window
        # - seen on one box - stop() falls back to one query right after the region
        self.smi_rows = []
        if self.source is None:
            try:
                self.proc = subprocess.Popen(
                    ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.QUERY,
                     '--format=csv,noheader,nounits', '-lms', '50'],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.smi_thread = threading.Thread(target=self._read, daemon=True)
                self.smi_thread.start()
                self.source = 'nvidia-smi'
            except OSError:
                self.proc = None
        self._index = index

    def _poll(self):
        nv = self._nvml
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                reasons = ['reasons_unavailable']
                for query in ('nvmlDeviceGetCurrentClocksEventReasons',
                              'nvmlDeviceGetCurrentClocksThrottleReasons'):
                    try:
                        mask = int(getattr(nv, query)(self._handle))
                        reasons = [name for name, bit in self.REASONS if mask & bit]
                        break
                    except Exception:
                        continue
                self.rows.append((time.time(), mhz, self._max, reasons))
            except Exception as err:
                self.last_error = repr(err)[:120]
            time.sleep(0.01)

    def _read(self):
        names = [n for n, _ in self.REASONS]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            try:
                self.smi_rows.append((time.time(), float(parts[0]), float(parts[1]),
                                      [n for n, f in zip(names, parts[3:7])
                                       if f.lower().startswith('active')]))
            except (ValueError, IndexError):
                continue

    def stop(self, t0, t1):
        if self.source is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no clock source'],
                    'samples': 0}
        time.sleep(0.06)
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()

        def window(rows):
            inside = [r for r in rows if t0 <= r[0] <= t1]
            if len(inside) < 5:   # widen slightly rather than report too few samples
                inside = [r for r in rows if t0 - 0.05 <= r[0] <= t1 + 0.05]
            return inside
        source = self.source
        inside = window(self.rows)
        if not inside and self.smi_rows:
            inside, source = window(self.smi_rows), 'nvidia-smi'
        if not inside:
            # last resort: one synchronous query right after the region (says so in `source`)
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self._index),
                                      '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=10).stdout
                parts = [p.strip() for p in out.strip().splitlines()[0].split(',')]
                names = [n for n, _ in self.REASONS]
                inside = [(t1, float(parts[0]), float(parts[1]),
                           [n for n, f in zip(names, parts[3:7]) if f.lower().startswith('active')])]
                source = 'nvidia-smi, one query right after the timed region'
            except Exception:
                return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples'],
                        'samples': 0, 'source': self.source,
                        'error': getattr(self, 'last_error', None)}
        reasons = sorted({name for r in inside for name in r[3]})
        return {'sm_mhz': float(np.median([r[1] for r in inside])),
                'sm_min_mhz': float(min(r[1] for r in inside)),
                'sm_max_mhz': float(max(r[2] for r in inside)),
                'reasons': reasons, 'samples': len(inside), 'source': source}


# ------------------------------------------------------------------------------- CPU arm
def _oracle_frame_fn(rng):
    """One fused RGB-D frame through the oracle (CPU restatement of the reference graph:
    experiments/timing.py:49-83 + base_model.py:140-151)."""
    import torch
    import oracle
    torch.set_num_threads(os.cpu_count())
    cms = _confusion_matrices(rng)
    tables = [cms[m].astype('float32').T for m in ('rgb', 'depth')]
    params = {}
    for m, cin in (('rgb', 3), ('depth', 1)):
        params.update(oracle.glorot_fcn_params(m, cin, NU, C, rng))
    rgb = rng.integers(0, 256, size=(1, H, W, 3)).astype(np.float32)
    depth = rng.integers(0, 65536, size=(1, H, W, 1)).astype(np.float32)
    labels = rng.integers(-1, C, size=(1, H, W)).astype(np.int32)

    def frame():
        cls = [oracle.test_pipeline(x, params, m, NU, C)['classification']
               for m, x in (('rgb', rgb), ('depth', depth))]
        fused = oracle.argmax_first(oracle.bayes_fusion(cls, tables, 'data')[0])
        return oracle.confusion_matrix(labels, fused, C)
    return frame, torch.get_num_threads()


def cpu_sweep_commands(num_units, num_classes, height, width, cms, dirichlet_params):
    """CPU leg of the batch-1 latency sweep (tools/timing.py --cpu): the oracle restatement of
    the commands of experiments/timing.py on the host cores.  Lives here because bench.py is the
    only non-test module that may execute oracle/."""
    import torch
    import oracle
    torch.set_num_threads(os.cpu_count())
    rng = np.random.default_rng(0)
    params = {}
    for m, cin in (('rgb', 3), ('depth', 1)):
        params.update(oracle.glorot_fcn_params(m, cin, num_units, num_classes, rng))
    rgb = np.ones((1, height, width, 3), np.float32)
    depth = np.ones((1, height, width, 1), np.float32)

    def expert(x, m):
        return oracle.test_pipeline(x, params, m, num_units, num_classes)

    def rgb_fcn():
        return expert(rgb, 'rgb')['classification']

    def bayes_fcn():
        cls = [expert(rgb, 'rgb')['classification'], expert(depth, 'depth')['classification']]
        return oracle.argmax_first(oracle.bayes_fusion(cls, cms)[0])

    def dirichlet_fcn():
        p = [expert(rgb, 'rgb')['prob'], expert(depth, 'depth')['prob']]
        return oracle.argmax_first(oracle.dirichlet_fusion(
            p, [dirichlet_params['rgb'], dirichlet_params['depth']],
            oracle.dirichlet_prior(dirichlet_params['class_counts'])))

    return {'rgb_fcn': rgb_fcn, 'bayes_fcn': bayes_fcn, 'dirichlet_fcn': dirichlet_fcn}


def cpu_baseline(rng, frames=2):
    frame, threads = _oracle_frame_fn(rng)
    frame()                                    # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    for _ in range(frames):
        frame()
    dt = time.perf_counter() - t0
    return {'value': frames / dt, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
            'sample': '%d fused 768x384 RGB-D frames (batch 1) through oracle/ (torch-CPU fp32 '
                      'restatement of the reference graph; TensorFlow 1.x is not installable), '
                      '%.2f s' % (frames, dt)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rng = np.random.default_rng(0)
    frame, threads = _oracle_frame_fn(rng)
    for _ in range(max(args.warmup, 1)):
        frame()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = ('each step = 1 fused 768x384 RGB-D frame (batch 1) through oracle/ on %d host '
              'threads; TensorFlow 1.x (the reference runtime) is not installable here' % threads)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': max(args.warmup, 1),
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}}))


# ------------------------------------------------------------------------------- GPU arm
def _event_ms(fn, iters, warm=3):
    """Mean CUDA-event milliseconds of fn() on the current stream after `warm` warm-up calls."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fusion_hbm_rows(peak_gbs):
    """HBM roofline of the per-pixel fusion / score kernels at the batch-16 768x384 size.
    Algorithmic bytes per pixel: SURVEY.md 8(d) / DESIGN.md 4.3; inputs (>= 226 MB per expert)
    are far larger than the 126 MB L2; CUDA events on the launching stream."""
    import torch
    from modular_semantic_segmentation_b200 import device as dev
    n, m = BATCH, 2
    npix = n * H * W
    g = torch.Generator(device='cuda').manual_seed(0)
    probs = [torch.softmax(2 * torch.randn((n, H, W, C), device='cuda', generator=g), -1)
             .contiguous() for _ in range(m)]
    score = torch.randn((n, H, W, C), device='cuda', generator=g)
    l64 = [torch.randint(0, C, (n, H, W), device='cuda', generator=g) for _ in range(m)]
    l8 = [t.to(torch.uint8) for t in l64]
    # segmentation-like maps: 32x32-pixel blocks of one class, prediction agrees on 90 %
    blocks = torch.randint(-1, C, (n, H // 32, W // 32), device='cuda', generator=g,
                           dtype=torch.int32)
    gt = blocks.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    gt_rand = torch.randint(-1, C, (n, H, W), device='cuda', generator=g, dtype=torch.int32)
    pred = torch.where(torch.rand((n, H, W), device='cuda', generator=g) < 0.9,
                       gt.clamp(min=0).to(torch.int64), l64[0])
    pred8 = pred.to(torch.uint8)
    lut = torch.randint(0, C, (C, C), device='cuda', generator=g, dtype=torch.int32)
    am1 = torch.rand((m, C, C), device='cuda', generator=g) * 3
    lognorm = torch.rand((m, C), device='cuda', generator=g)
    logprior = torch.log(torch.full((C,), 1.0 / C, device='cuda'))
    mags = dev.dirichlet_table_magnitudes(am1, lognorm, logprior)
    var = [torch.rand((n, H, W), device='cuda', generator=g) * 1e-2 for _ in range(m)]
    cm = torch.zeros((C, C), dtype=torch.int64, device='cuda')
    stats = torch.zeros((C, C), dtype=torch.float64, device='cuda')
    cnt = torch.zeros(C, dtype=torch.int64, device='cuda')
    cases = [
        ('softmax_argmax prob+i64', 8 * C + 8, lambda: dev.softmax_argmax(score)),
        ('softmax_argmax u8', 4 * C + 1,
         lambda: dev.softmax_argmax(score, want_prob=False, label_dtype=torch.uint8)),
        ('bayes_fuse_lut i64', 8 * m + 8, lambda: dev.bayes_fuse_lut(l64, lut, C)),
        ('bayes_fuse_lut u8', m + 1, lambda: dev.bayes_fuse_lut(l8, lut, C)),
        ('dirichlet_fuse fast u8', 4 * C * m + 1,
         lambda: dev.dirichlet_fuse(probs, am1, lognorm, logprior, label_dtype=torch.uint8)),
        ('dirichlet_fuse exact u8', 4 * C * m + 1,
         lambda: dev.dirichlet_fuse(probs, am1, lognorm, logprior, label_dtype=torch.uint8,
                                    exact=True, magnitudes=mags)),
        ('average_fuse i64', 4 * C * m + 8, lambda: dev.average_fuse(probs)),
        ('variance_fuse i64', 4 * C * m + 4 * m + 8, lambda: dev.variance_fuse(probs, var)),
        ('confusion i64 blocky', 12, lambda: dev.confusion_accumulate(pred, gt, cm)),
        ('confusion u8 blocky', 5, lambda: dev.confusion_accumulate(pred8, gt, cm)),
        ('confusion i64 random', 12, lambda: dev.confusion_accumulate(l64[0], gt_rand, cm)),
        ('suffstats blocky', 4 * C + 4, lambda: dev.dirichlet_suffstats(probs[0], gt, stats, cnt)),
        ('suffstats random', 4 * C + 4,
         lambda: dev.dirichlet_suffstats(probs[0], gt_rand, stats, cnt)),
    ]
    rows = {}
    for name, bytes_px, fn in cases:
        ms = _event_ms(fn, 10)
        gbs = bytes_px * npix / ms / 1e6
        rows[name] = {'bytes_per_pixel': bytes_px, 'ms': round(ms, 4), 'gbs': round(gbs, 1),
                      'frac': round(gbs / peak_gbs, 3)}
    return rows


def dirichlet_mc_bench(world, rank, steps=3, batch=8, num_samples=20):
    """BASELINE configs[2]: per modality T = 20 MC-dropout samples (dropout after pool3, all
    samples sharing one weight load) -> mean softmax -> Dirichlet fusion (exact argmax) ->
    confusion matrix; images sharded over the ranks (`batch` frames per GPU and step)."""
    import torch
    import torch.distributed as dist
    from xview.models import get_model
    rng = np.random.default_rng(77)
    params = {m: 1.0 + rng.gamma(2.0, 2.0, size=(C, C)) + 4.0 * np.eye(C) for m in ('rgb', 'depth')}
    params['class_counts'] = rng.integers(1000, 100000, size=C).astype(np.float64)
    net = get_model('dirichlet_mix')(
        data_description=_data_description(), modalities=['rgb', 'depth'], expert_model='fcn',
        num_units=NU, num_channels={'rgb': 3, 'depth': 1}, batchsize=batch, seed=7,
        dirichlet_params=params, num_samples=num_samples, dropout_rate=0.5,
        dropout_layers=['pool3'], shard_images=False)
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    sets = [{'rgb': torch.randint(0, 256, (batch, H, W, 3), device='cuda', generator=g).float(),
             'depth': torch.randint(0, 65536, (batch, H, W, 1), device='cuda', generator=g).float(),
             'labels': torch.randint(-1, C, (batch, H, W), device='cuda', generator=g,
                                     dtype=torch.int32)} for _ in range(2)]
    cm = torch.zeros((C, C), dtype=torch.int64, device='cuda')
    it = [0]

    def step():
        net.score_batch_on_device(sets[it[0] % 2], cm)
        it[0] += 1

    ms = _event_ms(step, steps, warm=2)
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    labelled = sum(int((s['labels'] >= 0).sum()) for s in sets)
    net.close()
    del sets
    torch.cuda.empty_cache()
    # T passes of conv4_1..conv5_3 + heads and one pass of the trunk, per modality (App. D)
    gflop = 2 * (109.7 + num_samples * 71.5)
    return {'value': world * batch / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms,
            'frames_per_gpu_per_step': batch, 'num_samples': num_samples, 'steps': steps,
            'tflops': gflop * batch / ms, 'fusion': 'dirichlet, exact argmax',
            'labelled_pixels_checksum_ok': bool(int(cm.sum()) > 0 and labelled > 0)}


def fit_bench(world, rank, steps=5, batch=16):
    """BASELINE configs[4]: data-parallel fit() of the VGG16-FCN depth stream on synthetic
    Cityscapes-shaped data (384x768, 12 classes), Adam lr 1e-4, `batch` frames per GPU: forward +
    backward + gradient all-reduce over NCCL + Adam, through SimpleFCN's training step."""
    import torch
    import torch.distributed as dist
    from xview.models import get_model
    desc = ({'depth': np.float32, 'labels': np.int32},
            {'depth': (None, None, 1), 'labels': (None, None)}, C)
    net = get_model('fcn')('depth', desc, 'depth', num_units=NU, batch_normalization=False,
                           learning_rate=1e-4, batchsize=batch, seed=1, shard_images=False)
    g = torch.Generator(device='cuda').manual_seed(rank)
    batch_dev = {'depth': torch.rand((batch, W, H, 1), device='cuda', generator=g) * 100.0,
                 'labels': torch.randint(-1, C, (batch, W, H), device='cuda', generator=g,
                                         dtype=torch.int32)}
    trainer = net.make_trainer()
    ms = _event_ms(lambda: trainer.step(batch_dev), steps, warm=2)
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss = trainer.last_loss()
    num_params = net._experts['depth'].num_params
    trainer.close()
    net.close()
    torch.cuda.empty_cache()
    return {'value': world * batch / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms,
            'frames_per_gpu_per_step': batch, 'steps': steps, 'trainer': 'adam',
            'allreduce_bytes_per_step': num_params * 4 + 16, 'final_loss': loss}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from modular_semantic_segmentation_b200 import device as dev
    from xview.models import get_model

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    _bind_to_local_numa_node(local_rank)
    warmup = max(args.warmup, 3)
    # sharding inside score() is by image over ranks; the bench gives every rank its own
    # batch-16 shard directly (weak scaling), so the model must not split it again
    rng = np.random.default_rng(1234 + rank)
    cms = _confusion_matrices(np.random.default_rng(0))

    def make_net(shard):
        return get_model('bayes_fusion')(
            confusion_matrices=cms, data_description=_data_description(),
            prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
            num_channels={'rgb': 3, 'depth': 1}, batchsize=BATCH, seed=7, class_prior='data',
            shard_images=shard)

    net = make_net(False)
    n_sets = 3    # rotate input sets: 3 x 94 MB of inputs + GBs of activations >> 126 MB L2
    host_sets = []
    for _ in range(n_sets):
        blob = {'rgb': torch.from_numpy(rng.integers(0, 256, size=(BATCH, H, W, 3))
                                        .astype(np.float32)).pin_memory(),
                'depth': torch.from_numpy(rng.integers(0, 65536, size=(BATCH, H, W, 1))
                                          .astype(np.float32)).pin_memory(),
                'labels': torch.from_numpy(rng.integers(-1, C, size=(BATCH, H, W))
                                           .astype(np.int32)).pin_memory()}
        host_sets.append(blob)
    dev_sets = [{k: v.cuda() for k, v in blob.items()} for blob in host_sets]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_sets[0].values())
    labelled = [int((b['labels'] >= 0).sum()) for b in host_sets]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device='cuda', dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sum_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device='cuda', dtype=torch.int64)
            dist.all_reduce(t)
            return int(t.item())
        return x

    cm = torch.zeros((C, C), dtype=torch.int64, device='cuda')

    def device_step(i):
        net.score_batch_on_device(dev_sets[i % n_sets], cm)

    # ---------------- device-resident throughput (value) + live roofline of the conv stack
    for i in range(warmup):
        device_step(i)
    barrier()
    # the clock sampler runs from the soak on: the soak is the same step under the same load, so
    # its samples (hundreds) back the few that fall into the ~0.1 s timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # untimed soak of the same step: clocks / power settle before anything is timed
    t_soak = time.time()
    i = 0
    while time.time() - t_soak < args.soak:
        for _ in range(10):
            device_step(i)
            i += 1
        torch.cuda.synchronize()
    soak_s = time.time() - t_soak
    barrier()
    cm.zero_()
    launches0 = dev.launch_count()
    dev.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.time()
    e0.record()
    for i in range(args.steps):
        device_step(i)
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    conv_ms, conv_flops, conv_launches = dev.profile_read()
    dev.profile_enable(False)
    launches = dev.launch_count() - launches0
    clocks = sampler.stop(t_start, t_end) if sampler else None
    if clocks is not None and sampler.source is not None:
        during_soak = [r for r in (sampler.rows or sampler.smi_rows)
                       if t_soak + 0.5 * soak_s <= r[0] < t_start]
        if during_soak:
            clocks['soak'] = {'sm_mhz': float(np.median([r[1] for r in during_soak])),
                              'samples': len(during_soak),
                              'reasons': sorted({n for r in during_soak for n in r[3]}),
                              'window': 'second half of the untimed soak (same step, same load)'}
    ms = max_over_ranks(ms)
    value = world * BATCH * args.steps / (ms * 1e-3)
    # checksum of the device-resident loop: every labelled pixel was counted exactly once
    want_px = sum(labelled[i % n_sets] for i in range(args.steps))
    assert int(cm.sum().item()) == want_px, (int(cm.sum().item()), want_px)

    # ---------------- the same K steps replayed from captured CUDA graphs (one graph per input
    # set): launch gaps between the ~33 kernels of a step disappear.  Reported beside `value`,
    # which stays the eager loop the per-kernel roofline events belong to.
    graph_info = None
    try:
        graphs = [net.capture_score_step(dev_sets[i], cm) for i in range(n_sets)]
        for i in range(warmup):
            graphs[i % n_sets][0].replay()
        barrier()
        cm.zero_()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(args.steps):
            graphs[i % n_sets][0].replay()
        g1.record()
        barrier()
        graph_ms = max_over_ranks(g0.elapsed_time(g1))
        assert int(cm.sum().item()) == want_px
        graph_info = {'value': world * BATCH * args.steps / (graph_ms * 1e-3),
                      'ms_per_step': graph_ms / args.steps,
                      'kernels_per_graph': int(graphs[0][1])}
        del graphs
    except Exception as err:                      # capture is an optimisation, never a requirement
        graph_info = {'unavailable': str(err)[:200]}
        torch.cuda.synchronize()

    # ---------------- end to end through the public API with host buffers (e2e)
    # each rank scores its own batch-16 shard (shard_images=False: the dict a rank passes IS
    # its shard); score() still all-reduces the confusion matrix over ranks at its end
    for i in range(2):
        net.score(host_sets[i % n_sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        measures, cm_host = net.score(host_sets[i % n_sets])
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * BATCH * args.steps / e2e_s
    # the all-reduced matrix counts the labelled pixels of every rank's last batch
    assert int(cm_host.sum()) == sum_over_ranks(labelled[(args.steps - 1) % n_sets])

    # one score() call over the whole data set (K x 16 frames, batchsize 16): the reference's
    # API shape (base_model.py:294-313); uploads of batch i+1 overlap the kernels of batch i
    def dataset(count=args.steps):
        for i in range(count):
            yield host_sets[i % n_sets]
    net.score(dataset(2 * n_sets))            # untimed: sizes the allocator for run-ahead uploads
    barrier()
    t0 = time.perf_counter()
    _, cm_ds = net.score(dataset())
    barrier()
    ds_s = max_over_ranks(time.perf_counter() - t0)
    assert int(cm_ds.sum()) == sum_over_ranks(want_px)

    # same calls with the sensors' own dtypes (uint8 rgb, uint16 depth): the float32 cast of
    # data_baseclass.py:77-78 then runs on the device after a 2.2x smaller copy
    raw_sets = [{'rgb': b['rgb'].to(torch.uint8).pin_memory(),
                 'depth': b['depth'].to(torch.uint16).pin_memory(),
                 'labels': b['labels']} for b in host_sets]
    raw_bytes = sum(v.numel() * v.element_size() for v in raw_sets[0].values())
    for i in range(2):
        net.score(raw_sets[i % n_sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        _, cm_raw = net.score(raw_sets[i % n_sets])
    barrier()
    raw_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(cm_raw, cm_host)

    def raw_dataset(count=args.steps):
        for i in range(count):
            yield raw_sets[i % n_sets]
    net.score(raw_dataset(2 * n_sets))
    barrier()
    t0 = time.perf_counter()
    _, cm_raw_ds = net.score(raw_dataset())
    barrier()
    raw_ds_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(cm_raw_ds, cm_ds)

    # ---------------- N > 1: the real sharded path, checked against a single-rank evaluation
    sharded_check = None
    if world > 1:
        # every rank passes the SAME global batch (rank 0's set 0); score() shards its images
        # over the ranks (image i -> rank i mod world) and all-reduces the matrix
        blob = [host_sets[0]] if rank == 0 else [None]
        dist.broadcast_object_list(blob, src=0)
        shared = blob[0]
        sharded = make_net(True)
        _, cm_sharded = sharded.score(shared)
        sharded.close()
        single = None
        if rank == 0:
            cm1 = torch.zeros((C, C), dtype=torch.int64, device='cuda')
            net.score_batch_on_device({k: v.cuda() for k, v in shared.items()}, cm1)
            single = cm1.cpu().numpy().astype(np.float64)
        ok = [bool(np.array_equal(cm_sharded, single))] if rank == 0 else [None]
        dist.broadcast_object_list(ok, src=0)
        assert ok[0], 'sharded score() differs from the single-rank confusion matrix'
        sharded_check = {'images': BATCH, 'ranks': world, 'equal_to_single_rank': True,
                         'labelled_pixels': int(cm_sharded.sum())}

    net.close()
    del dev_sets
    torch.cuda.empty_cache()

    # ---------------- secondary workloads (driver-visible figures for configs[2] and [4])
    peak_tf, peak_tf_burst, peak_gbs, peak_kind = _peaks()
    extra = {}
    if not args.no_extras:
        # secondary figures must never cost the headline line: a failure is reported in place
        # (single rank; with several ranks a rank that fails inside a collective cannot be
        # rescued here and the launcher's timeout applies)
        def guarded(fn, *fn_args):
            try:
                return fn(*fn_args)
            except Exception as err:          # noqa: BLE001
                torch.cuda.synchronize()
                return {'unavailable': repr(err)[:200]}
        extra['dirichlet_mc_T20'] = guarded(dirichlet_mc_bench, world, rank)
        extra['fit'] = guarded(fit_bench, world, rank)
        if rank == 0:
            extra['fusion_hbm'] = guarded(fusion_hbm_rows, peak_gbs)
            extra['fusion_hbm']['_peak_gbs'] = peak_gbs
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get('conv_igemm_dram_bytes_per_launch')
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic', 'soak_s': round(soak_s, 2), 'cuda_graph': graph_info,
        'config': {'workload': WORKLOAD, 'global_batch': world * BATCH,
                   'parallelism': 'images sharded over %d GPU(s), no data-path collective' % world,
                   'l2': 'inputs rotate over %d sets (%.0f MB) and every step streams >2 GB of '
                         'activations through the 126 MB L2' % (n_sets, n_sets * h2d_bytes / 1e6),
                   'weights': 'random init of the named architecture (Glorot kernels, zero '
                              'biases, bilinear transposed convs)'},
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': C * C * 8, 'ms_per_step': e2e_s / args.steps * 1e3,
                'api': "BayesFusion.score({'rgb','depth','labels'}) on pinned host arrays, one "
                       "call per batch of 16",
                'dataset_call': {'value': world * BATCH * args.steps / ds_s,
                                 'frames_per_call': world * BATCH * args.steps,
                                 'note': 'ONE score() call over steps x 16 frames per GPU, '
                                         'batchsize 16 (base_model.py:294-313)'},
                'raw_dtype_inputs': {'value': world * BATCH * args.steps / raw_s,
                                     'dataset_call': world * BATCH * args.steps / raw_ds_s,
                                     'h2d_bytes_per_step': raw_bytes,
                                     'note': 'uint8 rgb + uint16 depth + int32 labels, cast on '
                                             'the device; same confusion matrix'}},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf,
                     'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                     'peak_burst': peak_tf_burst, 'frac_burst': achieved / peak_tf_burst,
                     'traffic': traffic,
                     'traffic_source': 'static: ncu --set full capture summarised in '
                                       'profiles/roofline_traffic.json (not measured in this run)',
                     'kernel': 'conv_igemm_kernel (all %d tcgen05 conv launches of the timed '
                               'region)' % conv_launches,
                     'flops_per_step': conv_flops / args.steps,
                     'kernel_ms_per_step': conv_ms / args.steps,
                     'kernel_share_of_step': conv_ms / ms,
                     'peak_source': 'frac: bf16_tflops_sustained, frac_burst: bf16_tflops; %s'
                                    % peak_kind,
                     'frac_note': 'both the achieved rate and the sustained peak are power-capped '
                                  'figures (see clocks): frac slightly above 1 means this box '
                                  'throttles less than the one the peak was measured on; '
                                  'frac_burst is against the un-throttled cuBLAS figure',
                     'whole_step_tflops': GFLOP_PER_FRAME * BATCH * args.steps / ms},
    }
    if sharded_check is not None:
        line['sharded_check'] = sharded_check
    line.update(extra)
    if world == 1:
        line['cpu_baseline'] = cpu_baseline(np.random.default_rng(0))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _bind_to_local_numa_node(local_rank):
    """Pins this rank's host threads to the CPUs of the NUMA node its GPU hangs off (pinned
    staging buffers are then allocated node-locally by first touch).  Best effort: silently
    does nothing where sysfs / NVML do not tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = local_rank
        if visible:
            ids = [v for v in visible.split(',') if v.strip() != '']
            if local_rank < len(ids) and ids[local_rank].strip().isdigit():
                phys = int(ids[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node_path = '/sys/bus/pci/devices/%s/numa_node' % bus[-12:].lower()
        node = int(open(node_path).read().strip())
        if node < 0:
            return None
        cpus = open('/sys/devices/system/node/node%d/cpulist' % node).read().strip()
        ids = set()
        for part in cpus.split(','):
            lo, _, hi = part.partition('-')
            ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = ids & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--soak', type=float, default=2.0,
                    help='seconds of untimed steps before the timed region')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the secondary workloads (fusion kernels, configs[2] and [4])')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()

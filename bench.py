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
  value     device-resident throughput: inputs already in HBM when the timed region starts
  e2e       the same metric through the public API `net.score(host arrays)`: pinned host
            buffers, H2D of rgb/depth/labels every step, D2H of the confusion matrix
  roofline  tensor-core stack (all tcgen05 conv launches): algorithmic FLOPs / CUDA-event time
            measured live in the timed region, against MEASURED_PEAKS.json
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
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get('bf16_tflops_sustained', 1400.0), p.get('hbm_gbs', 6650.0), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for t, line in self.rows:
            if t < t0 - 0.2 or t > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(',')]
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples'], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)),
                'reasons': sorted(reasons), 'samples': len(sm)}


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
    # sharding inside score() is by image over ranks; the bench gives every rank its own
    # batch-16 shard directly (weak scaling), so the model must not split it again
    rng = np.random.default_rng(1234 + rank)
    cms = _confusion_matrices(np.random.default_rng(0))
    net = get_model('bayes_fusion')(
        confusion_matrices=cms, data_description=_data_description(),
        prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=NU,
        num_channels={'rgb': 3, 'depth': 1}, batchsize=BATCH, seed=7, class_prior='data',
        shard_images=False)

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cm = torch.zeros((C, C), dtype=torch.int64, device='cuda')

    def device_step(i):
        net.score_batch_on_device(dev_sets[i % n_sets], cm)

    # ---------------- device-resident throughput (value) + live roofline of the conv stack
    for i in range(max(args.warmup, 3)):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
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
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * BATCH * args.steps / (ms * 1e-3)

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
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * BATCH * args.steps / e2e_s
    assert int(cm_host.sum()) == int((host_sets[(args.steps - 1) % n_sets]['labels'] >= 0).sum()
                                     ) * world or world > 1

    # same call with the sensors' own dtypes (uint8 rgb, uint16 depth): the float32 cast of
    # data_baseclass.py:77-78 then runs on the device after a 2.2x smaller copy (not the headline)
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
    raw_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([raw_s], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        raw_s = float(t.item())
    assert np.array_equal(cm_raw, cm_host)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_gbs, peak_kind = _peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get('conv_igemm_dram_bytes_per_launch')
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': world * BATCH,
                   'parallelism': 'images sharded over %d GPU(s), no data-path collective' % world,
                   'l2': 'inputs rotate over %d sets (%.0f MB) and every step streams >2 GB of '
                         'activations through the 126 MB L2' % (n_sets, n_sets * h2d_bytes / 1e6),
                   'weights': 'random init of the named architecture (Glorot kernels, zero '
                              'biases, bilinear transposed convs)'},
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': C * C * 8, 'ms_per_step': e2e_s / args.steps * 1e3,
                'api': "BayesFusion.score({'rgb','depth','labels'}) on pinned host arrays",
                'raw_dtype_inputs': {'value': world * BATCH * args.steps / raw_s,
                                     'h2d_bytes_per_step': raw_bytes,
                                     'note': 'uint8 rgb + uint16 depth + int32 labels, cast on '
                                             'the device; same confusion matrix'}},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf,
                     'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': traffic,
                     'kernel': 'conv_igemm_kernel (all %d tcgen05 conv launches of the timed '
                               'region)' % conv_launches,
                     'flops_per_step': conv_flops / args.steps,
                     'kernel_ms_per_step': conv_ms / args.steps,
                     'kernel_share_of_step': conv_ms / ms,
                     'peak_source': 'bf16_tflops_sustained, %s' % peak_kind,
                     'whole_step_tflops': GFLOP_PER_FRAME * BATCH * args.steps / ms},
    }
    if world == 1:
        line['cpu_baseline'] = cpu_baseline(np.random.default_rng(0))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()

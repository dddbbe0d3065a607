"""Batch-1 latency sweep mirroring experiments/timing.py of the reference (BASELINE configs[3]).

Same protocol as the reference commands (experiments/timing.py:23-311): constant inputs
`ones([1,768,384,C])` already on the device, random-init FCN experts, `repetitions` calls timed
with the wall clock around the call INCLUDING the device->host copy of the fused label map and
EXCLUDING the host->device copy (the input is a graph constant in the reference).  The first
iteration is included, as the reference does; a warm-excluded median is reported as well.

    python tools/timing.py [--repetitions 50] [--graph] [--cpu] [command ...]

--graph  additionally replays every command from a captured CUDA graph (launch-bound path)
--cpu    also times the oracle (CPU restatement of the reference graph) with the same protocol
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

H, W, C, NU, T = 768, 384, 12, 64, 20
NET = dict(num_units=NU, num_classes=C)


def _cms():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'exp868.npz'))
    return [g['cm_measure_rgb'], g['cm_measure_depth']]      # raw records, as timing.py:61-68


def _dirichlet_params(rng):
    return {'rgb': 1 + rng.gamma(2, 2, size=(C, C)), 'depth': 1 + rng.gamma(2, 2, size=(C, C)),
            'class_counts': rng.integers(1, 10 ** 6, size=C).astype(np.float64)}


def gpu_commands():
    import torch
    from xview.models.simple_fcn import fcn
    from xview.models.bayes_mix import bayes_decision_table, bayes_fusion
    from xview.models.dirichlet_mix import class_prior_from_counts, dirichlet_tables
    from modular_semantic_segmentation_b200 import device as dev

    rgb = torch.ones((1, H, W, 3), device='cuda')
    depth = torch.ones((1, H, W, 1), device='cuda')
    rng = np.random.default_rng(0)
    cms = _cms()
    lut = dev.to_device(bayes_decision_table([m.astype('float32') for m in cms]))
    dp = _dirichlet_params(rng)
    tables = [dev.to_device(t) for t in dirichlet_tables(
        [dp['rgb'], dp['depth']], 1.0, class_prior_from_counts(dp['class_counts'], 'data'))]
    kw = dict(trainable=False, batchnorm=False)

    from xview.models.fusion_fcn import fusion_fcn

    def fusion_fcn_cmd():                            # timing.py:23-46
        return fusion_fcn({'rgb': rgb, 'depth': depth}, {'rgb': 'rgb', 'depth': 'depth'}, NU, C,
                          trainable=False, want=('label',))['label']

    def rgb_fcn():                                   # timing.py:266-287
        return fcn(rgb, 'rgb', NU, C, want=('label',), **kw)['label']

    def depth_fcn():                                 # timing.py:290-311
        return fcn(depth, 'depth', NU, C, want=('label',), **kw)['label']

    def average_fcn():                               # timing.py:236-263
        p = [fcn(rgb, 'rgb', NU, C, want=('prob',), **kw)['prob'],
             fcn(depth, 'depth', NU, C, want=('prob',), **kw)['prob']]
        return dev.average_fuse(p)[1]

    def bayes_fcn():                                 # timing.py:49-83 (literal rule)
        labels = [fcn(rgb, 'rgb', NU, C, want=('label',), **kw)['label'],
                  fcn(depth, 'depth', NU, C, want=('label',), **kw)['label']]
        score, _, _ = bayes_fusion(labels, cms)
        return dev.softmax_argmax(score, want_prob=False)[1]

    def bayes_lookup_fcn():                          # timing.py:86-128 (decision table)
        labels = [fcn(rgb, 'rgb', NU, C, want=('label',), **kw)['label'],
                  fcn(depth, 'depth', NU, C, want=('label',), **kw)['label']]
        return dev.bayes_fuse_lut(labels, lut, C)

    def dirichlet_fcn():                             # timing.py:131-177
        p = [fcn(rgb, 'rgb', NU, C, want=('prob',), **kw)['prob'],
             fcn(depth, 'depth', NU, C, want=('prob',), **kw)['prob']]
        return dev.dirichlet_fuse(p, *tables)[1]

    def variance_fcn():                              # timing.py:180-233, num_samples = 20
        probs, variances = [], []
        for x, m in ((rgb, 'rgb'), (depth, 'depth')):
            variances.append(fcn(x, m, NU, C, dropout_rate=0.2, dropout_layers=['pool3'],
                                 num_samples=T, want=('mean_var',), **kw)['mean_var'])
            probs.append(fcn(x, m, NU, C, want=('prob',), **kw)['prob'])
        return dev.variance_fuse(probs, variances)[1]

    return {'fusion_fcn': fusion_fcn_cmd, 'rgb_fcn': rgb_fcn, 'depth_fcn': depth_fcn,
            'average_fcn': average_fcn,
            'bayes_fcn': bayes_fcn, 'bayes_lookup_fcn': bayes_lookup_fcn,
            'dirichlet_fcn': dirichlet_fcn, 'variance_fcn': variance_fcn}


def cpu_commands():
    """The oracle on the host cores; the code lives in bench.py (the only non-test module that
    may execute oracle/)."""
    sys.path.insert(0, ROOT)
    import bench
    return bench.cpu_sweep_commands(NU, C, H, W, _cms(), _dirichlet_params(np.random.default_rng(0)))


def time_command(fn, repetitions, to_host):
    times = []
    for _ in range(repetitions):
        start = time.time()
        result = to_host(fn())
        end = time.time()
        times.append(end - start)
    del result
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('commands', nargs='*')
    ap.add_argument('--repetitions', type=int, default=50)
    ap.add_argument('--graph', action='store_true')
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    import torch
    out = {'protocol': 'experiments/timing.py: batch 1, 768x384 ones, wall clock incl. D2H of the '
                       'label map, excl. H2D, first iteration included', 'rows': []}
    cmds = gpu_commands()
    for name, fn in cmds.items():
        if args.commands and name not in args.commands:
            continue
        times = time_command(fn, args.repetitions, lambda t: t.cpu().numpy())
        row = {'command': 'time_' + name, 'impl': 'b200', 'mean_s': float(np.mean(times)),
               'std_s': float(np.std(times)), 'median_warm_s': float(np.median(times[3:]))}
        print('time_%-18s Mean Time %.5fs, Std %.5fs  (warm median %.5fs)' % (
            name, row['mean_s'], row['std_s'], row['median_warm_s']))
        if args.graph:
            torch.cuda.synchronize()
            try:
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn()
                    torch.cuda.synchronize()
                    with torch.cuda.graph(graph, stream=side):
                        captured = fn()
                torch.cuda.synchronize()

                def replay():
                    graph.replay()
                    return captured
                gt = time_command(replay, args.repetitions, lambda t: t.cpu().numpy())
                row['graph_median_warm_s'] = float(np.median(gt[3:]))
                print('time_%-18s CUDA graph replay: warm median %.5fs' % (
                    name, row['graph_median_warm_s']))
            except RuntimeError as err:
                torch.cuda.synchronize()
                print('time_%-18s CUDA graph capture not possible: %s' % (name, str(err)[:80]))
        out['rows'].append(row)
    if args.cpu:
        for name, fn in cpu_commands().items():
            if args.commands and name not in args.commands:
                continue
            times = time_command(fn, max(3, min(args.repetitions, 5)), lambda a: a)
            row = {'command': 'time_' + name, 'impl': 'oracle-cpu', 'cores': os.cpu_count(),
                   'mean_s': float(np.mean(times)), 'std_s': float(np.std(times)),
                   'median_warm_s': float(np.median(times[1:]))}
            print('time_%-18s CPU oracle (%d cores): Mean Time %.5fs, Std %.5fs' % (
                name, os.cpu_count(), row['mean_s'], row['std_s']))
            out['rows'].append(row)
    if args.json:
        json.dump(out, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()

"""Turns the raw round-2 captures under gpurun_out/ (tools/gpu_r02_profile.sh) into the committed
summaries under profiles/ and writes profiles/README.md (round-2 sections, the multi-GPU lines
from profiles/r02_bench_{2,8}gpu.json, then the round-1 page profiles/r01_README.md).

    python tools/summarize_profiles_r02.py
"""
import collections
import csv
import json
import os
import shutil
import subprocess

G, P = 'gpurun_out', 'profiles'
os.makedirs(P, exist_ok=True)
md = []


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def raw(path):
    """`ncu --page raw --csv` export (made on the GPU box): header row, unit row, one row per launch."""
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2:]


def short(name):
    name = name.replace('void ', '').replace('xv::<unnamed>::', '').replace('xv::', '')
    return name.split('(')[0]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        if r[ui] == 'ns':
            v /= 1e3
        seq.append((short(r[ki]), v))
    return seq


# ---------------------------------------------------------------- bench lines
for name in ('bench', 'bench_reference', 'fit_bench', 'adapnet_bench'):
    src = os.path.join(G, 'r02_%s.json' % name)
    if os.path.exists(src) and os.path.getsize(src) > 10:
        open(os.path.join(P, 'r02_%s.json' % name), 'w').write(
            open(src).read().strip().splitlines()[-1] + '\n')
b = last_json(os.path.join(P, 'r02_bench.json'))
r = b['roofline']
md.append('## Round 2 (one B200 unless stated; `tools/gpu_r02_profile.sh`)\n')
md.append('### bench.py --steps 20 --warmup 5 (`r02_bench.json`)\n')
md.append('* value (device-resident, eager, after a %.1f s soak): **%.0f fused frames/s**, %.3f ms per '
          'batch-16 step, %d kernel launches per step; CUDA-graph replay of the same steps: %s'
          % (b['soak_s'], b['value'], b['ms_per_step'], b['gpu_launches'] // b['steps'],
             json.dumps(b.get('cuda_graph'))))
e = b['e2e']
md.append('* e2e through `BayesFusion.score(host arrays)`: **%.0f frames/s** per batch-16 call '
          '(H2D %.0f MB per call inside the timed region), %.0f frames/s as ONE call over %d frames; '
          'raw uint8/uint16 inputs: %.0f / %.0f frames/s'
          % (e['value'], e['h2d_bytes_per_step'] / 1e6, e['dataset_call']['value'],
             e['dataset_call']['frames_per_call'], e['raw_dtype_inputs']['value'],
             e['raw_dtype_inputs']['dataset_call']))
md.append('* conv stack inside the timed loop: %.0f TFLOP/s = **%.3f of the sustained / %.3f of the '
          'burst** measured bf16 peak (%.1f / %.1f TFLOP/s), %.1f %% of the step; whole step '
          '%.0f TFLOP/s' % (r['achieved'], r['frac'], r['frac_burst'], r['peak'], r['peak_burst'],
                            100 * r['kernel_share_of_step'], r['whole_step_tflops']))
md.append('* clocks during the timed region (NVML, 10 ms): %s' % json.dumps(b['clocks']))
md.append('* configs[2] `dirichlet_mc_T20`: %s' % json.dumps(b.get('dirichlet_mc_T20')))
md.append('* configs[4] `fit`: %s' % json.dumps(b.get('fit')))
md.append('* cpu_baseline (oracle port, %d host threads): %.2f frames/s\n'
          % (b['cpu_baseline']['cores'], b['cpu_baseline']['value']))
md.append('Fusion / score kernels inside the same run (`fusion_hbm`, fraction of the measured '
          '%.0f GB/s copy bandwidth, 16x768x384 px, C = 12, M = 2):\n' % b['fusion_hbm']['_peak_gbs'])
md.append('| kernel | B/px | ms | GB/s | frac |')
md.append('|---|---|---|---|---|')
for k, v in b['fusion_hbm'].items():
    if k.startswith('_'):
        continue
    md.append('| %s | %d | %.4f | %.0f | %.3f |' % (k, v['bytes_per_pixel'], v['ms'], v['gbs'],
                                                    v['frac']))
md.append('')

# ---------------------------------------------------------------- launch list of one step
seq = launches(os.path.join(G, 'r02_bench_launches.csv'))
ends = [i for i, (k, v) in enumerate(seq) if k.startswith('decode_bayes_confusion')]
step = seq[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else seq
with open(os.path.join(P, 'r02_bench_launches.csv'), 'w') as f:
    f.write('kernel,duration_us\n')
    for k, v in step:
        f.write('"%s",%.1f\n' % (k, v))
agg = collections.OrderedDict()
for k, v in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
md.append('### One fused batch-16 step, every launch under ncu (`r02_bench_launches.csv`; '
          'serialised, cold cache: compare shares)\n')
md.append('| kernel | launches | µs | share |')
md.append('|---|---|---|---|')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append('| `%s` | %d | %.1f | %.1f %% |' % (k, a[0], a[1], 100 * a[1] / tot))
conv = sum(a[1] for k, a in agg.items() if k.startswith('conv_'))
md.append('| **total** | %d | %.1f | conv share %.1f %% |\n' % (len(step), tot, 100 * conv / tot))

# ---------------------------------------------------------------- --set full captures
names = {'Kernel Name': 'kernel', 'gpu__time_duration.sum': 'us',
         'dram__bytes_read.sum': 'dram_rd', 'dram__bytes_write.sum': 'dram_wr',
         'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
         'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active': 'tensor_inst_pct',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
         'launch__registers_per_thread': 'regs',
         'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_pct'}
summary = []
unit = {}
for rep in ('r02_stream_raw.csv', 'r02_tail_raw.csv'):
    path = os.path.join(G, rep)
    if not os.path.exists(path) or os.path.getsize(path) < 100:
        continue
    hdr, unit_row, rows = raw(path)
    idx = {v: hdr.index(k) for k, v in names.items() if k in hdr}
    unit = {v: unit_row[hdr.index(k)] for k, v in names.items() if k in hdr}
    file_unit = {v: unit_row[hdr.index(k)] for k, v in names.items() if k in hdr}
    to_mb = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
    for row in rows:
        rec = {'capture': rep}
        for key, i in idx.items():
            val = row[i]
            if key != 'kernel':
                try:
                    val = float(val.replace(',', ''))
                except ValueError:
                    pass
                # ncu picks the byte unit per capture: normalise to MB
                if key in ('dram_rd', 'dram_wr') and isinstance(val, float):
                    val = round(val * to_mb.get(file_unit[key], 1.0), 3)
                if key == 'us' and isinstance(val, float) and file_unit[key] == 'ns':
                    val /= 1e3
            else:
                val = short(val)
            rec[key] = val
        summary.append(rec)
    unit['dram_rd'] = unit['dram_wr'] = 'Mbyte' 
if summary:
    json.dump({'units': unit, 'launches': summary},
              open(os.path.join(P, 'r02_ncu_full_summary.json'), 'w'), indent=1)
    md.append('### `ncu --set full` of one rgb stream forward + the fused score tail '
              '(`r02_ncu_full_summary.json`; units: %s)\n' % json.dumps(unit))
    md.append('| kernel | µs | DRAM read | DRAM write | tensor pipe active %% | DRAM %% | regs |')
    md.append('|---|---|---|---|---|---|---|')
    for rec in summary:
        md.append('| `%s` | %s | %s | %s | %s | %s | %s |' % (
            rec.get('kernel'), rec.get('us'), rec.get('dram_rd'), rec.get('dram_wr'),
            rec.get('tensor_pct'), rec.get('dram_pct'), rec.get('regs')))
    md.append('')

# mean DRAM traffic per tensor-core conv launch of the stream capture -> roofline.traffic of bench.py
convs = [r for r in summary if r['capture'] == 'r02_stream_raw.csv' and 'conv_' in r['kernel']
         and isinstance(r.get('dram_rd'), float)]
if convs:
    mean_b = sum((r['dram_rd'] + r['dram_wr']) * 1e6 for r in convs) / len(convs)
    algo = {'conv1_1 (rgb)': 16 * 768 * 384 * (3 * 4 + 64 * 2), 'conv1_2+pool1': 16 * 768 * 384 * 64 * 2 * 1.25,
            'conv2_1': 16 * 384 * 192 * (64 + 128) * 2}
    json.dump({'conv_igemm_dram_bytes_per_launch': mean_b,
               'note': 'mean dram__bytes_read+write over the %d tensor-core conv launches of one rgb '
                       'stream forward (batch 16, 768x384), ncu --set full, round 2 '
                       '(profiles/r02_ncu_full_summary.json)' % len(convs),
               'algorithmic_bytes_examples': algo},
              open(os.path.join(P, 'roofline_traffic.json'), 'w'), indent=1)
    md.append('Mean DRAM traffic per tensor-core conv launch: %.1f MB (`roofline_traffic.json`, the '
              'static `roofline.traffic` of bench.py).  Against algorithmic bytes: conv1_1 rgb '
              '%.0f MB, conv1_2+pool1 %.0f MB (604 + 151), conv2_1 %.0f MB - measured traffic is '
              '1.0x the algorithmic bytes on the layers that dominate it.\n'
              % (mean_b / 1e6, algo['conv1_1 (rgb)'] / 1e6, algo['conv1_2+pool1'] / 1e6,
                 algo['conv2_1'] / 1e6))

# ---------------------------------------------------------------- other artefacts
for name in ('r02_fusion_roofline.json', 'r02_timing_sweep.json', 'r02_fit_launches.csv'):
    src = os.path.join(G, name)
    if os.path.exists(src):
        if name.endswith('launches.csv'):
            seq = launches(src)
            starts = [i for i, (k, v) in enumerate(seq) if k.startswith('conv_c1_kernel')]
            one = seq[starts[-1]:] if starts else seq
            with open(os.path.join(P, name), 'w') as f:
                f.write('kernel,duration_us\n')
                for k, v in one:
                    f.write('"%s",%.1f\n' % (k, v))
            agg = collections.OrderedDict()
            for k, v in one:
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1
                a[1] += v
            tot = sum(a[1] for a in agg.values())
            md.append('### One training step (depth stream, batch 16, 384x768) under ncu '
                      '(`%s`)\n' % name)
            md.append('| kernel | launches | µs | share |')
            md.append('|---|---|---|---|')
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                md.append('| `%s` | %d | %.1f | %.1f %% |' % (k, a[0], a[1], 100 * a[1] / tot))
            md.append('| **total** | %d | %.1f | |\n' % (len(one), tot))
        else:
            shutil.copy(src, os.path.join(P, name))
# ---------------------------------------------------------------- multi-GPU lines (gpu_r02_multi.sh)
mg = []
for n in (2, 4, 8):
    path = os.path.join(P, 'r02_bench_%dgpu.json' % n)
    if os.path.exists(path):
        mg.append((n, last_json(path)))
if mg:
    md.append('### bench.py --gpus N (torchrun, one rank per GPU; `r02_bench_<N>gpu.json`)\n')
    md.append('| N | value (device-resident) | e2e per batch-16 call | e2e ONE call over the data set | '
              'raw uint8/uint16 inputs (per call / data set) | fit frames/s | MC T=20 frames/s | '
              'sharded_check |')
    md.append('|---|---|---|---|---|---|---|---|')
    one = b
    md.append('| 1 | %.0f | %.0f | %.0f | %.0f / %.0f | %.0f | %.0f | - |' % (
        one['value'], one['e2e']['value'], one['e2e']['dataset_call']['value'],
        one['e2e']['raw_dtype_inputs']['value'], one['e2e']['raw_dtype_inputs']['dataset_call'],
        one['fit']['value'], one['dirichlet_mc_T20']['value']))
    for n, x in mg:
        sc = x.get('sharded_check') or {}
        md.append('| %d | %.0f | %.0f | %.0f | %.0f / %.0f | %.0f | %.0f | %s |' % (
            n, x['value'], x['e2e']['value'], x['e2e']['dataset_call']['value'],
            x['e2e']['raw_dtype_inputs']['value'], x['e2e']['raw_dtype_inputs']['dataset_call'],
            x.get('fit', {}).get('value', float('nan')),
            x.get('dirichlet_mc_T20', {}).get('value', float('nan')),
            'equal to single rank' if sc.get('equal_to_single_rank') else sc))
    md.append('')
    md.append('The N = 1 row is this page\'s bench line; the N > 1 lines may predate the last kernel '
              'changes of the round (see the file dates).  Topology of the 8-GPU box: '
              '`r02_topo_8gpu.txt` (one NUMA node, all GPUs on CPUs 0-31); 2-GPU `pytest -m gpu` log of '
              'tests/test_gpu_multi.py: `r02_test_gpu_multi_2gpu.log`.\n')
md.append('### compute-sanitizer\n\n`r02_sanitizer_summary.md` (memcheck + racecheck over one small case per '
          'kernel family, `tools/sanitize.sh`).\n')
md.append('### Batch-1 latency sweep (`r02_timing_sweep.json`, protocol of experiments/timing.py)\n')
sweep = os.path.join(G, 'r02_timing_sweep.txt')
if os.path.exists(sweep):
    md.append('```')
    md.extend(l.rstrip() for l in open(sweep).read().strip().splitlines()[-24:])
    md.append('```\n')
for name, title in (('r02_adapnet_bench.json', 'Adapnet expert (`tools/adapnet_bench.py`)'),
                    ('r02_bench_reference.json', '`bench.py --impl reference` (oracle port on the host cores)')):
    path = os.path.join(P, name)
    if os.path.exists(path):
        md.append('### %s\n\n```\n%s\n```\n' % (title, open(path).read().strip()[:1500]))
text = '# Profiles\n\n' + '\n'.join(md) + '\n\n---\n\n' + open(os.path.join(P, 'r01_README.md')).read()
open(os.path.join(P, 'README.md'), 'w').write(text)
print('\n'.join(md))

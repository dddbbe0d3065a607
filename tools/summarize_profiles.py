"""Turns the raw captures under gpurun_out/ into the committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else 'r01'
G, P = 'gpurun_out', 'profiles'
os.makedirs(P, exist_ok=True)
md = ['# Profiles - round %s (one B200, batch 16, 768x384)\n' % R[1:]]


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


def col(hdr, name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    return None


# ---- bench JSON lines
for name in ('bench', 'bench_reference', 'fit_bench'):
    path = os.path.join(G, '%s_%s.json' % (R, name))
    if os.path.exists(path):
        line = open(path).read().strip().splitlines()[-1]
        open(os.path.join(P, '%s_%s.json' % (R, name)), 'w').write(line + '\n')
b = json.load(open(os.path.join(P, '%s_bench.json' % R)))
md.append('## bench.py (--steps 50)\n')
md.append('* value (device-resident): **%.0f fused frames/s**, %.3f ms per batch-16 step' % (b['value'], b['ms_per_step']))
md.append('* e2e (`BayesFusion.score(host arrays)`, H2D %.0f MB/step inside the timed region): **%.0f frames/s**' % (b['e2e']['h2d_bytes_per_step'] / 1e6, b['e2e']['value']))
md.append('* conv stack inside the loop: %.0f TFLOP/s = %.2f of measured sustained bf16 peak (%.1f), share of step %.2f' % (b['roofline']['achieved'], b['roofline']['frac'], b['roofline']['peak'], b['roofline']['kernel_share_of_step']))
md.append('* clocks during the timed region: %s' % json.dumps(b['clocks']))
md.append('* cpu_baseline (oracle port, %d cores): %.2f frames/s\n' % (b['cpu_baseline']['cores'], b['cpu_baseline']['value']))

# ---- launch list of the bench command
rows = [r for r in csv.reader(open(os.path.join(G, '%s_bench_launches.csv' % R))) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, tot, seq = collections.OrderedDict(), 0.0, []
allrows = []
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    if r[ui] == 'ns':
        v /= 1e3
    name = r[ki].replace('void ', '').replace('<unnamed>::', '').replace('xv::', '').split('(')[0]
    allrows.append((name, v))
# the capture holds every launch of the process; a step ends with confusion_kernel, the run does
# 3 warm-up steps and then the 2 timed device-resident steps (bench.py --steps 2 --warmup 1)
steps, cur = [], []
for name, v in allrows:
    cur.append((name, v))
    if name.startswith('confusion_kernel'):
        steps.append(cur)
        cur = []
for name, v in steps[3] + steps[4]:
    seq.append((name, v))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
with open(os.path.join(P, '%s_bench_launches.csv' % R), 'w') as f:
    f.write('kernel,duration_us\n')
    for n, v in seq:
        f.write('"%s",%.1f\n' % (n, v))
md.append('## ncu launch list of `bench.py --steps 2 --warmup 1` (%d launches = the 2 timed device-resident steps; cold-cache, serialised)\n' % len(seq))
md.append('| kernel | launches | total us | share |\n|---|---|---|---|')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append('| `%s` | %d | %.1f | %.1f%% |' % (k, n, t, 100 * t / tot))
conv_share = sum(t for k, (n, t) in agg.items() if 'conv_igemm' in k or 'conv_c1' in k) / tot
md.append('\ntensor-core conv kernels: %.1f%% of the serialised step (bench.py live measurement: %.1f%%)\n' % (100 * conv_share, 100 * b['roofline']['kernel_share_of_step']))

# ---- ncu --set full of the conv kernels (one stream forward)
hdr, rows = raw(os.path.join(G, '%s_conv_igemm.ncu-rep' % R))
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'dram rd MB'),
        ('dram__bytes_write.sum', 'dram wr MB'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM GB'),
        ('l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'smem bank rd %'),
        ('l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed', 'smem bank wr %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
        ('launch__registers_per_thread', 'regs')]
idx = [(col(hdr, k), t) for k, t in want if col(hdr, k) is not None]
# the capture (-s 15 -c 15 over the conv launches, 15 per forward) is the 2nd forward
layer_names = ['conv1_1', 'conv1_2+pool1', 'conv2_1', 'conv2_2+pool2', 'conv3_1', 'conv3_2', 'conv3_3', 'conv4_1',
               'conv4_2', 'conv4_3', 'conv5_1', 'conv5_2', 'conv5_3', 'score_conv4', 'score_conv5']
md.append('## ncu --set full, tensor-core conv kernels of one stream forward\n')
md.append('| layer | ' + ' | '.join(t for _, t in idx) + ' |\n|' + '---|' * (len(idx) + 1))
traffic = []
for n, r in enumerate(rows):
    vals = []
    for i, t in idx:
        v = r[i]
        if t == 'kernel':
            v = '`' + v.replace('void ', '').replace('<unnamed>::', '').split('(')[0][-34:] + '`'
        else:
            try:
                v = '%.1f' % float(v)
            except ValueError:
                pass
        vals.append(v)
    md.append('| %s | ' % (layer_names[n] if n < len(layer_names) else '') + ' | '.join(vals) + ' |')
    ir, iw = col(hdr, 'dram__bytes_read.sum'), col(hdr, 'dram__bytes_write.sum')
    if n < 13:
        traffic.append((float(r[ir]) + float(r[iw])) * 1e6)
json.dump({'conv_igemm_dram_bytes_per_launch': sum(traffic) / len(traffic),
           'note': 'mean dram__bytes_read+write over the 13 conv layers of one stream forward (batch 16), ncu --set full'},
          open(os.path.join(P, 'roofline_traffic.json'), 'w'), indent=1)
md.append('\nmean DRAM traffic per 3x3/conv1_1 launch: %.1f MB -> `roofline.traffic` of bench.py\n' % (sum(traffic) / len(traffic) / 1e6))
md.append('(`--set full` replays every kernel ~40 times with all counters on: its durations run 5-10 % above the launch '
          'list. This capture predates the 16-byte fp32 epilogue stores and the 4-pixel label decode: score_conv4 42 -> 27 us, '
          'score_conv5 16 -> 11 us, decode 54 -> 43 us in the launch list above.)\n')

# ---- fusion kernels
f = json.load(open(os.path.join(G, '%s_fusion_roofline.json' % R)))
json.dump(f, open(os.path.join(P, '%s_fusion_roofline.json' % R), 'w'), indent=1)
md.append('## Fusion / score kernels vs measured HBM copy peak (%.1f GB/s), CUDA events (`tools/fusion_bench.py`)\n' % f['hbm_peak_gbs_measured'])
md.append('| kernel | algorithmic B/px | ms | GB/s | fraction |\n|---|---|---|---|---|')
for r in f['rows']:
    md.append('| %s | %.0f | %.3f | %.0f | %.2f |' % (r['kernel'], r['bytes_per_pixel'], r['ms'], r['achieved_gbs'], r['frac_of_measured_hbm']))
hdr, rows = raw(os.path.join(G, '%s_fusion.ncu-rep' % R))
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'dram rd MB'), ('dram__bytes_write.sum', 'dram wr MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak')]
idx = [(col(hdr, k), t) for k, t in want if col(hdr, k) is not None]
md.append('\nncu --set full (first launches of `tools/fusion_bench.py`):\n')
md.append('| ' + ' | '.join(t for _, t in idx) + ' |\n|' + '---|' * len(idx))
for r in rows:
    vals = []
    for i, t in idx:
        v = r[i]
        if t == 'kernel':
            v = '`' + v.replace('void ', '').replace('<unnamed>::', '').split('(')[0][-40:] + '`'
        vals.append(v)
    md.append('| ' + ' | '.join(vals) + ' |')

# ---- timing sweep, fit
t = json.load(open(os.path.join(G, '%s_timing_sweep.json' % R)))
json.dump(t, open(os.path.join(P, '%s_timing_sweep.json' % R), 'w'), indent=1)
ref = {'time_fusion_fcn': 0.0720, 'time_rgb_fcn': 0.0219, 'time_depth_fcn': 0.0218, 'time_average_fcn': 0.0432, 'time_bayes_fcn': 0.0461,
       'time_dirichlet_fcn': 0.0517, 'time_variance_fcn': 0.3064}
md.append('\n## Batch-1 latency sweep, protocol of experiments/timing.py (`tools/timing.py`)\n')
md.append('| command | impl | mean s (first call included) | warm median s | CUDA-graph replay s | reference (GTX 1080 Ti, TF 1.x) s |\n|---|---|---|---|---|---|')
for r in t['rows']:
    md.append('| %s | %s | %.5f | %.5f | %s | %s |' % (r['command'], r['impl'], r['mean_s'], r['median_warm_s'],
                                                  '%.5f' % r['graph_median_warm_s'] if 'graph_median_warm_s' in r else '',
                                                  ref.get(r['command'], '') if r['impl'] == 'b200' else ''))
fb = json.load(open(os.path.join(P, '%s_fit_bench.json' % R)))
md.append('\n## fit() (`tools/fit_bench.py`, depth stream 384x768, batch 16, Adam)\n')
md.append('* %.0f frames/s on 1 GPU, %.1f ms per step (forward + backward + Adam), %d parameters, %.0f MB gradient bucket' % (fb['value'], fb['ms_per_step'], fb['params'], fb['allreduce_bytes_per_step'] / 1e6))
# ---- Adapnet expert
ap = os.path.join(G, '%s_adapnet_bench.json' % R)
if os.path.exists(ap):
    a = json.loads(open(ap).read().strip().splitlines()[-1])
    json.dump(a, open(os.path.join(P, '%s_adapnet_bench.json' % R), 'w'), indent=1)
    md.append('\n## Adapnet expert (`tools/adapnet_bench.py`, rgb 768x384, batch 16, num_units 64, 12 classes)\n')
    md.append('* %.0f frames/s per expert, %.2f ms per forward, %d launches; tensor-core convs %.0f GFLOP/frame at %.0f TFLOP/s' % (
        a['frames_per_s'], a['ms_per_forward'], a['launches_per_forward'], a['conv_gflop_per_frame'], a['conv_tflops']))
    lp = os.path.join(G, '%s_adapnet_launches.csv' % R)
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
        hdr = rows[0]
        ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
        agg, tot = collections.OrderedDict(), 0.0
        with open(os.path.join(P, '%s_adapnet_launches.csv' % R), 'w') as f:
            f.write('kernel,duration_us\n')
            for r in rows[1:]:
                try:
                    v = float(r[vi].replace(',', ''))
                except ValueError:
                    continue
                if r[ui] == 'ns':
                    v /= 1e3
                name = r[ki].replace('void ', '').replace('<unnamed>::', '').replace('xv::', '').split('(')[0]
                f.write('"%s",%.1f\n' % (name, v))
                e = agg.setdefault(name, [0, 0.0])
                e[0] += 1
                e[1] += v
                tot += v
        md.append('\n| kernel (ncu launch list of one forward) | launches | total us | share |\n|---|---|---|---|')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append('| `%s` | %d | %.1f | %.1f%% |' % (k, n, t, 100 * t / tot))
for extra in ('bench_2gpu', 'bench_4gpu', 'bench_8gpu'):
    ep = os.path.join(P, '%s_%s.json' % (R, extra))
    if os.path.exists(ep):
        e = json.loads(open(ep).read().strip().splitlines()[-1])
        md.append('\n## bench.py --gpus %d\n' % e['n_gpus'])
        md.append('* value %.0f frames/s (%.3f ms per step), e2e %.0f frames/s' % (e['value'], e['ms_per_step'], e['e2e']['value']))
        if extra == 'bench_8gpu':
            md.append('* captured earlier in the round, before the conv1_1 / halo kernels (that build: 2816 frames/s on 1 GPU, '
                      'i.e. 0.994 weak-scaling efficiency); the 8-GPU box costs 8x GPU-minutes, so it was not repeated')
open(os.path.join(P, 'README.md'), 'w').write('\n'.join(md) + '\n')
print('\n'.join(md))

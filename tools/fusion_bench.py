"""HBM roofline of the per-pixel fusion + score kernels at the batch-16 768x384 workload size.

Algorithmic bytes per pixel follow SURVEY.md 8(d) / DESIGN.md 4.4.  Timing: CUDA events on the
launching stream, 3 warm-ups, inputs (>= 226 MB per expert) far larger than the 126 MB L2.
"""
import json, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from modular_semantic_segmentation_b200 import device as dev

dev.init()
N, H, W, C, M, T = 16, 768, 384, 12, 2, 20
npix = N * H * W
peak = 6554.9
if os.path.exists('MEASURED_PEAKS.json'):
    peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
g = torch.Generator(device='cuda').manual_seed(0)


def probs():
    return torch.softmax(2 * torch.randn((N, H, W, C), device='cuda', generator=g), -1).contiguous()


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = []


def report(name, bytes_per_px, ms, px=npix):
    gbs = bytes_per_px * px / ms / 1e6
    rows.append({'kernel': name, 'bytes_per_pixel': bytes_per_px, 'ms': round(ms, 4),
                 'achieved_gbs': round(gbs, 1), 'frac_of_measured_hbm': round(gbs / peak, 3)})
    print('%-34s %6.1f B/px  %8.3f ms  %8.1f GB/s  %5.1f%% of %.0f' % (name, bytes_per_px, ms, gbs, 100 * gbs / peak, peak))


p = [probs() for _ in range(M)]
score = torch.randn((N, H, W, C), device='cuda', generator=g)
l64 = [torch.randint(0, C, (N, H, W), device='cuda', generator=g) for _ in range(M)]
l64_placeholder = l64[0]
l8 = [t.to(torch.uint8) for t in l64]
gt_rand = torch.randint(-1, C, (N, H, W), device='cuda', generator=g, dtype=torch.int32)
# segmentation-like label maps: 32x32-pixel blocks of one class (ground truth and prediction agree
# on most blocks), which is what score() sees on real data
blocks = torch.randint(-1, C, (N, H // 32, W // 32), device='cuda', generator=g, dtype=torch.int32)
gt = blocks.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
pred_blocky = torch.where(torch.rand((N, H, W), device='cuda', generator=g) < 0.9,
                          gt.clamp(min=0).to(torch.int64), l64_placeholder)
lut = torch.randint(0, C, (C, C), device='cuda', generator=g, dtype=torch.int32)
am1 = torch.rand((M, C, C), device='cuda', generator=g) * 3
lognorm = torch.rand((M, C), device='cuda', generator=g)
logprior = torch.log(torch.full((C,), 1.0 / C, device='cuda'))
var = [torch.rand((N, H, W), device='cuda', generator=g) * 1e-2 for _ in range(M)]

report('softmax_argmax (prob + i64 label)', 4 * C * 2 + 8, timeit(lambda: dev.softmax_argmax(score)))
report('softmax_argmax (u8 label only)', 4 * C + 1,
       timeit(lambda: dev.softmax_argmax(score, want_prob=False, label_dtype=torch.uint8)))
report('bayes_fuse_lut int64', 8 * M + 8, timeit(lambda: dev.bayes_fuse_lut(l64, lut, C)))
report('bayes_fuse_lut uint8', M + 1, timeit(lambda: dev.bayes_fuse_lut(l8, lut, C)))
report('dirichlet_fuse (i64 label)', 4 * C * M + 8, timeit(lambda: dev.dirichlet_fuse(p, am1, lognorm, logprior)))
report('dirichlet_fuse (u8 label)', 4 * C * M + 1,
       timeit(lambda: dev.dirichlet_fuse(p, am1, lognorm, logprior, label_dtype=torch.uint8)))
report('average_fuse (i64 label)', 4 * C * M + 8, timeit(lambda: dev.average_fuse(p)))
report('variance_fuse (i64 label)', 4 * C * M + 4 * M + 8, timeit(lambda: dev.variance_fuse(p, var)))
cm = torch.zeros((C, C), dtype=torch.int64, device='cuda')
pb8 = pred_blocky.to(torch.uint8)
report('confusion int64 pred (blocky maps)', 12, timeit(lambda: dev.confusion_accumulate(pred_blocky, gt, cm)))
report('confusion uint8 pred (blocky maps)', 5, timeit(lambda: dev.confusion_accumulate(pb8, gt, cm)))
report('confusion int64 pred (random maps)', 12, timeit(lambda: dev.confusion_accumulate(l64[0], gt_rand, cm)))
stats = torch.zeros((C, C), dtype=torch.float64, device='cuda')
cnt = torch.zeros(C, dtype=torch.int64, device='cuda')
report('dirichlet_suffstats (blocky maps)', 4 * C + 4, timeit(lambda: dev.dirichlet_suffstats(p[0], gt, stats, cnt)))
report('dirichlet_suffstats (random maps)', 4 * C + 4, timeit(lambda: dev.dirichlet_suffstats(p[0], gt_rand, stats, cnt)))
C13 = 13
p13 = [torch.softmax(torch.randn((N, H, W, C13), device='cuda', generator=g), -1).contiguous() for _ in range(M)]
report('average_fuse C=13 (i64 label)', 4 * C13 * M + 8, timeit(lambda: dev.average_fuse(p13)))
report('softmax_argmax C=13 (prob+i64)', 4 * C13 * 2 + 8, timeit(lambda: dev.softmax_argmax(p13[0])))
del p13
# MC moments over T materialised samples of 2 frames (T*2 frames of probabilities = 1.1 GB)
n2 = 2
samples = torch.softmax(torch.randn((T, n2, H, W, C), device='cuda', generator=g), -1).contiguous()
report('mc_moments T=20 (mean,var,mean_var)', 4 * C * T + 8 * C + 4,
       timeit(lambda: dev.mc_moments(samples), 5), px=n2 * H * W)
os.makedirs('gpurun_out', exist_ok=True)
json.dump({'hbm_peak_gbs_measured': peak, 'workload': '%dx%dx%d px, C=%d, M=%d' % (N, H, W, C, M),
           'rows': rows}, open('gpurun_out/fusion_roofline.json', 'w'), indent=1)

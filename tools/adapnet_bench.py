"""Adapnet expert throughput probe: frames/s and tensor-core TFLOP/s of the conv launches at the
bench resolution (768x384), random weights.  usage: adapnet_bench.py [batch] [iters] [cin]"""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modular_semantic_segmentation_b200 import device as dev            # noqa: E402
from modular_semantic_segmentation_b200.models.adapnet import build_adapnet   # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cin = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    h, w, nu, c = 384, 768, 64, 12
    dev.init()
    expert, variables = build_adapnet('rgb', cin, nu, c, rng=np.random.default_rng(0))
    expert.set_params({k[4:]: v for k, v in variables.items()})
    x = torch.rand((batch, h, w, cin), device='cuda') * 255
    for _ in range(3):
        expert.forward(x, want=('label',), label_dtype=torch.uint8)
    torch.cuda.synchronize()
    l0 = dev.launch_count()
    dev.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        expert.forward(x, want=('label',), label_dtype=torch.uint8)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    conv_ms, conv_flops, conv_launches = dev.profile_read()
    dev.profile_enable(False)
    print(json.dumps({
        'batch': batch, 'ms_per_forward': ms, 'frames_per_s': batch / ms * 1e3,
        'conv_ms_per_forward': conv_ms / iters, 'conv_gflop_per_frame': conv_flops / iters / batch / 1e9,
        'conv_tflops': conv_flops / conv_ms / 1e9 if conv_ms else 0,
        'conv_launches_per_forward': conv_launches / iters,
        'launches_per_forward': (dev.launch_count() - l0) / iters}))


if __name__ == '__main__':
    main()

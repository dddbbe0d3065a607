"""Quick perf probe: one FCN stream at batch N, 768x384, nu=64, C=12 (random weights)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from modular_semantic_segmentation_b200 import device as dev

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cin = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev.init()
rng = np.random.default_rng(0)
from modular_semantic_segmentation_b200.models.simple_fcn import init_fcn_variables
params = init_fcn_variables('m', cin, 64, 12, rng=rng)
net = dev.FcnExpert(cin, 64, 12, precision='bf16')
net.set_params({k.split('/', 1)[1]: v for k, v in params.items()})
x = torch.rand((N, 768, 384, cin), device='cuda')
for _ in range(2):
    out = net.forward(x, want=('label',), label_dtype=torch.uint8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = net.forward(x, want=('label',), label_dtype=torch.uint8)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
gflop = 181.23 if cin == 3 else 180.55
print('N=%d cin=%d: %.3f ms/forward, %.1f frames/s/stream, %.1f TFLOP/s' % (N, cin, ms, N / ms * 1e3, N * gflop / ms))

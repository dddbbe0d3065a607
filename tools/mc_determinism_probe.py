"""Debug probe: the MC moments of two identical calls, under the default kernels and with the
scalar dropout (debug flag 256) / per-sample-barrier decode (512) kernels."""
import numpy as np
import torch
from modular_semantic_segmentation_b200 import device as dev
from modular_semantic_segmentation_b200.models.simple_fcn import build_expert

rng = np.random.default_rng(9)
expert, variables = build_expert('depth', 1, 64, 12, rng=rng)
variables['depth/conv1_1/kernel'] = variables['depth/conv1_1/kernel'] / np.float32(65535.0)
expert.set_params({k[6:]: v for k, v in variables.items()})
x = torch.from_numpy(rng.integers(0, 65536, size=(1, 768, 384, 1)).astype(np.float32)).cuda()
drop = {'rate': 0.5, 'layers': ['pool3'], 'num_samples': 20, 'seed': 4}
res = {}
for flags in (0, 256, 512, 768):
    dev.set_debug_flags(flags)
    a = expert.forward(x, want=('prob', 'mean_prob', 'var_prob', 'mean_var'), dropout=drop)
    a = {k: v.clone() for k, v in a.items()}
    b = expert.forward(x, want=('mean_var',), dropout=drop)['mean_var'].clone()
    c = expert.forward(x, want=('mean_var',), dropout=drop)['mean_var'].clone()
    res[flags] = a
    print(flags, 'a==b', torch.equal(a['mean_var'], b), 'b==c', torch.equal(b, c),
          'max|a-b|', (a['mean_var'] - b).abs().max().item(), 'max', b.abs().max().item())
dev.set_debug_flags(0)
for f in (256, 512, 768):
    print(f, 'prob equal to default', torch.equal(res[f]['prob'], res[0]['prob']),
          'mean_var diff', (res[f]['mean_var'] - res[0]['mean_var']).abs().max().item())

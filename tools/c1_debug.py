import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modular_semantic_segmentation_b200 import device as dev
import oracle
dev.init()
rng = np.random.default_rng(0)
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 3
h, w = 32, 48
params = oracle.glorot_fcn_params('m', cin, 8, 5, rng, gain=1.4, bias_scale=0.05)
net = dev.FcnExpert(cin, 8, 5, precision='bf16')
net.set_params({k.split('/', 1)[1]: v for k, v in params.items()})
x = rng.uniform(0, 1, size=(2, h, w, cin)).astype(np.float32)
out = net.forward(torch.from_numpy(x).cuda(), want=('score',))
torch.cuda.synchronize()
got = net.layer('conv1_1')
ref = oracle.test_pipeline(x, params, 'm', 8, 5)['conv1_1']
print('conv1_1 max err', np.abs(got - ref).max(), 'scale', np.abs(ref).max())

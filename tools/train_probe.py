import sys
sys.path.insert(0,'.')
import numpy as np, torch
from modular_semantic_segmentation_b200 import device as dev
dev.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng=np.random.default_rng(0)
from modular_semantic_segmentation_b200.models.simple_fcn import init_fcn_variables
params=init_fcn_variables('m',1,64,12,rng=rng)
net=dev.FcnExpert(1,64,12,precision='bf16'); net.set_params({k.split('/',1)[1]:v for k,v in params.items()})
net.train_begin()
x=torch.rand((N,384,768,1),device='cuda'); lab=torch.randint(0,12,(N,384,768),device='cuda',dtype=torch.int32)
g=l=None
for _ in range(2):
    g,l=net.train_gradients(x,lab,grads=g,loss=l); net.adam_step(g)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    g,l=net.train_gradients(x,lab,grads=g,loss=l); net.adam_step(g)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/steps
print('fit step N=%d 384x768 depth stream: %.1f ms/step, %.1f frames/s'%(N,ms,N/ms*1e3))

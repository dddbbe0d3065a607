"""One step of BASELINE configs[2] (T = 20 MC-dropout samples per modality -> Dirichlet fusion ->
confusion matrix) for profiling:  python tools/mc_probe.py [batch] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
out = bench.dirichlet_mc_bench(1, 0, steps=steps, batch=batch)
print(out)

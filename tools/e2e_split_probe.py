"""Which upload schedule gives the best per-batch score() rate?  (kernel study, not a bench)
    python tools/e2e_split_probe.py
Times BayesFusion.score(host batch of 16) for upload_split = 1 (whole batch, smallest modality
first) and 2 ([4, 12] pieces) and explicit piece lists."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xview.models import get_model

torch.cuda.set_device(0)
rng = np.random.default_rng(0)
cms = bench._confusion_matrices(rng)
H, W, C, B = bench.H, bench.W, bench.C, bench.BATCH
sets = [{'rgb': torch.from_numpy(rng.integers(0, 256, size=(B, H, W, 3)).astype(np.float32)).pin_memory(),
         'depth': torch.from_numpy(rng.integers(0, 65536, size=(B, H, W, 1)).astype(np.float32)).pin_memory(),
         'labels': torch.from_numpy(rng.integers(-1, C, size=(B, H, W)).astype(np.int32)).pin_memory()}
        for _ in range(3)]
for label, extra in (('split=2 [4,12]', {}), ('split=1 (whole)', {'upload_split': 1}),
                     ('pieces [8,8]', {'upload_pieces': [8, 8]}),
                     ('pieces [2,14]', {'upload_pieces': [2, 14]})):
    net = get_model('bayes_fusion')(
        confusion_matrices=cms, data_description=bench._data_description(),
        prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=bench.NU,
        num_channels={'rgb': 3, 'depth': 1}, batchsize=B, seed=7, shard_images=False, **extra)
    for i in range(3):
        net.score(sets[i % 3])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        net.score(sets[i % 3])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print('%-18s %.3f ms per score()  %.0f frames/s' % (label, dt * 1e3, B / dt), flush=True)
    net.close()

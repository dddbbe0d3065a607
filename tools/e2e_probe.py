"""score() on pinned host batches of 16 frames: device-resident step time vs batch size, and the
end-to-end time for different upload piece lists."""
import sys, os, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xview.models import get_model

rng = np.random.default_rng(0)
cms = bench._confusion_matrices(np.random.default_rng(0))
B, H, W, C = 16, bench.H, bench.W, bench.C


def model(**extra):
    return get_model('bayes_fusion')(
        confusion_matrices=cms, data_description=bench._data_description(),
        prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn', num_units=bench.NU,
        num_channels={'rgb': 3, 'depth': 1}, batchsize=B, seed=7, class_prior='data',
        shard_images=False, **extra)


host = {'rgb': torch.from_numpy(rng.integers(0, 256, size=(B, H, W, 3)).astype(np.float32)).pin_memory(),
        'depth': torch.from_numpy(rng.integers(0, 65536, size=(B, H, W, 1)).astype(np.float32)).pin_memory(),
        'labels': torch.from_numpy(rng.integers(-1, C, size=(B, H, W)).astype(np.int32)).pin_memory()}
devb = {k: v.cuda() for k, v in host.items()}
net = model()
cm = torch.zeros((C, C), dtype=torch.int64, device='cuda')
for n in (1, 2, 3, 4, 6, 8, 12, 16):
    part = {k: v[:n].contiguous() for k, v in devb.items()}
    for _ in range(3):
        net.score_batch_on_device(part, cm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        net.score_batch_on_device(part, cm)
    torch.cuda.synchronize()
    print('device step, batch %2d: %.3f ms' % (n, (time.perf_counter() - t0) * 100))
net.close()
for pieces in ([16], [2, 6, 8], [4, 12], [3, 13], [2, 14], [4, 4, 8], [1, 3, 12], [2, 4, 10], [6, 10]):
    net = model(upload_pieces=pieces)
    for _ in range(3):
        net.score(host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        net.score(host)
    dt = (time.perf_counter() - t0) / 10
    print('e2e pieces %-12s %.3f ms  %.0f frames/s' % (pieces, dt * 1e3, B / dt))
    net.close()

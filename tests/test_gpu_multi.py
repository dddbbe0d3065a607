"""Multi-GPU parity (needs >= 2 GPUs): images sharded over ranks, NCCL all-reduce of the int64
confusion matrix, all-gather of the label maps - results must equal the single-GPU run exactly."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ['XV_ROOT'])
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
from xview.models import get_model
data = dict(np.load(os.environ['XV_DATA']))
c = int(data.pop('c'))
desc = ({'rgb': np.float32, 'depth': np.float32, 'labels': np.int32},
        {'rgb': (None, None, 3), 'depth': (None, None, 1), 'labels': (None, None)}, c)
cms = {'rgb': data.pop('cm_rgb'), 'depth': data.pop('cm_depth')}
with get_model('bayes_fusion')(confusion_matrices=cms, data_description=desc,
                               prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn',
                               num_units=8, num_channels={'rgb': 3, 'depth': 1}, batchsize=2,
                               seed=3) as net:
    measures, cm = net.score(data)
    pred = net.predict({'rgb': data['rgb'], 'depth': data['depth']})
# shard_images=False: the dict a rank passes IS its share (here every rank passes all images, so
# the all-reduced matrix counts every image `world` times)
with get_model('bayes_fusion')(confusion_matrices=cms, data_description=desc,
                               prefixes={'rgb': 'rgb', 'depth': 'depth'}, expert_model='fcn',
                               num_units=8, num_channels={'rgb': 3, 'depth': 1}, batchsize=2,
                               seed=3, shard_images=False) as net2:
    _, cm_rep = net2.score(data)
replica_ok = bool(np.array_equal(cm_rep, world * cm))
# data-parallel fit(): every rank trains on its share of each batch, gradients are summed
fdesc = ({'rgb': np.float32, 'labels': np.int32}, {'rgb': (None, None, 3), 'labels': (None, None)}, c)
train = {'rgb': data['rgb'][:6] / 255.0, 'labels': np.clip(data['labels'][:6], 0, None)}
with get_model('fcn')('rgb', fdesc, 'rgb', num_units=8, batch_normalization=False,
                      learning_rate=1e-4, batchsize=4 // world, seed=9) as fnet:  # same global batch
    fnet.fit(train, 6)
    trained = fnet.variables['rgb/conv4_2/kernel'].copy()
    trained_head = fnet.variables['rgb/score/kernel'].copy()
    final_loss = fnet.loss
# batch-1 latency mode: the T MC-dropout samples (not the images) are split over the ranks and the
# per-rank moments merged over NCCL; every rank must end with the same fused labels
split_ok, split_shape = True, (2, 32, 48)
if world > 1:
    with get_model('variance_fusion')(data_description=desc, prefixes={'rgb': 'rgb', 'depth': 'depth'},
                                      expert_model='fcn', num_units=8,
                                      num_channels={'rgb': 3, 'depth': 1}, batchsize=1,
                                      num_samples=7, dropout_rate=0.3, seed=4, shard_images=False,
                                      split_samples=True, deterministic_dropout=True) as vnet:
        vpred = vnet.predict({'rgb': data['rgb'][:2], 'depth': data['depth'][:2]})
    mine = torch.from_numpy(vpred).cuda()
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    split_ok = bool(all(torch.equal(b, both[0]) for b in both))
    split_shape = tuple(vpred.shape)
if rank == 0:
    np.savez(os.environ['XV_OUT'], cm=cm, pred=pred, miou=measures['mean_IoU'], trained=trained,
             trained_head=trained_head, final_loss=final_loss, split_ok=split_ok,
             split_shape=np.asarray(split_shape), replica_ok=replica_ok)
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_score_and_predict_equal_single_gpu(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(0)
    c, n, h, w = 6, 7, 32, 48                       # 7 images: uneven split over 2 ranks
    np.savez(tmp_path / 'data.npz', c=c,
             rgb=rng.integers(0, 256, size=(n, h, w, 3)).astype(np.float32),
             depth=rng.integers(0, 65536, size=(n, h, w, 1)).astype(np.float32),
             labels=rng.integers(-1, c, size=(n, h, w)).astype(np.int32),
             cm_rgb=rng.integers(1, 50, size=(c, c)).astype(np.float64) + 100 * np.eye(c),
             cm_depth=rng.integers(1, 50, size=(c, c)).astype(np.float64) + 100 * np.eye(c))
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    results = {}
    for world in (1, 2):
        env = dict(os.environ, XV_ROOT=root, XV_DATA=str(tmp_path / 'data.npz'),
                   XV_OUT=str(tmp_path / ('out%d.npz' % world)))
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
               '--master-port', str(_free_port()), str(script)]
        proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
        results[world] = dict(np.load(tmp_path / ('out%d.npz' % world)))
    np.testing.assert_array_equal(results[2]['cm'], results[1]['cm'])
    np.testing.assert_array_equal(results[2]['pred'], results[1]['pred'])
    assert results[2]['miou'] == results[1]['miou']
    assert results[1]['pred'].shape == (n, h, w)
    # MC samples split over the ranks, merged moments: every rank ends with the same label maps
    assert tuple(results[2]['split_shape']) == (2, h, w)
    assert bool(results[2]['split_ok'])
    assert bool(results[2]['replica_ok']) and bool(results[1]['replica_ok'])
    # 2-rank data-parallel training follows the 1-rank trajectory (same global batches; the
    # gradient sums differ only by fp32 accumulation order)
    np.testing.assert_allclose(results[2]['final_loss'], results[1]['final_loss'], rtol=2e-2)
    # Adam normalises every step to ~lr, so entries whose gradient is fp32 noise may move in
    # opposite directions: bound by 2 * lr * steps, and require the bulk to agree closely
    for key in ('trained_head', 'trained'):
        diff = np.abs(results[2][key] - results[1][key])
        assert diff.max() <= 2 * 1e-4 * 6 + 1e-6, (key, diff.max())
        assert np.median(diff) < 2e-5, (key, np.median(diff))

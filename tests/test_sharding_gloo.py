"""world_size-2 gloo test (CPU) of the image sharding + confusion-matrix all-reduce +
prediction gather used by score()/predict() when one process drives each GPU."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from modular_semantic_segmentation_b200 import sharding
    rng = np.random.default_rng(0)                 # same data on every rank
    c, n_img = 5, 7                                 # uneven split: 4 + 3 images
    labels = rng.integers(-1, c, size=(n_img, 6, 8))
    pred = rng.integers(0, c, size=(n_img, 6, 8))
    r, w = sharding.rank_world()
    assert (r, w) == (rank, world)
    mine = []
    for start in range(0, n_img, 3):                # batches of 3 images
        count = min(3, n_img - start)
        mine += [start + i for i in sharding.rows_for_rank(start, count, rank, world)]
    assert mine == list(range(rank, n_img, world))
    cm = torch.from_numpy(oracle.confusion_matrix(labels[mine], pred[mine], c))
    sharding.allreduce_sum_(cm)
    gathered = sharding.gather_interleaved(torch.from_numpy(pred[mine]), 'cpu')
    np.save(os.path.join(out_dir, 'cm%d.npy' % rank), cm.numpy())
    np.save(os.path.join(out_dir, 'pred%d.npy' % rank), gathered.numpy())
    np.save(os.path.join(out_dir, 'ref_cm.npy'), oracle.confusion_matrix(labels, pred, c))
    np.save(os.path.join(out_dir, 'ref_pred.npy'), pred)
    dist.destroy_process_group()


def test_two_rank_score_and_predict_plumbing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref_cm = np.load(tmp_path / 'ref_cm.npy')
    ref_pred = np.load(tmp_path / 'ref_pred.npy')
    for rank in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / ('cm%d.npy' % rank)), ref_cm)
        np.testing.assert_array_equal(np.load(tmp_path / ('pred%d.npy' % rank)), ref_pred)


def _fit_worker(rank, world, port, out_dir):
    """DirichletFusion.fit plumbing: per-rank float64 sufficient statistics + int64 class counts
    summed over ranks, and predict() when one rank has no image at all."""
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from modular_semantic_segmentation_b200 import sharding
    rng = np.random.default_rng(1)
    c, n_img = 4, 5
    logits = rng.normal(size=(2, n_img, 6, 8, c))
    prob = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    labels = rng.integers(-1, c, size=(n_img, 6, 8))
    mine = list(range(rank, n_img, world))
    stats = torch.zeros((2, c, c), dtype=torch.float64)
    counts = torch.zeros(c, dtype=torch.int64)
    for m in range(2):
        s, n = oracle.sufficient_statistics(prob[m][mine], labels[mine], c)
        stats[m] = torch.from_numpy(s)
        counts = torch.from_numpy(n)
    sharding.allreduce_sum_(stats)
    sharding.allreduce_sum_(counts)
    ref = [oracle.sufficient_statistics(prob[m], labels, c) for m in range(2)]
    np.testing.assert_allclose(stats.numpy(), np.stack([r[0] for r in ref]), rtol=1e-12)
    np.testing.assert_array_equal(counts.numpy(), ref[0][1])
    # a single image: rank 1 has nothing to contribute, every rank still gets the full result
    one = torch.from_numpy(labels[:1]) if rank == 0 else None
    gathered = sharding.gather_interleaved(one, 'cpu')
    np.testing.assert_array_equal(gathered.numpy(), labels[:1])
    open(os.path.join(out_dir, 'ok%d' % rank), 'w').write('ok')
    dist.destroy_process_group()


def test_two_rank_dirichlet_fit_statistics_and_empty_rank(tmp_path):
    world = 2
    mp.spawn(_fit_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ('ok%d' % r)).exists() for r in range(world))


def _moments_worker(rank, world, port, out_dir):
    """MC samples split over the ranks (batch-1 latency mode): every rank draws its share of the
    T samples, the per-rank population moments are merged by all-reduce."""
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from modular_semantic_segmentation_b200 import sharding
    rng = np.random.default_rng(2)
    t_total, c = 7, 5                                    # uneven: 4 + 3 samples
    samples = rng.random((t_total, 3, 4, 6, c))
    counts = [sharding.samples_for_rank(t_total, r, world) for r in range(world)]
    assert counts == [4, 3] and sum(counts) == t_total
    lo = sum(counts[:rank])
    mine = samples[lo:lo + counts[rank]]
    mean = torch.from_numpy(mine.mean(0))
    var = torch.from_numpy(mine.var(0))
    total = sharding.combine_moments_(mean, var, counts[rank])
    assert total == t_total
    np.testing.assert_allclose(mean.numpy(), samples.mean(0), rtol=1e-12)
    np.testing.assert_allclose(var.numpy(), samples.var(0), rtol=1e-9, atol=1e-15)
    open(os.path.join(out_dir, 'ok%d' % rank), 'w').write('ok')
    dist.destroy_process_group()


def test_two_rank_mc_sample_split_moments(tmp_path):
    world = 2
    mp.spawn(_moments_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ('ok%d' % r)).exists() for r in range(world))

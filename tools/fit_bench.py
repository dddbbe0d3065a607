"""Data-parallel fit() throughput of the VGG16-FCN depth stream (BASELINE configs[4]).

    python tools/fit_bench.py [--steps 10] [--batch 16]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/fit_bench.py # N GPUs, NCCL

Synthetic Cityscapes-shaped data (384x768 depth maps, 12 classes), Adam lr 1e-4, batch `--batch`
per GPU (weak scaling).  Every step = forward + backward + one flat-gradient all-reduce + Adam.
Timed with CUDA events (max over ranks).  Prints one JSON line on rank 0.
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=16)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from modular_semantic_segmentation_b200 import device as dev, sharding
    from xview.models import get_model
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    H, W, C = 384, 768, 12
    desc = ({'depth': np.float32, 'labels': np.int32},
            {'depth': (None, None, 1), 'labels': (None, None)}, C)
    net = get_model('fcn')('depth', desc, 'depth', num_units=64, batch_normalization=False,
                           learning_rate=1e-4, batchsize=args.batch, seed=1, shard_images=False)
    expert = net._experts['depth']
    expert.train_begin()
    g = torch.Generator(device='cuda').manual_seed(rank)
    x = torch.rand((args.batch, H, W, 1), device='cuda', generator=g) * 100.0
    labels = torch.randint(-1, C, (args.batch, H, W), device='cuda', generator=g, dtype=torch.int32)
    grads = loss = None

    def step():
        nonlocal grads, loss
        grads, loss = expert.train_gradients(x, labels, normalize=False, grads=grads, loss=loss)
        sharding.allreduce_sum_(grads)
        sharding.allreduce_sum_(loss)
        dev.scale_by_count(grads, loss)
        expert.adam_step(grads, learning_rate=1e-4)

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    l = loss.cpu().numpy()
    if rank == 0:
        print(json.dumps({'metric': 'fit_frames_per_s_384x768_depth_stream',
                          'value': world * args.batch * args.steps / (ms * 1e-3),
                          'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
                          'ms_per_step': ms / args.steps, 'scaling': 'weak', 'dtype': 'bf16',
                          'global_batch': world * args.batch, 'params': expert.num_params,
                          'allreduce_bytes_per_step': expert.num_params * 4 + 16,
                          'final_loss': float(l[0] / (1e-20 + l[1]))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

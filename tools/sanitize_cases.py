"""Small cases, one per kernel family, for compute-sanitizer (tools/sanitize.sh).

    python tools/sanitize_cases.py [family ...]     # default: all families

Families: conv2cta (CTA-pair cta_group::2 kernel, odd last tile pair), convt (transposed-role
kernel with and without the fused pool), conv1 (conv1_1 operand packing), fcn (whole expert incl.
decoder, MC dropout), tails (row-pair conv1_2 + pool1, staged MC decode, vector dropout, fused
Bayes / Dirichlet score tails), wgrad (training step: tensor-core weight gradient, pool/ReLU backward,
optimizers), fusion (softmax, Bayes, Dirichlet fast + exact, average, variance, moments,
sufficient statistics), confusion.  Shapes are tiny so that the ~50x slowdown of the tools stays
in seconds; every case still covers ragged tiles."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modular_semantic_segmentation_b200 import device as dev  # noqa: E402
from modular_semantic_segmentation_b200.models.simple_fcn import build_expert  # noqa: E402


def conv_case(cin, cout, h, w, n=1, k=3):
    rng = np.random.default_rng(cin + cout)
    x = torch.from_numpy(rng.standard_normal((n, h, w, cin)).astype(np.float32)).cuda()
    kern = (rng.standard_normal((k, k, cin, cout)) * 0.05).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    return dev.conv2d(x, kern, bias, relu=True, precision='bf16')


def conv2cta():
    conv_case(128, 256, 40, 24)        # 3 pixel tiles per image: odd last pair
    conv_case(256, 512, 16, 16)


def convt():
    conv_case(64, 64, 48, 40)          # ragged 16x16 tiles
    conv_case(64, 128, 32, 32)


def conv1():
    conv_case(3, 64, 40, 56)
    conv_case(1, 64, 32, 48)


def expert(cin=3, nu=8, c=5):
    e, variables = build_expert('m', cin, nu, c, rng=np.random.default_rng(0))
    e.set_params({k[2:]: v for k, v in variables.items()})
    return e


def fcn():
    e = expert()
    x = torch.rand((2, 32, 48, 3), device='cuda')
    e.forward(x, want=('prob', 'label', 'score'))
    e.forward(x, want=('label',), label_dtype=torch.uint8)
    e.forward(x, want=('prob', 'mean_prob', 'var_prob', 'mean_var'),
              dropout={'rate': 0.3, 'layers': ['pool3', 'conv4_3', 'features'], 'num_samples': 3,
                       'seed': 1, 'with_deterministic': True})
    e.close()


def tails():
    """Round-2 kernels: row-pair conv1_2 + pool1 (height not a multiple of its 32-row tile),
    8-element dropout + staged MC decode (T > 8: two staging chunks), the fused Bayes and
    Dirichlet score tails."""
    c = 5
    experts = [expert(3, 8, c), expert(1, 8, c)]
    g = torch.Generator(device='cuda').manual_seed(2)
    xs = [torch.rand((2, 48, 32, 3), device='cuda', generator=g),
          torch.rand((2, 48, 32, 1), device='cuda', generator=g)]
    gt = torch.randint(-1, c, (2, 48, 32), device='cuda', generator=g, dtype=torch.int32)
    cm = torch.zeros((c, c), dtype=torch.int64, device='cuda')
    for e, x in zip(experts, xs):
        e.forward(x, want=('mean_prob', 'var_prob', 'mean_var'),
                  dropout={'rate': 0.5, 'layers': ['pool3'], 'num_samples': 11, 'seed': 3})
        e.forward(x, want=())
    lut = torch.randint(0, c, (c, c), device='cuda', generator=g, dtype=torch.int32)
    dev.bayes_decode_score(experts, lut, c, gt, cm, want_fused=True)
    am1 = torch.rand((2, c, c), device='cuda', generator=g) * 3
    norm = torch.rand((2, c), device='cuda', generator=g)
    prior = torch.log(torch.full((c,), 1.0 / c, device='cuda'))
    mags = dev.dirichlet_table_magnitudes(am1, norm, prior)
    dev.dirichlet_decode_score(experts, am1, norm, prior, c, mags, gt_labels=gt, cm=cm)
    dev.dirichlet_decode_score(experts, am1, norm, prior, c, mags, exact=False,
                               label_dtype=torch.uint8)
    assert int(cm.sum()) == 2 * int((gt >= 0).sum())
    for e in experts:
        e.close()


def wgrad():
    e = expert()
    e.train_begin()
    x = torch.rand((2, 32, 48, 3), device='cuda')
    labels = torch.randint(-1, 5, (2, 32, 48), device='cuda', dtype=torch.int32)
    for trainer in ('adam', 'adagrad', 'rmsprop'):
        grads, _ = e.train_gradients(x, labels)
        e.optimizer_step(grads, trainer, 1e-3)
    e.close()


def fusion():
    g = torch.Generator(device='cuda').manual_seed(0)
    for c in (12, 13):
        shape = (2, 37, 29, c)
        probs = [torch.softmax(torch.randn(shape, device='cuda', generator=g), -1) for _ in range(2)]
        dev.softmax_argmax(torch.randn(shape, device='cuda', generator=g))
        labels = [torch.randint(0, c, shape[:-1], device='cuda', generator=g) for _ in range(2)]
        lut = torch.randint(0, c, (c, c), device='cuda', generator=g, dtype=torch.int32)
        dev.bayes_fuse_lut(labels, lut, c)
        dev.bayes_fuse_lut([t.to(torch.uint8) for t in labels], lut, c)
        cond = torch.randn((2, c, c), device='cuda', generator=g)
        dev.bayes_fuse_score(labels, cond, cond[0, 0].contiguous())
        am1 = torch.rand((2, c, c), device='cuda', generator=g) * 3
        norm = torch.rand((2, c), device='cuda', generator=g)
        prior = torch.log(torch.full((c,), 1.0 / c, device='cuda'))
        dev.dirichlet_fuse(probs, am1, norm, prior, want_score=True)
        dev.dirichlet_fuse(probs, am1, norm, prior, exact=True)
        dev.dirichlet_fuse(probs, am1, norm, prior, want_score=True, exact=True)
        dev.average_fuse(probs, want_score=True)
        var = [torch.rand(shape[:-1], device='cuda', generator=g) for _ in range(2)]
        dev.variance_fuse(probs, var, want_score=True)
        dev.mc_moments(torch.stack(probs + probs), want=('mean', 'var', 'mean_var', 'entropy',
                                                           'cond_entropy', 'sum_var'))
        dev.dirichlet_fit_samples(torch.stack(probs + probs + probs), maxiter=20)
        stats = torch.zeros((c, c), dtype=torch.float64, device='cuda')
        counts = torch.zeros(c, dtype=torch.int64, device='cuda')
        gt = torch.randint(-1, c, shape[:-1], device='cuda', generator=g, dtype=torch.int32)
        dev.dirichlet_suffstats(probs[0], gt, stats, counts)


def confusion():
    g = torch.Generator(device='cuda').manual_seed(1)
    c = 12
    for npix in (5000, 4096 * 3):
        pred = torch.randint(0, c, (npix,), device='cuda', generator=g)
        gt = torch.randint(-1, c, (npix,), device='cuda', generator=g, dtype=torch.int32)
        cm = torch.zeros((c, c), dtype=torch.int64, device='cuda')
        dev.confusion_accumulate(pred, gt, cm)
        dev.confusion_accumulate(pred.to(torch.uint8), gt, cm)
        assert int(cm.sum()) == 2 * int((gt >= 0).sum())


FAMILIES = {'conv2cta': conv2cta, 'convt': convt, 'conv1': conv1, 'fcn': fcn, 'tails': tails,
            'wgrad': wgrad,
            'fusion': fusion, 'confusion': confusion}

if __name__ == '__main__':
    dev.init()
    for name in (sys.argv[1:] or list(FAMILIES)):
        FAMILIES[name]()
        torch.cuda.synchronize()
        print('case %s done' % name, flush=True)

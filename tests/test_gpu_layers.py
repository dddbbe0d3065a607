"""GPU parity: single layers (tcgen05 implicit-GEMM conv, fp32 validation conv, transposed conv,
pooling) against the oracle, through the C ABI."""
import numpy as np
import pytest
import torch

import oracle
from util import bf16_round, cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from modular_semantic_segmentation_b200 import device
    device.init()
    return device


def _conv_case(rng, n, h, w, cin, cout, k, in_scale=1.0):
    x = (rng.standard_normal((n, h, w, cin)) * in_scale).astype(np.float32)
    kern = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    return x, kern, bias


BF16_CASES = [
    # n, h, w, cin, cout, k
    (1, 16, 8, 64, 64, 3),        # exactly one 128-pixel tile
    (1, 16, 16, 64, 64, 3),       # exactly one 256-pixel tile of the transposed-role kernel
    (2, 40, 24, 128, 128, 3),     # transposed-role kernel, ragged 16x16 tiles, 2 K chunks
    (2, 32, 48, 64, 64, 3),       # conv1_2-like: un-pooled row-pair kernel, one 32-row tile
    (1, 41, 23, 64, 64, 3),       # same, odd height and width: ragged 32 x 16 tiles
    (3, 80, 40, 64, 64, 3),       # same, three row tiles with a ragged last one
    (1, 24, 40, 64, 128, 3),      # ragged tiles in both directions
    (1, 16, 24, 128, 256, 3),     # BLOCK_N = 256
    (2, 8, 16, 256, 512, 3),      # two N blocks
    (1, 6, 3, 512, 512, 3),       # tiny 1/16-resolution map (conv5-like), many K blocks
    (1, 12, 6, 512, 8, 1),        # 1x1 head, fp32 epilogue, Cout < 64
    (1, 12, 6, 512, 64, 1),       # 1x1 head, bf16 epilogue
]


@pytest.mark.parametrize('n,h,w,cin,cout,k', BF16_CASES)
@pytest.mark.parametrize('relu', [True, False])
def test_conv2d_tcgen05_matches_oracle(dev, n, h, w, cin, cout, k, relu):
    rng = np.random.default_rng(hash((n, h, w, cin, cout, k)) % 2**31)
    x, kern, bias = _conv_case(rng, n, h, w, cin, cout, k)
    xb, kb = bf16_round(x), bf16_round(kern)
    ref = oracle.conv2d(xb, {'l/kernel': kb, 'l/bias': bias}, 'l', activation=relu)
    got = dev.conv2d(cuda(xb), kb, bias, relu=relu, precision='bf16').cpu().numpy()
    scale = np.abs(ref).max()
    # identical bf16 operands, fp32 accumulation; the bf16 epilogue rounds the result once
    tol = (2.0 ** -8 if cout % 64 == 0 else 1e-5) * scale
    np.testing.assert_allclose(got, ref, rtol=0, atol=tol)
    if cout >= 256:
        # nine-shifted-tiles variant of the single-CTA kernel (debug bit6): same numbers up to the
        # summation order (tap-major there, (chunk, column shift, row shift) in the halo variant)
        # and one bf16 rounding
        dev.set_debug_flags(64)
        alt = dev.conv2d(cuda(xb), kb, bias, relu=relu, precision='bf16').cpu().numpy()
        # single-CTA halo kernel (debug bit7) vs the default CTA-pair (cta_group::2) kernel: same
        # operands, same summation order
        dev.set_debug_flags(128)
        alt3 = dev.conv2d(cuda(xb), kb, bias, relu=relu, precision='bf16').cpu().numpy()
        dev.set_debug_flags(0)
        np.testing.assert_allclose(alt, got, rtol=0, atol=tol)
        np.testing.assert_array_equal(alt3, got)
    if cout <= 128 and cout % 64 == 0 and k == 3:
        # same layer through the pixel-major kernel (debug bit1) must agree to bf16 rounding
        dev.set_debug_flags(2)
        alt = dev.conv2d(cuda(xb), kb, bias, relu=relu, precision='bf16').cpu().numpy()
        dev.set_debug_flags(0)
        np.testing.assert_allclose(alt, ref, rtol=0, atol=tol)


@pytest.mark.parametrize('cin', [3, 1])
def test_conv1_1_operand_packing_keeps_16_bits(dev, cin):
    """Raw uint8 rgb / uint16 depth inputs survive the hi+lo bf16 split (SURVEY.md 'Hard parts')."""
    rng = np.random.default_rng(cin)
    hi = 255 if cin == 3 else 65535
    x = rng.integers(0, hi + 1, size=(1, 32, 48, cin)).astype(np.float32)
    kern = (rng.standard_normal((3, 3, cin, 64)) / (3 * np.sqrt(cin) * hi)).astype(np.float32)
    bias = np.zeros(64, np.float32)
    kb = bf16_round(kern)
    ref = oracle.conv2d(x.astype(np.float64), {'l/kernel': kb.astype(np.float64),
                                               'l/bias': bias.astype(np.float64)}, 'l')
    got = dev.conv2d(cuda(x), kb, bias, relu=True, precision='bf16').cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=2.0 ** -8 * np.abs(ref).max())
    dev.set_debug_flags(4)          # materialised operand buffer: same numbers
    alt = dev.conv2d(cuda(x), kb, bias, relu=True, precision='bf16').cpu().numpy()
    dev.set_debug_flags(0)
    np.testing.assert_array_equal(alt, got)


@pytest.mark.parametrize('n,h,w,cin,cout,k', [(1, 16, 24, 3, 64, 3), (2, 8, 8, 64, 20, 3),
                                              (1, 9, 7, 17, 5, 1)])
def test_conv2d_fp32_validation_mode(dev, n, h, w, cin, cout, k):
    rng = np.random.default_rng(cin * cout)
    x, kern, bias = _conv_case(rng, n, h, w, cin, cout, k)
    ref = oracle.conv2d(x, {'l/kernel': kern, 'l/bias': bias}, 'l')
    got = dev.conv2d(cuda(x), kern, bias, relu=True, precision='fp32').cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('k,stride', [(4, 2), (16, 8)])
def test_deconv2d_matches_oracle(dev, k, stride):
    rng = np.random.default_rng(k)
    nu = 6
    x = rng.standard_normal((2, 5, 7, nu)).astype(np.float32)
    for kern in (oracle.bilinear_filter((k, k, nu, nu)),
                 rng.standard_normal((k, k, nu, nu)).astype(np.float32)):
        ref = oracle.deconv2d(x, {'l/kernel': kern}, 'l', stride)
        got = dev.deconv2d(cuda(x), kern, stride).cpu().numpy()
        assert got.shape == (2, 5 * stride, 7 * stride, nu)
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)


def test_maxpool(dev):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 8, 12, 10)).astype(np.float32)
    np.testing.assert_array_equal(dev.maxpool2x2(cuda(x)).cpu().numpy(), oracle.max_pool2x2(x))


@pytest.mark.parametrize('n,h,w,c,relu', [(2, 9, 13, 64, True), (3, 16, 8, 5, False),
                                          (1, 24, 24, 320, True)])
def test_batchnorm_training_layer(dev, n, h, w, c, relu):
    """tf.layers.batch_normalization(training=True) + ReLU (custom_layers.py:116,132-134): forward,
    batch statistics, moving-average update and the full backward pass against torch autograd."""
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 100 + c)
    x = (rng.standard_normal((n, h, w, c)) * rng.uniform(0.5, 3, size=c) +
         rng.uniform(-2, 2, size=c)).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, size=c).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, size=c).astype(np.float32)
    dy = rng.standard_normal((n, h, w, c)).astype(np.float32)
    mm0 = rng.standard_normal(c).astype(np.float32)
    mv0 = rng.uniform(0.5, 2, size=c).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    gt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    mean = xt.mean(dim=(0, 1, 2))
    var = xt.var(dim=(0, 1, 2), unbiased=False)
    yt = (xt - mean) * torch.rsqrt(var + 1e-3) * gt + bt
    if relu:
        yt = F.relu(yt)
    (yt * torch.tensor(dy, dtype=torch.float64)).sum().backward()
    mm, mv = cuda(mm0), cuda(mv0)
    y, bmean, bvar = dev.batchnorm_train(cuda(x), cuda(gamma), cuda(beta), relu=relu,
                                         moving_mean=mm, moving_var=mv)
    np.testing.assert_allclose(y.cpu().numpy(), yt.detach().numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(bmean.cpu().numpy(), mean.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bvar.cpu().numpy(), var.detach().numpy(), rtol=1e-4, atol=1e-6)
    count = n * h * w
    np.testing.assert_allclose(mm.cpu().numpy(), 0.99 * mm0 + 0.01 * mean.detach().numpy(),
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mv.cpu().numpy(),
                               0.99 * mv0 + 0.01 * var.detach().numpy() * count / (count - 1),
                               rtol=1e-5, atol=1e-6)
    dx, dgamma, dbeta = dev.batchnorm_train_backward(cuda(x), y, cuda(dy), cuda(gamma), relu=relu)
    scale = np.abs(xt.grad.numpy()).max()
    np.testing.assert_allclose(dx.cpu().numpy(), xt.grad.numpy(), rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(dgamma.cpu().numpy(), gt.grad.numpy(), rtol=1e-4,
                               atol=1e-4 * np.abs(gt.grad.numpy()).max())
    np.testing.assert_allclose(dbeta.cpu().numpy(), bt.grad.numpy(), rtol=1e-4,
                               atol=1e-4 * np.abs(bt.grad.numpy()).max())

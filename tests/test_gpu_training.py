"""GPU parity: SimpleFCN.fit() - gradients and Adam steps against the oracle (torch autograd)."""
import numpy as np
import pytest
import torch

import oracle
from oracle.training import adam_update, loss_and_grads
from util import cuda

pytestmark = pytest.mark.gpu
NU, C = 8, 5


@pytest.fixture(scope='module')
def dev():
    from modular_semantic_segmentation_b200 import device
    device.init()
    return device


def _setup(dev, rng, cin=3, n=2, h=32, w=48):
    params = oracle.glorot_fcn_params('m', cin, NU, C, rng, gain=1.45, bias_scale=0.05)
    net = dev.FcnExpert(cin, NU, C, precision='bf16')
    net.set_params({k.split('/', 1)[1]: v for k, v in params.items()})
    x = rng.uniform(0, 1, size=(n, h, w, cin)).astype(np.float32)
    labels = rng.integers(-1, C, size=(n, h, w)).astype(np.int32)
    return net, params, x, labels


@pytest.mark.parametrize('cin', [3, 1])
def test_gradients_match_autograd(dev, cin):
    rng = np.random.default_rng(40 + cin)
    net, params, x, labels = _setup(dev, rng, cin)
    total = net.train_begin()
    grads, loss = net.train_gradients(cuda(x), cuda(labels))
    grads = grads.cpu().numpy()
    loss = loss.cpu().numpy()
    ref_loss, ref = loss_and_grads(params, 'm', x, labels, C)
    assert loss[1] == (labels >= 0).sum()
    assert abs(loss[0] / loss[1] - ref_loss) < 2e-2 * abs(ref_loss)
    covered = 0
    report = []
    for name, g_ref in ref.items():
        off, size = net.param_span(name.split('/', 1)[1])
        g = grads[off:off + size].reshape(g_ref.shape)
        covered += size
        denom = np.linalg.norm(g_ref) + 1e-12
        rel = np.linalg.norm(g - g_ref) / denom
        cos = float((g * g_ref).sum() / (np.linalg.norm(g) * denom + 1e-20))
        report.append('%s rel=%.3f cos=%.4f' % (name, rel, cos))
        assert cos > 0.98 and rel < 0.2, report[-1]
    assert covered == total
    print('\n'.join(report))


@pytest.mark.parametrize('cin', [3, 1])
def test_gradients_match_bf16_emulating_autograd_tightly(dev, cin):
    """The same comparison against autograd on the graph that restates the device's storage
    precision (bf16 weights / activations / data gradients, fp32 heads): what is left is
    accumulation order and double rounding (measured: 0.3 - 13 % relative L2, against 13 - 20 %
    versus the pure fp32 graph), so every gradient tensor must agree to cosine > 0.99."""
    rng = np.random.default_rng(140 + cin)
    net, params, x, labels = _setup(dev, rng, cin)
    net.train_begin()
    grads = net.train_gradients(cuda(x), cuda(labels))[0].cpu().numpy()
    _, ref = loss_and_grads(params, 'm', x, labels, C, emulate_bf16=True)
    report = []
    worst = 0.0
    for name, g_ref in ref.items():
        off, size = net.param_span(name.split('/', 1)[1])
        g = grads[off:off + size].reshape(g_ref.shape)
        denom = np.linalg.norm(g_ref) + 1e-12
        rel = np.linalg.norm(g - g_ref) / denom
        cos = float((g * g_ref).sum() / (np.linalg.norm(g) * denom + 1e-20))
        report.append('%s rel=%.4f cos=%.5f' % (name, rel, cos))
        worst = max(worst, rel)
    print('\n'.join(report))
    for line in report:
        rel, cos = float(line.split('rel=')[1].split()[0]), float(line.split('cos=')[1])
        assert rel < 0.15 and cos > 0.99, line


def test_tensor_core_weight_gradient_equals_cuda_core_reference(dev):
    """Same bf16 operands, fp32 accumulation: the tcgen05 wgrad kernel and the CUDA-core kernel
    must agree to accumulation-order noise."""
    rng = np.random.default_rng(5)
    net, params, x, labels = _setup(dev, rng, n=3, h=48, w=80)
    net.train_begin()
    g_tc = net.train_gradients(cuda(x), cuda(labels))[0].cpu().numpy()
    dev.set_debug_flags(8)
    g_cc = net.train_gradients(cuda(x), cuda(labels))[0].cpu().numpy()
    dev.set_debug_flags(0)
    for layer in ('conv1_2', 'conv2_1', 'conv3_2', 'conv4_3', 'conv5_1'):
        off, size = net.param_span(layer + '/kernel')
        a, b = g_tc[off:off + size], g_cc[off:off + size]
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-3 * np.abs(b).max(), err_msg=layer)


def test_head_only_training_leaves_encoder_gradients_zero(dev):
    rng = np.random.default_rng(7)
    net, params, x, labels = _setup(dev, rng)
    net.train_begin()
    grads, _ = net.train_gradients(cuda(x), cuda(labels), train_encoder=False)
    grads = grads.cpu().numpy()
    off, size = net.param_span('conv3_2/kernel')
    assert not grads[off:off + size].any()
    off, size = net.param_span('score_conv4/kernel')
    assert grads[off:off + size].any()


def test_adam_steps_follow_the_oracle_and_reduce_the_loss(dev):
    rng = np.random.default_rng(11)
    net, params, x, labels = _setup(dev, rng)
    net.train_begin()
    lr = 1e-4
    p_ref = {k: v.copy() for k, v in params.items()}
    m, v = {}, {}
    losses, ref_losses = [], []
    for step in range(1, 9):
        grads, loss = net.train_gradients(cuda(x), cuda(labels))
        l = loss.cpu().numpy()
        losses.append(l[0] / l[1])
        net.adam_step(grads, learning_rate=lr)
        ref_loss, g_ref = loss_and_grads(p_ref, 'm', x, labels, C)
        ref_losses.append(ref_loss)
        adam_update(p_ref, g_ref, m, v, step, learning_rate=lr)
    assert losses[-1] < losses[0]
    np.testing.assert_allclose(losses, ref_losses, rtol=5e-2)
    flat = net.get_params()
    off, size = net.param_span('score/kernel')
    np.testing.assert_allclose(flat[off:off + size].reshape(1, 1, NU, C), p_ref['m/score/kernel'],
                               atol=2.5 * lr * 5)      # |Adam step| <= ~lr per iteration


def test_simple_fcn_fit_api(tmp_path):
    from xview.models import get_model
    rng = np.random.default_rng(3)
    desc = ({'rgb': np.float32, 'labels': np.int32}, {'rgb': (None, None, 3),
                                                      'labels': (None, None)}, C)
    data = {'rgb': rng.uniform(0, 1, size=(4, 32, 32, 3)).astype(np.float32),
            'labels': rng.integers(0, C, size=(4, 32, 32)).astype(np.int32)}
    with get_model('fcn')('rgb', desc, 'rgb', num_units=NU, batch_normalization=False,
                          learning_rate=1e-3, batchsize=2, seed=5,
                          output_dir=str(tmp_path)) as net:
        before = {k: v.copy() for k, v in net.variables.items()}
        m0, _ = net.score(data)
        net.fit(data, 12, validation_dataset=data, validation_interval=4, output=False)
        assert net.global_step == 12
        assert net.loss_history[-1] < net.loss_history[0]
        assert not np.array_equal(net.variables['rgb/conv1_1/kernel'], before['rgb/conv1_1/kernel'])
        np.testing.assert_array_equal(net.variables['rgb/upscore/kernel'],
                                      before['rgb/upscore/kernel'])      # never trained
        path = net.export_weights()
        assert path.endswith('SimpleFCN_weights_12.npz')
        pred = net.predict({'rgb': data['rgb']})
    # a fresh model loaded from the exported file predicts the same
    with get_model('fcn')('rgb', desc, 'rgb', num_units=NU, batch_normalization=False) as net2:
        net2.import_weights(path, warnings=False)
        np.testing.assert_array_equal(net2.predict({'rgb': data['rgb']}), pred)
        assert net2.global_step == 12


@pytest.mark.parametrize('trainer', ['adagrad', 'rmsprop'])
def test_adagrad_and_rmsprop_steps_follow_the_oracle(dev, trainer):
    """base_model.py:157-159: the other two trainers, TensorFlow 1.x defaults.  The device applies
    its own bf16-path gradient; the update rule itself is checked by replaying the oracle's rule on
    the device's gradients (tight), the trajectory against autograd gradients (loose)."""
    from oracle.training import adagrad_update, rmsprop_update
    rng = np.random.default_rng(21)
    net, params, x, labels = _setup(dev, rng)
    total = net.train_begin()
    lr = 1e-3
    names = [n for n in params if 'upscore' not in n]
    spans = {n: net.param_span(n.split('/', 1)[1]) for n in names}
    flat0 = net.get_params()
    p_dev = {n: flat0[o:o + s].reshape(params[n].shape).copy() for n, (o, s) in spans.items()}
    slot_a, slot_b = {}, {}
    losses = []
    for step in range(4):
        grads, loss = net.train_gradients(cuda(x), cuda(labels))
        l = loss.cpu().numpy()
        losses.append(l[0] / l[1])
        g = grads.cpu().numpy()
        g_dev = {n: g[o:o + s].reshape(params[n].shape) for n, (o, s) in spans.items()}
        net.optimizer_step(grads, trainer, lr)
        if trainer == 'adagrad':
            adagrad_update(p_dev, g_dev, slot_a, learning_rate=lr)
        else:
            rmsprop_update(p_dev, g_dev, slot_a, slot_b, learning_rate=lr)
        flat = net.get_params()
        for n, (o, s) in spans.items():
            np.testing.assert_allclose(flat[o:o + s].reshape(params[n].shape), p_dev[n],
                                       rtol=1e-5, atol=1e-7, err_msg='%s step %d' % (n, step))
    assert sum(s for _, s in spans.values()) == total
    assert losses[-1] < losses[0]


def test_fit_twice_continues_and_writes_summaries(tmp_path):
    """A second fit() continues from the trained weights and optimizer state (the reference keeps
    its session); summaries.jsonl receives loss / accuracy / IoU / additional data sets every
    validation_interval steps (base_model.py:191-251); one-hot labels with all-zero rows
    (negative labels through tf.one_hot, base_model.py:198-201) are ignored, not class 0."""
    import json
    from xview.models import get_model
    rng = np.random.default_rng(13)
    desc = ({'rgb': np.float32, 'labels': np.int32}, {'rgb': (None, None, 3),
                                                      'labels': (None, None)}, C)
    labels = rng.integers(-1, C, size=(4, 32, 32)).astype(np.int32)
    onehot = (labels[..., None] == np.arange(C)).astype(np.int32)      # -1 -> all-zero row
    data = {'rgb': rng.uniform(0, 1, size=(4, 32, 32, 3)).astype(np.float32), 'labels': labels}
    data_onehot = {'rgb': data['rgb'], 'labels': onehot}
    common = dict(num_units=NU, batch_normalization=False, learning_rate=1e-3, batchsize=4, seed=5)
    with get_model('fcn')('rgb', desc, 'rgb', output_dir=str(tmp_path), **common) as net:
        net.fit(data, 6, validation_dataset=data, validation_interval=2, output=False,
                additional_eval_datasets={'extra_set': data})
        first = {k: v.copy() for k, v in net.variables.items()}
        loss_6 = net.loss
        net.fit(data, 6, output=False)
        assert net.global_step == 12
        twelve = {k: v.copy() for k, v in net.variables.items()}
    with get_model('fcn')('rgb', desc, 'rgb', **common) as ref:
        # one-hot labels (all-zero rows for -1) reach the device as the same class ids
        as_ids = ref._to_device({'labels': onehot})['labels'].cpu().numpy()
        np.testing.assert_array_equal(as_ids, labels)
        ref.fit(data_onehot, 12, output=False)       # 12 steps in one go, one-hot labels
        # same trajectory as 6 + 6 steps (kept optimizer state).  Weights are compared through the
        # loss: Adam moves parameters whose gradient is accumulation-order noise by +-lr per step,
        # so element-wise equality of two runs is not a property of the algorithm
        assert abs(ref.loss - net.loss) < 0.03 * abs(ref.loss), (ref.loss, net.loss)
        for name in ('rgb/conv3_2/kernel', 'rgb/score/bias'):
            assert not np.array_equal(first[name], twelve[name])
            moved_ref = ref.variables[name] - first[name]
            moved = twelve[name] - first[name]
            cos = float((moved * moved_ref).sum() /
                        (np.linalg.norm(moved) * np.linalg.norm(moved_ref) + 1e-20))
            assert cos > 0.5, (name, cos)
    records = [json.loads(line) for line in open(tmp_path / 'summaries.jsonl')]
    assert [r['step'] for r in records] == [0, 2, 4]
    assert set(records[0]) == {'step', 'loss', 'accuracy', 'IoU', 'extra_set'}
    with get_model('fcn')('rgb', desc, 'rgb', trainer='rmsprop', **common) as net:
        net.fit(data, 3, output=False)
        assert np.isfinite(net.loss)


def _setup_bn(dev, rng, cin=3, n=2, h=32, w=48):
    params = oracle.glorot_fcn_params('m', cin, NU, C, rng, gain=1.45, bias_scale=0.05,
                                      batchnorm=True)
    net = dev.FcnExpert(cin, NU, C, batchnorm=True, precision='bf16')
    net.set_params({k.split('/', 1)[1]: v for k, v in params.items()})
    x = rng.uniform(0, 1, size=(n, h, w, cin)).astype(np.float32)
    labels = rng.integers(-1, C, size=(n, h, w)).astype(np.int32)
    return net, params, x, labels


@pytest.mark.parametrize('cin', [3, 1])
def test_batchnorm_training_gradients_and_moving_statistics(dev, cin):
    """fit() with batch_normalization=True (custom_layers.py:112-119,127-136 in training mode):
    loss, every gradient tensor (kernels, gamma, beta) and the moving-average update against torch
    autograd on the same graph (oracle.training.loss_and_grads_bn)."""
    from oracle.training import loss_and_grads_bn, moving_average_update
    rng = np.random.default_rng(60 + cin)
    net, params, x, labels = _setup_bn(dev, rng, cin)
    total = net.train_begin()
    grads, loss = net.train_gradients(cuda(x), cuda(labels))
    grads = grads.cpu().numpy()
    loss = loss.cpu().numpy()
    # reference: autograd on the graph that restates the device's storage precision (bf16 encoder
    # tensors).  In a randomly initialised batch-normalised net the gradients are very sensitive
    # to that rounding - the bf16-emulating graph itself differs from the pure fp32 graph by up
    # to 60 % in relative L2 norm at conv1_1 - so fp32 autograd is only a loose cross-check.
    ref_loss, ref, stats = loss_and_grads_bn(params, 'm', x, labels, C, emulate_bf16=True)
    _, ref32, _ = loss_and_grads_bn(params, 'm', x, labels, C)
    assert loss[1] == (labels >= 0).sum()
    assert abs(loss[0] / loss[1] - ref_loss) < 1e-2 * abs(ref_loss), (loss[0] / loss[1], ref_loss)
    report = []
    for name, g_ref in ref.items():
        off, size = net.param_span(name.split('/', 1)[1])
        g = grads[off:off + size].reshape(g_ref.shape)
        denom = np.linalg.norm(g_ref) + 1e-12
        rel = np.linalg.norm(g - g_ref) / denom
        cos = float((g * g_ref).sum() / (np.linalg.norm(g) * denom + 1e-20))
        rel32 = np.linalg.norm(g - ref32[name]) / (np.linalg.norm(ref32[name]) + 1e-12)
        report.append((name, rel, cos, rel32))
    print('\n'.join('%s rel=%.3f cos=%.4f (vs fp32 graph rel=%.3f)' % r for r in report))
    # Measured: a 1e-6 relative perturbation of the input changes the bf16-emulating graph's own
    # conv1_1 gradient by ~40 % (rounding decisions, ReLU masks and pool winners flip and batch
    # norm re-scales the result), so layer-wise agreement deep in the encoder is bounded by that
    # chaos, not by the kernels; the batch-norm kernels themselves are pinned to 1e-5 by
    # test_gpu_layers.py::test_batchnorm_training_layer.  Here: the decoder side (few roundings
    # away from the loss) must be close, everything must point in the same direction.
    for name, rel, cos, rel32 in report:
        if name.endswith('/bias'):
            continue            # batch norm cancels the bias: its true gradient is zero
        scope = name.split('/')[1]
        if scope in ('score', 'upscore'):
            assert cos > 0.97 and rel < 0.25, (name, rel, cos)
        elif scope in ('score_conv4', 'score_conv5', 'upscore_conv5'):
            assert cos > 0.9, (name, rel, cos)
        else:
            assert cos > 0.8, (name, rel, cos)
    # conv biases: batch norm subtracts the batch mean, so their gradient is identically zero;
    # the device writes exact zeros, fp32 autograd returns rounding noise
    for name, g_ref in ref32.items():
        if name.endswith('/bias'):
            off, size = net.param_span(name.split('/', 1)[1])
            assert not grads[off:off + size].any()
            assert np.abs(g_ref).max() < 1e-3 * max(np.abs(ref32[name[:-4] + 'kernel']).max(), 1e-6)
    # moving statistics after this one step
    flat = net.get_params()
    for scope, (mean, var, count) in stats.items():
        mm_ref, mv_ref = moving_average_update(params['m/%s/moving_mean' % scope],
                                               params['m/%s/moving_variance' % scope], mean, var,
                                               count)
        off, size = net.param_span(scope + '/moving_mean')
        np.testing.assert_allclose(flat[off:off + size], mm_ref, rtol=0,
                                   atol=2e-3 * max(np.abs(mean).max(), 1e-3), err_msg=scope)
        off, size = net.param_span(scope + '/moving_variance')
        np.testing.assert_allclose(flat[off:off + size], mv_ref, rtol=2e-3, atol=1e-5,
                                   err_msg=scope)
    assert sum(net.param_span(n.split('/', 1)[1])[1] for n in params
               if 'upscore' not in n or not n.endswith('kernel')) == total


def test_simple_fcn_fit_with_batch_normalization(tmp_path):
    """SimpleFCN.fit() with batch_normalization=True: the loss falls, gamma / beta / moving
    statistics change, and a fresh model importing the exported weights predicts the same
    (test-time graph: batch norm on the moving statistics, folded into the convolutions)."""
    from xview.models import get_model
    rng = np.random.default_rng(23)
    desc = ({'rgb': np.float32, 'labels': np.int32}, {'rgb': (None, None, 3),
                                                      'labels': (None, None)}, C)
    data = {'rgb': rng.uniform(0, 1, size=(4, 32, 32, 3)).astype(np.float32),
            'labels': rng.integers(0, C, size=(4, 32, 32)).astype(np.int32)}
    common = dict(num_units=NU, batch_normalization=True, learning_rate=1e-3, batchsize=4, seed=5)
    with get_model('fcn')('rgb', desc, 'rgb', output_dir=str(tmp_path), **common) as net:
        before = {k: v.copy() for k, v in net.variables.items()}
        net.fit(data, 10, validation_dataset=data, validation_interval=3, output=False)
        assert net.loss_history[-1] < net.loss_history[0]
        for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            name = 'rgb/conv2_1/' + leaf
            assert not np.array_equal(net.variables[name], before[name]), name
        assert not np.array_equal(net.variables['rgb/upscore/gamma'], before['rgb/upscore/gamma'])
        np.testing.assert_array_equal(net.variables['rgb/upscore/kernel'],
                                      before['rgb/upscore/kernel'])
        path = net.export_weights()
        pred = net.predict({'rgb': data['rgb']})
    with get_model('fcn')('rgb', desc, 'rgb', num_units=NU, batch_normalization=True) as net2:
        net2.import_weights(path, warnings=False)
        np.testing.assert_array_equal(net2.predict({'rgb': data['rgb']}), pred)


@pytest.mark.parametrize('n,h,w,cin,cout', [(2, 32, 48, 64, 64), (1, 48, 32, 64, 128),
                                            (2, 16, 32, 128, 256), (1, 16, 16, 512, 512),
                                            (3, 24, 40, 128, 256), (1, 40, 24, 256, 512)])
def test_conv_layer_gradients_within_1e3_of_float64(dev, n, h, w, cin, cout):
    """The backward kernels of one 3x3 layer on their own, away from the chaos of a deep
    random net: weight gradient (tcgen05 kernel and CUDA-core reference, xv_conv2d_weight_gradient)
    and data gradient (the forward conv kernels on the flipped, transposed filter - exactly what
    fit() launches) against float64 autograd of the SAME bf16-rounded operands.  fp32
    accumulation: 1e-3 of the largest entry for dW; dX is stored as bf16 (2^-8 relative)."""
    rng = np.random.default_rng(cin + cout + h)
    bf = lambda a: torch.from_numpy(a).bfloat16().float()          # noqa: E731
    x = bf(rng.standard_normal((n, h, w, cin)).astype(np.float32))
    dy = bf(rng.standard_normal((n, h, w, cout)).astype(np.float32))
    kern = bf((rng.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32))
    # float64 reference
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = kern.double().permute(3, 2, 0, 1).requires_grad_(True)
    out = torch.nn.functional.conv2d(xr, wr, padding=1)
    out.backward(dy.double().permute(0, 3, 1, 2))
    dw_ref = wr.grad.permute(2, 3, 1, 0).numpy()                    # -> HWIO
    dx_ref = xr.grad.permute(0, 2, 3, 1).numpy()
    for tensor_cores in (True, False):
        dw = dev.conv2d_weight_gradient(x.cuda(), dy.cuda(), tensor_cores=tensor_cores)
        np.testing.assert_allclose(dw.cpu().numpy(), dw_ref, rtol=0,
                                   atol=1e-3 * np.abs(dw_ref).max(), err_msg=str(tensor_cores))
    flipped = np.ascontiguousarray(kern.numpy()[::-1, ::-1].transpose(0, 1, 3, 2))
    dx = dev.conv2d(dy.cuda(), flipped, None, relu=False, precision='bf16').cpu().numpy()
    np.testing.assert_allclose(dx, dx_ref, rtol=2.0 ** -8, atol=1e-3 * np.abs(dx_ref).max())


@pytest.mark.parametrize('n,h,w', [(2, 32, 48), (1, 48, 80), (3, 16, 16)])
def test_fused_loss_head_equals_three_kernel_form(dev, n, h, w):
    """loss_lowres_grad_kernel (decode + softmax - onehot + loss + transposed upsampling from the
    1/8-resolution scores, nothing at full resolution in HBM) against the decode / ce_grad /
    upsample8_transpose sequence it replaces (debug bit12): same loss, same gradients up to the
    order of the float32 sums."""
    rng = np.random.default_rng(h * w)
    net, params, x, labels = _setup(dev, rng, n=n, h=h, w=w)
    labels[0, :5] = -1                                   # unlabelled pixels
    net.train_begin()
    g_new, loss_new = net.train_gradients(cuda(x), cuda(labels))
    g_new = g_new.cpu().numpy()
    dev.set_debug_flags(4096)
    g_old, loss_old = net.train_gradients(cuda(x), cuda(labels))
    dev.set_debug_flags(0)
    g_old = g_old.cpu().numpy()
    loss_new, loss_old = loss_new.cpu().numpy(), loss_old.cpu().numpy()     # {sum -log p, #valid}
    assert loss_new[1] == loss_old[1] == (labels >= 0).sum()
    assert abs(loss_new[0] - loss_old[0]) <= 1e-5 * abs(loss_old[0])
    for name in ('score/kernel', 'score/bias', 'score_conv4/kernel', 'conv5_3/kernel',
                 'conv3_1/kernel', 'conv1_1/kernel'):
        off, size = net.param_span(name)
        a, b = g_new[off:off + size], g_old[off:off + size]
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-3 * np.abs(b).max() + 1e-12, err_msg=name)
    off, size = net.param_span('score/bias')
    np.testing.assert_allclose(g_new[off:off + size], g_old[off:off + size], rtol=1e-4, atol=1e-7)
    net.close()

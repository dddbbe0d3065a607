"""Tensor-level Python wrappers over the C ABI.

torch is used for plumbing only: device memory (torch.empty on cuda), the current CUDA stream
and (elsewhere) torch.distributed.  Every function below launches hand-written kernels from
libxview_b200.so through ctypes; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._abi import call

_inited = set()


def init(device=None):
    """Bind the library to a CUDA device (xv_init) once per device."""
    if not torch.cuda.is_available():
        raise _abi.XViewError('no CUDA device: xview_b200 has no CPU fallback')
    if device is None:
        device = torch.cuda.current_device()
    device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    torch.cuda.set_device(idx)
    if idx not in _inited:
        torch.zeros(1, device=device)          # make sure the primary context exists
        call('xv_init', idx)
        _inited.add(idx)
    return torch.device('cuda', idx)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'device wrappers need contiguous CUDA tensors'
    return C.c_void_p(t.data_ptr())


def to_device(a, dtype=None):
    """numpy / torch (any device) -> contiguous CUDA tensor on the current device."""
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        a = a.to(dtype)
    return a.cuda(non_blocking=True).contiguous()


_RAW_DTYPES = {torch.uint8: 0, torch.uint16: 1, torch.int16: 2, torch.int32: 3}


def convert_to_f32(t):
    """uint8 / uint16 / int16 / int32 CUDA tensor -> float32 CUDA tensor of the same shape."""
    init()
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    call('xv_convert_to_f32', ptr(t), _RAW_DTYPES[t.dtype], t.numel(), ptr(out), stream_ptr())
    return out


def _label_bytes(t):
    if t.dtype == torch.int64:
        return 8
    if t.dtype == torch.uint8:
        return 1
    raise TypeError('labels must be int64 or uint8, got %s' % t.dtype)


def _new_label(shape, label_dtype, device):
    return torch.empty(shape, dtype=label_dtype, device=device)


# --------------------------------------------------------------------------- FCN expert
class FcnExpert(object):
    """One VGG16-FCN expert on the device (xv_fcn handle): the `fcn()` of
    xview/models/simple_fcn.py:137-170 plus softmax/argmax of basic_fusion_model.py:21-22."""

    ROLES = {'expert': 0, 'encoder': 1, 'head': 2}

    def __init__(self, cin, num_units, num_classes, batchnorm=False, precision='bf16',
                 role='expert', head_cin=512, arch='fcn'):
        """role 'encoder' / 'head' are the pieces of the mid-level fusion net (fusion_fcn.py):
        one VGG16 tower per modality and one head on the concatenated conv4_3 / conv5_3.
        arch 'adapnet' builds the Adapnet expert (adapnet.py:99-173) behind the same calls."""
        init()
        self.cin, self.num_units, self.num_classes = cin, num_units, num_classes
        self.batchnorm = bool(batchnorm) or arch == 'adapnet'
        # 'decoder': batch norm on the decoder's upscore / score layers only (fusion_fcn head)
        bn_flag = 2 if batchnorm == 'decoder' else int(self.batchnorm)
        self.precision = precision
        self.role = role
        self.arch = arch
        handle = C.c_void_p()
        prec = {'bf16': _abi.XV_PRECISION_BF16, 'fp32': _abi.XV_PRECISION_FP32}[precision]
        if arch == 'adapnet':
            call('xv_adapnet_create', C.byref(handle), cin, num_units, num_classes, prec)
        elif arch == 'fcn':
            call('xv_fcn_create_ex', C.byref(handle), cin, num_units, num_classes,
                 bn_flag, prec, self.ROLES[role], head_cin)
        else:
            raise ValueError('unknown expert architecture %r' % (arch,))
        self._h = handle
        self._dirty = True
        self._train_live = False

    def set_param(self, name, array, keep_train_state=False):
        """keep_train_state: the array IS the current device master copy (pulled after
        training), so the optimizer state stays valid."""
        live = self._train_live and keep_train_state
        array = np.ascontiguousarray(array, dtype=np.float32)
        shape = (C.c_int64 * array.ndim)(*array.shape)
        call('xv_fcn_set_param_host', self._h, name.encode(), array.ctypes.data_as(C.c_void_p),
             shape, array.ndim)
        self._dirty = True
        self._train_live = live       # new weights: the optimizer state starts over

    def set_params(self, params, keep_train_state=False):
        for name, array in params.items():
            self.set_param(name, array, keep_train_state)

    def finalize(self):
        call('xv_fcn_finalize', self._h)
        self._dirty = False

    def forward(self, x, want=('label',), dropout=None, label_dtype=torch.int64):
        """x: float32 CUDA tensor [N,H,W,cin].  want: any of 'score','prob','label',
        'mean_prob','var_prob','mean_var'.  dropout: None or dict(rate, layers, num_samples,
        seed, masks={site: uint8 CUDA tensor}, with_deterministic=False).  With
        `with_deterministic` the dropout-free network runs as an extra leading sample sharing the
        trunk: 'score'/'prob'/'label' are then its [N,...] outputs and the moments cover the
        num_samples dropout passes (variance_mix.py:62-69 in one call).  Returns a dict of CUDA
        tensors."""
        if self._dirty:
            self.finalize()
        assert x.dtype == torch.float32 and x.dim() == 4 and x.shape[3] == self.cin
        x = x.contiguous()
        n, h, w, _ = x.shape
        cfg = None
        t_samples = 1
        keep_alive = []
        if dropout is not None and dropout.get('layers'):
            cfg = _abi.DropoutCfg()
            cfg.rate = float(dropout.get('rate', 0.0))
            cfg.sites = 0
            for site in dropout['layers']:
                cfg.sites |= _abi.DROPOUT_SITES[site]
            cfg.num_samples = int(dropout.get('num_samples', 1))
            cfg.seed = int(dropout.get('seed', 0))
            cfg.flags = (_abi.XV_DROP_FLAG_KEEP_FIRST if dropout.get('with_deterministic')
                         else 0)
            masks = dropout.get('masks') or {}
            for i, site in enumerate(_abi.MASK_ORDER):
                m = masks.get(site)
                if m is not None:
                    m = to_device(m, torch.uint8)
                    keep_alive.append(m)
                    cfg.ext_mask[i] = m.data_ptr()
                else:
                    cfg.ext_mask[i] = None
            t_samples = cfg.num_samples
        lead = cfg is not None and bool(cfg.flags & _abi.XV_DROP_FLAG_KEEP_FIRST)
        b = n if lead else n * t_samples
        c = self.num_classes
        dev = x.device
        out = {}
        o = _abi.FcnOutputs()
        if 'score' in want:
            out['score'] = torch.empty((b, h, w, c), dtype=torch.float32, device=dev)
            o.score = out['score'].data_ptr()
        if 'prob' in want:
            out['prob'] = torch.empty((b, h, w, c), dtype=torch.float32, device=dev)
            o.prob = out['prob'].data_ptr()
        if 'label' in want:
            out['label'] = torch.empty((b, h, w), dtype=label_dtype, device=dev)
            if label_dtype == torch.int64:
                o.label_i64 = out['label'].data_ptr()
            else:
                o.label_u8 = out['label'].data_ptr()
        if t_samples > 1:
            if 'mean_prob' in want:
                out['mean_prob'] = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
                o.mean_prob = out['mean_prob'].data_ptr()
            if 'var_prob' in want:
                out['var_prob'] = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
                o.var_prob = out['var_prob'].data_ptr()
            if 'mean_var' in want:
                out['mean_var'] = torch.empty((n, h, w), dtype=torch.float32, device=dev)
                o.mean_var = out['mean_var'].data_ptr()
        call('xv_fcn_forward', self._h, ptr(x), n, h, w, C.byref(cfg) if cfg is not None else None,
             C.byref(o), stream_ptr())
        self._last_forward_shape = (b, h, w)
        if keep_alive:
            torch.cuda.current_stream().synchronize()
        return out

    # ------------------------------------------------------------------ mid-level fusion net
    def forward_encoder(self, x):
        """Encoder-only handle: runs conv1_1..conv5_3 on x (float32 CUDA [N,H,W,cin])."""
        if self._dirty:
            self.finalize()
        n, h, w, _ = x.shape
        call('xv_fcn_forward_encoder', self._h, ptr(x.contiguous()), n, h, w, stream_ptr())
        self._last_shape = (n, h, w)

    def forward_head(self, towers, want=('label',), label_dtype=torch.int64):
        """Head handle: fuses the towers' last encoder passes (fusion_fcn.py:24-39)."""
        if self._dirty:
            self.finalize()
        n, h, w = towers[0]._last_shape
        c = self.num_classes
        dev = torch.device('cuda', torch.cuda.current_device())
        out = {}
        o = _abi.FcnOutputs()
        if 'score' in want:
            out['score'] = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
            o.score = out['score'].data_ptr()
        if 'prob' in want:
            out['prob'] = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
            o.prob = out['prob'].data_ptr()
        if 'label' in want:
            out['label'] = torch.empty((n, h, w), dtype=label_dtype, device=dev)
            if label_dtype == torch.int64:
                o.label_i64 = out['label'].data_ptr()
            else:
                o.label_u8 = out['label'].data_ptr()
        handles = (C.c_void_p * len(towers))(*[t._h for t in towers])
        call('xv_fcn_forward_head', self._h, C.cast(handles, C.POINTER(C.c_void_p)), len(towers),
             C.byref(o), stream_ptr())
        return out

    # ------------------------------------------------------------------ training
    def train_begin(self):
        """Creates the device-side fp32 master parameters + Adam state; returns their length."""
        if self._dirty:
            self.finalize()
        n = C.c_int64()
        call('xv_fcn_train_begin', self._h, C.byref(n))
        self.num_params = n.value
        self._train_live = True
        return n.value

    def param_span(self, name):
        off, size = C.c_int64(), C.c_int64()
        call('xv_fcn_param_span', self._h, name.encode(), C.byref(off), C.byref(size))
        return off.value, size.value

    def grad_buckets(self):
        """[(start, stop)] ranges of the flat gradient in the order the backward pass completes
        them (the ranges the overlapped all-reduce works on)."""
        offsets = (C.c_int64 * 8)()
        count = C.c_int()
        call('xv_fcn_grad_buckets', self._h, offsets, 8, C.byref(count))
        bounds = [offsets[i] for i in range(count.value + 1)]
        return [(bounds[i], bounds[i + 1]) for i in range(count.value - 1, -1, -1)]

    def train_gradients(self, x, labels, train_encoder=True, normalize=True, grads=None,
                        loss=None, bucket_events=None):
        """One forward + backward pass.  Returns (grads float32 CUDA [num_params],
        loss float64 CUDA [2] = {sum of -log p, number of valid pixels}).  bucket_events: list
        of torch.cuda.Event, one per grad_buckets() entry, recorded on the current stream as
        each bucket becomes final (needs normalize=False)."""
        n, h, w, _ = x.shape
        if grads is None:
            grads = torch.empty(self.num_params, dtype=torch.float32, device=x.device)
        if loss is None:
            loss = torch.empty(2, dtype=torch.float64, device=x.device)
        if bucket_events:
            for e in bucket_events:
                if not e.cuda_event:        # torch creates the handle lazily, on first record
                    e.record()
            arr = (C.c_void_p * len(bucket_events))(*[e.cuda_event for e in bucket_events])
            call('xv_fcn_train_gradients_ex', self._h, ptr(x.contiguous()),
                 ptr(labels.contiguous()), n, h, w, int(train_encoder), int(normalize),
                 ptr(grads), ptr(loss), C.cast(arr, C.POINTER(C.c_void_p)), len(bucket_events),
                 stream_ptr())
        else:
            call('xv_fcn_train_gradients', self._h, ptr(x.contiguous()), ptr(labels.contiguous()),
                 n, h, w, int(train_encoder), int(normalize), ptr(grads), ptr(loss), stream_ptr())
        return grads, loss

    def adam_step(self, grads, learning_rate=1e-4, beta1=0.9, beta2=0.999, epsilon=1e-8):
        call('xv_fcn_adam_step', self._h, ptr(grads), C.c_float(learning_rate), C.c_float(beta1),
             C.c_float(beta2), C.c_float(epsilon), stream_ptr())

    OPTIMIZERS = {'adam': 0, 'adagrad': 1, 'rmsprop': 2}

    def optimizer_step(self, grads, trainer='adam', learning_rate=1e-4):
        """One step of tf.train.{Adam,Adagrad,RMSProp}Optimizer with the TF 1.x defaults
        (base_model.py:157-162)."""
        call('xv_fcn_optimizer_step', self._h, ptr(grads), self.OPTIMIZERS[trainer],
             C.c_float(learning_rate), stream_ptr())

    def get_params(self):
        """Current master parameters as one flat float32 numpy vector."""
        out = np.empty(self.num_params, dtype=np.float32)
        call('xv_fcn_get_params_host', self._h, out.ctypes.data_as(C.c_void_p), out.size,
             stream_ptr())
        return out

    def layer(self, name):
        """Activation of a named layer of the last forward call as float32 numpy NHWC."""
        shape = (C.c_int64 * 4)()
        call('xv_fcn_get_layer_host', self._h, name.encode(), None, 0, shape, stream_ptr())
        out = np.empty(tuple(shape), dtype=np.float32)
        call('xv_fcn_get_layer_host', self._h, name.encode(), out.ctypes.data_as(C.c_void_p),
             out.size, shape, stream_ptr())
        return out

    def close(self):
        if getattr(self, '_h', None):
            _abi.load().xv_fcn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- single layers
def conv2d(x, kernel, bias=None, relu=True, precision='bf16'):
    """custom_layers.py:124-139 without batch norm; x CUDA float32 NHWC, kernel numpy HWIO."""
    init()
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    k, _, cin, cout = kernel.shape
    n, h, w, _ = x.shape
    bias_p = None
    if bias is not None:
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        bias_p = bias.ctypes.data_as(C.c_void_p)
    out = torch.empty((n, h, w, cout), dtype=torch.float32, device=x.device)
    call('xv_conv2d', ptr(x.contiguous()), kernel.ctypes.data_as(C.c_void_p), bias_p, n, h, w, cin,
         cout, k, int(relu), {'bf16': 0, 'fp32': 1}[precision], ptr(out), stream_ptr())
    return out


def conv2d_weight_gradient(x, dy, tensor_cores=True):
    """Gradient wrt the 3x3 kernel of a 'same' convolution (the kernels fit() uses): x
    [N,H,W,cin], dy [N,H,W,cout] CUDA float32 (rounded to bf16 inside) -> dw [3,3,cin,cout]."""
    init()
    n, h, w, cin = x.shape
    cout = dy.shape[-1]
    dw = torch.empty((3, 3, cin, cout), dtype=torch.float32, device=x.device)
    call('xv_conv2d_weight_gradient', ptr(x.contiguous()), ptr(dy.contiguous()), n, h, w, cin, cout,
         1 if tensor_cores else 0, ptr(dw), stream_ptr())
    return dw


def deconv2d(x, kernel, stride, relu=True):
    """custom_layers.py:71-121 without batch norm; kernel numpy [kh,kw,Cout,Cin]."""
    init()
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    k, _, cout, cin = kernel.shape
    n, h, w, _ = x.shape
    out = torch.empty((n, h * stride, w * stride, cout), dtype=torch.float32, device=x.device)
    call('xv_deconv2d', ptr(x.contiguous()), kernel.ctypes.data_as(C.c_void_p), n, h, w, cin, cout,
         k, stride, int(relu), ptr(out), stream_ptr())
    return out


def batchnorm_train(x, gamma, beta, relu=True, moving_mean=None, moving_var=None):
    """tf.layers.batch_normalization(training=True) (+ ReLU) on a float32 CUDA NHWC tensor.
    Returns (y, batch mean, biased batch variance); the optional moving statistics (float32
    CUDA [c]) are updated in place."""
    init()
    n, h, w, c = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    var = torch.empty(c, dtype=torch.float32, device=x.device)
    call('xv_batchnorm_train', ptr(x.contiguous()), ptr(gamma), ptr(beta), n, h, w, c, int(relu),
         ptr(y), ptr(mean), ptr(var), ptr(moving_mean), ptr(moving_var), stream_ptr())
    return y, mean, var


def batchnorm_train_backward(x, y, dy, gamma, relu=True):
    """Backward of batchnorm_train: returns (dx, dgamma, dbeta)."""
    init()
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    call('xv_batchnorm_train_backward', ptr(x.contiguous()), ptr(y), ptr(dy.contiguous()),
         ptr(gamma), n, h, w, c, int(relu), ptr(dx), ptr(dgamma), ptr(dbeta), stream_ptr())
    return dx, dgamma, dbeta


def maxpool2x2(x):
    init()
    n, h, w, c = x.shape
    out = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
    call('xv_maxpool2x2', ptr(x.contiguous()), n, h, w, c, ptr(out), stream_ptr())
    return out


# --------------------------------------------------------------------------- fusion stage
def softmax_argmax(score, want_prob=True, label_dtype=torch.int64):
    init()
    c = score.shape[-1]
    npix = score.numel() // c
    prob = torch.empty_like(score) if want_prob else None
    label = _new_label(score.shape[:-1], label_dtype, score.device)
    call('xv_softmax_argmax', ptr(score), npix, c, ptr(prob), ptr(label), _label_bytes(label),
         stream_ptr())
    return prob, label


def bayes_fuse_lut(labels, lut, num_classes):
    """labels: list of int64/uint8 CUDA tensors of equal shape; lut: int32 CUDA tensor [C]*M."""
    init()
    arr, keep = _abi.ptr_array([t.data_ptr() for t in labels])
    out = torch.empty_like(labels[0])
    call('xv_bayes_fuse_lut', arr, len(labels), _label_bytes(labels[0]), ptr(lut), num_classes,
         labels[0].numel(), ptr(out), stream_ptr())
    del keep
    return out


def bayes_decode_score(experts, lut, num_classes, gt_labels, cm, want_fused=False):
    """One kernel behind the experts' last forward calls: per-expert label decode -> decision
    table -> confusion-matrix accumulation into cm (xv_bayes_decode_score).  experts: FcnExpert
    handles whose forward() ran on the batch; gt_labels int32 CUDA [N,H,W]; returns the fused
    uint8 label map if `want_fused`, else None."""
    init()
    assert gt_labels.dtype == torch.int32 and cm.dtype == torch.int64
    handles = (C.c_void_p * len(experts))(*[e._h for e in experts])
    fused = (torch.empty(gt_labels.shape, dtype=torch.uint8, device=gt_labels.device)
             if want_fused else None)
    call('xv_bayes_decode_score', C.cast(handles, C.POINTER(C.c_void_p)), len(experts), ptr(lut),
         num_classes, ptr(gt_labels.contiguous()), ptr(cm), ptr(fused), stream_ptr())
    return fused


def bayes_fuse_score(labels, log_cond, log_prior, want_score=True):
    """log_cond float32 CUDA [M,C,C], log_prior float32 CUDA [C]."""
    init()
    c = log_prior.numel()
    arr, keep = _abi.ptr_array([t.data_ptr() for t in labels])
    score = (torch.empty(labels[0].shape + (c,), dtype=torch.float32, device=labels[0].device)
             if want_score else None)
    out = torch.empty_like(labels[0])
    call('xv_bayes_fuse_score', arr, len(labels), _label_bytes(labels[0]), ptr(log_cond),
         ptr(log_prior), c, labels[0].numel(), ptr(score), ptr(out), stream_ptr())
    del keep
    return score, out


def dirichlet_table_magnitudes(alpha_m1, log_norm, log_prior):
    """The two table magnitudes that scale the |fast - exact| bound of xv_dirichlet_fuse_exact:
    max_c sum_{m,k} |alpha_m1[m][k][c]| and max_c sum_m |log_norm[m][c]| + max_c |log_prior[c]|
    (accepts numpy arrays or tensors; evaluated on the host)."""
    a, n, p = (np.abs(np.asarray(t.detach().cpu() if isinstance(t, torch.Tensor) else t,
                                 dtype=np.float64)) for t in (alpha_m1, log_norm, log_prior))
    return float(a.sum(axis=(0, 1)).max()), float(n.sum(axis=0).max() + p.max())


def dirichlet_fuse(probs, alpha_m1, log_norm, log_prior, want_score=False,
                   label_dtype=torch.int64, exact=False, magnitudes=None, num_exact=None):
    """dirichlet_mix.py:14-36 on the device.  exact=True: argmax (and score) bit-exact against
    the fixed-order float32 statement (oracle.dirichlet_fusion_f32); `magnitudes` =
    dirichlet_table_magnitudes(...) (computed here if omitted - a device->host copy of the
    tables); `num_exact`: optional int64 CUDA scalar that counts the re-evaluated pixels."""
    init()
    c = probs[0].shape[-1]
    npix = probs[0].numel() // c
    arr, keep = _abi.ptr_array([t.data_ptr() for t in probs])
    score = torch.empty_like(probs[0]) if want_score else None
    label = _new_label(probs[0].shape[:-1], label_dtype, probs[0].device)
    if exact:
        if magnitudes is None:
            magnitudes = dirichlet_table_magnitudes(alpha_m1, log_norm, log_prior)
        call('xv_dirichlet_fuse_exact', arr, len(probs), ptr(alpha_m1), ptr(log_norm),
             ptr(log_prior), c, npix, C.c_float(magnitudes[0]), C.c_float(magnitudes[1]),
             ptr(score), ptr(label), _label_bytes(label), ptr(num_exact), stream_ptr())
    else:
        call('xv_dirichlet_fuse', arr, len(probs), ptr(alpha_m1), ptr(log_norm), ptr(log_prior),
             c, npix, ptr(score), ptr(label), _label_bytes(label), stream_ptr())
    del keep
    return score, label


def dirichlet_decode_score(experts, alpha_m1, log_norm, log_prior, num_classes, magnitudes,
                           gt_labels=None, cm=None, want_label=True, label_dtype=torch.int64,
                           exact=True, num_exact=None):
    """One kernel behind the two experts' last forward calls: decoder tail (x8 upsampling +
    softmax) of both -> Dirichlet fusion (bit-exact argmax if `exact`) -> fused labels and / or
    confusion-matrix accumulation into cm (xv_dirichlet_decode_score).  No probability tensor is
    written.  Returns the label map (or None)."""
    init()
    handles = (C.c_void_p * len(experts))(*[e._h for e in experts])
    label = None
    if want_label:
        n, h, w = experts[0]._last_forward_shape
        label = torch.empty((n, h, w), dtype=label_dtype, device=alpha_m1.device)
    amax, tail = magnitudes if exact else (-1.0, 0.0)
    call('xv_dirichlet_decode_score', C.cast(handles, C.POINTER(C.c_void_p)), len(experts),
         ptr(alpha_m1), ptr(log_norm), ptr(log_prior), num_classes, C.c_float(amax),
         C.c_float(tail), ptr(gt_labels), ptr(cm), ptr(label),
         _label_bytes(label) if label is not None else 8, ptr(num_exact), stream_ptr())
    return label


def average_fuse(probs, want_score=False, label_dtype=torch.int64):
    init()
    c = probs[0].shape[-1]
    npix = probs[0].numel() // c
    arr, keep = _abi.ptr_array([t.data_ptr() for t in probs])
    score = torch.empty_like(probs[0]) if want_score else None
    label = _new_label(probs[0].shape[:-1], label_dtype, probs[0].device)
    call('xv_average_fuse', arr, len(probs), c, npix, ptr(score), ptr(label), _label_bytes(label),
         stream_ptr())
    del keep
    return score, label


def variance_fuse(probs, variances, want_score=False, label_dtype=torch.int64):
    init()
    c = probs[0].shape[-1]
    npix = probs[0].numel() // c
    arr, keep = _abi.ptr_array([t.data_ptr() for t in probs])
    varr, vkeep = _abi.ptr_array([t.data_ptr() for t in variances])
    score = torch.empty_like(probs[0]) if want_score else None
    label = _new_label(probs[0].shape[:-1], label_dtype, probs[0].device)
    call('xv_variance_fuse', arr, varr, len(probs), c, npix, ptr(score), ptr(label),
         _label_bytes(label), stream_ptr())
    del keep, vkeep
    return score, label


def mc_moments(samples, want=('mean', 'var', 'mean_var')):
    """samples: float32 CUDA [T, ..., C] -> dict of requested statistics."""
    init()
    t = samples.shape[0]
    c = samples.shape[-1]
    npix = samples[0].numel() // c
    dev = samples.device
    out = {}
    shape_c, shape_1 = samples.shape[1:], samples.shape[1:-1]
    for key in ('mean', 'var'):
        out[key] = torch.empty(shape_c, dtype=torch.float32, device=dev) if key in want else None
    for key in ('mean_var', 'entropy', 'cond_entropy', 'sum_var'):
        out[key] = torch.empty(shape_1, dtype=torch.float32, device=dev) if key in want else None
    call('xv_mc_moments', ptr(samples), t, npix, c, ptr(out['mean']), ptr(out['var']),
         ptr(out['mean_var']), ptr(out['entropy']), ptr(out['cond_entropy']),
         ptr(out['sum_var']), stream_ptr())
    return {k: v for k, v in out.items() if v is not None}


def dirichlet_fit_samples(samples, tol=1e-7, maxiter=100, want_iterations=False):
    """samples: float32 CUDA [T, ..., C] softmax samples -> per-pixel Dirichlet alpha [..., C]."""
    init()
    t, c = samples.shape[0], samples.shape[-1]
    npix = samples[0].numel() // c
    alpha = torch.empty(samples.shape[1:], dtype=torch.float32, device=samples.device)
    iters = (torch.empty(samples.shape[1:-1], dtype=torch.int32, device=samples.device)
             if want_iterations else None)
    call('xv_dirichlet_fit_samples', ptr(samples), t, npix, c, C.c_float(tol), maxiter, ptr(alpha),
         ptr(iters), stream_ptr())
    return (alpha, iters) if want_iterations else alpha


def dirichlet_uncertainty_fuse(probs, variances, cond_params, log_prior, want_score=False,
                               label_dtype=torch.int64):
    """uncertainty_dirichlet_mix.py:18-52; variances: list of float32 CUDA [..., C]."""
    init()
    c = probs[0].shape[-1]
    npix = probs[0].numel() // c
    arr, keep = _abi.ptr_array([t.data_ptr() for t in probs])
    varr, vkeep = _abi.ptr_array([t.data_ptr() for t in variances])
    score = torch.empty_like(probs[0]) if want_score else None
    label = _new_label(probs[0].shape[:-1], label_dtype, probs[0].device)
    call('xv_dirichlet_uncertainty_fuse', arr, varr, len(probs), ptr(cond_params), ptr(log_prior),
         c, npix, ptr(score), ptr(label), _label_bytes(label), stream_ptr())
    del keep, vkeep
    return score, label


def dirichlet_suffstats(prob, labels, stats, counts):
    """Accumulates into stats (float64 CUDA [C,C]) and counts (int64 CUDA [C])."""
    init()
    c = prob.shape[-1]
    assert labels.dtype == torch.int32 and stats.dtype == torch.float64
    call('xv_dirichlet_suffstats', ptr(prob), ptr(labels), prob.numel() // c, c, ptr(stats),
         ptr(counts), stream_ptr())


def confusion_accumulate(pred, labels, cm):
    """Accumulates into cm (int64 CUDA [C,C]); rows = ground truth, cols = prediction."""
    init()
    assert labels.dtype == torch.int32 and cm.dtype == torch.int64
    call('xv_confusion_accumulate', ptr(pred), _label_bytes(pred), ptr(labels), pred.numel(),
         cm.shape[0], ptr(cm), stream_ptr())


# --------------------------------------------------------------------------- measurement
def set_debug_flags(flags):
    """bit0: never fuse pooling, bit1: never use the transposed conv kernel (tests only)."""
    call('xv_set_debug_flags', int(flags))


def scale_by_count(grads, loss):
    """grads *= 1 / (1e-20 + loss[1]) on the device."""
    call('xv_scale_by_count', ptr(grads), grads.numel(), ptr(loss), stream_ptr())


def launch_count():
    n = C.c_int64()
    call('xv_launch_count', C.byref(n))
    return n.value


def profile_enable(on=True):
    call('xv_profile_enable', int(on))


def profile_read():
    """(device ms, algorithmic FLOPs, launches) summed over the tensor-core conv launches
    recorded since profile_enable(True)."""
    ms, flops, n = C.c_double(), C.c_double(), C.c_int64()
    call('xv_profile_read', C.byref(ms), C.byref(flops), C.byref(n))
    return ms.value, flops.value, n.value

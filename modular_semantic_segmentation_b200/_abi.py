"""ctypes binding of libxview_b200.so (the C ABI declared in include/xview_b200.h).

There is no CPU fallback: if the shared library is missing, loading raises; if no CUDA device is
present every compute entry point fails with the library's own error message.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libxview_b200.so')
HEADER_PATH = os.path.join(HERE, '..', 'include', 'xview_b200.h')

XV_PRECISION_BF16 = 0
XV_PRECISION_FP32 = 1
XV_DROP_POOL3 = 1
XV_DROP_CONV4_3 = 2
XV_DROP_CONV5_3 = 4
XV_DROP_FEATURES = 8
XV_DROP_FLAG_KEEP_FIRST = 1
DROPOUT_SITES = {'pool3': XV_DROP_POOL3, 'conv4_3': XV_DROP_CONV4_3,
                 'conv5_3': XV_DROP_CONV5_3, 'features': XV_DROP_FEATURES}
# order of xv_dropout_cfg.ext_mask
MASK_ORDER = ('pool3', 'pool4', 'conv4_3', 'conv5_3', 'features')


class XViewError(RuntimeError):
    """A non-zero return code of the C ABI, carrying xv_last_error()."""


class DropoutCfg(C.Structure):
    _fields_ = [('rate', C.c_float), ('sites', C.c_uint32), ('num_samples', C.c_int32),
                ('seed', C.c_uint64), ('ext_mask', C.c_void_p * 5), ('flags', C.c_uint32)]


class FcnOutputs(C.Structure):
    _fields_ = [('score', C.c_void_p), ('prob', C.c_void_p), ('label_i64', C.c_void_p),
                ('label_u8', C.c_void_p), ('mean_prob', C.c_void_p), ('var_prob', C.c_void_p),
                ('mean_var', C.c_void_p)]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_Z = C.c_size_t
_PP = C.POINTER(C.c_void_p)

# name -> argtypes; every function returns int except the two noted below
PROTOTYPES = {
    'xv_abi_version': [],
    'xv_last_error': [],
    'xv_init': [_I],
    'xv_device_sm_count': [C.POINTER(_I)],
    'xv_set_debug_flags': [_I],
    'xv_launch_count': [C.POINTER(_L)],
    'xv_profile_enable': [_I],
    'xv_profile_read': [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_L)],
    'xv_bench_conv_igemm': [_I, _I, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_float)],
    'xv_malloc': [_PP, _Z],
    'xv_free': [_P],
    'xv_malloc_host': [_PP, _Z],
    'xv_free_host': [_P],
    'xv_memcpy_h2d': [_P, _P, _Z, _P],
    'xv_memcpy_d2h': [_P, _P, _Z, _P],
    'xv_memset': [_P, _I, _Z, _P],
    'xv_stream_sync': [_P],
    'xv_convert_to_f32': [_P, _I, _L, _P, _P],
    'xv_fcn_create': [_PP, _I, _I, _I, _I, _I],
    'xv_fcn_create_ex': [_PP, _I, _I, _I, _I, _I, _I, _I],
    'xv_adapnet_create': [_PP, _I, _I, _I, _I],
    'xv_fcn_forward_encoder': [_P, _P, _I, _I, _I, _P],
    'xv_fcn_forward_head': [_P, _PP, _I, C.POINTER(FcnOutputs), _P],
    'xv_fcn_destroy': [_P],
    'xv_fcn_set_param_host': [_P, C.c_char_p, _P, C.POINTER(_L), _I],
    'xv_fcn_finalize': [_P],
    'xv_fcn_forward': [_P, _P, _I, _I, _I, C.POINTER(DropoutCfg), C.POINTER(FcnOutputs), _P],
    'xv_fcn_get_layer_host': [_P, C.c_char_p, _P, _Z, C.POINTER(_L), _P],
    'xv_fcn_train_begin': [_P, C.POINTER(_L)],
    'xv_fcn_param_span': [_P, C.c_char_p, C.POINTER(_L), C.POINTER(_L)],
    'xv_fcn_train_gradients': [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    'xv_fcn_train_gradients_ex': [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _PP, _I, _P],
    'xv_fcn_grad_buckets': [_P, C.POINTER(_L), _I, C.POINTER(_I)],
    'xv_fcn_optimizer_step': [_P, _P, _I, C.c_float, _P],
    'xv_scale_by_count': [_P, _L, _P, _P],
    'xv_fcn_adam_step': [_P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P],
    'xv_fcn_get_params_host': [_P, _P, _L, _P],
    'xv_fcn_train_end': [_P],
    'xv_conv2d': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'xv_conv2d_weight_gradient': [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    'xv_deconv2d': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    'xv_maxpool2x2': [_P, _I, _I, _I, _I, _P, _P],
    'xv_batchnorm_train': [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    'xv_batchnorm_train_backward': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    'xv_softmax_argmax': [_P, _L, _I, _P, _P, _I, _P],
    'xv_bayes_fuse_lut': [_PP, _I, _I, _P, _I, _L, _P, _P],
    'xv_bayes_decode_score': [_PP, _I, _P, _I, _P, _P, _P, _P],
    'xv_bayes_fuse_score': [_PP, _I, _I, _P, _P, _I, _L, _P, _P, _P],
    'xv_dirichlet_fuse': [_PP, _I, _P, _P, _P, _I, _L, _P, _P, _I, _P],
    'xv_dirichlet_fuse_exact': [_PP, _I, _P, _P, _P, _I, _L, C.c_float, C.c_float, _P, _P, _I, _P,
                                _P],
    'xv_dirichlet_decode_score': [_PP, _I, _P, _P, _P, _I, C.c_float, C.c_float, _P, _P, _P, _I,
                                  _P, _P],
    'xv_average_fuse': [_PP, _I, _I, _L, _P, _P, _I, _P],
    'xv_variance_fuse': [_PP, _PP, _I, _I, _L, _P, _P, _I, _P],
    'xv_mc_moments': [_P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P],
    'xv_dirichlet_suffstats': [_P, _P, _L, _I, _P, _P, _P],
    'xv_dirichlet_fit_samples': [_P, _I, _L, _I, C.c_float, _I, _P, _P, _P],
    'xv_dirichlet_uncertainty_fuse': [_PP, _PP, _I, _P, _P, _I, _L, _P, _P, _I, _P],
    'xv_confusion_accumulate': [_P, _I, _P, _L, _I, _P, _P],
}

_lib = None


def load():
    """Returns the loaded library (ctypes.CDLL), loading and typing it on first use."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XViewError(
            'libxview_b200.so is not built (%s). Run `python -m '
            'modular_semantic_segmentation_b200.build`; there is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == 'xv_last_error' else C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().xv_last_error()
        raise XViewError(msg.decode() if msg else 'xview_b200 call failed (%d)' % rc)


def call(name, *args):
    """Calls an int-returning ABI function and raises XViewError on failure."""
    check(getattr(load(), name)(*args))


def ptr_array(ptrs):
    """Host array of device pointers (for the `*_host` pointer-list parameters)."""
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return C.cast(arr, _PP), arr

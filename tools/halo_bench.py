"""Transposed-role kernel: halo variant (3 patch copies) vs nine shifted tiles, batch 16."""
import ctypes as C, sys
sys.path.insert(0, '.')
from modular_semantic_segmentation_b200 import _abi, device
device.init()
LAYERS = [('conv1_2+pool', 768, 384, 64, 64, 1024), ('conv2_1', 384, 192, 64, 128, 0),
          ('conv2_2+pool', 384, 192, 128, 128, 1024)]
for name, h, w, cin, cout, pool in LAYERS:
    for dbg in (0, 64):
        device.set_debug_flags(dbg)
        ms = C.c_float()
        _abi.call('xv_bench_conv_igemm', 16, h, w, cin, cout, 3, 5, 512 | pool, C.byref(ms))
        gf = 2.0 * 16 * h * w * cout * 9 * cin / 1e9
        print('%-13s %s  %8.1f us  %7.1f TFLOP/s' % (name, 'nine tiles ' if dbg else 'halo copies',
                                                     ms.value * 1e3, gf / ms.value))
device.set_debug_flags(0)

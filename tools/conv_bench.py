"""Times single tensor-core conv layers of the batch-16 768x384 workload in isolation."""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
from modular_semantic_segmentation_b200 import _abi, device
device.init()
LAYERS = [('conv1_1', 768, 384, 3, 64), ('conv1_2', 768, 384, 64, 64), ('conv2_1', 384, 192, 64, 128), ('conv2_2', 384, 192, 128, 128),
          ('conv3_1', 192, 96, 128, 256), ('conv3_2', 192, 96, 256, 256), ('conv4_2', 96, 48, 512, 512),
          ('conv5_1', 48, 24, 512, 512)]
flags_list = [int(f) for f in sys.argv[1].split(',')] if len(sys.argv) > 1 else [0]
names = sys.argv[2].split(',') if len(sys.argv) > 2 else None
for name, h, w, cin, cout in LAYERS:
    if names and name not in names: continue
    for flags in flags_list:
        ms = C.c_float()
        _abi.call('xv_bench_conv_igemm', 16, h, w, cin, cout, 3, 5, flags, C.byref(ms))
        gf = 2.0 * 16 * h * w * cout * 9 * cin / 1e9
        print('%-8s flags=%2d  %8.1f us  %7.1f TFLOP/s' % (name, flags, ms.value * 1e3, gf / ms.value))

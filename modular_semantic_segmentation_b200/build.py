"""Builds libxview_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object loaded through ctypes)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libxview_b200.so')
SOURCES = ['abi.cu', 'conv_igemm_sm100.cu', 'conv_igemm_t_sm100.cu', 'layers.cu', 'fusion.cu',
           'mc_dirichlet.cu', 'train.cu', 'bn_train.cu', 'decode_dirichlet.cu',
           'conv_wgrad_sm100.cu', 'conv_wgrad_2cta_sm100.cu', 'adapnet.cu', 'adapnet_kernels.cu',
           'conv_c1_sm100.cu', 'conv_igemm_2cta_sm100.cu', 'conv_igemm_rowpair_sm100.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--use_fast_math=false', '-Xptxas', '-v',
         '-I', os.path.join(HERE, '..', 'include')]
FLAGS = [f for f in FLAGS if f != '--use_fast_math=false']


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    """Compile every .cu to an object (in parallel) and link the shared library."""
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(HERE, '..', 'include', 'xview_b200.h'))
    procs = []
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        objs.append(obj)
        stale = force or _newer(path, obj) or any(_newer(h, obj) for h in headers)
        if stale:
            cmd = [NVCC] + FLAGS + ['-c', path, '-o', obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT, text=True)))
    relink = force or not os.path.exists(LIB) or bool(procs)
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % src)
        if verbose:
            sys.stderr.write(out)
        with open(os.path.join(HERE, 'build', src + '.ptxas.log'), 'w') as f:
            f.write(out)
    if relink:
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + \
              ['-cudart', 'static']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

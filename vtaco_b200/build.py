"""Build libvtaco_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m vtaco_b200.build [--force]

The shared library exposes only the C ABI declared in include/vtaco_b200.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libvtaco_b200.so')
SOURCES = ['core.cu', 'pack.cu', 'decoder.cu', 'decoder_tc.cu', 'decoder_tc4.cu', 'decoder_bwd.cu', 'encoder.cu', 'mcubes.cu', 'exchange.cu', 'tactile.cu', 'metrics.cu', 'emd.cu', 'conv3d.cu', 'plane_ops.cu']
# debug / measurement kernels live in their own library (tools/ only), not in the product ABI
BENCH_LIB = os.path.join(HERE, 'lib', 'libvtaco_microbench.so')
BENCH_SOURCES = ['tc_microbench.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2',
              '--fmad=true', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES + BENCH_SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(BENCH_LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h', '.inl'))]
    deps.append(os.path.join(ROOT, 'include', 'vtaco_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(HERE, 'lib', os.path.basename(src)[:-3] + '.o')
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false'] + ['-c', src, '-o', obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append('== %s\n%s' % (os.path.basename(src), out))
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, out))
        objs.append(obj)
    bench_objs = [o for o in objs if os.path.basename(o)[:-2] + '.cu' in BENCH_SOURCES]
    objs = [o for o in objs if o not in bench_objs]
    for target, members in ((LIB, objs), (BENCH_LIB, bench_objs + [o for o in objs if os.path.basename(o) == 'core.o'])):
        cmd = [_nvcc(), '-shared', '-o', target] + members + ['-lcudart']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout)
    with open(os.path.join(HERE, 'lib', 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

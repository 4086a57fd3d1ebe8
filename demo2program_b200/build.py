"""Builds libd2p.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the repo snapshot to the GPU box)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libd2p.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
         '-std=c++17', '-Xcompiler', '-fPIC', '--use_fast_math=false'][:-1]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in _sources():
        obj = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if (not force and os.path.exists(obj) and
                os.path.getmtime(obj) > max(os.path.getmtime(src), *[
                    os.path.getmtime(p) for p in
                    glob.glob(os.path.join(CSRC, '*.cuh')) +
                    glob.glob(os.path.join(HERE, '..', 'include', '*.h'))])):
            continue
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
            ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcuda']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

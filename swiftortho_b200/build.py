"""Builds libswiftortho_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m swiftortho_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libswiftortho_b200.so')
SOURCES = ['api.cpp', 'fasta.cpp', 'host_algos.cpp', 'align.cu', 'search.cu', 'select.cu', 'dedupe.cu', 'orth.cu', 'cluster.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC,-O3,-Wall,-pthread', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _nvcc():
    for p in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), 'include', 'swiftortho_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src + '.o')
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + ['-x', 'cu', '-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append('== %s\n%s' % (src, out))
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-lcudart', '-Xcompiler', '-pthread']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    with open(os.path.join(HERE, 'build', 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print('built', LIB)

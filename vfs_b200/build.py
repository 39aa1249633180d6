"""Build libvfs_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m vfs_b200.build [--force] [--verbose]

The library has no torch / Python dependency; it is loaded with ctypes (vfs_b200/_native.py).
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libvfs_b200.so')
STAMP = os.path.join(LIB_DIR, 'build.stamp')

SOURCES = ['api.cu', 'conv_tc.cu', 'layout.cu', 'stem.cu', 'affinity.cu', 'head.cu', 'xcorr.cu', 'bn.cu', 'wgrad_tc.cu', 'train.cu', 'dense.cu', 'post.cu', 'comm.cu', 'bn_stream.cu', 'siamfc_train.cu', 'linear_mma.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ['../../include/vfs_b200.h']
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, 'rb') as fh:
                h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into one shared library. Returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
            print(' '.join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f'--- nvcc failed for {src} ---\n{out}\n')
        elif verbose or out.strip():
            print(out)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    link = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(link)
    with open(STAMP, 'w') as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))

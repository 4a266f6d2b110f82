"""Build libdreamgnn.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dreamgnn_b200.build [--force] [--verbose]

Outputs dreamgnn_b200/lib/libdreamgnn.so (git-ignored; travels to the GPU box with the snapshot).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
OBJ_DIR = os.path.join(PKG, 'build')
LIB_PATH = os.path.join(LIB_DIR, 'libdreamgnn.so')
INCLUDE = os.path.join(os.path.dirname(PKG), 'include')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC,-fvisibility=hidden', '--expt-relaxed-constexpr',
]


def _nvcc():
    return os.environ.get('NVCC') or shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/dreamgnn.h']:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(f.encode() + b'\0' + fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def kernel_digest(files):
    """Digest of the named csrc files + compile flags: what an ncu capture of one kernel stays valid for."""
    h = hashlib.sha256()
    for f in sorted(files):
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(f.encode() + b'\0' + fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, 'libdreamgnn.sha256')
    digest = _digest()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    extra = ['-Xptxas', '-v'] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *extra, '-I', INCLUDE, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, r in ex.map(compile_one, sources()):
            if verbose or r.returncode:
                sys.stderr.write('== %s ==\n%s%s' % (src, r.stdout, r.stderr))
            if r.returncode:
                raise RuntimeError('nvcc failed on %s' % src)
            objs.append(obj)
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH, *objs,
           '-Xlinker', '--no-undefined', '-cudart', 'shared', '-Xlinker', '-rpath,/usr/local/cuda/lib64',
           '-ldl', '-lrt', '-lpthread']
    # the SHARED CUDA runtime: inside a torch process libcudart.so.12 is already loaded (torch's own copy) and is the
    # one the loader binds; the rpath covers a plain C host. (A static runtime would embed its whole API name table
    # in the artefact that ships to the GPU box.)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    with open(stamp, 'w') as fh:
        fh.write(digest + '\n')
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))

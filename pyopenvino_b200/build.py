"""In-tree build of libb200ov.so (the C-ABI CUDA library) for sm_100a.

    python -m pyopenvino_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting `.so` sits next to this file so it travels to
the GPU box with the repo snapshot.  Objects are cached under `csrc/build/` and rebuilt when a
source or header is newer.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(PKG, 'libb200ov.so')
INCLUDE = os.path.join(os.path.dirname(PKG), 'include')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ARCH + ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
                     '-Xptxas', '-v', '-I', INCLUDE, '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found; libb200ov.so cannot be built')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    m = 0.0
    for d in (CSRC, INCLUDE):
        for f in os.listdir(d):
            if f.endswith(('.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libb200ov.so.  Returns the library path."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr = _deps_mtime()
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        stale = force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr)
        jobs.append((src, obj, stale))

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return ''
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('B200OV_EXTRA_NVCC_FLAGS', '').split() + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for {}:\n{}\n{}'.format(src, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(jobs) or 1)) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for log in logs:
            if log:
                sys.stderr.write(log)
    objs = [j[1] for j in jobs]
    need_link = force or not os.path.isfile(LIB) or any(j[2] for j in jobs) or \
        os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs)
    if need_link:
        cmd = [nvcc] + ARCH + ['-shared', '-o', LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n{}\n{}'.format(r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

"""Build libs3grl_b200.so in-tree with nvcc for sm_100a (and only sm_100a).

    python -m s3grl_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc
cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, '..', 'build')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIB_DIR, 'libs3grl_b200.so')
SOURCES = ['pair.cu', 'expand.cu', 'peer.cu', 'probe.cu', 'pool.cu', 'extract.cu', 'extract_sorted.cu', 'walks.cu', 'plan.cu', 'diffuse.cu', 'sign_full.cu', 'ccn_chain.cu', 'loader.cu', 'head.cu', 'gather.cu', 'gather_sc1_lo.cu', 'gather_sc1_mid.cu',
           'gather_sc1_hi.cu', 'gather_sc2_lo.cu', 'gather_sc2_mid.cu', 'gather_sc2_hi.cu', 'gather_sc8_k2.cu', 'gather_sc8_k3.cu',
           'gather_sc8_k4.cu', 'gather_sc8_k5.cu', 'gather_sc8_k6.cu', 'c_abi.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isfile(cand) or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 's3grl_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu to an object in parallel, link the shared library. Returns its path."""
    if not force and not _stale():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(BUILD, src.replace('.cu', '.log')), 'w') as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)

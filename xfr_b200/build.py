"""Builds libxfr_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libxfr_b200.so')
SOURCES = ['api.cu', 'conv_simt.cu', 'conv_tc.cu', 'conv_tc_pair.cu', 'stages.cu', 'lightcnn.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'xfrb.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: A/B variants of the library (tuning knobs of csrc/conv_tc.cuh) next to the product build:
    python -m xfr_b200.build --define XFRB_PAIRA_EW=12 --out libxfr_b200_ew12.so ; XFRB_LIB=... selects one at run time"""
    lib = LIB if out is None else os.path.join(HERE, out)
    if out is None and not defines and not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    bdir = os.path.join(HERE, 'build' if out is None else 'build_' + os.path.splitext(out)[0])
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('--- nvcc %s\n%s\n' % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-o', lib] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    return lib


if __name__ == '__main__':
    defs = [sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == '--define']
    outs = [sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == '--out']
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs, out=outs[0] if outs else None))

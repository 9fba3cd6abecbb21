"""In-tree build of libscab200.so (nvcc, sm_100a only).  `python -m sca_b200.build [--force] [-v]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libscab200.so')
SOURCES = ['scab_api.cu', 'scab_grid.cu', 'scab_neighbors.cu', 'scab_policy_sampled.cu', 'scab_policy_lp.cu',
           'scab_status.cu', 'scab_halo.cu', 'scab_p2p.cu', 'scab_dubins.cu']
# -fmad=false: the reference's decision scalars are reproduced op by op in fp64; the only fused
# multiply-adds are the explicit ones (NumPy's 3-element dot, the fp32 candidate loop).
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-fmad=false',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '--threads', '0']


def nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'scab200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """`defines` / `out`: A/B builds (e.g. defines=['FG_MIN_CTAS=5'], out='_ab/libscab200_ctas5.so'; load with SCAB_LIB_PATH)."""
    if out is None and not force and not needs_build():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, 'build' if out is None else 'build_' + os.path.basename(out))
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace('.cu', '.o'))
        cmd = [nvcc()] + NVCC_FLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write('== %s\n%s\n' % (src, log))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    lib = LIB if out is None else out
    os.makedirs(os.path.dirname(os.path.abspath(lib)), exist_ok=True)
    subprocess.check_call([nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return lib


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith('--out=')]
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs, out=outs[0] if outs else None))

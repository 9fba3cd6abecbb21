"""TEST INFRASTRUCTURE: build tests/hostsim/libscab200_hostsim.so -- sca_b200/csrc/*.cu compiled for the HOST.

    python tests/hostsim/build.py [--force]

The sources are used as they are; the only transformation is textual and happens on a scratch copy: every kernel launch
`kernel<<<grid, block, smem, stream>>>(args);` becomes `hostsim::launch(hostsim::cfg(grid, block, smem, stream), [=]() {
kernel(args); });` (g++ has no launch syntax).  cuda_host_shim.h supplies the device environment, hostsim_runtime.cpp the
execution engine (one host thread per CUDA thread of a block) and stand-ins for the CUDA runtime calls.  The result
exports the same C ABI as libscab200.so, so tests can drive it through sca_b200.capi (tests/hostsim_util.py)."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'sca_b200', 'csrc')
GEN = os.path.join(HERE, '_gen')
LIB = os.path.join(HERE, 'libscab200_hostsim.so')
SOURCES = ['scab_api.cu', 'scab_grid.cu', 'scab_neighbors.cu', 'scab_policy_sampled.cu', 'scab_policy_lp.cu', 'scab_status.cu',
           'scab_halo.cu', 'scab_p2p.cu', 'scab_dubins.cu']
LAUNCH = re.compile(r'([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);', re.S)
FLAGS = ['-O1', '-g', '-std=c++17', '-ffp-contract=off', '-fPIC', '-w', '-pthread', '-D_GNU_SOURCE', '-I/usr/local/cuda/include',
         '-I' + CSRC, '-I' + HERE, '-include', os.path.join(HERE, 'cuda_host_shim.h')]


def transform(text):
    text = text.replace('extern __shared__', 'extern')      # dynamic shared memory: defined once in hostsim_runtime.cpp
    text = re.sub(r'\b__noinline__\b', '__attribute__((noinline))', text)   # (libstdc++ uses the bare token itself)
    text = re.sub(r'asm\s+volatile\s*\("prefetch[^;]*;[^;]*\);', '(void)0;', text)   # inline PTX prefetch hints: no-ops here
    return LAUNCH.sub(lambda m: 'hostsim::launch(hostsim::cfg(%s), [=]() { %s(%s); });' % (m.group(2), m.group(1), m.group(3)), text)


def deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ('cuda_host_shim.h', 'hostsim_runtime.cpp', 'build.py')]
    return d + [os.path.join(ROOT, 'include', 'scab200.h')]


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in deps()):
        return LIB
    os.makedirs(GEN, exist_ok=True)
    objs, procs = [], []
    for f in os.listdir(CSRC):
        if f.endswith('.cuh'):       # headers may hold launches too (none today) -- same transformation
            open(os.path.join(GEN, f), 'w').write(transform(open(os.path.join(CSRC, f)).read()))
    for src in SOURCES:
        text = transform(open(os.path.join(CSRC, src)).read())
        assert '<<<' not in text, 'untransformed launch in ' + src
        cpp = os.path.join(GEN, src.replace('.cu', '.cpp'))
        open(cpp, 'w').write(text)
        obj = cpp.replace('.cpp', '.o')
        # headers: the scratch copies first (so that a launch inside a header is transformed too)
        procs.append((src, subprocess.Popen(['g++'] + FLAGS[:-4] + ['-I' + GEN, '-I' + CSRC, '-I' + HERE, '-I' + os.path.join(ROOT, 'include'),
                                                                    '-include', os.path.join(HERE, 'cuda_host_shim.h'), '-c', cpp, '-o', obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    rt = os.path.join(GEN, 'hostsim_runtime.o')
    procs.append(('runtime', subprocess.Popen(['g++'] + FLAGS[:-2] + ['-c', os.path.join(HERE, 'hostsim_runtime.cpp'), '-o', rt],
                                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bad = False
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write('== %s\n%s\n' % (name, out[-6000:]))
            bad = True
    if bad:
        raise RuntimeError('hostsim build failed')
    subprocess.check_call(['g++', '-shared', '-pthread', '-Wl,-Bsymbolic', '-o', LIB] + objs + [rt])   # bind the runtime stand-ins locally (torch has libcudart loaded)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))

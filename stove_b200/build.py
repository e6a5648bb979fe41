"""Build the C-ABI library (stove_b200/csrc/libstove_b200.so) with nvcc for sm_100a.

    python -m stove_b200.build [--force]

In-tree on purpose: the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libstove_b200.so')
SOURCES = ['api.cu', 'spn_pack.cu', 'spn_obj.cu', 'spn_bg.cu', 'scene.cu', 'scene_ll.cu', 'scene_ll_bwd.cu', 'gnn.cu', 'dynloop.cu', 'glue.cu', 'lstm_tc.cu', 'enc_head.cu', 'optim.cu', 'microbench.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.inc'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'stove_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, src[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', path, '-o', obj]
        if verbose:
            print(' '.join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out.decode()))
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

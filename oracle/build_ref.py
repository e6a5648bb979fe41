"""ORACLE (test / benchmark infrastructure only): stage the reference's own implementation of the hot path
under oracle/_ref/ so it can travel to the GPU box.

The reference is pure Python: "building" it means copying, UNMODIFIED, the handful of modules the path
imports from /root/reference into the git-ignored directory oracle/_ref/
(listed in .gitignore, not in .gpurunignore: like a compiled .so it ships with the gpurun snapshot but
never enters the history).  `bench.py --impl reference` and the `cpu_baseline` leg then time the stock
`Stove.forward` + backward from there; tests use it as a second pin of the oracle port.

    python oracle/build_ref.py            # no-op when /root/reference is absent (the GPU box)
"""
import os
import shutil
import sys

SRC = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = [
    'model/__init__.py',
    'model/video_prediction/__init__.py', 'model/video_prediction/config.py', 'model/video_prediction/stove.py',
    'model/video_prediction/supair.py', 'model/video_prediction/dynamics.py', 'model/video_prediction/encoder.py',
    'model/spn/__init__.py', 'model/spn/rat_torch.py', 'model/spn/region_graph.py', 'model/spn/probabilistic_models.py',
    'model/utils/__init__.py', 'model/utils/utils.py',
    'LICENSE',
]


def build(verbose=False):
    """-> path of the staged reference, or None when there is nothing to stage from."""
    if not os.path.isdir(os.path.join(SRC, 'model', 'video_prediction')):
        return DST if os.path.isdir(os.path.join(DST, 'model', 'video_prediction')) else None
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
            if verbose:
                print('staged', rel)
    return DST


if __name__ == '__main__':
    print(build(verbose=True))

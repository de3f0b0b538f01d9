"""ORACLE tooling (test / baseline infrastructure, not product code): stage the reference's OWN Python
implementation of the hot path under ``oracle/_ref/`` so that it can travel to the GPU box.

    python -m oracle.build_ref            (also called by __graft_entry__.build() when /root/reference exists)

``/root/reference`` exists only in the authoring container.  ``oracle/_ref/`` is a BUILD OUTPUT: it is
git-ignored (never part of the repository history) but not gpurun-ignored, like the compiled ``.so``.
It holds byte-for-byte copies of the files of the reference that the embedding path executes
(models/, the config + model cfg files, misc/{utils,torch_utils}.py, eval/{pnv_evaluate,utils}.py and the
dataset helpers they import) -- nothing is edited.  They run on the CPU over the stand-ins of
oracle/ocnn_standin.py for the third-party packages that are not installable offline (ocnn, dwconv's CUDA
extension, open3d, matplotlib) and serve as
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg (kind = "reference": the reference's own
    eval loop -- per-submap build_octree, merge_octrees + construct_all_neigh, model forward), and
  * a second pin for oracle/model_ref.py.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('HFL_REFERENCE_ROOT', '/root/reference')

FILES = [
    'LICENSE',
    'misc/utils.py', 'misc/torch_utils.py',
    'eval/pnv_evaluate.py', 'eval/utils.py',
    'datasets/augmentation.py', 'datasets/coordinate_utils.py', 'datasets/base_datasets.py',
    'datasets/pointnetvlad/pnv_raw.py', 'datasets/CSWildPlaces/CSWildPlaces_raw.py',
]
DIRS = ['models', 'config']          # every .py / .txt below


def build(verbose: bool = True) -> str:
    if not os.path.isdir(SRC):
        if os.path.isdir(DST):
            return DST               # GPU box: use the staged copy
        raise FileNotFoundError(f'{SRC} is not available and {DST} has not been staged')
    files = list(FILES)
    for d in DIRS:
        for root, _, names in os.walk(os.path.join(SRC, d)):
            for n in names:
                if n.endswith(('.py', '.txt')):
                    files.append(os.path.relpath(os.path.join(root, n), SRC))
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    if verbose:
        print(f'staged {len(files)} reference files under {DST}')
    return DST


def available() -> bool:
    return os.path.isfile(os.path.join(DST, 'models', 'model_factory.py'))


if __name__ == '__main__':
    build()
    sys.exit(0)

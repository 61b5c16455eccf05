"""Stage the UNMODIFIED reference files of the hot path under `oracle/_ref/` -- TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/stage_ref.py          (also run by __graft_entry__.build() when /root/reference is mounted)

The reference is pure Python (nothing to compile), and `/root/reference` does not exist on the GPU box.
`oracle/_ref/` is git-ignored (no reference source enters the history) but not gpurun-ignored, so the staged
files travel to the box like our own built `.so`.  There `oracle/ref_loader.py` imports them by file path, and
`bench.py --impl reference` / `cpu_baseline` time the reference's own `apgd_train`
(/root/reference/autopgd_train_clean.py:123-371) on the reference's own ConvNeXt-T-CvSt
(/root/reference/models/convnext.py + utils_architecture.py ConvBlock1, timm stubbed: SURVEY F8) on the box's
host cores: `cpu_baseline.kind == "reference"`.  Nothing in the product package reads this directory.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('B200AT_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')
# the files of the path: the attack, the vendored ConvNeXt, the CvSt stems + normaliser
FILES = ('autopgd_train_clean.py', os.path.join('models', 'convnext.py'), 'utils_architecture.py')


def stage(verbose=False):
    """Copy FILES from the mounted reference; returns the staged directory, or None when no reference is mounted
    (the GPU box: it uses what travelled with the snapshot)."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return DST if os.path.isfile(os.path.join(DST, FILES[0])) else None
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
            if verbose:
                print('staged', rel)
    return DST


if __name__ == '__main__':
    print(stage(verbose=True))

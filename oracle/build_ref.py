"""Recipe for oracle/_ref: the UNMODIFIED reference model code, copied from the read-only reference tree so that it can
travel to the GPU box (oracle/_ref/ is git-ignored — no reference source enters the history — but not gpurun-ignored).

Copied: the import closure of `models.detection.yolox_extension.models.detector` and `models.detection.yolox.utils.boxes`
(the whole `models/` tree plus three leaf modules), and this repo's omegaconf stand-in (tests/golden/_refshim/omegaconf, the
reference's only non-installed import on that path).  Used by:
  * bench.py --impl reference      (the reference's own PyTorch path on the host cores, `kind: "reference"`)
  * bench.py --impl reference-gpu  (the same stock-PyTorch path on the B200 — the "real bar" of SURVEY.md §8d)
TEST / MEASUREMENT INFRASTRUCTURE — nothing under leod_b200/ imports it.

    python oracle/build_ref.py            # no-op with a message when /root/reference is absent (GPU box)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('LEOD_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '_ref')
EXTRA = ['data/genx_utils/labels.py', 'data/utils/types.py', 'utils/timers.py']


def build(verbose=True):
    if not os.path.isdir(os.path.join(REF, 'models')):
        if verbose:
            print(f'oracle/build_ref.py: {REF} not present, keeping the existing oracle/_ref ({"found" if os.path.isdir(OUT) else "absent"})')
        return os.path.isdir(OUT)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    shutil.copytree(os.path.join(REF, 'models'), os.path.join(OUT, 'models'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for rel in EXTRA:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(os.path.join(REF, rel), dst)
    shutil.copytree(os.path.join(HERE, '..', 'tests', 'golden', '_refshim', 'omegaconf'), os.path.join(OUT, 'omegaconf'),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    n = sum(len(f) for _, _, f in os.walk(OUT))
    if verbose:
        print(f'oracle/build_ref.py: copied {n} files from {REF} to {OUT}')
    return True


def import_reference():
    """-> (YoloXDetector class, postprocess) of the reference, imported from oracle/_ref.  Raises if it was not built."""
    if not os.path.isdir(os.path.join(OUT, 'models')):
        raise RuntimeError('oracle/_ref is absent: run `python oracle/build_ref.py` where /root/reference exists')
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    from models.detection.yolox_extension.models.detector import YoloXDetector
    from models.detection.yolox.utils.boxes import postprocess
    return YoloXDetector, postprocess


if __name__ == '__main__':
    build()

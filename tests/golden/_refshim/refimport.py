"""Import helper for the read-only reference tree (/root/reference), build container only.

Puts the omegaconf stand-in and the reference on sys.path and registers inert stand-ins for the
third-party packages the reference's *orchestration* modules import but which are absent here
(pytorch_lightning, nerv, h5py, pycocotools, ...). Nothing under tests/ that runs on the GPU box
imports this file; it exists for tests/golden/make_golden.py."""
import os
import sys
import types

REF = os.environ.get('LEOD_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def available():
    return os.path.isdir(os.path.join(REF, 'models'))


def setup(stub_orchestration=True):
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    for p in (REF, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not stub_orchestration:
        return
    import torch

    class LightningModule(torch.nn.Module):
        """Only what modules/detection.py touches outside logging paths."""
        trainer = types.SimpleNamespace(global_step=0, world_size=1)

        @property
        def dtype(self):
            return next(self.parameters()).dtype

        @property
        def device(self):
            return next(self.parameters()).device

        def log_dict(self, *a, **k):
            return None

        def log(self, *a, **k):
            return None

    if 'pytorch_lightning' not in sys.modules:
        pl = _mod('pytorch_lightning', LightningModule=LightningModule)
        _mod('pytorch_lightning.utilities')
        _mod('pytorch_lightning.utilities.types', STEP_OUTPUT=object)
        pl.utilities = sys.modules['pytorch_lightning.utilities']

    class AverageMeter:
        def __init__(self, *a, **k):
            self.sum, self.count = 0., 0

        def update(self, v, n=1):
            self.sum += float(v) * n
            self.count += n

        @property
        def avg(self):
            return self.sum / max(self.count, 1)

    if 'nerv' not in sys.modules:
        _mod('nerv')
        _mod('nerv.utils', AverageMeter=AverageMeter, load_obj=None, dump_obj=None, glob_all=None)
    for name in ('h5py', 'hdf5plugin'):
        if name not in sys.modules:
            _mod(name)
    # pycocotools is imported by utils/evaluation/prophesee/metrics/coco_eval.py, which also calls
    # torch.cuda.get_device_name() at import time: replace that leaf module wholesale.
    if 'utils.evaluation.prophesee.metrics.coco_eval' not in sys.modules:
        import importlib
        importlib.import_module('utils.evaluation.prophesee.metrics')
        _mod('utils.evaluation.prophesee.metrics.coco_eval', evaluate_detection=None)

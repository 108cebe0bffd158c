"""Minimal stand-in for `omegaconf`, used ONLY to import the reference (/root/reference) when
generating golden fixtures in the build container. Test infrastructure, not product code."""
from contextlib import nullcontext


def _wrap(v):
    if isinstance(v, DictConfig):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(_wrap(e) for e in v)
    return v


class ListConfig(list):
    pass


class DictConfig(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            dict.__setitem__(self, k, _wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        dict.__setitem__(self, k, _wrap(v))

    def __setitem__(self, k, v):
        dict.__setitem__(self, k, _wrap(v))


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(e) for k, e in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(e) for e in v]
    return v


class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True, throw_on_missing=True):
        return _plain(cfg)

    @staticmethod
    def is_config(obj):
        return isinstance(obj, (DictConfig, ListConfig))

    @staticmethod
    def create(d=None):
        return _wrap(d or {})

    @staticmethod
    def set_struct(cfg, flag):
        return None


def open_dict(cfg):
    return nullcontext(cfg)

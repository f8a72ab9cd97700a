"""Test-side shim that lets the *unmodified* reference (neu-vi/ezflow) import in
this image, where two of its third-party deps are absent:

* ``fvcore.common.config.CfgNode`` (ezflow/config/config.py:8)
* ``omegaconf.DictConfig``         (ezflow/config/config.py:156)

Only ``tools/make_golden.py`` and the optional ``reference``-marked tests use it.
It is never imported by the product package.
"""

import os
import sys
import types

import yaml

_VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
# EZB_REFERENCE_ROOT pins one location (e.g. the vendored copy, to exercise what the GPU box will see)
REF_CANDIDATES = [os.environ["EZB_REFERENCE_ROOT"]] if os.environ.get("EZB_REFERENCE_ROOT") else ["/root/reference", _VENDORED]


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=True):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode._wrap(v)

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, CfgNode):
            return CfgNode(v)
        return v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = CfgNode._wrap(value)

    def merge_from_file(self, path, allow_unsafe=False):
        with open(path) as f:
            self.merge_from_other_cfg(CfgNode(yaml.safe_load(f)))

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = CfgNode._wrap(v)

    def dump(self, *a, **k):
        return yaml.safe_dump(dict(self))

    def clone(self):
        import copy

        return copy.deepcopy(self)


def install_shim():
    if "fvcore" not in sys.modules:
        fv = types.ModuleType("fvcore")
        fvc = types.ModuleType("fvcore.common")
        fvcc = types.ModuleType("fvcore.common.config")
        fvcc.CfgNode = CfgNode
        fv.common = fvc
        fvc.config = fvcc
        sys.modules.update({"fvcore": fv, "fvcore.common": fvc, "fvcore.common.config": fvcc})
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class DictConfig:  # never instantiated; only used in isinstance checks
            pass

        oc.DictConfig = DictConfig
        sys.modules["omegaconf"] = oc


def reference_root():
    for p in REF_CANDIDATES:
        if os.path.isdir(os.path.join(p, "ezflow")):
            return p
    return None


def import_reference():
    """Returns the imported reference ``ezflow`` package, or None if it is not on this machine."""
    root = reference_root()
    if root is None:
        return None
    install_shim()
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import ezflow  # noqa: F401

    return ezflow


def reference_config_path(name):
    root = reference_root()
    for sub in ("configs/models", "ezflow/model_zoo/configs/models"):
        p = os.path.join(root, sub, name)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)

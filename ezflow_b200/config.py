"""Host-side mirror of the reference's config plumbing for this path.

``configurable`` behaves like ezflow.config.configurable (reference ezflow/config/config.py:49-163):
a class's ``__init__`` may be called with explicit arguments or with a config node as first argument,
in which case the ``from_config`` classmethod translates it.  It recognises the reference's own node
types (fvcore / yacs ``CfgNode``, omegaconf ``DictConfig``) when those packages are loaded, plus the
small ``CfgNode`` below so the package is usable without ezflow installed (the GPU box has no ezflow).
"""

import functools
import inspect
import sys


class CfgNode(dict):
    """Minimal attribute-access config node (stand-in for fvcore's CfgNode)."""

    def __init__(self, init_dict=None, key_list=None, new_allowed=True):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def _cfg_types():
    types = [CfgNode]
    for mod, name in (("fvcore.common.config", "CfgNode"), ("yacs.config", "CfgNode"), ("omegaconf", "DictConfig")):
        m = sys.modules.get(mod)
        t = getattr(m, name, None) if m is not None else None
        if isinstance(t, type):
            types.append(t)
    return tuple(types)


def _called_with_cfg(*args, **kwargs):
    types = _cfg_types()
    if len(args) and isinstance(args[0], types):
        return True
    if isinstance(kwargs.get("cfg", None), types):
        return True
    return False


def _get_args_from_config(from_config_func, *args, **kwargs):
    signature = inspect.signature(from_config_func)
    if list(signature.parameters.keys())[0] != "cfg":
        raise TypeError(f"{from_config_func.__self__}.from_config must take 'cfg' as the first argument!")
    support_var_arg = any(
        p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in signature.parameters.values()
    )
    if support_var_arg:
        return from_config_func(*args, **kwargs)
    supported = set(signature.parameters.keys())
    extra = {k: kwargs.pop(k) for k in list(kwargs.keys()) if k not in supported}
    ret = from_config_func(*args, **kwargs)
    ret.update(extra)
    return ret


def configurable(init_func):
    """The reference's own decorator when ``ezflow`` is already loaded in this process (the drop-in classes then go
    through exactly the reference's config plumbing); otherwise — on a machine without ezflow, e.g. the GPU box's tests that
    never import it — the equivalent below."""
    ref = sys.modules.get("ezflow.config")
    if ref is not None and callable(getattr(ref, "configurable", None)) and "omegaconf" in sys.modules:
        return ref.configurable(init_func)
    assert inspect.isfunction(init_func) and init_func.__name__ == "__init__", "Incorrect use of @configurable."

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        try:
            from_config_func = type(self).from_config
        except AttributeError as e:
            raise AttributeError("Class with @configurable must have a 'from_config' classmethod.") from e
        if not inspect.ismethod(from_config_func):
            raise TypeError("Class with @configurable must have a 'from_config' classmethod.")
        if _called_with_cfg(*args, **kwargs):
            init_func(self, **_get_args_from_config(from_config_func, *args, **kwargs))
        else:
            init_func(self, *args, **kwargs)

    return wrapped

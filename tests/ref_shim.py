"""Import the *reference* PaSST modules from /root/reference without its non-numeric dependencies.

Only used by CPU tests and by tests/golden/make_golden.py in the build container (the GPU box has no
/root/reference).  Nothing is copied: the reference tree is put on sys.path and five absent third-party
module names are stubbed: ba3l.ingredients.ingredient (sacred/munch glue) and
timm.models._hub (checkpoint download).
"""
import contextlib
import io
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("PASST_REF_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "passt.py"))


class _Ingredient:
    def __init__(self, *a, **k):
        pass

    def command(self, fn=None, **k):
        if fn is None:
            return lambda f: f
        return fn

    def add_config(self, *a, **k):
        return None

    config = add_config


def _install_stubs():
    for name in ("ba3l", "ba3l.ingredients"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    ing = types.ModuleType("ba3l.ingredients.ingredient")
    ing.Ingredient = _Ingredient
    sys.modules["ba3l.ingredients.ingredient"] = ing
    for name in ("timm", "timm.models"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    hub = types.ModuleType("timm.models._hub")

    def download_cached_file(*a, **k):
        raise RuntimeError("no network: pretrained weights unavailable")

    hub.download_cached_file = download_cached_file
    sys.modules["timm.models._hub"] = hub


_cache = {}


def load_reference():
    """Returns (ref_passt_module, ref_preprocess_module)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _install_stubs()
    # the reference's top-level package is called `models`; import it under a private alias so that it cannot
    # collide with this repo's drop-in `models` shim
    import importlib.util

    saved = {k: sys.modules.get(k) for k in ("models", "models.helpers", "models.helpers.vit_helpers")}
    sys.path.insert(0, REF_ROOT)
    try:
        for k in list(saved):
            sys.modules.pop(k, None)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import models.passt as ref_passt          # noqa
            import models.preprocess as ref_pre       # noqa
    finally:
        sys.path.remove(REF_ROOT)
        for k in ("models", "models.helpers", "models.helpers.vit_helpers", "models.passt", "models.preprocess"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _cache["mods"] = (ref_passt, ref_pre)
    return _cache["mods"]


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield

"""TEST / MEASUREMENT INFRASTRUCTURE ONLY - imports the unmodified reference modules staged by oracle/build_ref.py.

The repository root holds drop-in shims called `Model.py` / `loss.py` (the reference's own module names), so the
reference's files are imported under PRIVATE names (`_ref_Model`, `_ref_loss`, `_ref_Trainer`) straight from their
file paths; nothing is put on sys.path.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.path.join(HERE, "_ref"), os.environ.get("B200UNET_REFERENCE", "/root/reference")]
_cache = {}


def ref_dir():
    for d in _CANDIDATES:
        if d and os.path.exists(os.path.join(d, "Model.py")) and os.path.exists(os.path.join(d, "loss.py")):
            return d
    return None


def available() -> bool:
    return ref_dir() is not None


def _verify(d):
    """The staged copy must still be the files the recipe hashed."""
    mf = os.path.join(d, "MANIFEST.json")
    if not os.path.exists(mf):
        return
    with open(mf) as f:
        want = json.load(f)["files"]
    for name, digest in want.items():
        with open(os.path.join(d, name), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                raise RuntimeError(f"oracle/_ref/{name} differs from the staged reference file (re-run oracle/build_ref.py)")


def _load(name):
    key = "_ref_" + name
    if key in _cache:
        return _cache[key]
    d = ref_dir()
    if d is None:
        raise RuntimeError("the reference modules are not staged: run `python oracle/build_ref.py` where /root/reference exists")
    _verify(d)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only; keep oracle/_ref/ source-only
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = importlib.util.spec_from_file_location(key, os.path.join(d, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[key] = mod
            spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = old
    _cache[key] = mod
    return mod


def load():
    """(Model, loss) modules of the unmodified reference."""
    return _load("Model"), _load("loss")


def load_trainer(model_module, loss_module):
    """The reference's Trainer.py, unchanged, with `Model` / `loss` resolving to the GIVEN modules (the product shims
    for the integration test) and `matplotlib` stubbed (absent from this image; only used for plots after training)."""
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            from unittest import mock

            mpl = types.ModuleType("matplotlib")
            plt = mock.MagicMock(name="matplotlib.pyplot")
            ax = mock.MagicMock(name="axes")
            ax.get_legend_handles_labels.return_value = ([], [])
            ax.twinx.return_value = ax
            plt.subplots.return_value = (mock.MagicMock(name="figure"), ax)
            mpl.pyplot = plt
            mpl.use = lambda *a, **k: None
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    saved = {k: sys.modules.get(k) for k in ("Model", "loss")}
    sys.modules["Model"], sys.modules["loss"] = model_module, loss_module
    try:
        _cache.pop("_ref_Trainer", None)
        return _load("Trainer")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

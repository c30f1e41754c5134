"""Loader for the UNMODIFIED reference modules (test infrastructure only).

TEST INFRASTRUCTURE - not part of the product path.  Only ``tests/``,
``oracle/make_golden.py``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import this package.

The reference (``/root/reference``) is a flat directory of scripts.  Two
obstacles stand between it and ``import``:

* ``matplotlib`` is imported at module scope (``Training_VAE.py:100``,
  ``Tools.py:7-9``) but is not installed in this image -> empty stub modules
  are placed in ``sys.modules`` first.
* ``Training_VAE`` and ``Tools`` import each other (``Training_VAE.py:102``,
  ``Tools.py:14``); importing ``Tools`` first is the order that works.

The reference exists only in the build container, never on the GPU box, so
everything that must run there goes through ``oracle.vae_oracle`` (the
restatement) and the committed fixtures under ``tests/golden``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DMVAE_REFERENCE_ROOT", "/root/reference")

_BOUNDARY_NAMES = ("Tools", "Training_VAE", "Driver_Models", "Traj_Data_Process")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Training_VAE.py"))


def _stub_matplotlib() -> None:
    if "matplotlib" in sys.modules:
        return
    try:  # a real matplotlib is fine too
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []  # mark as package
    pyplot = types.ModuleType("matplotlib.pyplot")
    pyplot.rcParams = {}
    patches = types.ModuleType("matplotlib.patches")
    animation = types.ModuleType("matplotlib.animation")
    colors = types.ModuleType("matplotlib.colors")

    class _FuncAnimation:  # only the name has to exist at import time
        def __init__(self, *a, **k):
            raise RuntimeError("matplotlib is stubbed in the oracle loader")

    animation.FuncAnimation = _FuncAnimation
    mpl.pyplot, mpl.patches, mpl.animation, mpl.colors = pyplot, patches, animation, colors
    mpl.use = lambda *a, **k: None
    for mod in (mpl, pyplot, patches, animation, colors):
        sys.modules[mod.__name__] = mod


class ReferenceModules:
    """Namespace holding the reference's own modules."""

    def __init__(self, tools, training_vae, driver_models, traj_data_process):
        self.Tools = tools
        self.Training_VAE = training_vae
        self.Driver_Models = driver_models
        self.Traj_Data_Process = traj_data_process


_cached: ReferenceModules | None = None


def load_reference() -> ReferenceModules:
    """Import the reference's modules under their own names, then remove them
    from ``sys.modules`` again so that the repo's same-named drop-in modules
    (``/root/repo/Training_VAE.py`` ...) can still be imported afterwards."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _stub_matplotlib()
    saved = {n: sys.modules.pop(n) for n in _BOUNDARY_NAMES if n in sys.modules}
    saved_path = list(sys.path)
    saved_dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        tools = importlib.import_module("Tools")  # must come first (circular import)
        tvae = importlib.import_module("Training_VAE")
        drv = importlib.import_module("Driver_Models")
        tdp = importlib.import_module("Traj_Data_Process")
        for m in (tools, tvae, drv, tdp):
            assert os.path.dirname(os.path.abspath(m.__file__)) == os.path.abspath(REFERENCE_ROOT), m.__file__
    finally:
        sys.path[:] = saved_path
        sys.dont_write_bytecode = saved_dont_write
        for n in _BOUNDARY_NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    _cached = ReferenceModules(tools, tvae, drv, tdp)
    return _cached

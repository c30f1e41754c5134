"""Drop-in for the reference's ``MPC/MPC_Tracking.py``: ``PathTracker`` on the GPU.

``from MPC.MPC_Tracking import PathTracker`` (``Distribution.py:9``, ``Traj_Tracking_Intact.py:4``,
``MPC_test1.py:3``) keeps working: same constructor, ``run_simulation``, ``step``, recorded ``trajectory`` /
``controls`` / ``times``.  The tracking loop - one SLSQP solve per time step in the reference
(``MPC_Tracking.py:311-415``, about a second each) - runs in ``libdmvae.so`` (``dmvae_mpc_prepare`` /
``dmvae_mpc_track``, one GPU thread per trajectory; ``dmvae/tracker.py``).

The rest of the reference's file (``VehicleModel``, ``PathInterpolator``, ``MPCController``, the ``plot_results``
figure, ``create_test_path``, ``main``) is host-side glue around that loop and is NOT re-implemented: keep the
reference's own file next to this one as ``MPC_Tracking_host.py`` and every public name this module does not define is
served from there; ``PathTracker`` then also carries the reference's ``path_interp`` (for callers that read
``tracker.path_interp.get_reference``) and ``plot_results``.  Without that file ``PathTracker`` still tracks, and the
glue names raise an ImportError that says so.
"""
import importlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PKG = os.path.join(_ROOT, "defensive-model-vae_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dmvae.tracker import PathTracker as _GpuPathTracker  # noqa: E402
from dmvae.tracker import track_batch  # noqa: E402,F401  (the batched call underneath)


def _load_host():
    for name in ("MPC.MPC_Tracking_host", "MPC_Tracking_host"):      # imported as a package member or from inside MPC/
        try:
            return importlib.import_module(name)
        except ImportError:
            continue
    return None


_host = _load_host()
_HOST_MERGED = _host is not None

if _HOST_MERGED:
    for _name, _value in vars(_host).items():
        if not _name.startswith("_") and _name not in globals():
            globals()[_name] = _value

    class PathTracker(_host.PathTracker):
        """The reference's ``PathTracker`` with its tracking loop on the GPU.  Construction is the reference's own
        (vehicle model, ``path_interp``, controller object, record lists: ``MPC_Tracking.py:421-452``); ``step`` and
        ``run_simulation`` (``:454-523``) hand the work to ``dmvae.tracker.PathTracker`` and keep ``current_state``,
        ``trajectory``, ``controls``, ``times`` and ``mpc.last_control`` as the reference would, so ``plot_results``
        and everything else that reads them works unchanged."""

        def __init__(self, waypoints, initial_state, wheelbase=2.8, prediction_horizon=10, control_horizon=5, dt=0.01):
            super().__init__(waypoints, initial_state, wheelbase, prediction_horizon, control_horizon, dt)
            # initial_state's heading has been wrapped in place by the reference's constructor (:435-436)
            self._gpu = _GpuPathTracker(waypoints, np.array(initial_state, dtype=np.float64), wheelbase,
                                        prediction_horizon, control_horizon, dt)

        def _sync(self):
            g = self._gpu
            self.current_state = g.current_state.copy()
            self.trajectory = [s.copy() for s in g.trajectory]
            self.controls = [c.copy() for c in g.controls]
            self.times = list(g.times)
            if g.controls:
                self.mpc.last_control = g.controls[-1].copy()

        def step(self, current_time):
            state, control = self._gpu.step(current_time)
            self._sync()
            return state, control

        def run_simulation(self, total_time):
            out = self._gpu.run_simulation(total_time)
            self._sync()
            return out
else:
    PathTracker = _GpuPathTracker

    _HOST_NAMES = ("VehicleModel", "PathInterpolator", "MPCController", "create_test_path", "main")

    def __getattr__(name):
        if name in _HOST_NAMES:
            raise ImportError(
                f"MPC_Tracking.{name} is host-side glue of the reference's MPC/MPC_Tracking.py that the B200 drop-in does "
                "not re-implement: keep the reference's file next to this one as MPC/MPC_Tracking_host.py (INTEGRATION.md section 2)")
        raise AttributeError(f"module 'MPC_Tracking' has no attribute {name!r}")

"""INTEGRATION.md section 2, checked: with the reference's own Tools.py kept as Tools_host.py and this repo's
drop-ins swapped in, every import-guarded caller of the reference imports exactly as before.  The layout is built
in a temporary directory from symlinks into /root/reference (nothing of the reference is copied into the repo), so
the test runs only where the reference is mounted (the build container)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DMVAE_REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "Tools.py")), reason="reference not mounted")


def test_reference_callers_import_against_the_drop_ins(tmp_path):
    # the reference tree, with Tools.py under its new name and the four boundary modules replaced
    for name in os.listdir(REF):
        if name in ("Tools.py", "Training_VAE.py", "Driver_Models.py", "Traj_Data_Process.py", "MPC", "__pycache__", ".git", ".idea"):
            continue
        os.symlink(os.path.join(REF, name), tmp_path / name)
    os.symlink(os.path.join(REF, "Tools.py"), tmp_path / "Tools_host.py")
    # MPC/: the reference's files, its MPC_Tracking.py under the name MPC_Tracking_host.py, the drop-in beside it
    os.mkdir(tmp_path / "MPC")
    for name in os.listdir(os.path.join(REF, "MPC")):
        if name not in ("MPC_Tracking.py", "__pycache__"):
            os.symlink(os.path.join(REF, "MPC", name), tmp_path / "MPC" / name)
    os.symlink(os.path.join(REF, "MPC", "MPC_Tracking.py"), tmp_path / "MPC" / "MPC_Tracking_host.py")
    os.symlink(os.path.join(ROOT, "MPC", "MPC_Tracking.py"), tmp_path / "MPC" / "MPC_Tracking.py")
    for name in ("Tools.py", "Training_VAE.py", "Driver_Models.py", "Traj_Data_Process.py", "defensive-model-vae_b200"):
        os.symlink(os.path.join(ROOT, name), tmp_path / name)
    prog = textwrap.dedent(f"""
        import os, sys
        sys.dont_write_bytecode = True
        sys.path.insert(0, {ROOT!r})
        from oracle.ref_loader import _stub_matplotlib
        _stub_matplotlib()                       # matplotlib is not installed in this image
        import types
        for extra in ("matplotlib.lines", "matplotlib.collections", "matplotlib.ticker", "mpl_toolkits", "mpl_toolkits.mplot3d"):
            m = types.ModuleType(extra)
            m.PathCollection = m.FuncFormatter = m.Axes3D = object
            sys.modules[extra] = m
        sys.modules["matplotlib.colors"].LinearSegmentedColormap = object
        sys.path.remove({ROOT!r})
        sys.path.insert(0, os.getcwd())
        import Tools, Training_VAE
        here = os.path.realpath(os.getcwd())
        assert os.path.realpath(Tools.__file__) == os.path.join({ROOT!r}, "Tools.py"), Tools.__file__
        assert Tools._HOST_MERGED
        # the accelerated entry points are the drop-in's, the glue is the reference's own
        assert Tools.load_model_and_generate_trajectory.__module__ == "Tools"
        assert Tools.get_human_and_bv_trajectories.__module__ == "Tools_host"
        assert Tools.ConditionalTrajectoryVAE is Training_VAE.ConditionalTrajectoryVAE
        assert Training_VAE.ConditionalTrajectoryVAE.__module__ == "dmvae.model"
        # names the reference's star-import hands on (Training_VAE.py:102)
        for n in ("nn", "optim", "Dataset", "DataLoader", "tqdm", "pd", "np", "torch", "plot_losses", "visualize_trajectories"):
            assert hasattr(Training_VAE, n), n
        import Distribution, Plot_case, Plot_Gif, Get_THW          # the import-guarded callers (INTEGRATION.md table)
        assert Distribution.load_model_and_generate_trajectory is Tools.load_model_and_generate_trajectory
        assert Distribution.get_human_and_bv_trajectories is Tools.get_human_and_bv_trajectories
        assert Plot_Gif.create_smooth_curve is Tools.create_smooth_curve
        assert Plot_Gif.ConditionalTrajectoryVAE is Training_VAE.ConditionalTrajectoryVAE
        from Tools import plot_gif_human_vs_model, save_animation_as_gif, process_model_trajectory   # Traj_Tracking_Intact.py:6
        # the tracker: Distribution.py:9 imports the drop-in, whose PathTracker is the reference's class with the loop replaced
        import MPC.MPC_Tracking as MT
        import MPC.MPC_Tracking_host as MH
        import dmvae.tracker
        assert os.path.realpath(MT.__file__) == os.path.join({ROOT!r}, "MPC", "MPC_Tracking.py"), MT.__file__
        assert MT._HOST_MERGED and Distribution.PathTracker is MT.PathTracker
        assert issubclass(MT.PathTracker, MH.PathTracker) and MT.PathTracker.plot_results is MH.PathTracker.plot_results
        assert MT.PathTracker.run_simulation is not MH.PathTracker.run_simulation
        assert MT.VehicleModel is MH.VehicleModel and MT.create_test_path is MH.create_test_path
        import numpy as np, torch
        way = MT.create_test_path()
        if not torch.cuda.is_available():        # the host half of the constructor runs, the GPU half says what is missing
            try:
                MT.PathTracker(way, np.array([0.0, 0.0, 0.0, 5.0, 0.0]), 2.8, 30, 20, 0.02)
                raise SystemExit("constructed without a CUDA device")
            except dmvae.DmvaeError as e:
                assert "no CPU path" in str(e)
        print("swap ok")
    """)
    out = subprocess.run([sys.executable, "-c", prog], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "swap ok" in out.stdout, out.stdout + out.stderr


def test_tracker_drop_in_without_the_host_module():
    prog = ("import sys; sys.path.insert(0, %r)\nimport MPC.MPC_Tracking as MT, dmvae.tracker\n"
            "assert not MT._HOST_MERGED and MT.PathTracker is dmvae.tracker.PathTracker\n"
            "try:\n    from MPC.MPC_Tracking import VehicleModel\nexcept ImportError as e:\n    assert 'MPC_Tracking_host.py' in str(e); print('ok')\n") % ROOT
    out = subprocess.run([sys.executable, "-c", prog], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_without_the_host_module_the_glue_names_say_what_to_do():
    prog = ("import sys; sys.path.insert(0, %r); import Tools; assert not Tools._HOST_MERGED\n"
            "try:\n    from Tools import create_smooth_curve\nexcept ImportError as e:\n    assert 'Tools_host.py' in str(e); print('ok')\n") % ROOT
    out = subprocess.run([sys.executable, "-c", prog], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr

"""Host-only check of the packed-arena mapping (defensive-model-vae_b200/csrc/dmvae_pack.cuh): the optimizer
kernels update the kernel-layout weight arena parameter by parameter (scatter_param); that must reproduce,
bit for bit, the full repack (pack_element) for every layout family in the envelope.  The mapping functions
are __host__ __device__, so nvcc builds a CPU executable from the same header the kernels include."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "defensive-model-vae_b200", "csrc")
SRC = os.path.join(ROOT, "tests", "native", "pack_scatter_check.cu")

CONFIGS = [(10, 8), (12, 8), (2, 1), (21, 16), (30, 24), (42, 64), (10, 5), (7, 3), (43, 8), (100, 16), (400, 64)]


def test_scatter_reproduces_gather(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "pack_scatter_check")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-I", CSRC, SRC, "-o", exe], check=True, capture_output=True, timeout=600)
    args = [str(v) for tl in CONFIGS for v in tl]
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("mismatches 0") == len(CONFIGS), out.stdout

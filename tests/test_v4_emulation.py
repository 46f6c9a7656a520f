"""Host emulation of the radix-16 pass kernels (pfx_pruned_v4.cuh): tests/v4_emul.cu replays the kernels' index
arithmetic thread by thread with their own helpers (radix plan, digit reversal, padding, dft16, tw_dft) and
compares every FFT length 32..2048 on padded axes of 2048 and 8192 samples, wrapped and unwrapped supports, full
and half outputs, with the direct sum.  CPU only (nvcc compiles the host side; no GPU is touched)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_v4_line_emulation_matches_direct_sum(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "v4_emul")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "v4_emul.cu")], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "v4 emulation: ok" in out.stdout, out.stdout[-2000:]


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_flex_model_host_build_matches_the_literal_formulas(tmp_path):
    """tests/flex_emul.cu: the device flexure model (pfx_flex.cuh, written in the four amplitudes mu_h, mu_w t, nu_h,
    nu_w t) compiled for the host against flex.f90:74-169 restated literally with its quotients and square roots
    (1e-11), its analytic Jacobian against central differences of that restatement, and the two-wavenumber form
    against the scalar one bit for bit."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "flex_emul")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "flex_emul.cu")], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "flex emulation: ok" in out.stdout, out.stdout[-2000:]

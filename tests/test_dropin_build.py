"""The drop-in build of INTEGRATION.md, exercised: the reference's own main() and kept translation units + this repository's
replacement units + libpbte_cuda (recipe: tests/native/Makefile).

  CPU (-m "not gpu"): where /root/reference exists, the recipe compiles and links.
  GPU (-m gpu):       the resulting binary, and the package's own bin/pbte_demo, run the shipped default configuration
                      (config/config.yaml + config/si.yaml: 101 source iterations) and write output/log/{Tc_all,coeff_all,
                      integrals_all}.txt; compared with the logs the reference ships (copies under tests/golden/reference_output_log/).
The logs hold 6 significant digits, so the comparison is numeric: |a - b| <= 1e-5 * max|b| per file (text rounding), and the
share of values whose printed text is identical is reported.
"""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
DROPIN = os.path.join(NATIVE, "_dropin", "pbte_demo_dropin")
DEMO = os.path.join(ROOT, "dg-solver-for-pbte-with-mfem_b200", "bin", "pbte_demo")
GOLD = os.path.join(ROOT, "tests", "golden", "reference_output_log")
REF = "/root/reference"


def test_dropin_recipe_links():
    if not os.path.exists(os.path.join(REF, "src", "PhononBTE.cpp")):
        pytest.skip("needs /root/reference (build container only)")
    subprocess.check_call(["make", "-C", NATIVE, "dropin"], stdout=subprocess.DEVNULL)
    assert os.path.exists(DROPIN)
    # the binary really is the reference's driver on top of the CUDA library
    needed = subprocess.check_output(["ldd", DROPIN], text=True)
    assert "libpbte_cuda.so" in needed


def numbers(path):
    vals, texts = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("#") or line.startswith("elem") or not line.strip():
                continue
            for tok in line.split():
                vals.append(float(tok))
                texts.append(tok)
    return np.array(vals), texts


def run_driver(binary, tmp_path, extra_env=None):
    shutil.copytree(os.path.join(ROOT, "config"), os.path.join(tmp_path, "config"))
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([binary, "-c", "config/config.yaml", "-o", "1", "-r", "0"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def check_logs(tmp_path, stdout):
    # the reference's per-iteration line: iteration 1 has residual exactly 1 (cold start), 101 iterations in all (max_iter of the config)
    lines = re.findall(r"\[Serial\] iter (\d+), residual = ([0-9.e+-]+)", stdout)
    assert len(lines) == 101 and lines[0] == ("1", "1.000000e+00") and lines[-1][0] == "101"
    for name in ("Tc_all.txt", "coeff_all.txt"):
        got, gt = numbers(os.path.join(tmp_path, "output", "log", name))
        want, wt = numbers(os.path.join(GOLD, name))
        assert got.shape == want.shape, name
        err = np.abs(got - want).max() / np.abs(want).max()
        same = sum(a == b for a, b in zip(gt, wt)) / len(wt)
        print(f"{name}: {len(wt)} values, max |diff| / max |ref| = {err:.2e}, identical text {100 * same:.2f} %")
        assert err <= 1e-5, (name, err)
        assert same >= 0.99, (name, same)
    with open(os.path.join(tmp_path, "output", "log", "integrals_all.txt")) as a, open(os.path.join(GOLD, "integrals_all.txt")) as b:
        assert a.read().split() == b.read().split()  # element matrices: produced by kept code, identical text


@pytest.mark.gpu
def test_dropin_binary_reproduces_the_shipped_logs(tmp_path):
    if not os.path.exists(DROPIN):
        pytest.skip("tests/native/_dropin/pbte_demo_dropin missing (it is built where /root/reference exists)")
    check_logs(str(tmp_path), run_driver(DROPIN, str(tmp_path)))


@pytest.mark.gpu
def test_package_driver_reproduces_the_shipped_logs(tmp_path):
    check_logs(str(tmp_path), run_driver(DEMO, str(tmp_path)))


@pytest.mark.gpu
def test_package_driver_on_two_gpus(tmp_path):
    """PBTE_DEVICES=0,1: directions split over two GPUs of one process, NCCL all-reduce inside the library, same logs."""
    import pbte_b200 as pb
    if pb.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    check_logs(str(tmp_path), run_driver(DEMO, str(tmp_path), {"PBTE_DEVICES": "0,1"}))

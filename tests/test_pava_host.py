"""The thread-per-row PAVA routine of csrc/pava.cuh (pava_block_runs: mask-driven replay of the reference's
sweeps, table-based small-integer division, forced run starts for rows of several blocks) compiled FOR THE
HOST and run against the oracle, bit for bit: values, pool sizes, stale interior entries, head masks; cold
and warm starts; ties, zeros, the reference's own generator and its worst case.  No GPU involved: this pins
the kernel's arithmetic and control flow on the CPU (tools/pava_host_check.cu)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_pava_block_runs_on_host(tmp_path):
    from oracle import cpu
    cpu.port()  # builds oracle/liboracle.so if needed
    exe = str(tmp_path / "pava_host_check")
    cmd = [NVCC, "-O2", "-std=c++17", "-fmad=false", "-Xcompiler", "-ffp-contract=off", "-o", exe,
           os.path.join(ROOT, "tools", "pava_host_check.cu"), "-L" + os.path.join(ROOT, "oracle"), "-loracle",
           "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe, "1500"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout[-500:] + run.stderr[-2000:]
    assert run.stdout.startswith("ok ")


def test_small_integer_division_is_correctly_rounded(tmp_path):
    """div_small (reciprocal table + two FMA corrections) == num / den bit for bit on 3*10^7 quotients
    (tools/divtest.c; the full 10^9-case run takes 15 s)."""
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    exe = str(tmp_path / "divtest")
    res = subprocess.run([gcc, "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "divtest.c"), "-lm"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe, "20000000"], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "bad(two-step)=0" in run.stdout, run.stdout + run.stderr


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_cluster_solver_share_arithmetic_on_host(tmp_path):
    """The share arithmetic of the cluster solver (csrc/solver_cluster.cuh: cluster_share / cluster_share_end, and the
    block search of the kernel) on 20,000 random layouts: the columns a CTA owns stay inside the groups its shared memory
    is sized for, shares tile all blocks and all rows (tools/cluster_share_check.cu)."""
    exe = str(tmp_path / "cluster_share_check")
    cmd = [NVCC, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
           os.path.join(ROOT, "block-simplex-least-squares_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tools", "cluster_share_check.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout[-500:] + run.stderr[-2000:]
    assert run.stdout.startswith("ok ")

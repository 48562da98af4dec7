"""CPU: the exact fp.cuh / ec.cuh template code the kernels use, compiled with g++ (PTX carry chains emulated and the
fast host multiply), checked against the oracle: field ops on all four fields, XYZZ mixed/full adds incl. doubling,
cancellation and infinity cases."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flags", [["-DSB_HOST_EMULATE_PTX"], []])
def test_host_fp_check(tmp_path, flags):
    from oracle import oracle as O
    so = O.build()
    exe = str(tmp_path / "host_fp_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", *flags, "-o", exe, os.path.join(ROOT, "tests", "host", "host_fp_check.cpp"), "-ldl"])
    out = subprocess.run([exe, so], capture_output=True, text=True)
    assert out.returncode == 0 and "HOST CHECK PASSED" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("flags", [["-DSB_HOST_EMULATE_PTX"], []])
def test_host_lane_pair_check(tmp_path, flags):
    """Fp2L (ec.cuh): the lane-pair form of Fq2 that k_accumulate_pair runs, two host threads in lockstep as the two lanes,
    against the plain Fp2 / XYZZ<Fp2> code on both curves (multiplies, squarings, votes, mixed additions incl. doubling,
    cancellation, y = 0 and infinity)."""
    exe = str(tmp_path / "host_pair_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-DSB_PAIR_HOST_EMULATE", *flags, "-o", exe,
                           os.path.join(ROOT, "tests", "host", "host_pair_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PAIR CHECK PASSED" in out.stdout, out.stdout + out.stderr

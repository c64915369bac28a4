"""The host-side band scheduler's policy (mdz_b200/csrc/band_grants.h, used by mdzcuda_render and the
rth_* layer for several devices -- SURVEY 8e "Partitioning") against simulated devices, on CPU: every
band is handed out exactly once and the render ends close to total work / total speed even when one
device runs at a fifth of the others' pace and a tenth of the image costs twelve times the rest.
A fixed interleave would end when the slow device is through its eighth: 5x later."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sched_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sched") / "sched_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_emu", "sched_test.cpp")])
    return exe


@pytest.mark.parametrize("total,speeds,slack", [
    (2160, [1] * 8, 1.05),
    (2160, [1] * 7 + [0.2], 1.10),          # one busy device
    (2160, [1, 0.5], 1.05),
    (1080, [1, 1, 1, 0.1], 1.20),
    (90, [1, 1], 1.35),                     # a preview: hardly more bands than one grid holds
    (7, [1, 1, 1, 1, 1, 1, 1, 1], 3.0),     # fewer bands than devices
])
def test_guided_chunks_balance_unequal_devices(sched_exe, total, speeds, slack):
    r = subprocess.run([sched_exe, str(total), "15", "4"] + [str(s) for s in speeds], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr
    makespan, ideal, grants = (float(x) for x in r.stdout.split()[1:4])
    assert makespan <= ideal * slack + 13.0, r.stdout       # + one expensive band: the indivisible tail
    assert grants <= 60 * len(speeds)

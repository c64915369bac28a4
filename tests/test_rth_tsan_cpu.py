"""The rth_* protocol layer (mdz_b200/csrc/rth.cpp: watch thread, render thread, in-order publication of bands,
stop / restart-while-rendering / quit) under ThreadSanitizer, on CPU: rth.cpp is compiled with -fsanitize=thread
against a stub of the CUDA side (tests/host_emu/rth_stub_backend.cpp) and driven by the same program that drives
the real library on the GPU box (tests/host_emu/rth_protocol.c, shaped after render.c:49-92 and the Julia
preview's restarts, main_gui.c:786-793).  Any data race report fails the test.  The reference's own pool has
unsynchronised reads (render_threads.c:495) and a known intermittent hang (BUGS:1-6); the replacement must not."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rth_protocol_is_race_free(tmp_path):
    exe = str(tmp_path / "rth_tsan")
    obj = str(tmp_path / "proto.o")
    flags = ["-fsanitize=thread", "-g", "-O1"]
    try:
        subprocess.check_call(["gcc", "-std=gnu99"] + flags + ["-c", "-o", obj, os.path.join(ROOT, "tests", "host_emu", "rth_protocol.c")])
        subprocess.check_call(["g++", "-std=c++17"] + flags + ["-o", exe, obj,
                               os.path.join(ROOT, "mdz_b200", "csrc", "rth.cpp"),
                               os.path.join(ROOT, "tests", "host_emu", "rth_stub_backend.cpp"),
                               "-l:libmpfr.so.6", "-l:libgmp.so.10", "-lpthread"])
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime for this compiler")
    for rep in range(5):
        r = subprocess.run([exe, "128"], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 second_deadlock_stack=1"))
        assert "ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
        assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr[-2000:]


def test_preview_restarts_are_race_free(tmp_path):
    """The Julia preview's pattern (tests/host_emu/rth_preview.c: a restart every few hundred microseconds, most of
    them over a render in progress) on the same sanitised build."""
    exe = str(tmp_path / "preview_tsan")
    obj = str(tmp_path / "preview.o")
    flags = ["-fsanitize=thread", "-g", "-O1"]
    try:
        subprocess.check_call(["gcc", "-std=gnu99"] + flags + ["-c", "-o", obj, os.path.join(ROOT, "tests", "host_emu", "rth_preview.c")])
        subprocess.check_call(["g++", "-std=c++17"] + flags + ["-o", exe, obj,
                               os.path.join(ROOT, "mdz_b200", "csrc", "rth.cpp"),
                               os.path.join(ROOT, "tests", "host_emu", "rth_stub_backend.cpp"),
                               "-l:libmpfr.so.6", "-l:libgmp.so.10", "-lpthread", "-lm"])
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime for this compiler")
    out = str(tmp_path / "frames.bin")
    r = subprocess.run([exe, "0", "150", "6000", out], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0"))
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr[-2000:]
    done = int(r.stdout.split()[1])
    assert 0 < done < 150, r.stdout          # some frames completed, some were cut short by the next restart
    import numpy as np
    frames = np.fromfile(out, dtype=np.int32).reshape(-1, 1 + 320 * 180)
    want = np.repeat(np.arange(1, 181, dtype=np.int32), 320)
    for f in frames:
        assert np.array_equal(f[1:], want)      # the stub writes line + 1: a completed frame is complete

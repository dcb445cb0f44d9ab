"""The device feeder's read-ahead ring (regtools_b200/csrc/stage_pipe.h) on the CPU: tools/stage_pipe_check.cc asks it for
windows the way run_device does (sequential runs, new ranges, abandoned ranges, the short last window) and compares the bytes;
run plain and under ThreadSanitizer."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tsan", [False, True])
def test_stage_pipe_windows_are_the_file_bytes(tmp_path, tsan):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "stage_pipe_check")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread"] + (["-fsanitize=thread"] if tsan else []) + [os.path.join(ROOT, "tools", "stage_pipe_check.cc"), "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and tsan:
        pytest.skip("ThreadSanitizer runtime not available: " + b.stderr[-200:])
    assert b.returncode == 0, b.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 errors" in r.stdout
    assert "ThreadSanitizer" not in r.stderr

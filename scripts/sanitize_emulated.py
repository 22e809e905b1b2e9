"""Runs the emulated kernel tests (tests/test_*_emulated.py: the library's .cu sources compiled for the CPU emulation) under a gcc sanitizer.

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python scripts/sanitize_emulated.py address
    LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS=report_signal_unsafe=0 python scripts/sanitize_emulated.py thread 2> tsan.log; grep -c "data race" tsan.log

address: out-of-bounds accesses to global / shared (static) / local memory.  thread: the emulation's barriers and atomics are mutex /
condition-variable based, so a missing __syncthreads() or an unsynchronised shared-memory update is reported as a data race.
The sanitizer flag is injected into the g++ command tests/cuda_emu/__init__.py issues; nothing else changes."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
kind = sys.argv[1] if len(sys.argv) > 1 else "address"
_run = subprocess.run


def run(cmd, *a, **k):
    if isinstance(cmd, list) and cmd and cmd[0] == "g++":
        cmd = cmd[:1] + [f"-fsanitize={kind}", "-fno-omit-frame-pointer", "-g"] + cmd[1:]
    return _run(cmd, *a, **k)


subprocess.run = run
import pytest  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
files = sys.argv[2:] or sorted(os.path.join(root, "tests", f) for f in os.listdir(os.path.join(root, "tests")) if f.endswith("_emulated.py") and "reference" not in f)
sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider"] + files))

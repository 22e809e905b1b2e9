"""TEST INFRASTRUCTURE ONLY: builds a CPU-executable twin of a csrc/*.cu file on top of cuda_emu.h (see its header)."""
import ctypes as C
import os
import re
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_emulated(cu_name: str, expected_launches: int, extra_pre: str = "") -> C.CDLL:
    csrc = os.path.join(ROOT, "bundlefusion_b200", "csrc")
    src = open(os.path.join(csrc, cu_name)).read()
    # the library's own device headers are spliced in (they include bf_common.cuh, which the shim replaces)
    src = re.sub(r'#include "(\w+\.cuh)"', lambda m: m.group(0) if m.group(1) == "bf_common.cuh" else open(os.path.join(csrc, m.group(1))).read().replace("#pragma once", ""), src)
    src = src.replace('#include "bf_common.cuh"', "")
    src = re.sub(r"extern __shared__ float (\w+)\[\];", lambda m: "" if m.group(1) == "sm" else f"float* {m.group(1)} = sm;", src)      # dynamic shared memory: the shim's sm[]
    src = re.sub(r"extern __shared__ (unsigned|int) (\w+)\[\];", lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(sm);", src)
    src, n = re.subn(r"(\w+(?:<\w+>)?)<<<\s*([^,]+),\s*([^,]+),\s*[^,]+,\s*[^>]+>>>\(", r"EMU_LAUNCH(\1, \2, \3, ", src)
    assert n == expected_launches, (cu_name, n)
    pre = ('#include "%s"\n' % os.path.join(ROOT, "tests", "cuda_emu", "cuda_emu.h") +
           "#define BF_CHECK(e) do { int _e = (int)(e); if (_e) return _e; } while (0)\n#define BF_SAFE(e) do { (void)(e); } while (0)\n"
           "namespace bf { unsigned long long g_launchCount = 0; static inline cudaStream_t stream() { return nullptr; } static inline int num_sms() { return 2; } }\n")
    d = tempfile.mkdtemp(prefix="bf_emu_")
    cpp = os.path.join(d, cu_name.replace(".cu", "_emu.cpp"))
    open(cpp, "w").write(pre + extra_pre + src)
    so = os.path.join(d, "lib" + cu_name.replace(".cu", "_emu.so"))
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "bundlefusion_b200", "csrc"),
                        cpp, "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    shutil.rmtree(d, ignore_errors=True)               # the mapping outlives the file
    return lib

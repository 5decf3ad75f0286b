"""Build libmdqt_b200.so (CUDA kernels + C ABI + host IO) and the mdqt_run driver in-tree with nvcc for sm_100a.
No JIT cache: the built files travel with the repo snapshot."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmdqt_b200.so")
DRIVER = os.path.join(HERE, "mdqt_run")
SOURCES = ["mdqt_force.cu", "mdqt_qt.cu", "mdqt_diag.cu", "mdqt_capi.cu", "mdqt_comm.cu", "mdqt_io.cpp"]
HEADERS = ["mdqt_internal.h", "mdqt_handle.h", "mdqt_fixed.cuh", "mdqt_qtconsts.h", os.path.join("..", "..", "include", "mdqt.h"),
           os.path.join("..", "..", "include", "mdqt_io.h")]
DRIVER_SOURCES = ["mdqt_driver.cpp", "mdqt_programs.cpp"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in sources)


def needs_build():
    return _stale(LIB, SOURCES + HEADERS) or _stale(DRIVER, DRIVER_SOURCES + HEADERS)


def build(force=False, verbose=False, defines=(), out=None):
    if not force and not needs_build() and out is None:
        return LIB
    target = out or LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++"] + ["-D" + d for d in defines] + [
           "-o", target] + [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    if out is None:
        # the C++ host driver (the reference's main loop on top of the C ABI); finds the library next to itself
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-o", DRIVER] + [os.path.join(CSRC, f) for f in DRIVER_SOURCES] + [
            "-L" + HERE, "-lmdqt_b200", "-Wl,-rpath,$ORIGIN", "-pthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ (mdqt_run) failed:\n" + res.stdout + res.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

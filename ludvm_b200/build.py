"""In-tree build of libludvm_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libludvm_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]
# Per-file flags.  sim.cu holds the scalar phases of the time step, written as plain C expressions that
# must not be contracted into FMAs (bit parity with numpy).
PER_FILE = {"sim.cu": ["-fmad=false"]}


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header():
    t = os.path.getmtime(os.path.join(HERE, "..", "include", "ludvm_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _newest_header()
    objs, rebuilt = [], False
    for f in sources():
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + ARCH + COMMON + PER_FILE.get(f, []) + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed on %s" % f)
            with open(obj + ".ptxas.txt", "w") as fh:
                fh.write(r.stderr)
            rebuilt = True
    if rebuilt or not os.path.isfile(LIB):
        cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

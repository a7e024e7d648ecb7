#!/usr/bin/env python
"""Build libddl_b200.so (sm_100a) in-tree with nvcc.  Called by __graft_entry__.build().

    python dedalus-1.0_b200/build.py            # CUDA library (cross-compiles without a GPU)
    python dedalus-1.0_b200/build.py --emul OUT # host-emulation build for tests/host (g++ only)
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "dedalus", "_lib")
LIB = os.path.join(LIBDIR, "libddl_b200.so")
SIZES = [0, 8, 16, 32, 64, 128, 256, 512, 1024, 2048]      # 0: the runtime-length (mixed radix) instantiation
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
EXTRA_CU = ["api.cu", "p2p.cu", "microbench.cu"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + ["../../include/ddl.h"]


def _digest(extra=""):
    h = hashlib.sha256(extra.encode())
    for f in _sources():
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


class _BuildLock:
    """Exclusive lock on <dir>/.lock for the duration of a build: several processes asking for the same stale target (the ranks of a
    gloo test, pytest-xdist workers) must not compile into the same object files at once; the late-comers find it fresh."""

    def __init__(self, directory):
        self.path = os.path.join(directory, ".lock")

    def __enter__(self):
        import fcntl
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


def build_cuda(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    with _BuildLock(objdir):
        return _build_cuda_locked(objdir, force, verbose)


def _build_cuda_locked(objdir, force, verbose):
    stamp = os.path.join(LIBDIR, ".digest")
    dig = _digest("cuda")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    jobs = []
    for n in SIZES:
        obj = os.path.join(objdir, "tile_inst_%d.o" % n)
        jobs.append((NVCC_FLAGS_CMD(["-DDDL_N=%d" % n, "-c", os.path.join(CSRC, "tile_inst.cu"), "-o", obj]), obj + ".log", obj))
    for cu in EXTRA_CU:
        obj = os.path.join(objdir, cu.replace(".cu", ".o"))
        jobs.append((NVCC_FLAGS_CMD(["-c", os.path.join(CSRC, cu), "-o", obj]), obj + ".log", obj))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        outs = list(ex.map(lambda j: _run(j[0], j[1]), jobs))
    if verbose:
        for o in outs:
            sys.stdout.write(o)
    objs = [j[2] for j in jobs]
    _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
         os.path.join(objdir, "link.log"))
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


def NVCC_FLAGS_CMD(rest):
    return [NVCC] + NVCC_FLAGS + rest


def build_emul(outdir, sanitize=False):
    """g++-only build of the same sources with DDL_HOST_EMUL (tests/host only; never shipped).
    sanitize: AddressSanitizer + UBSan instrumentation (the process needs libasan preloaded; tests/test_host_sanitizer.py)."""
    os.makedirs(outdir, exist_ok=True)
    with _BuildLock(outdir):
        return _build_emul_locked(outdir, sanitize)


def _build_emul_locked(outdir, sanitize):
    lib = os.path.join(outdir, "libddl_emul.so")
    stamp = os.path.join(outdir, ".digest")
    dig = _digest("emul-asan" if sanitize else "emul")
    if os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    cxx = os.environ.get("CXX", "g++")
    flags = ["-x", "c++", "-std=c++17", "-O2", "-fPIC", "-DDDL_HOST_EMUL", "-w"]
    link = []
    if sanitize:
        flags = [f for f in flags if f != "-O2"] + ["-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=address",
                                                     "-fsanitize=bounds,shift,integer-divide-by-zero,null", "-fno-sanitize-recover=all"]
        link = ["-fsanitize=address", "-fsanitize=undefined", "-L/usr/lib/gcc/x86_64-linux-gnu/13"]
    jobs = []
    for n in SIZES:
        obj = os.path.join(outdir, "tile_inst_%d.o" % n)
        jobs.append(([cxx] + flags + ["-DDDL_N=%d" % n, "-c", os.path.join(CSRC, "tile_inst.cu"), "-o", obj], obj + ".log", obj))
    for cu in EXTRA_CU:
        obj = os.path.join(outdir, cu.replace(".cu", ".o"))
        jobs.append(([cxx] + flags + ["-c", os.path.join(CSRC, cu), "-o", obj], obj + ".log", obj))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda j: _run(j[0], j[1]), jobs))
    _run([cxx, "-shared", "-o", lib] + link + [j[2] for j in jobs], os.path.join(outdir, "link.log"))
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


if __name__ == "__main__":
    if "--emul" in sys.argv:
        print(build_emul(sys.argv[sys.argv.index("--emul") + 1]))
    else:
        print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))

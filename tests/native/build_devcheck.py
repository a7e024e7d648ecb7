"""Builds tests/native/devcheck.cu (test infrastructure): `device()` with nvcc for sm_100a against the in-tree
libddl_b200.so, `emul()` with g++ against the host-emulation build.  Binaries land in tests/native/_build/
(git-ignored; they travel to the GPU box with the snapshot like the library itself)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "devcheck.cu")
OUT = os.path.join(HERE, "_build")
LIBDIR = os.path.join(ROOT, "dedalus-1.0_b200", "dedalus", "_lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def device():
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "devcheck")
    lib = os.path.join(LIBDIR, "libddl_b200.so")
    if _stale(exe, [SRC, lib, os.path.join(ROOT, "include", "ddl.h")]):
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O2", "-std=c++17", "-cudart", "shared",
              "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR, "-lddl_b200",
              "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../../dedalus-1.0_b200/dedalus/_lib", "-o", exe])
    return exe


def emul():
    sys.path.insert(0, os.path.join(ROOT, "dedalus-1.0_b200"))
    import build as ddl_build
    libdir = os.path.join(ROOT, "tests", "host", "_build")
    lib = ddl_build.build_emul(libdir)
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "devcheck_emul")
    if _stale(exe, [SRC, lib, os.path.join(ROOT, "include", "ddl.h")]):
        _run([os.environ.get("CXX", "g++"), "-x", "c++", "-std=c++17", "-O2", "-DDEVCHECK_EMUL", "-I", os.path.join(ROOT, "include"),
              SRC, "-L", libdir, "-lddl_emul", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def cuda_env():
    """Environment in which the binary finds libcudart (the image keeps it in the venv's nvidia wheels)."""
    env = dict(os.environ)
    extra = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime as m
        extra.insert(0, os.path.join(list(m.__path__)[0], "lib"))
    except Exception:
        pass
    env["LD_LIBRARY_PATH"] = ":".join(extra + [env.get("LD_LIBRARY_PATH", "")])
    return env


if __name__ == "__main__":
    print(emul() if "--emul" in sys.argv else device())

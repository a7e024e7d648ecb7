"""CPU: resource limits that only a real launch would otherwise reveal.  The host-emulation build cannot see a kernel
whose registers x threads exceed an SM's 64 K registers (the launch fails on the device with "too many resources
requested"); `-Xptxas -v` in the build logs can.  Checked per kernel family against the largest CTA its launcher uses."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "dedalus-1.0_b200", "build")

# kernel family -> largest CTA its launcher asks for (csrc/api.cu round32, pointwise.cuh, reduce.cuh, p2p.cu)
MAX_THREADS = {"ddl::tile_kernel": 768, "ddl::items_kernel": 256, "ddl::items_kernel_b": 256, "ddl::reduce_kernel": 256, "ddl::reduce_final_kernel": 256,
               "ddl::p2p_wait_kernel": 32, "ddl::p2p_signal_kernel": 32, "ddl::fp64_rate_kernel": 256, "ddl::p2p_push_kernel": 1024, "ddl::p2p_tma_push_kernel": 32}


def _entries():
    if not glob.glob(os.path.join(BUILD, "*.o.log")):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    text = "".join(open(p).read() for p in sorted(glob.glob(os.path.join(BUILD, "*.o.log"))))
    found = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'.*?(\d+) bytes stack frame.*?Used (\d+) registers", text, re.S)
    names = subprocess.run(["c++filt"], input="\n".join(f[0] for f in found), capture_output=True, text=True).stdout.split("\n")
    return [(n.replace("void ", ""), int(f[1]), int(f[2])) for n, f in zip(names, found)]


def test_every_kernel_fits_its_largest_launch():
    entries = _entries()
    assert len(entries) > 250
    seen = collections.Counter()
    for name, stack, regs in entries:
        family = re.sub(r"[<(].*", "", name)
        seen[family] += 1
        if family in MAX_THREADS:
            warps = (MAX_THREADS[family] + 31) // 32
            per_warp = (regs * 32 + 255) // 256 * 256          # register allocation granularity: 256 per warp
            assert warps * per_warp <= 65536, (name, regs, MAX_THREADS[family])
        else:
            # xfused_kernel / strided_fast carry __maxnreg__ / __launch_bounds__ derived from their own CTA shape
            assert family in ("ddl::xfused_kernel", "ddl::xfused_persist_kernel", "ddl::xfused_rot_kernel", "ddl::strided_fast", "ddl::strided_staged", "ddl::strided_two"), name
    for family in MAX_THREADS:
        assert seen[family] > 0, family


def test_runtime_length_kernels_are_built():
    names = [n for n, _, _ in _entries() if n.startswith("ddl::tile_kernel<0,")]
    assert len(names) == 19, names          # C2C fwd / inv, C2R, R2C and the fifteen fused physics policies (six solenoidal, six advective, three traceless-flux)

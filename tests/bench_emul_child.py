"""Child process of tests/test_bench_contract.py (TEST INFRASTRUCTURE): the GPU arm of bench.py, line for line, in the GPU-less
build container.  tests/conftest.py's opt-in harness points the drop-in package at the g++ host-emulation build; on top of it the
handful of torch.cuda calls bench.py makes itself (events, side streams, pinned buffers, the device generator) get CPU stand-ins
IN THIS PROCESS ONLY.  It checks that the script still runs against the package as it is now and prints the contract line;
the numbers mean nothing (wall-clock 'events', one CPU thread)."""
import argparse
import contextlib
import os
import sys
import time

os.environ["DDL_TEST_HOST_EMUL"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import conftest  # noqa: F401,E402   (enables the harness)
import torch  # noqa: E402


class _Event(object):
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)

    def synchronize(self):
        pass


class _Stream(object):
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass

    def synchronize(self):
        pass


def _on_cpu(fn):
    def wrapped(*a, **kw):
        if str(kw.get("device", "")).startswith("cuda"):
            kw["device"] = "cpu"
        kw.pop("pin_memory", None)
        return fn(*a, **kw)
    return wrapped


_generator = torch.Generator
torch.Generator = lambda device=None: _generator()
for name in ("randn", "empty", "tensor", "zeros"):
    setattr(torch, name, _on_cpu(getattr(torch, name)))
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.empty_cache = lambda: None
torch.cuda.Event = _Event
torch.cuda.Stream = _Stream
_main = _Stream()
torch.cuda.current_stream = lambda *a: _main
torch.cuda.stream = lambda s: contextlib.nullcontext()

import torch.distributed as dist  # noqa: E402

_init = dist.init_process_group
dist.init_process_group = lambda backend=None, device_id=None, **kw: _init("gloo", **kw)     # torchrun children: gloo for nccl
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["DEDALUS_SLAB_EXCHANGE"] = "collective"          # peer memory needs the device

import bench  # noqa: E402

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1].endswith(".py"):
    # any other measurement script of the repo (profiles/*.py) under the same stand-ins:  bench_emul_child.py <script> [args]
    import runpy
    sys.argv = sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name="__main__")
elif __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    bench.cpu_baseline = lambda args: {"value": None, "unit": bench.UNIT, "cores": 1, "kind": "reference", "sample": "skipped in the emulated run"}
    bench.run_ours(argparse.Namespace(n=n, steps=2, warmup=1, quick=False, cpu_n=16, cpu_steps=1, gpus=1, impl="ours",
                                     e2e_full="--e2e-full" in sys.argv, skip_e2e=False))

"""Cumulative wall-clock timers (reference: dedalus/utils/timer.py:29-65)."""
import functools
import time

from .parallelism import com_sys


class Timer(object):
    """Decorator object: accumulates wall time per decorated function name in `timers`.
    GPU work is asynchronous; times are host-side unless `sync` is set (then the current CUDA
    stream is synchronised around the call, for profiling runs)."""

    sync = False

    def __init__(self):
        self.timers = {}

    def __call__(self, func):
        name = func.__name__

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if Timer.sync:
                import torch
                torch.cuda.synchronize()
            t0 = time.time()
            out = func(*args, **kwargs)
            if Timer.sync:
                import torch
                torch.cuda.synchronize()
            self.timers[name] = self.timers.get(name, 0.0) + (time.time() - t0)
            return out
        return wrapper

    def print_stats(self, proc=0):
        if com_sys.myproc == proc:
            print()
            print("---Timings (proc %i)---" % proc)
            for name, sec in self.timers.items():
                print("%s: %10.5f sec" % (name, sec))
            print()


timer = Timer()

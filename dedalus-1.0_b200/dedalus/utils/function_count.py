"""Call counters (reference: dedalus/utils/function_count.py:2-35)."""
import functools

from .parallelism import com_sys


class countcalls(object):
    def __init__(self):
        self.counters = {}

    def __call__(self, func):
        name = func.__name__

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            out = func(*args, **kwargs)
            self.counters[name] = self.counters.get(name, 0) + 1
            return out
        return wrapper

    def print_stats(self, proc=0):
        if com_sys.myproc == proc:
            print()
            print("---Call counts (proc %i)---" % proc)
            for name, n in self.counters.items():
                print("%s: %i calls" % (name, n))
            print()


counts = countcalls()

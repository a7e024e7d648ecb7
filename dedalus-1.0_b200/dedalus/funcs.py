"""Small developer helpers scripts import by name (reference: dedalus/funcs.py:41-110)."""
import sys
import traceback


def insert_ipython(num_up=1):
    """Drop into an embedded IPython shell at the caller's frame (funcs.py:41-63)."""
    import inspect
    try:
        from IPython import embed
    except ImportError:
        raise RuntimeError("insert_ipython needs IPython")
    frame = inspect.stack()[num_up]
    embed(user_ns=dict(frame[0].f_globals, **frame[0].f_locals))


def signal_print_traceback(signo, frame):
    """SIGHUP handler of dedalus/mods.py:82-86: print the current stack."""
    print(traceback.print_stack(frame), file=sys.stderr)

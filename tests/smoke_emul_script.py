"""Run by tests/bench_emul_child.py (TEST INFRASTRUCTURE): the body of __graft_entry__.smoke() against the host-emulation build."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

g._paths()
g._smoke_cases()

"""Logger with the rank in every record (reference: dedalus/utils/logger.py:49-72)."""
import logging
import sys

from ..config import decfg
from .parallelism import com_sys


class _RankFilter(logging.Filter):
    def filter(self, record):
        record.proc = com_sys.myproc
        return True


_base = logging.getLogger("Dedalus")
_base.setLevel(getattr(logging, decfg.get("utils", "loglevel").upper(), logging.WARNING))
if not _base.handlers:
    _fmt = logging.Formatter("%(asctime)s %(name)-3s: [%(levelname)-9s] %(proc)i %(message)s")
    _err = logging.StreamHandler(sys.stderr)      # DEBUG from every rank
    _err.setLevel(logging.DEBUG)
    _err.addFilter(lambda r: r.levelno == logging.DEBUG)
    _err.setFormatter(_fmt)
    _out = logging.StreamHandler(sys.stdout)      # INFO+ from rank 0 only
    _out.setLevel(logging.INFO)
    _out.addFilter(lambda r: com_sys.myproc == 0)
    _out.setFormatter(_fmt)
    _base.addFilter(_RankFilter())
    _base.addHandler(_err)
    _base.addHandler(_out)

mylog = _base

"""Per-rank logger (``VGG/settings.py:27-38``, ``VGG/main_trainer.py:165-176``): stream handler on every
rank, optional ``<log_dir>/<host>-<rank>.log`` file."""
from __future__ import annotations

import logging
import os
import socket
from typing import Optional

_LOGGERS = {}


def get_logger(rank: int = 0, log_dir: Optional[str] = None, level: int = logging.INFO) -> logging.Logger:
    key = (rank, log_dir)
    if key in _LOGGERS:
        return _LOGGERS[key]
    host = socket.gethostname()
    lg = logging.getLogger("oktopk.%s.%d" % (host, rank))
    lg.setLevel(level if rank == 0 else logging.WARNING)
    lg.propagate = False
    fmt = logging.Formatter("%(asctime)s [" + host + "-%d" % rank + "] %(levelname)s %(message)s")
    if not lg.handlers:
        sh = logging.StreamHandler()
        sh.setFormatter(fmt)
        lg.addHandler(sh)
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
        fh = logging.FileHandler(os.path.join(log_dir, "%s-%d.log" % (host, rank)))
        fh.setFormatter(fmt)
        fh.setLevel(level)
        lg.addHandler(fh)
    _LOGGERS[key] = lg
    return lg

"""`hetu.logger`: rank-prefixed Python logging whose level follows HETU_INTERNAL_LOG_LEVEL like the native logger
(ref: python/hetu/logger.py, hetu/common/logging.cc)."""
from __future__ import annotations

import logging
import os

_LEVELS = {"TRACE": 5, "DEBUG": logging.DEBUG, "INFO": logging.INFO, "WARN": logging.WARNING, "WARNING": logging.WARNING, "ERROR": logging.ERROR,
           "FATAL": logging.CRITICAL}


class _RankFilter(logging.Filter):
    def filter(self, record):
        record.rank = os.environ.get("RANK", "0")
        return True


def get_logger(name: str = "hetu") -> logging.Logger:
    lg = logging.getLogger(name)
    if not lg.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("[%(asctime)s] [%(levelname)s] [rank %(rank)s] %(message)s", "%H:%M:%S"))
        h.addFilter(_RankFilter())
        lg.addHandler(h)
        lg.propagate = False
    lg.setLevel(_LEVELS.get(os.environ.get("HETU_INTERNAL_LOG_LEVEL", "INFO").upper(), logging.INFO))
    return lg


logger = get_logger()
debug, info, warning, error = logger.debug, logger.info, logger.warning, logger.error
warn = logger.warning

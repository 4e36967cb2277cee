"""Small logging helpers with the call signatures the reference code uses (src/utils/logger.py)."""
import logging

_loggers = {}


class _Indented(logging.LoggerAdapter):
    def __init__(self, logger, spaces=2):
        super().__init__(logger, {})
        self._spaces, self._depth = spaces, 0

    def add(self, n=1):
        self._depth += n
        return self

    def sub(self, n=1):
        self._depth = max(0, self._depth - n)
        return self

    def process(self, msg, kwargs):
        return " " * (self._spaces * self._depth) + str(msg), kwargs


def get_root_logger(logger_name="basiccomp", log_level=logging.INFO, log_file=None):
    if logger_name in _loggers:
        return _loggers[logger_name]
    if isinstance(log_level, str):
        log_level = getattr(logging, log_level)
    base = logging.getLogger(logger_name)
    base.setLevel(logging.DEBUG)
    base.propagate = False
    h = logging.StreamHandler()
    h.setLevel(log_level)
    h.setFormatter(logging.Formatter("[%(levelname)-7s] %(message)s"))
    base.addHandler(h)
    if log_file is not None:
        fh = logging.FileHandler(log_file, "w")
        fh.setFormatter(logging.Formatter("%(asctime)s %(levelname)-8s: %(message)s"))
        base.addHandler(fh)
    _loggers[logger_name] = _Indented(base)
    return _loggers[logger_name]


def log_dict_items(dic, level="INFO", indent=True, key_color=None, val_color=None):
    lg = get_root_logger()
    lvl = getattr(logging, level) if isinstance(level, str) else level
    if indent:
        lg.add()
    for k, v in dic.items():
        lg.log(lvl, f"{k}: {v}")
    if indent:
        lg.sub()

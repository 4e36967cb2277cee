"""ORACLE shim: IndentedLoggerAdapter(logger, spaces) with add()/sub()."""
import logging


class IndentedLoggerAdapter(logging.LoggerAdapter):
    def __init__(self, logger, extra=None, auto_add=True, **kwargs):
        super().__init__(logger, extra or {})
        self._spaces = int(kwargs.get("spaces", 4))
        self._level = 0

    def add(self, n=1):
        self._level += n
        return self

    def sub(self, n=1):
        self._level = max(0, self._level - n)
        return self

    def push(self):
        return self

    def pop(self):
        return self

    def process(self, msg, kwargs):
        return " " * (self._spaces * self._level) + str(msg), kwargs

"""``wrap_exceptions`` convention of the reference (calculations/exceptions.py:12-37):
inside a daemon worker process an exception is re-raised as ``WrapException`` carrying
the formatted traceback, so it survives the trip through multiprocessing / Pyro."""
import multiprocessing
import sys
import traceback
from functools import wraps


class WrapException(Exception):
    def __init__(self):
        kind, value, tb = sys.exc_info()
        super(WrapException, self).__init__()
        self.exception = value
        self.formatted = "".join(traceback.format_exception(kind, value, tb))

    def __str__(self):
        return "%s\nOriginal traceback:\n%s" % (Exception.__str__(self), self.formatted)


def wrap_exceptions(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        try:
            return func(*args, **kwargs)
        except Exception:
            if multiprocessing.current_process().daemon:
                raise WrapException()
            raise
    return wrapper

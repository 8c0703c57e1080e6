"""An async-server provider with the interface of the reference's providers
(generic/asynchronous/dummy_async_provider.py:10-34; selected through
``settings.ASYNC_SERER_PROVIDERS``, data/settings.py:221-224) that turns "submit all, then
fetch all" (refinement/async_evaluatable.py:38-50) into ONE batched device evaluation.

``submit(func)`` only records the callable.  The first ``.get()`` of any pending result
flushes the queue: every ``functools.partial(get_optimized_residual, mixture_data)`` in it is
evaluated together by ``pyxrd_b200.batched``; anything else is simply called.
"""
import functools
import threading


class B200AsyncResult(object):
    def __init__(self, server, func):
        self._server, self._func = server, func
        self._done, self._value, self._error = False, None, None

    def get(self):
        if not self._done:
            self._server.flush()
        if self._error is not None:
            raise self._error
        return self._value


class B200AsyncServer(object):
    def __init__(self, batch_evaluator=None):
        self._pending = []
        self._lock = threading.RLock()
        self._batch_evaluator = batch_evaluator

    def loopCondition(self):
        return True

    def submit(self, func):
        result = B200AsyncResult(self, func)
        with self._lock:
            self._pending.append(result)
        return result

    @staticmethod
    def _trial_mixture(func):
        """the MixtureData of a ``partial(get_optimized_residual, data)``, else None"""
        if isinstance(func, functools.partial) and len(func.args) == 1 and not func.keywords:
            if getattr(func.func, "__name__", "") == "get_optimized_residual":
                return func.args[0]
        return None

    def flush(self):
        with self._lock:
            pending, self._pending = self._pending, []
        trials = [(r, self._trial_mixture(r._func)) for r in pending]
        batch = [(r, m) for r, m in trials if m is not None]
        if batch:
            evaluator = self._batch_evaluator
            if evaluator is None:
                from .batched import get_optimized_residuals as evaluator
            try:
                values = evaluator([m for _, m in batch])
                for (r, _), v in zip(batch, values):
                    r._value, r._done = v, True
            except Exception as err:  # delivered to every waiting caller
                for r, _ in batch:
                    r._error, r._done = err, True
        for r, m in trials:
            if m is None:
                try:
                    r._value = r._func()
                except Exception as err:
                    r._error = err
                r._done = True

    def shutdown(self):
        self.flush()


class B200AsyncServerProvider(object):
    """drop-in for ``settings.ASYNC_SERER_PROVIDERS`` (class methods, like the reference's)"""
    _server = None

    @classmethod
    def get_status(cls):
        return ("#00FF00", "Connected (B200)", "Batched device evaluation through pyxrd_b200")

    @classmethod
    def get_server(cls):
        if cls._server is None:
            cls._server = B200AsyncServer()
        return cls._server

    @classmethod
    def launch_server(cls):
        cls.get_server()

    @classmethod
    def stop_server(cls):
        if cls._server is not None:
            cls._server.shutdown()
            cls._server = None

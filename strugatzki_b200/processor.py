"""Minimal mirror of the `de.sciss.processor` contract the reference's processors are written against
(third-party, un-vendored; call sites Strugatzki.scala:95-99,177-211, FeatureCorrelationImpl.scala:164,402):

    proc = Factory(config)            # ProcessorFactory.apply -> prepared, not started
    proc.add_listener(observer)       # observer receives Progress(p) and Result(Success(v) | Failure(e))
    proc.start()                      # body() runs on ONE worker thread
    proc.abort()                      # cooperative: body sees checkAborted() -> Aborted
    Factory.run(config)(observer)     # = apply + add_listener + start

`progress` is only dispatched when it rose by >= 1 % (ProcessorImpl behaviour noted in SURVEY.md section 5).
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Any, Callable, List, Optional


class Aborted(Exception):
    """Processor.Aborted"""


@dataclass
class Progress:
    source: Any
    amount: float


@dataclass
class Success:
    value: Any


@dataclass
class Failure:
    exception: BaseException


@dataclass
class Result:
    source: Any
    value: Any  # Success | Failure


class ProcessorImpl:
    """Base of the three processors; subclasses implement body()."""

    def __init__(self, config):
        self.config = config
        self._listeners: List[Callable] = []
        self._thread: Optional[threading.Thread] = None
        self._aborted = threading.Event()
        self._done = threading.Event()
        self._result = None
        self._progress = 0.0
        self._last_dispatched = -1.0

    # ---- observer side ----
    def add_listener(self, fn: Callable):
        self._listeners.append(fn)
        return fn

    def remove_listener(self, fn: Callable):
        self._listeners.remove(fn)

    def _dispatch(self, msg):
        for fn in list(self._listeners):
            fn(msg)

    # ---- control ----
    def start(self):
        if self._thread is not None:
            raise RuntimeError("processor already started")
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _run(self):
        try:
            v = self.body()
            self._result = Success(v)
        except BaseException as e:  # noqa: BLE001 - mirrors Future failure propagation
            self._result = Failure(e)
        self._done.set()
        self._dispatch(Result(self, self._result))

    def abort(self):
        self._aborted.set()
        self._on_abort()

    def _on_abort(self):
        pass

    @property
    def aborted(self) -> bool:
        return self._aborted.is_set()

    def check_aborted(self):
        if self._aborted.is_set():
            raise Aborted()

    @property
    def progress(self) -> float:
        return self._progress

    @progress.setter
    def progress(self, v: float):
        self._progress = v
        if v - self._last_dispatched >= 0.01 or v >= 1.0 > self._last_dispatched:
            self._last_dispatched = v
            self._dispatch(Progress(self, v))

    def await_result(self, timeout: Optional[float] = None):
        """Await.result(processor, Duration.Inf): returns the product or raises the failure."""
        if not self._done.wait(timeout):
            raise TimeoutError("processor still running")
        if isinstance(self._result, Failure):
            raise self._result.exception
        return self._result.value

    @property
    def is_completed(self) -> bool:
        return self._done.is_set()

    def body(self):  # pragma: no cover - abstract
        raise NotImplementedError


class ProcessorFactory:
    """object X extends ProcessorFactory.WithDefaults"""
    Impl = None

    @classmethod
    def apply(cls, config=None):
        return cls.Impl(config if config is not None else cls.default_config())

    @classmethod
    def run(cls, config=None, observer: Optional[Callable] = None):
        p = cls.apply(config)
        if observer is not None:
            p.add_listener(observer)
        return p.start()

    @classmethod
    def default_config(cls):
        raise NotImplementedError

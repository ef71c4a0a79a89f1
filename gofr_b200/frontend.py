"""Batching front-end: single requests in from many threads, batches out to the GPU (gofr_frontend_*, include/gofr_b200.h).

The reference serves one request per goroutine (net/http conn.serve → Router.ServeHTTP, pkg/gofr/httpServer.go:29-33);
`Frontend.serve` is that per-request call: it blocks until the batch the request joined has come back.  ctypes drops
the GIL for the duration of the call, so Python threads can stand in for connection goroutines.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from . import _abi
from .engine import Engine


class Frontend:
    def __init__(self, engine: Engine, max_batch: int = 4096, max_wait_us: int = 200, slot_bytes: int = 1024,
                 max_request_bytes: int = 4096):
        self._f = C.c_void_p()
        self.engine = engine
        self.slot_bytes = slot_bytes
        _abi.check(_abi.lib().gofr_frontend_create(C.byref(self._f), engine._e, max_batch, max_wait_us, slot_bytes,
                                                   max_request_bytes), "gofr_frontend_create")

    def close(self):
        if self._f:
            _abi.lib().gofr_frontend_destroy(self._f)
            self._f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_clock(self, unix_seconds: int) -> None:
        """pin the Date header of every batch (tests); 0 = wall clock"""
        _abi.check(_abi.lib().gofr_frontend_set_clock(self._f, unix_seconds), "gofr_frontend_set_clock")

    def stats(self) -> Tuple[int, int]:
        b, r = C.c_uint64(), C.c_uint64()
        _abi.check(_abi.lib().gofr_frontend_stats(self._f, C.byref(b), C.byref(r)), "gofr_frontend_stats")
        return int(b.value), int(r.value)

    def serve(self, method: int, path: bytes, query: bytes = b"", data: bytes = b"", trace_id: bytes = b"\0" * 16,
              flags: int = 0, resp_cap: int = 0) -> Tuple[bytes, int]:
        """One request → (response bytes, meta).  Blocks until the batch it joined has been served.  resp_cap: size of
        the caller's buffer (default: one slot); a response longer than the slot but within resp_cap is served alone."""
        assert len(trace_id) == 16
        cap = resp_cap or self.slot_bytes
        buf = C.create_string_buffer(cap)
        n, meta = C.c_uint32(), C.c_uint32()
        _abi.check(_abi.lib().gofr_frontend_serve(self._f, method, path, len(path), query, len(query), flags, data, len(data),
                                                  trace_id, buf, cap, C.byref(n), C.byref(meta)),
                   "gofr_frontend_serve")
        return buf.raw[:n.value], int(meta.value)

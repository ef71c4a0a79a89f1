"""gofr_frontend_*: many threads hand in single requests, each gets the reference's response for ITS request back.

The per-request call of the reference is Router.ServeHTTP under net/http's conn goroutine (pkg/gofr/httpServer.go:29-33);
the checker is the oracle serving the same requests as one batch."""
import threading

import numpy as np
import pytest

from gofr_b200 import spec as S
from gofr_b200 import synth
from gofr_b200.table import Table
from tests import oracle as O

CLOCK = 1_700_000_000
DATE = S.http_date(CLOCK)


def _requests(batch):
    """(method, path, query, data, trace_id, flags) per request of a RequestBatch"""
    ar = batch.arena.tobytes()
    out = []
    for i in range(batch.n):
        d = batch.desc[i]
        a, pl, ql, dl = int(d["arena_off"]), int(d["path_len"]), int(d["query_len"]), int(d["data_len"])
        b = (a + pl + ql + 3) & ~3
        out.append((int(d["method"]), ar[a:a + pl], ar[a + pl:a + pl + ql], ar[b:b + dl], batch.trace_ids[i].tobytes(),
                    int(d["flags"])))
    return out


def _run(fe, reqs, n_threads):
    got = [None] * len(reqs)
    errs = []

    def work(t):
        try:
            for i in range(t, len(reqs), n_threads):
                m, p, q, d, tid, fl = reqs[i]
                got[i] = fe.serve(m, p, q, d, tid, fl)
        except Exception as e:  # pragma: no cover - reported below
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th), "front-end hung"
    assert not errs, errs
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("which,max_batch,wait_us,threads", [
    ("config2", 64, 200, 16),       # batches close on the timer (16 in flight < 64)
    ("config2", 8, 100000, 32),     # batches close because they are full (the timer is far away)
    ("config4", 256, 500, 24),
    ("config1", 1, 0, 4),           # degenerate: one request per batch
])
def test_frontend_matches_oracle(which, max_batch, wait_us, threads):
    import torch
    assert torch.cuda.is_available()
    from gofr_b200.engine import Engine
    from gofr_b200.frontend import Frontend
    spec, batch = {"config1": (synth.config1_spec(), synth.config1_batch(120)),
                   "config2": (synth.config2_spec(), synth.config2_batch(600, escape_every=5)),
                   "config4": (synth.config4_spec(), synth.config4_batch(1200))}[which]
    o1, f1, m1 = O.OracleTable(spec).serve(batch, DATE)
    ob = o1.tobytes()
    eng = Engine(Table(spec), 0)
    fe = Frontend(eng, max_batch=max_batch, max_wait_us=wait_us, slot_bytes=1024, max_request_bytes=2048)
    fe.set_clock(CLOCK)
    got = _run(fe, _requests(batch), threads)
    for i, (resp, meta) in enumerate(got):
        assert resp == ob[int(f1[i]):int(f1[i + 1])], i
        assert meta == int(m1[i]), i
    batches, served = fe.stats()
    assert served == batch.n and 1 <= batches <= batch.n
    if max_batch == 1:
        assert batches == batch.n
    fe.close()
    eng.close()


@pytest.mark.gpu
def test_frontend_oversize_responses_and_reuse():
    """a response longer than its slot is served again alone through the packed call (same bytes as the oracle); only a
    response longer than the caller's buffer is an error; the front-end keeps serving afterwards"""
    import ctypes as C
    import torch
    assert torch.cuda.is_available()
    from gofr_b200 import _abi
    from gofr_b200.engine import Engine
    from gofr_b200.frontend import Frontend
    spec, batch = synth.config2_spec(), synth.config2_batch(40)
    o1, f1, m1 = O.OracleTable(spec).serve(batch, DATE)
    ob = o1.tobytes()
    eng = Engine(Table(spec), 0)
    reqs = _requests(batch)
    fe = Frontend(eng, max_batch=4, max_wait_us=50, slot_bytes=64, max_request_bytes=2048)   # every response is longer than 64
    fe.set_clock(CLOCK)
    # Frontend.serve hands in a buffer of slot_bytes: too small for the caller as well → CAPACITY with the length set
    with pytest.raises(_abi.GofrError):
        fe.serve(*reqs[0])
    # with room on the caller's side the oversize response arrives through the packed path
    got = _run_raw(fe, reqs[:12], 4, resp_cap=2048)
    for i, (resp, meta) in enumerate(got):
        assert resp == ob[int(f1[i]):int(f1[i + 1])] and meta == int(m1[i]), i
    with pytest.raises(_abi.GofrError):           # request larger than max_request_bytes: refused before it is queued
        Frontend(eng, max_batch=4, max_wait_us=50, slot_bytes=1024, max_request_bytes=16).serve(0, b"/" + b"a" * 200)
    fe.close()
    fe = Frontend(eng, max_batch=4, max_wait_us=50, slot_bytes=1024, max_request_bytes=2048)
    fe.set_clock(CLOCK)
    for i in (0, 1, 2):
        resp, meta = fe.serve(*reqs[i])
        assert resp == ob[int(f1[i]):int(f1[i + 1])] and meta == int(m1[i])
    fe.close()
    eng.close()


def _run_raw(fe, reqs, n_threads, resp_cap):
    """like _run, with a caller buffer larger than the slot (calls the C entry point directly)"""
    import ctypes as C
    from gofr_b200 import _abi
    got = [None] * len(reqs)
    errs = []

    def work(t):
        try:
            for i in range(t, len(reqs), n_threads):
                m, p, q, d, tid, fl = reqs[i]
                buf = C.create_string_buffer(resp_cap)
                n, meta = C.c_uint32(), C.c_uint32()
                _abi.check(_abi.lib().gofr_frontend_serve(fe._f, m, p, len(p), q, len(q), fl, d, len(d), tid, buf, resp_cap,
                                                          C.byref(n), C.byref(meta)), "gofr_frontend_serve")
                got[i] = (buf.raw[:n.value], int(meta.value))
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in th) and not errs, errs
    return got


def test_frontend_argument_checks():
    """no GPU: the entry points refuse bad arguments instead of touching them"""
    import ctypes as C
    from gofr_b200 import _abi
    L = _abi.lib()
    f = C.c_void_p()
    assert L.gofr_frontend_create(C.byref(f), None, 16, 100, 1024, 4096) == 1
    assert L.gofr_frontend_create(None, None, 16, 100, 1024, 4096) == 1
    n = C.c_uint32()
    assert L.gofr_frontend_serve(None, 0, b"/", 1, b"", 0, 0, b"", 0, b"\0" * 16, None, 0, C.byref(n), None) == 1
    assert L.gofr_frontend_stats(None, None, None) == 1
    assert L.gofr_frontend_set_clock(None, 1) == 1
    L.gofr_frontend_destroy(None)


@pytest.mark.parametrize("args", ["16 200 8 100", "32 100 64 50", "8 100 1 0", "64 50 16 100000", "4 50 1000 200", "16 100 8 100 64"])
def test_frontend_fan_in_fan_out_tsan(args):
    """no GPU: the batching logic itself (frontend.cpp) against a stub engine, under ThreadSanitizer — every producer
    gets the response built from ITS request, batches never exceed max_batch, nothing races."""
    import json
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
    exe = os.path.join(here, "frontend_stress")
    srcs = [os.path.join(here, "frontend_stress.cpp"),
            os.path.join(os.path.dirname(here), "..", "gofr_b200", "csrc", "frontend.cpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-Wall", srcs[0], "-o", exe],
                              cwd=here)
    r = subprocess.run([exe] + args.split(), capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    t, n, max_batch = (int(v) for v in args.split()[:3])
    assert res["bad"] == 0 and res["requests"] == t * n and res["largest"] <= max_batch
    if len(args.split()) > 4:                       # 64-byte slots: the responses that carry a body do not fit and went the packed way, alone
        assert res["alone"] >= t * n // 3

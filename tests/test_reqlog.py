"""RequestLog JSON line: middleware.Logging → logger.Log
(pkg/gofr/http/middleware/logger.go:24-33,41-84; pkg/gofr/logging/logger.go:37-74; logging/level.go:64-70).

CPU: the oracle (oracle/orc_reqlog.c) against the reference's own test cases for getIPAddress, against an independent
Python restatement of Go's time formatting, and against json.loads; the kernel's per-record code (reqlog_device.cuh)
compiled for the host against the oracle.  GPU: gofr_requestlog_device against the oracle, byte for byte."""
import datetime
import json

import numpy as np
import pytest

from tests import oracle as O
from tests.emu import emu
from gofr_b200 import spec as S
from gofr_b200 import synth


def _lines(out, off):
    return O.responses(out, off)


def _one(**kw):
    rec = S.LogRec(1_700_000_000_123_456_789, 1_234_567, 1_700_000_000_125_000_000, **kw)
    out, off = O.request_log(S.LogBatch.pack([rec]))
    return json.loads(_lines(out, off)[0])


# ---- pins from the reference's own tests (middleware/logger_test.go:11-37) ----
def test_get_ip_address_remote_addr():
    assert _one(remote_addr=b"0.0.0.0:8080")["message"]["ip"] == "0.0.0.0:8080"


def test_get_ip_address_forwarded_for():
    assert _one(xff=b"192.168.0.1:8080", remote_addr=b"10.1.1.1:1")["message"]["ip"] == "192.168.0.1:8080"


def test_get_ip_address_rules():
    # ips := strings.Split(xff, ","); ips[0]; "" → RemoteAddr; strings.TrimSpace (logger.go:72-84)
    assert _one(xff=b" 1.2.3.4 , 5.6.7.8", remote_addr=b"r")["message"]["ip"] == "1.2.3.4"
    assert _one(xff=b",5.6.7.8", remote_addr=b" r\t")["message"]["ip"] == "r"
    assert "ip" not in _one(xff=b" ", remote_addr=b"r")["message"]          # " " != "" → trimmed to "" → omitted
    assert _one(xff="  x　".encode(), remote_addr=b"r")["message"]["ip"] == "x"
    assert _one(xff=b"\xa0x", remote_addr=b"r")["message"]["ip"] == "�x"  # a lone continuation byte is no space


def test_line_shape_and_omitempty():
    m = _one(method=b"GET", user_agent=b"ua", remote_addr=b"1.1.1.1:5", uri=b"/x?y=1", status=201)
    assert list(m.keys()) == ["Level", "time", "message"] and m["Level"] == "INFO"
    assert list(m["message"].keys()) == ["id", "start_time", "response_time", "method", "user_agent", "ip", "uri", "response"]
    assert m["message"]["response_time"] == 1234 and m["message"]["response"] == 201
    rec = S.LogRec(1_700_000_000_000_000_000, 999, 1_700_000_000_000_000_000, method=b"", uri=b"", status=0)
    out, off = O.request_log(S.LogBatch.pack([rec]))
    line = _lines(out, off)[0]
    assert line.endswith(b'+00:00"}}\n') and list(json.loads(line)["message"].keys()) == ["id", "start_time"]
    assert json.loads(line)["time"] == "2023-11-14T22:13:20Z"


# ---- Go time formatting, restated independently with datetime ----
def _go_time(unix_ns: int, off: int, zulu: bool) -> str:
    sec, ns = divmod(unix_ns, 1_000_000_000)
    t = datetime.datetime(1970, 1, 1) + datetime.timedelta(seconds=sec + off)
    s = t.strftime("%Y-%m-%dT%H:%M:%S")
    if len(s) < 19:
        s = s.rjust(19, "0")
    if ns:
        s += "." + ("%09d" % ns).rstrip("0")
    if zulu and off == 0:
        return s + "Z"
    zone = int(off / 60)  # truncation toward zero
    return s + ("-" if zone < 0 else "+") + "%02d:%02d" % (abs(zone) // 60, abs(zone) % 60)


def test_time_formats_against_datetime():
    rng = np.random.default_rng(7)
    recs, want = [], []
    offs = [0, 3600, -3600, 19800, -12600, 14 * 3600, -12 * 3600, 20745, -1, 59, -59]
    for k in range(3000):
        ns = int(rng.integers(-(2 ** 62), 2 ** 62)) if k % 3 else int(rng.integers(0, 4_200_000_000)) * 1_000_000_000
        if k % 5 == 0:
            ns -= ns % 1000
        if k % 11 == 0:
            ns -= ns % 1_000_000_000
        ln = int(rng.integers(-(2 ** 62), 2 ** 62))
        off = offs[k % len(offs)]
        recs.append(S.LogRec(ns, 0, ln, method=b"", uri=b"", status=0, tz_offset_s=off))
        want.append((_go_time(ln, off, True), _go_time(ns, off, False)))
    out, off_ = O.request_log(S.LogBatch.pack(recs))
    for line, (t, st) in zip(_lines(out, off_), want):
        m = json.loads(line)
        assert m["time"] == t and m["message"]["start_time"] == st


def test_json_round_trip_of_strings():
    b = synth.reqlog_batch(512, hostile_every=3)
    out, off = O.request_log(b)
    arena = b.arena.tobytes()
    for i, line in enumerate(_lines(out, off)):
        m = json.loads(line)["message"]
        d = b.desc[i]
        o = int(d["arena_off"])
        method = arena[o:o + int(d["method_len"])]
        assert m.get("method", "") == method.decode("utf-8", "replace")
        uri = arena[o + int(d["method_len"]) + int(d["ua_len"]) + int(d["xff_len"]) + int(d["remote_len"]):][:int(d["uri_len"])]
        assert m.get("uri", "") == uri.decode("utf-8", "replace")
        assert m["id"] == b.trace_ids[i].tobytes().hex()
        assert m.get("response_time", 0) == int(int(d["elapsed_ns"]) / 1000)
        assert m.get("response", 0) == int(d["status"])


# ---- the gRPC interceptor's line (pkg/gofr/grpc/log.go:15-50): message = the STRING json.Marshal(RPCLog) ----
def test_rpclog_inner_document_matches_reference_golden():
    # grpc/log_test.go:21-31 pins RPCLog.String(); here the same document must come back out of the outer string
    rec = S.LogRec(1_577_880_732_000_000_000, 0, 1_577_880_732_000_000_000, method=b"GET", kind=S.LOG_RPC,
                   trace_id=bytes.fromhex("b00ff8de800911ec8f6502bfe7568078"))
    out, off = O.request_log(S.LogBatch.pack([rec]))
    entry = json.loads(_lines(out, off)[0])
    assert entry["Level"] == "INFO" and isinstance(entry["message"], str)
    assert entry["message"] == ('{"id":"b00ff8de800911ec8f6502bfe7568078","startTime":"2020-01-01T12:12:12+00:00",'
                                '"responseTime":0,"method":"GET"}')
    want = O.rpclog_string("b00ff8de800911ec8f6502bfe7568078", "2020-01-01T12:12:12+00:00", 0, "GET")
    assert entry["message"].encode() == want


def test_rpclog_double_escaping():
    m = b'/a"b\\<>&\n\x01\xff\xe2\x80\xa8\xc3\xa9'
    rec = S.LogRec(1_700_000_000_000_000_001, 5_999, 1_700_000_000_000_007_000, method=m, kind=S.LOG_RPC, tz_offset_s=3600)
    out, off = O.request_log(S.LogBatch.pack([rec]))
    line = _lines(out, off)[0]
    inner = json.loads(json.loads(line)["message"])
    assert inner["method"] == m.decode("utf-8", "replace") and inner["responseTime"] == 5
    assert inner["startTime"] == "2023-11-14T23:13:20.000000001+01:00"
    assert b'\\\\u003c' in line and b'\\\\\\"' in line  # \\u003c and \\\" on the wire


# ---- the kernel's per-record code on the CPU ----
@pytest.mark.parametrize("mis", [0, 1, 7, 15])
def test_emu_matches_oracle(mis):
    b = synth.reqlog_batch(700, hostile_every=2, tz_offset_s=19800, rpc_every=3)
    o1, f1 = O.request_log(b)
    o2, f2 = emu.request_log(b, mis)
    assert np.array_equal(f1 + mis, f2)
    assert o1[:f1[-1]].tobytes() == o2[mis:f2[-1]].tobytes()
    assert (o2[:mis] == 0xEE).all() and (o2[f2[-1]:f2[-1] + 64] == 0xEE).all()  # nothing written outside the lines


def test_emu_empty_and_single():
    e = S.LogBatch.pack([])
    o, f = emu.request_log(e)
    assert f.tolist() == [0]
    b = synth.reqlog_batch(1)
    o1, f1 = O.request_log(b)
    o2, f2 = emu.request_log(b, 3)
    assert o1[:f1[-1]].tobytes() == o2[3:f2[-1]].tobytes()


from hypothesis import given, settings, strategies as st  # noqa: E402

_bytes = st.one_of(st.binary(max_size=24), st.text(alphabet=' \t,"\\<&>\u00a0\u2028\u3000\u0085abc:[]1.', max_size=16).map(lambda t: t.encode()),
                   st.sampled_from([b"", b" ", b",", b"\xc2", b"\xe2\x80", b"\xe2\x80\xa8", b" \xe3\x80\x80x\xe1\x9a\x80 "]))


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(_bytes, _bytes, _bytes, _bytes, _bytes, st.integers(-(2 ** 62), 2 ** 62), st.integers(-10 ** 12, 10 ** 12),
                          st.integers(0, 999), st.integers(-50400, 50400), st.booleans()), min_size=1, max_size=6),
       st.integers(0, 15))
def test_emu_random_records_property(recs, mis):
    batch = S.LogBatch.pack([S.LogRec(t, el, t + 7, method=m, user_agent=ua, xff=x, remote_addr=ra, uri=u, status=stt, tz_offset_s=tz,
                                      kind=S.LOG_RPC if rpc else S.LOG_REQUEST)
                             for (m, ua, x, ra, u, t, el, stt, tz, rpc) in recs])
    o1, f1 = O.request_log(batch)
    o2, f2 = emu.request_log(batch, mis)
    assert np.array_equal(f1 + mis, f2)
    assert o1[:f1[-1]].tobytes() == o2[mis:f2[-1]].tobytes()
    for line in _lines(o1, f1):
        json.loads(line)  # every line is valid JSON whatever the input bytes


# ---- GPU ----
def _gpu_lines(eng, b):
    d_out, d_off = eng.request_log_device(b)
    off = d_off.cpu().numpy().view(np.uint32)
    return d_out[:int(off[-1])].cpu().numpy(), off


@pytest.fixture(scope="module")
def eng():
    from gofr_b200.engine import Engine
    from gofr_b200.table import Table
    t = Table(synth.config1_spec())
    e = Engine(t, 0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,hostile", [(1, 0), (127, 1), (1000, 2), (4097, 5), (200_000, 97)])
def test_gpu_matches_oracle(eng, n, hostile):
    b = synth.reqlog_batch(n, hostile_every=hostile, tz_offset_s=-12600 if n % 2 else 0, rpc_every=3 if n > 100 else 0)
    o1, f1 = O.request_log(b)
    o2, f2 = _gpu_lines(eng, b)
    assert np.array_equal(f1, f2)
    assert o1[:f1[-1]].tobytes() == o2.tobytes()
    assert not eng.overflowed()


@pytest.mark.gpu
def test_gpu_long_strings_unstaged_tiles(eng):
    """records whose strings exceed the kernel's 24 KB staging budget per tile are read straight from HBM"""
    recs = []
    for k in range(700):
        ua = (b"Mozilla/5.0 <" + b"x" * (300 + k % 200) + b"> \xe2\x80\xa8" + b"\xff" * (k % 3))
        recs.append(S.LogRec(1_700_000_000_000_000_000 + k * 1_000_003, 1000 + k, 1_700_000_000_500_000_000 + k,
                             method=b"POST", user_agent=ua, xff=b" " * (k % 5) + b"10.0.%d.%d , 1.1.1.1" % (k % 256, k // 256),
                             remote_addr=b"[2001:db8::%x]:443" % k, uri=b"/upload/" + b"a" * (k % 300) + b"?q=\"%d\"" % k,
                             status=200 + k % 300, tz_offset_s=(k % 27 - 13) * 1800, kind=S.LOG_RPC if k % 11 == 0 else S.LOG_REQUEST))
    b = S.LogBatch.pack(recs)
    o1, f1 = O.request_log(b)
    o2, f2 = _gpu_lines(eng, b)
    assert np.array_equal(f1, f2)
    assert o1[:f1[-1]].tobytes() == o2.tobytes()


@pytest.mark.gpu
def test_gpu_empty_and_overflow(eng):
    d_out, d_off = eng.request_log_device(S.LogBatch.pack([]))
    assert d_off.cpu().numpy().tolist() == [0]
    b = synth.reqlog_batch(300)
    eng.request_log_device(b, out_cap=1000)
    import torch
    torch.cuda.synchronize()
    assert eng.overflowed()

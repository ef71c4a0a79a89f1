"""GOFR_H_RESULT — stage 2 of the split API: the closure ran on the host, Responder.Respond runs on its (data, err)
(pkg/gofr/http/responder.go:19-62; handler.ServeHTTP pkg/gofr/handler.go:32-36).  Stage 1 is gofr_route_device
(tests/test_route.py)."""
import os

import numpy as np
import pytest

from gofr_b200 import spec as S
from gofr_b200 import synth
from gofr_b200.table import Table
from tests import oracle as O
from tests.emu import emu

DATE = S.http_date(1_700_000_000)
ITEM = S.Schema(7, "main.Item", [S.Field("SKU", S.F_STRING, "sku"), S.Field("Qty", S.F_INT32, "qty"),
                                 S.Field("Note", S.F_STRING, "note", omitempty=True)])


def _spec(frame=S.FRAME_WIRE) -> S.TableSpec:
    return S.TableSpec(frame_mode=frame, schemas=[ITEM, synth.C2_SCHEMA], routes=[
        S.Route(S.M_GET, "/items/{id}", S.H_RESULT, schema_id=7),
        S.Route(S.M_POST, "/items", S.H_RESULT, schema_id=7),
        S.Route(S.M_GET, "/profile", S.H_RESULT, schema_id=1),
    ])


def _batch() -> S.RequestBatch:
    row = ITEM.encode_row(["A-1", 3, "fragile <glass>"])
    prof = synth.C2_SCHEMA.encode_row([9484229377066216, "n" * 10, "e@x", True, 5])
    R, rec = S.Req, S.result_record
    reqs = [R(S.M_GET, b"/items/1", data=rec(S.RESULT_DATA, row)),
            R(S.M_GET, b"/items/2", data=rec(S.RESULT_ERROR, b"db: connection refused")),
            R(S.M_GET, b"/items/3", data=rec(S.RESULT_NIL)),
            R(S.M_GET, b"/items/4", data=rec(S.RESULT_MISSING, b"http: no such file")),
            R(S.M_POST, b"/items", data=rec(S.RESULT_ERROR, b'bad "json" <\xff> \xe2\x80\xa8 line\n2')),
            R(S.M_POST, b"/items", data=rec(S.RESULT_DATA, ITEM.encode_row(["", 0, ""]))),
            R(S.M_GET, b"/profile", data=rec(S.RESULT_DATA, prof)),
            R(S.M_HEAD, b"/items/9", data=rec(S.RESULT_ERROR, b"x")),      # GET-only route: 405 before the handler
            R(S.M_GET, b"/items/5", data=b"\x09\x00\x00\x00"),              # unknown outcome → answered like a panic
            R(S.M_GET, b"/items/6", data=b"\x01\x00\x00\x00\x10\x00\x00\x00ab"),  # message longer than the record
            R(S.M_GET, b"/items/7", data=b""),
            R(S.M_GET, b"/items/8", data=rec(S.RESULT_DATA, row[:6])),      # truncated row
            R(S.M_GET, b"/items/10", data=rec(S.RESULT_ERROR, b"")),
            R(S.M_OPTIONS, b"/items/1", data=rec(S.RESULT_NIL)),
            R(S.M_GET, b"/items/11", data=S.result_both(ITEM, ["A-2", -7, ""], b'partial: 2 of 3 "shards" <down>')),
            R(S.M_GET, b"/profile", data=S.result_both(synth.C2_SCHEMA, [5, "n", "e", False, -1], b"")),
            R(S.M_GET, b"/items/12", data=S.result_both(ITEM, ["A", 1, "n"], b"msg")[:14])]   # truncated → panic
    return S.RequestBatch.pack(reqs)


def test_oracle_bodies():
    spec = _spec(S.FRAME_BODY)
    out, off, meta = O.OracleTable(spec).serve(_batch(), DATE)
    r = O.responses(out, off)
    st = [int(m) & 0xFFFF for m in meta]
    assert r[0] == b'{"data":{"sku":"A-1","qty":3,"note":"fragile \\u003cglass\\u003e"}}\n' and st[0] == 200
    assert r[1] == b'{"error":{"message":"db: connection refused"}}\n' and st[1] == 500
    assert r[2] == b"{}\n" and st[2] == 200
    assert r[3] == b'{"error":{"message":"http: no such file"}}\n' and st[3] == 404
    assert r[4] == b'{"error":{"message":"bad \\"json\\" \\u003c\\ufffd\\u003e \\u2028 line\\n2"}}\n' and st[4] == 500
    assert r[5] == b'{"data":{"sku":"","qty":0}}\n'
    assert st[8] == st[9] == st[10] == st[11] == 500 and b"Some unexpected error" in r[8]
    assert r[12] == b'{"error":{"message":""}}\n' and st[12] == 500
    assert st[13] == 200 and r[13] == b""          # CORS answers OPTIONS (catch-all matches), handler never runs
    # (data, err) both non-nil: response{Error, Data} carries both members, status from the error (responder.go:19-62)
    assert r[14] == (b'{"error":{"message":"partial: 2 of 3 \\"shards\\" \\u003cdown\\u003e"},'
                     b'"data":{"sku":"A-2","qty":-7}}\n') and st[14] == 500
    assert r[15] == b'{"error":{"message":""},"data":{"id":5,"name":"n","email":"e","active":false,"count":-1}}\n'
    assert st[16] == 500 and b"Some unexpected error" in r[16]


@pytest.mark.parametrize("frame", [S.FRAME_WIRE, S.FRAME_INTENDED, S.FRAME_BODY])
@pytest.mark.parametrize("mis", [0, 5])
def test_emu_matches_oracle(frame, mis):
    spec, b = _spec(frame), _batch()
    o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
    o2, f2, m2 = emu.serve(Table(spec).serialize(), b, DATE, misalign=mis)
    assert np.array_equal(m1, m2) and np.array_equal(f1 + mis, f2)
    assert o1[:f1[-1]].tobytes() == o2[mis:f2[-1]].tobytes()


@pytest.mark.gpu
def test_gpu_split_api_round_trip():
    """route on the GPU → closures on the host (Python here) → Respond + framing on the GPU; compared with the oracle
    serving the same result records."""
    from gofr_b200.engine import Engine
    spec = _spec()
    eng = Engine(Table(spec), 0)
    n = 20000
    rng = np.random.default_rng(3)
    paths = [b"/items/%d" % k if k % 3 else (b"/profile" if k % 2 else b"/nothing/%d" % k) for k in range(n)]
    probe = S.RequestBatch.pack([S.Req(S.M_GET, p) for p in paths])
    meta, vars_ = eng.route_device(eng.upload(probe))
    meta, vars_ = meta.cpu().numpy().view(np.uint32), vars_.cpu().numpy().view(np.uint32)
    reqs = []
    for k in range(n):
        st, route = int(meta[k]) & 0xFFFF, int(meta[k]) >> 16
        data = b""
        if st == 0 and route == 0:      # closure of /items/{id}: reads c.PathParam("id")
            off, ln = int(vars_[k, 0]) & 0xFFFF, int(vars_[k, 0]) >> 16
            ident = int(paths[k][off:off + ln])
            if ident % 5 == 0:
                data = S.result_record(S.RESULT_ERROR, b"item %d is out of stock" % ident)
            elif ident % 7 == 0:
                data = S.result_record(S.RESULT_NIL)
            else:
                data = S.result_record(S.RESULT_DATA, ITEM.encode_row(["SKU-%d" % ident, ident % 100, "" if ident % 2 else "n&b"]))
        elif st == 0 and route == 2:    # closure of /profile
            data = S.result_record(S.RESULT_DATA, synth.C2_SCHEMA.encode_row([k, "name%d" % k, "e%d@x" % k, k % 2 == 0, k % 1000]))
        reqs.append(S.Req(S.M_GET, paths[k], data=data))
    b = S.RequestBatch.pack(reqs)
    b.trace_ids[:] = probe.trace_ids
    o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
    resp = eng.alloc_responses(n, int(f1[-1]) + 4096)
    eng.serve_device(eng.upload(b), DATE, resp)
    out, off, m2 = resp.to_host()
    assert np.array_equal(m1, m2) and np.array_equal(f1, off)
    assert o1[:f1[-1]].tobytes() == out.tobytes()
    assert (m1 & 0xFFFF == 500).sum() > 100 and (m1 & 0xFFFF == 404).sum() > 100 and (m1 & 0xFFFF == 200).sum() > 100
    eng.close()


# ---- GOFR_RESULT_STRING: the closure returned a Go string (most handlers of the reference's examples do,
#      examples/http-server/main.go:29-41; gofr_test.go:95-105 expects {"data":"Hello World!"} etc.) ----

def _string_spec(frame=S.FRAME_WIRE) -> S.TableSpec:
    return S.TableSpec(frame_mode=frame, schemas=[ITEM], routes=[
        S.Route(S.M_GET, "/hello", S.H_RESULT),                    # no schema: strings, errors and nil only
        S.Route(S.M_GET, "/items/{id}", S.H_RESULT, schema_id=7),  # the same outcome on a route that also has a schema
    ])


def _string_batch() -> S.RequestBatch:
    R, rec = S.Req, S.result_record
    reqs = [R(S.M_GET, b"/hello", data=rec(S.RESULT_STRING, b"Hello World!")),
            R(S.M_GET, b"/hello", b"name=Vikash", data=rec(S.RESULT_STRING, b"Hello Vikash!")),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_STRING, b"")),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_STRING, b'<b>"q"</b> & \\ \x01\x7f \xe2\x80\xa8\xe2\x80\xa9 \xff\xfe caf\xc3\xa9\n')),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_STRING, b"x" * 3000)),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_ERROR, b"some error occurred")),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_NIL)),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_DATA, ITEM.encode_row(["A", 1, ""]))),   # no schema on this route → panic
            R(S.M_GET, b"/hello", data=S.result_both(ITEM, ["A", 1, ""], b"m")),             # likewise
            R(S.M_GET, b"/hello", data=b"\x05\x00\x00\x00\x09\x00\x00\x00abc"),              # string longer than the record
            R(S.M_GET, b"/hello", data=b"\x05\x00\x00\x00"),                                 # no length word
            R(S.M_HEAD, b"/hello", data=rec(S.RESULT_STRING, b"Hello")),                     # GET-only route, the catch-all answers
            R(S.M_GET, b"/items/1", data=rec(S.RESULT_STRING, b"one")),
            R(S.M_GET, b"/items/2", data=rec(S.RESULT_DATA, ITEM.encode_row(["A-2", 2, "n"]))),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_STRING, "Success ✓ – naïve".encode()))]
    return S.RequestBatch.pack(reqs)


def test_string_outcome_oracle_bodies():
    out, off, meta = O.OracleTable(_string_spec(S.FRAME_BODY)).serve(_string_batch(), DATE)
    r = O.responses(out, off)
    st = [int(m) & 0xFFFF for m in meta]
    assert r[0] == b'{"data":"Hello World!"}\n' and st[0] == 200          # gofr_test.go:95-105
    assert r[1] == b'{"data":"Hello Vikash!"}\n' and st[1] == 200
    assert r[2] == b'{"data":""}\n' and st[2] == 200
    assert r[3] == (b'{"data":"\\u003cb\\u003e\\"q\\"\\u003c/b\\u003e \\u0026 \\\\ \\u0001\x7f \\u2028\\u2029 \\ufffd\\ufffd caf\xc3\xa9\\n"}\n')
    assert r[4] == b'{"data":"' + b"x" * 3000 + b'"}\n'
    assert r[5] == b'{"error":{"message":"some error occurred"}}\n' and st[5] == 500
    assert r[6] == b"{}\n" and st[6] == 200
    assert st[7] == st[8] == st[9] == st[10] == 500 and all(b"Some unexpected error" in r[k] for k in (7, 8, 9, 10))
    assert st[11] == 404     # method mismatch on /hello, then the catch-all of App.Run matches (gofr.go:104)
    assert r[12] == b'{"data":"one"}\n' and r[13] == b'{"data":{"sku":"A-2","qty":2,"note":"n"}}\n'
    assert r[14] == b'{"data":"Success \xe2\x9c\x93 \xe2\x80\x93 na\xc3\xafve"}\n'
    # the same strings through Python's json module (ensure_ascii off, HTML-safe escapes applied by hand)
    import json
    want = {0: "Hello World!", 1: "Hello Vikash!", 2: "", 4: "x" * 3000, 12: "one", 14: "Success ✓ – naïve"}
    for k, v in want.items():
        assert json.loads(r[k])["data"] == v
    assert json.loads(r[3])["data"] == '<b>"q"</b> & \\ \x01\x7f \u2028\u2029 \ufffd\ufffd café\n'


@pytest.mark.parametrize("frame", [S.FRAME_WIRE, S.FRAME_INTENDED, S.FRAME_BODY])
@pytest.mark.parametrize("mis", [0, 3])
def test_string_outcome_emu_matches_oracle(frame, mis):
    spec, b = _string_spec(frame), _string_batch()
    o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
    o2, f2, m2 = emu.serve(Table(spec).serialize(), b, DATE, misalign=mis)
    assert np.array_equal(m1, m2) and np.array_equal(f1 + mis, f2)
    assert o1[:f1[-1]].tobytes() == o2[mis:f2[-1]].tobytes()


def test_string_outcome_emu_slots():
    spec, b = _string_spec(), _string_batch()
    o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
    out, ln, meta = emu.serve_slots(Table(spec).serialize(), b, DATE, 4096)
    assert np.array_equal(meta, m1) and np.array_equal(ln, np.diff(f1.astype(np.int64)).astype(np.uint32))
    ob = o1.tobytes()
    for i in range(b.n):
        assert out[i, :int(ln[i])].tobytes() == ob[int(f1[i]):int(f1[i + 1])], i


@pytest.mark.gpu
def test_gpu_string_outcome():
    from gofr_b200.engine import Engine
    spec = _string_spec()
    eng = Engine(Table(spec), 0)
    base = _string_batch()
    rng = np.random.default_rng(11)
    reqs = []
    alphabet = np.frombuffer(b'abcXYZ019 <>&"\\\n\t\x01\x7f\xc3\xa9\xe2\x80\xa8\xff', dtype=np.uint8)
    for k in range(20000):
        ln = int(rng.integers(0, 120))
        s = alphabet[rng.integers(0, len(alphabet), ln)].tobytes()
        path = b"/hello" if k % 2 else b"/items/%d" % k
        reqs.append(S.Req(S.M_GET, path, data=S.result_record(S.RESULT_STRING if k % 5 else S.RESULT_ERROR, s)))
    for b in (base, S.RequestBatch.pack(reqs)):
        o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
        resp = eng.alloc_responses(b.n, int(f1[-1]) + 4096)
        eng.serve_device(eng.upload(b), DATE, resp)
        out, off, m2 = resp.to_host()
        assert np.array_equal(m1, m2) and np.array_equal(f1, off)
        assert o1[:f1[-1]].tobytes() == out.tobytes()
        o_s, ln_s, m_s = eng.serve_device_slots(eng.upload(b), DATE, 4096)
        assert np.array_equal(m_s.cpu().numpy().view(np.uint32), m1)
        assert np.array_equal(ln_s.cpu().numpy().view(np.uint32), np.diff(f1.astype(np.int64)).astype(np.uint32))
    eng.close()


# ---- property test: any byte string a closure may return, three ways (oracle, device code on the CPU, Python's json) ----
from hypothesis import given, settings, strategies as st  # noqa: E402

from tests.test_independent_checks import _go_json_string_from_python  # noqa: E402


@settings(max_examples=120, deadline=None)
@given(st.lists(st.one_of(st.binary(max_size=60),
                          st.text(alphabet=st.characters(blacklist_categories=("Cs",)), max_size=40).map(lambda s: s.encode("utf-8")),
                          st.sampled_from([b"<>&", b"\xe2\x80\xa8", b"\xff\xfe", b"\\\"", b"\x00\x1f\x7f", b"x" * 700])),
                min_size=1, max_size=12), st.integers(0, 15))
def test_string_outcome_property(strings, mis):
    spec = _string_spec(S.FRAME_BODY)
    b = S.RequestBatch.pack([S.Req(S.M_GET, b"/hello", data=S.result_record(S.RESULT_STRING if k % 4 else S.RESULT_ERROR, s))
                             for k, s in enumerate(strings)])
    o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
    o2, f2, m2 = emu.serve(Table(spec).serialize(), b, DATE, misalign=mis)
    assert np.array_equal(m1, m2) and np.array_equal(f1 + mis, f2)
    assert o1[:f1[-1]].tobytes() == o2[mis:f2[-1]].tobytes()
    for k, (s, r) in enumerate(zip(strings, O.responses(o1, f1))):
        try:
            text = s.decode("utf-8")
        except UnicodeDecodeError:
            continue                      # invalid UTF-8 becomes U+FFFD per byte sequence: covered by the oracle's own vectors
        if "\\" in text:
            continue                      # see test_independent_checks: the textual replacement trick needs no backslashes
        want = _go_json_string_from_python(text)
        assert r == (b'{"data":' if k % 4 else b'{"error":{"message":') + want + (b"}\n" if k % 4 else b"}}\n"), k


# ---- response.Raw (pkg/gofr/http/response/raw.go:3-5): Respond encodes Raw.Data bare — no envelope — and the error, if any,
#      only picks the status code (responder.go:19-26).  responder_test.go:21 pins Raw{} -> Content-Type application/json ----

def _raw_batch() -> S.RequestBatch:
    R, rec = S.Req, S.result_record
    row = ITEM.encode_row(["A-1", 3, "n<1>"])
    reqs = [R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_NIL)),                                    # response.Raw{}
            R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_STRING, b"plain text")),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_STRING, b'<b>"q"</b> & \\ \x01 \xe2\x80\xa8 \xff caf\xc3\xa9\n')),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_STRING, b"")),
            R(S.M_GET, b"/items/1", data=rec(S.RESULT_RAW_DATA, row)),
            R(S.M_GET, b"/items/1", data=rec(S.RESULT_RAW_DATA, row, S.RAW_ERR)),                  # (Raw{...}, err): 500, same body
            R(S.M_GET, b"/items/1", data=rec(S.RESULT_RAW_NIL, b"", S.RAW_MISSING)),               # (Raw{}, ErrMissingFile): 404 null
            R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_STRING, b"x", S.RAW_ERR)),
            R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_DATA, row)),                               # no schema on this route: panic
            R(S.M_GET, b"/hello", data=(S.RESULT_RAW_NIL | 3 << 8).to_bytes(4, "little")),         # unknown error selector: panic
            R(S.M_GET, b"/hello", data=(S.RESULT_STRING | 1 << 8).to_bytes(4, "little") + b"\x01\0\0\0x"),  # selector on a non-Raw outcome
            R(S.M_GET, b"/hello", data=(9).to_bytes(4, "little")),                                 # unknown outcome
            R(S.M_GET, b"/items/2", data=rec(S.RESULT_RAW_DATA, row[:5])),                         # truncated row
            R(S.M_GET, b"/hello", data=rec(S.RESULT_RAW_STRING, b"y" * 2000)),
            R(S.M_HEAD, b"/hello", data=rec(S.RESULT_RAW_NIL)),
            R(S.M_OPTIONS, b"/items/7", data=rec(S.RESULT_RAW_NIL))]
    return S.RequestBatch.pack(reqs)


def test_raw_oracle_bodies():
    out, off, meta = O.OracleTable(_string_spec(S.FRAME_BODY)).serve(_raw_batch(), DATE)
    r = O.responses(out, off)
    st = [int(m) & 0xFFFF for m in meta]
    assert r[0] == b"null\n" and st[0] == 200                      # json.Encoder.Encode(nil interface)
    assert r[1] == b'"plain text"\n' and st[1] == 200
    assert r[2] == b'"\\u003cb\\u003e\\"q\\"\\u003c/b\\u003e \\u0026 \\\\ \\u0001 \\u2028 \\ufffd caf\xc3\xa9\\n"\n'
    assert r[3] == b'""\n'
    assert r[4] == b'{"sku":"A-1","qty":3,"note":"n\\u003c1\\u003e"}\n' and st[4] == 200
    assert r[5] == r[4] and st[5] == 500                             # the error object is computed and dropped
    assert r[6] == b"null\n" and st[6] == 404
    assert r[7] == b'"x"\n' and st[7] == 500
    assert all(st[k] == 500 and b"Some unexpected error" in r[k] for k in (8, 9, 10, 11, 12))
    assert r[13] == b'"' + b"y" * 2000 + b'"\n'
    assert st[14] == 404 and st[15] == 200 and r[15] == b""         # HEAD: GET-only route -> catch-all; OPTIONS: CORS answers
    import json
    assert json.loads(r[0]) is None and json.loads(r[1]) == "plain text" and json.loads(r[4]) == {"sku": "A-1", "qty": 3, "note": "n<1>"}


def test_raw_content_type_pin():
    """pkg/gofr/http/responder_test.go:21 — `{"raw response type", resTypes.Raw{}, "application/json"}` read from the
    live recorder map (GOFR_FRAME_INTENDED); on the wire net/http sniffs the body instead (WriteHeader came first)."""
    import json as _json
    pins = _json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_pins.json")))
    pin = [c for c in pins["responder_content_type"]["cases"] if c["kind"] == "raw"][0]
    b = S.RequestBatch.pack([S.Req(S.M_GET, b"/hello", data=S.result_record(S.RESULT_RAW_NIL))])
    for frame, want in ((S.FRAME_INTENDED, pin["content_type"]), (S.FRAME_WIRE, "text/plain; charset=utf-8")):
        out, off, _ = O.OracleTable(_string_spec(frame)).serve(b, DATE)
        head = O.responses(out, off)[0].split(b"\r\n\r\n")[0].decode().lower()
        assert "content-type: " + want in head, head


@pytest.mark.parametrize("frame", [S.FRAME_WIRE, S.FRAME_INTENDED, S.FRAME_BODY])
@pytest.mark.parametrize("mis", [0, 5])
def test_raw_emu_matches_oracle(frame, mis):
    spec, b = _string_spec(frame), _raw_batch()
    o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
    o2, f2, m2 = emu.serve(Table(spec).serialize(), b, DATE, misalign=mis)
    assert np.array_equal(m1, m2) and np.array_equal(f1 + mis, f2)
    assert o1[:f1[-1]].tobytes() == o2[mis:f2[-1]].tobytes()
    for stage in (0, 1):            # slot layout: general interpreter and the fast path (all requests "staged")
        emu.set_stage_mode(stage)
        try:
            out, ln, meta = emu.serve_slots(Table(spec).serialize(), b, DATE, 4096)
        finally:
            emu.set_stage_mode(0)
        assert np.array_equal(meta, m1) and np.array_equal(ln, np.diff(f1.astype(np.int64)).astype(np.uint32))
        ob = o1.tobytes()
        for i in range(b.n):
            assert out[i, :int(ln[i])].tobytes() == ob[int(f1[i]):int(f1[i + 1])], i


@pytest.mark.gpu
def test_gpu_raw_outcomes():
    from gofr_b200.engine import Engine
    spec = _string_spec()
    eng = Engine(Table(spec), 0)
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b'abcXYZ019 <>&"\\\n\x01\xc3\xa9\xe2\x80\xa8\xff', dtype=np.uint8)
    reqs = []
    for k in range(20000):
        s = alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 90)))].tobytes()
        es = (S.RAW_OK, S.RAW_OK, S.RAW_ERR, S.RAW_MISSING)[k % 4]
        if k % 3 == 0:
            reqs.append(S.Req(S.M_GET, b"/items/%d" % k, data=S.result_record(S.RESULT_RAW_DATA, ITEM.encode_row([s[:20], k, s[20:]]), es)))
        elif k % 3 == 1:
            reqs.append(S.Req(S.M_GET, b"/hello", data=S.result_record(S.RESULT_RAW_STRING, s, es)))
        else:
            reqs.append(S.Req(S.M_GET, b"/hello", data=S.result_record(S.RESULT_RAW_NIL, b"", es)))
    for b in (_raw_batch(), S.RequestBatch.pack(reqs)):
        o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
        resp = eng.alloc_responses(b.n, int(f1[-1]) + 4096)
        eng.serve_device(eng.upload(b), DATE, resp)
        out, off, m2 = resp.to_host()
        assert np.array_equal(m1, m2) and np.array_equal(f1, off)
        assert o1[:f1[-1]].tobytes() == out.tobytes()
        o_s, ln_s, m_s = eng.serve_device_slots(eng.upload(b), DATE, 4096)
        ln = ln_s.cpu().numpy().view(np.uint32)
        assert np.array_equal(m_s.cpu().numpy().view(np.uint32), m1) and np.array_equal(ln, np.diff(f1.astype(np.int64)).astype(np.uint32))
        so, ob = o_s.cpu().numpy().reshape(b.n, 4096), o1.tobytes()
        for i in range(0, b.n, 3):
            assert so[i, :int(ln[i])].tobytes() == ob[int(f1[i]):int(f1[i + 1])], i
    eng.close()


def test_file_responses_are_sniffed_like_detect_content_type():
    """response.File on the wire: the handler's Content-Type is set after WriteHeader and never leaves, net/http sniffs the
    first 512 bytes (http.DetectContentType, the WHATWG MIME Sniffing tables).  Known answers for every signature family, the
    product's seal-time sniffer (table_build.cpp) and the oracle's agreeing byte for byte on the whole response."""
    kat = [(b"", "text/plain; charset=utf-8"), (b"hello", "text/plain; charset=utf-8"), (b"  \n<html><body>", "text/html; charset=utf-8"),
           (b"<HTML lang=en>", "text/html; charset=utf-8"), (b"<htmlx>", "text/plain; charset=utf-8"), (b"<!-- c -->", "text/html; charset=utf-8"),
           (b"<!DOCTYPE html>", "text/html; charset=utf-8"), (b"<p>x", "text/html; charset=utf-8"), (b"<p", "text/plain; charset=utf-8"),
           (b"\t<?xml version='1.0'?>", "text/xml; charset=utf-8"), (b"%PDF-1.7", "application/pdf"), (b"%!PS-Adobe-3.0", "application/postscript"),
           (b"\xfe\xff\x00h", "text/plain; charset=utf-16be"), (b"\xff\xfeh\x00", "text/plain; charset=utf-16le"), (b"\xef\xbb\xbfhi", "text/plain; charset=utf-8"),
           (b"\xfe\xff", "text/plain; charset=utf-8"),                                 # a BOM signature needs four bytes; 0xFE 0xFF are not "binary" bytes
           (b"\x00\x00\x01\x00\x01", "image/x-icon"), (b"\x00\x00\x02\x00", "image/x-icon"), (b"BM6\x00", "image/bmp"),
           (b"GIF87a..", "image/gif"), (b"GIF89a..", "image/gif"), (b"RIFF\x24\x00\x00\x00WEBPVP8 ", "image/webp"),
           (b"\x89PNG\r\n\x1a\n\x00\x00", "image/png"), (b"\xff\xd8\xff\xe0", "image/jpeg"),
           (b"FORM\x00\x00\x10\x00AIFFCOMM", "audio/aiff"), (b"ID3\x03\x00", "audio/mpeg"), (b"OggS\x00\x02", "application/ogg"),
           (b"MThd\x00\x00\x00\x06\x00\x01", "audio/midi"), (b"RIFF\x10\x00\x00\x00AVI LIST", "video/avi"), (b"RIFF\x10\x00\x00\x00WAVEfmt ", "audio/wave"),
           (b"\x00\x00\x00\x18ftypmp42\x00\x00\x00\x00mp42isom", "video/mp4"), (b"\x00\x00\x00\x18ftypisom\x00\x00\x02\x00isomiso2", "application/octet-stream"),
           (b"\x00\x00\x00\x14ftypisom\x00\x00\x02\x00mp41", "video/mp4"), (b"\x00\x00\x00\x10ftypqt  mp4 ", "application/octet-stream"),  # the version word is skipped
           (b"\x1a\x45\xdf\xa3\x9f", "video/webm"), (b"\x01" * 34 + b"LP", "application/vnd.ms-fontobject"), (b"\x00\x01\x00\x00\x00\x0c", "font/ttf"),
           (b"OTTO\x00", "font/otf"), (b"ttcf\x00", "font/collection"), (b"wOFF\x00", "font/woff"), (b"wOF2\x00", "font/woff2"),
           (b"\x1f\x8b\x08\x00", "application/x-gzip"), (b"PK\x03\x04\x14", "application/zip"), (b"Rar!\x1a\x07\x00\xcf", "application/x-rar-compressed"),
           (b"Rar!\x1a\x07\x01\x00", "application/x-rar-compressed"), (b"\x00asm\x01\x00\x00\x00", "application/wasm"),
           (b"text with \x1b escape", "text/plain; charset=utf-8"), (b"bin\x00ary", "application/octet-stream"), (b"\x7f\x80\xff", "text/plain; charset=utf-8"),
           (b"x" * 600 + b"\x00", "text/plain; charset=utf-8"), (b"x" * 511 + b"\x00", "application/octet-stream")]   # only the first 512 bytes are looked at
    for k0 in range(0, len(kat), 10):
        part = kat[k0:k0 + 10]
        spec = S.TableSpec(frame_mode=S.FRAME_WIRE, routes=[S.Route(S.M_GET, "/f%d" % k, S.H_FILE, s0=b"application/json", blob=blob)
                                                           for k, (blob, _) in enumerate(part)])
        b = S.RequestBatch.pack([S.Req(S.M_GET, b"/f%d" % k) for k in range(len(part))])
        o1, f1, m1 = O.OracleTable(spec).serve(b, DATE)
        o2, f2, m2 = emu.serve(Table(spec).serialize(), b, DATE)
        r1, r2 = O.responses(o1, f1), O.responses(o2, f2)
        for (blob, want), a, c in zip(part, r1, r2):
            assert a == c, (blob[:40], a[:300], c[:300])
            head, _, body = a.partition(b"\r\n\r\n")
            assert body == blob
            if blob:
                assert ("Content-Type: " + want).encode() in head, (blob[:40], want, head)
            else:   # an empty file: Content-Length: 0 and nothing to sniff
                assert b"Content-Type" not in head and b"Content-Length: 0" in head

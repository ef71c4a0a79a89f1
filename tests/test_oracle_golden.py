"""Pins the CPU oracle against every result the reference's own tests hold for the hot path
(tests/golden/reference_pins.json), plus known-answer vectors for the upstream arithmetic the path
relies on (Go 1.21 encoding/json escaping, strconv, path.Clean, url.ParseQuery / url escaping).
"""
import json
import os

import pytest

from gofr_b200 import spec as S
from tests import oracle as O

PINS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_pins.json")))
DATE = S.http_date(1789974595)


def _split(resp: bytes):
    head, _, body = resp.partition(b"\r\n\r\n")
    lines = head.split(b"\r\n")
    status = int(lines[0].split(b" ")[1])
    headers = {}
    for ln in lines[1:]:
        k, _, v = ln.partition(b": ")
        headers[k.decode().lower()] = v.decode()
    return status, headers, body


def _spec_from_pin(pin, frame_mode):
    routes = []
    for r in pin["routes"]:
        m = S.method_code(r["method"])
        if r["handler"] == "static_string":
            routes.append(S.Route(m, r["pattern"], S.H_STATIC_STRING, s0=r["text"].encode()))
        else:
            routes.append(S.Route(m, r["pattern"], S.H_PARAM_FORMAT, s0=r["key"].encode(), s1=r["default"].encode(),
                                  s2=r["prefix"].encode(), s3=r["suffix"].encode()))
    return S.TableSpec(frame_mode=frame_mode, routes=routes, default_routes=pin["default_routes"])


@pytest.mark.parametrize("mode", [S.FRAME_WIRE, S.FRAME_INTENDED])
def test_server_routes_pin(mode):
    """pkg/gofr/gofr_test.go:40-107 — data after json.Unmarshal, and (recorder view) content-type application/json."""
    pin = PINS["server_routes"]
    t = O.OracleTable(_spec_from_pin(pin, mode))
    reqs = []
    for c in pin["cases"]:
        path, _, q = c["target"].partition("?")
        reqs.append(S.Req(S.method_code(c["method"]), path.encode(), q.encode()))
    out, off, meta = t.serve(S.RequestBatch.pack(reqs), DATE)
    for c, r, m in zip(pin["cases"], O.responses(out, off), meta):
        status, headers, body = _split(r)
        assert status == 200 and (m & 0xFFFF) == 200
        assert json.loads(body)["data"] == c["data"]
        if mode == S.FRAME_INTENDED:
            assert headers[pin["header"][0]] == pin["header"][1]
        else:  # what net/http really sends, because Respond calls WriteHeader before Header().Set
            assert headers["content-type"] == "text/plain; charset=utf-8"


def test_example_server_status_pin():
    """examples/http-server/main_test.go:21-29 (paths that need no redis/mysql)."""
    from gofr_b200 import synth
    t = O.OracleTable(synth.config1_spec())
    reqs = []
    for c in PINS["example_server_status"]["cases"]:
        path, _, q = c["path"].partition("?")
        reqs.append(S.Req(S.M_GET, path.encode(), q.encode()))
    out, off, meta = t.serve(S.RequestBatch.pack(reqs), DATE)
    for c, r, m in zip(PINS["example_server_status"]["cases"], O.responses(out, off), meta):
        assert _split(r)[0] == c["status"] == (m & 0xFFFF), c


def test_status_from_error_and_handler_status_pins():
    """responder_test.go:43-47 and handler_test.go:29-30."""
    spec = S.TableSpec(frame_mode=S.FRAME_BODY, routes=[
        S.Route(S.M_GET, "/nil", S.H_NIL),
        S.Route(S.M_GET, "/missing", S.H_MISSING_FILE),
        S.Route(S.M_GET, "/timeout", S.H_STATIC_ERROR, s0=b"http: Handler timeout"),
        S.Route(S.M_GET, "/err", S.H_STATIC_ERROR, s0=b"some error")], default_routes=False)
    t = O.OracleTable(spec)
    out, off, meta = t.serve(S.RequestBatch.pack([S.Req(S.M_GET, p) for p in (b"/nil", b"/missing", b"/timeout", b"/err")]), DATE)
    r = O.responses(out, off)
    assert [int(m & 0xFFFF) for m in meta] == [200, 404, 500, 500]
    assert r[0] == PINS["survey_sizes"]["nil_body"].encode()
    assert json.loads(r[1])["error"] == PINS["status_from_error"]["cases"][1]["errObj"]
    assert json.loads(r[2])["error"] == PINS["status_from_error"]["cases"][2]["errObj"]
    assert r[1] == PINS["survey_sizes"]["missing_file_body"].encode()


def test_responder_content_type_pin():
    """responder_test.go:21-23,28: recorder view — File keeps its given type, everything else application/json."""
    spec = S.TableSpec(frame_mode=S.FRAME_INTENDED, routes=[
        S.Route(S.M_GET, "/map", S.H_HEALTH), S.Route(S.M_GET, "/file", S.H_FILE, s0=b"image/png", blob=b"")],
        default_routes=False)
    t = O.OracleTable(spec)
    out, off, _ = t.serve(S.RequestBatch.pack([S.Req(S.M_GET, b"/map"), S.Req(S.M_GET, b"/file")]), DATE)
    r = O.responses(out, off)
    assert _split(r[0])[1]["content-type"] == "application/json"
    assert _split(r[1])[1]["content-type"] == "image/png"


def test_cors_pin():
    """middleware/cors_test.go:30-46."""
    pin = PINS["cors"]
    t = O.OracleTable(S.TableSpec(routes=[S.Route(S.M_GET, "/hello", S.H_STATIC_STRING, s0=b"x")]))
    out, off, _ = t.serve(S.RequestBatch.pack([S.Req(S.M_GET, b"/hello"), S.Req(S.M_OPTIONS, b"/hello")]), DATE)
    for i, r in enumerate(O.responses(out, off)):
        status, h, body = _split(r)
        assert h["access-control-allow-origin"] == pin["allow_origin"]
        assert h["access-control-allow-methods"] == pin["allow_methods"]
        if i == 1:
            assert status == pin["options_status"] and body == pin["options_body"].encode()


def test_param_pin():
    for c in PINS["param"]["cases"]:
        q = c["target"].partition("?")[2]
        assert O.query_get(q.encode(), c["key"].encode()) == c["value"].encode()


def test_bind_pins():
    """request_test.go:17-30, context_test.go:23-49."""
    kinds = {"string": S.F_STRING, "int": S.F_INT}
    for n, c in enumerate(PINS["bind"]["cases"]):
        sc = S.Schema(10 + n, "main.T", [S.Field(g, kinds[k], j) for g, j, k in c["fields"]])
        t = O.OracleTable(S.TableSpec(schemas=[sc], routes=[]))
        ok, row = t.bind(sc.id, c["body"].encode())
        assert ok
        assert row == sc.encode_row(c["values"])


def test_rpclog_json_pin():
    """pkg/gofr/grpc/log_test.go:28 — byte-exact encoding/json output."""
    p = PINS["rpclog_json"]
    assert O.rpclog_string(p["id"], p["startTime"], p["responseTime"], p["method"]) == p["expect"].encode()


def test_say_hello_pin_and_frame():
    import numpy as np
    frames = bytearray()
    offs = [0]
    for name, _ in PINS["say_hello"]["cases"]:
        msg = (b"\x0a" + bytes([len(name)]) + name.encode()) if name else b""
        frames += b"\x00" + len(msg).to_bytes(4, "big") + msg
        offs.append(len(frames))
    out, off, meta = O.grpc_hello(np.frombuffer(bytes(frames), dtype=np.uint8).copy(), np.array(offs, dtype=np.uint32))
    for (name, want), r, m in zip(PINS["say_hello"]["cases"], O.responses(out, off), meta):
        assert m == 0
        assert r[0] == 0 and int.from_bytes(r[1:5], "big") == len(r) - 5
        assert r[5] == 0x0A and r[6] == len(want) and r[7:] == want.encode()
    assert O.responses(out, off)[2].hex() == PINS["survey_sizes"]["grpc_hello_world_frame_hex"]


def test_survey_sizes():
    from gofr_b200 import synth
    s = PINS["survey_sizes"]
    t = O.OracleTable(synth.config1_spec())
    out, off, _ = t.serve(S.RequestBatch.pack([S.Req(S.M_GET, b"/hello"), S.Req(S.M_GET, b"/error"),
                                               S.Req(S.M_GET, b"/.well-known/health")]), DATE)
    r = O.responses(out, off)
    assert len(r[0]) == s["c1_wire_bytes"] and r[0].endswith(s["c1_body"].encode())
    assert json.loads(_split(r[0])[2]) == PINS["documented_shape"]["body_json"]
    assert r[1].endswith(s["error_body"].encode())
    assert r[2].endswith(s["health_body"].encode())
    t2 = O.OracleTable(synth.config2_spec())
    out, off, _ = t2.serve(synth.config2_batch(64), DATE)
    assert {len(x) for x in O.responses(out, off)} == {s["c2_wire_bytes"]}
    tp = O.OracleTable(S.TableSpec(frame_mode=S.FRAME_BODY, routes=[S.Route(S.M_GET, "/p", S.H_PANIC)]))
    out, off, meta = tp.serve(S.RequestBatch.pack([S.Req(S.M_GET, b"/p")]), DATE)
    assert len(O.responses(out, off)[0]) == s["panic_body_len"] and (meta[0] & 0xFFFF) == 500
    assert json.loads(O.responses(out, off)[0]) == {"code": 500, "message": "Some unexpected error has occurred", "status": "ERROR"}


# ---------------------------------------------------------------------------------------------------------------
# known-answer vectors for the upstream arithmetic (hand-derived from the published behaviour of the pinned versions)
# ---------------------------------------------------------------------------------------------------------------
JSON_STRING_KAT = [
    (b"", b'""'), (b"Hello World!", b'"Hello World!"'),
    (b'a"b\\c', b'"a\\"b\\\\c"'), (b"\n\r\t", b'"\\n\\r\\t"'),
    (b"\x00\x01\x1f", b'"\\u0000\\u0001\\u001f"'),
    (b"\x08\x0c", b'"\\u0008\\u000c"'),            # Go 1.21: no \b \f short forms (go.mod:3)
    (b"<script>&", b'"\\u003cscript\\u003e\\u0026"'),  # Encoder escapes HTML by default
    (b"\x7f", b'"\x7f"'),                           # DEL is copied
    ("é€😀".encode(), '"é€😀"'.encode()),           # valid UTF-8 is copied
    ("  ".encode(), b'"\\u2028\\u2029"'),
    (b"\xff", b'"\\ufffd"'), (b"\xc3", b'"\\ufffd"'), (b"\xe2\x82", b'"\\ufffd\\ufffd"'),
    (b"\xc0\xaf", b'"\\ufffd\\ufffd"'),             # overlong
    (b"\xed\xa0\x80", b'"\\ufffd\\ufffd\\ufffd"'),  # surrogate half
    (b"\xf4\x90\x80\x80", b'"\\ufffd\\ufffd\\ufffd\\ufffd"'),  # > U+10FFFF
    ("�".encode(), "\"�\"".encode()),     # a real U+FFFD is valid and copied
]


@pytest.mark.parametrize("src,want", JSON_STRING_KAT)
def test_json_string_kat(src, want):
    assert O.json_string(src) == want


@pytest.mark.parametrize("v", [0, 1, -1, 9, 10, 99, 100, 12345, -12345, 2 ** 31 - 1, -2 ** 31, 2 ** 63 - 1, -2 ** 63,
                               10 ** 15, 10 ** 16 - 1, 10 ** 18])
def test_json_int_kat(v):
    assert O.json_int(v) == str(v).encode()


CLEAN_KAT = [  # mux cleanPath (path.Clean + trailing slash restore)
    (b"", b"/"), (b"/", b"/"), (b"/a", b"/a"), (b"/a/", b"/a/"), (b"//a", b"/a"), (b"/a//b", b"/a/b"), (b"/a/./b", b"/a/b"),
    (b"/a/../b", b"/b"), (b"/../a", b"/a"), (b"/a/b/..", b"/a"), (b"/a/b/../", b"/a/"), (b"/a/..", b"/"), (b"/a/../", b"/"),
    (b"a", b"/a"), (b"/.", b"/"), (b"/..", b"/"), (b"/a/b/../../c/./d//", b"/c/d/"), (b"/...", b"/..."), (b"/a/.b", b"/a/.b"),
]


@pytest.mark.parametrize("src,want", CLEAN_KAT)
def test_clean_path_kat(src, want):
    assert O.clean_path(src) == want


QUERY_KAT = [  # url.ParseQuery(...).Get(key)
    (b"a=b", b"a", b"b"), (b"a=b&a=c", b"a", b"b"), (b"x=1&name=gofr", b"name", b"gofr"), (b"name", b"name", b""),
    (b"name=a+b%20c", b"name", b"a b c"), (b"na%6De=v", b"name", b"v"), (b"name=%zz&name=ok", b"name", b"ok"),
    (b"name=a;b&name=c", b"name", b"c"), (b"&&name=x", b"name", b"x"), (b"name=%4", b"name", b""), (b"name=a=b", b"name", b"a=b"),
    (b"Name=x", b"name", b""), (b"name=%e2%82%ac", b"name", "€".encode()),
]


@pytest.mark.parametrize("q,key,want", QUERY_KAT)
def test_query_get_kat(q, key, want):
    assert O.query_get(q, key) == want


def test_escape_path_kat():
    assert O.escape_path(b"/a b/c?d/%/\xc3\xa9/$&+,:;=@~") == b"/a%20b/c%3Fd/%25/%C3%A9/$&+,:;=@~"


def test_http_date():
    assert O.http_date(1789974595) == b"Mon, 21 Sep 2026 07:09:55 GMT" == S.http_date(1789974595)
    assert O.http_date(0) == b"Thu, 01 Jan 1970 00:00:00 GMT"


def test_mux_method_mismatch_rules():
    """gorilla/mux v1.8.1 Route.Match bookkeeping (marked 'unverified against Go' in DESIGN.md)."""
    spec = S.TableSpec(default_routes=False, routes=[
        S.Route(S.M_GET, "/a", S.H_NIL), S.Route(S.M_POST, "/b", S.H_NIL), S.Route(S.M_GET, "/u/{id:[0-9]+}", S.H_NIL),
        S.Route(S.M_GET, "/f/{name}.json", S.H_NIL), S.Route(S.M_GET, "noslash", S.H_NIL)])
    t = O.OracleTable(spec)
    assert t.match(S.M_GET, b"/a") == 0
    assert t.match(S.M_DELETE, b"/a") == -2          # path matches, method does not → 405
    assert t.match(S.M_POST, b"/a") == -1            # a later route's method matcher succeeds → stale mismatch cleared → 404
    assert t.match(S.M_GET, b"/zzz") == -1
    assert t.match(S.M_GET, b"/u/123") == 2 and t.match(S.M_GET, b"/u/12a") == -1
    assert t.match(S.M_GET, b"/f/x.json") == 3 and t.match(S.M_GET, b"/f/a.json.json") == 3 and t.match(S.M_GET, b"/f/.json") == -1
    assert t.match(S.M_GET, b"noslash") == -3 and t.match(S.M_GET, b"/noslash") == -1  # route.err → never matches
    assert t.match(S.M_GET, b"//a") == -3
    assert t.match(S.M_OTHER, b"/a") == -2
    with_catchall = O.OracleTable(S.TableSpec(routes=[S.Route(S.M_GET, "/a", S.H_NIL)]))
    assert with_catchall.match(S.M_POST, b"/a") == 3  # health, favicon, then the PathPrefix("/") catch-all


BIND_ERR_KAT = [  # err.Error() of json.Unmarshal into struct{ID int64 `json:"id"`; Name string `json:"name"`; OK bool `json:"ok"`; N int32 `json:"n"`}
    (b"", b"unexpected end of JSON input"),
    (b"{", b"unexpected end of JSON input"),
    (b"x", b"invalid character 'x' looking for beginning of value"),
    (b'{"id":1,}', b"invalid character '}' looking for beginning of object key string"),
    (b'{"id" 1}', b"invalid character '1' after object key"),
    (b'{"id":1 "n":2}', b"invalid character '\"' after object key:value pair"),
    (b'{"id":1}x', b"invalid character 'x' after top-level value"),
    (b'{"name":"a\nb"}', b"invalid character '\\n' in string literal"),
    (b'{"name":"\\q"}', b"invalid character 'q' in string escape code"),
    (b'{"name":"\\u12g4"}', b"invalid character 'g' in \\u hexadecimal character escape"),
    (b'{"id":-}', b"invalid character '}' in numeric literal"),
    (b'{"id":1.}', b"invalid character '}' after decimal point in numeric literal"),
    (b'{"id":1e}', b"invalid character '}' in exponent of numeric literal"),
    (b'{"ok":tru}', b"invalid character '}' in literal true (expecting 'e')"),
    (b'{"ok":t', b"invalid character ' ' in literal true (expecting 'r')"),
    (b"[1,2", b"unexpected end of JSON input"),
    (b"[1 2]", b"invalid character '2' after array element"),
    (b'{"id":"7"}', b"json: cannot unmarshal string into Go struct field T.id of type int64"),
    (b'{"id":1.5}', b"json: cannot unmarshal number 1.5 into Go struct field T.id of type int64"),
    (b'{"n":3000000000}', b"json: cannot unmarshal number 3000000000 into Go struct field T.n of type int32"),
    (b'{"name":5}', b"json: cannot unmarshal number into Go struct field T.name of type string"),
    (b'{"ok":"yes"}', b"json: cannot unmarshal string into Go struct field T.ok of type bool"),
    (b'{"name":true}', b"json: cannot unmarshal bool into Go struct field T.name of type string"),
    (b'{"name":{"a":1}}', b"json: cannot unmarshal object into Go struct field T.name of type string"),
    (b'{"id":[1]}', b"json: cannot unmarshal array into Go struct field T.id of type int64"),
    (b"[1]", b"json: cannot unmarshal array into Go value of type main.T"),
    (b'"s"', b"json: cannot unmarshal string into Go value of type main.T"),
    (b"12", b"json: cannot unmarshal number into Go value of type main.T"),
    (b"true", b"json: cannot unmarshal bool into Go value of type main.T"),
    (b'{"id":"x","name":5}', b"json: cannot unmarshal string into Go struct field T.id of type int64"),  # first error wins
]
BIND_SCHEMA = S.Schema(7, "main.T", [S.Field("ID", S.F_INT64, "id"), S.Field("Name", S.F_STRING, "name"),
                                     S.Field("OK", S.F_BOOL, "ok"), S.Field("N", S.F_INT32, "n")])


@pytest.mark.parametrize("body,want", BIND_ERR_KAT)
def test_bind_error_kat(body, want):
    t = O.OracleTable(S.TableSpec(schemas=[BIND_SCHEMA], routes=[]))
    ok, msg = t.bind(7, body)
    assert not ok and msg == want


BIND_OK_KAT = [
    (b'{"id":1,"name":"Bob","ok":true,"n":-5}', [1, "Bob", True, -5]),
    (b' { "ID" : 2 , "NAME" : "x" } ', [2, "x", False, 0]),                 # case-insensitive fallback
    (b'{"id":1,"id":9,"zzz":{"a":[1,2,{"b":null}]},"name":null}', [9, "", False, 0]),  # last wins, unknown skipped, null no-op
    (b'{"name":"a\\u00e9\\ud83d\\ude00\\ud800x\\n\\/"}', [0, "aé😀�x\n/", False, 0]),
    (b'{"name":"\xff"}', [0, "�", False, 0]),
    (b"null", [0, "", False, 0]),
    (b'{"id":-0}', [0, "", False, 0]),
    (b'{"\\u212aey":1}', [0, "", False, 0]),
]


@pytest.mark.parametrize("body,values", BIND_OK_KAT)
def test_bind_ok_kat(body, values):
    t = O.OracleTable(S.TableSpec(schemas=[BIND_SCHEMA], routes=[]))
    ok, row = t.bind(7, body)
    assert ok, row
    assert row == BIND_SCHEMA.encode_row(values)

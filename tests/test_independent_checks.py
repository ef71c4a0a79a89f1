"""Third-party cross-checks of the oracle's building blocks against Python's standard library, on the input subsets
where the Python and Go behaviours are known to coincide (the differences are handled explicitly below).  The oracle is
a restatement written for this repo; these tests tie it to independently written code."""
import json
import posixpath
import urllib.parse

from hypothesis import given, settings, strategies as st

from tests import oracle as O


def _go_json_string_from_python(s: str) -> bytes:
    """encoding/json (Go 1.21, HTML escaping on) derived from json.dumps: differences are <, >, &, U+2028/9 and the
    short escapes \\b \\f, which Go writes as \\u0008 / \\u000c."""
    out = json.dumps(s, ensure_ascii=False)
    out = out.replace("\\b", "\\u0008").replace("\\f", "\\u000c")
    out = out.replace("<", "\\u003c").replace(">", "\\u003e").replace("&", "\\u0026")
    out = out.replace("\u2028", "\\u2028").replace("\u2029", "\\u2029")
    return out.encode("utf-8")


@settings(max_examples=400, deadline=None)
@given(st.text(alphabet=st.characters(blacklist_categories=("Cs",)), max_size=40))
def test_json_string_against_json_dumps(s):
    # json.dumps writes "\\b" for a backspace: a literal backslash followed by b cannot appear otherwise (a backslash in
    # the input is always doubled), so the textual replacement above is exact
    if "\\" in s:
        s = s.replace("\\", "/")  # keep the replacement trick unambiguous
    assert O.json_string(s.encode("utf-8")) == _go_json_string_from_python(s)


@settings(max_examples=200, deadline=None)
@given(st.integers(-(2 ** 63), 2 ** 63 - 1))
def test_json_int_against_str(v):
    assert O.json_int(v) == str(v).encode()


def _mux_clean_path_from_python(p: str) -> str:
    """mux cleanPath (mux.go): "" → "/", ensure a leading "/", path.Clean, keep one trailing slash."""
    if p == "":
        return "/"
    if p[0] != "/":
        p = "/" + p
    np_ = posixpath.normpath(p)
    if np_.startswith("//"):          # POSIX keeps exactly two leading slashes; Go's path.Clean does not
        np_ = "/" + np_.lstrip("/")
    if p.endswith("/") and np_ != "/":
        np_ += "/"
    return np_


@settings(max_examples=400, deadline=None)
@given(st.text(alphabet="/.ab-", max_size=20))
def test_clean_path_against_posixpath(p):
    assert O.clean_path(p.encode()) == _mux_clean_path_from_python(p).encode()


@settings(max_examples=400, deadline=None)
@given(st.lists(st.tuples(st.text(alphabet="abk =+%2Fe", max_size=6), st.text(alphabet="abv +%41z&", max_size=8)), max_size=5),
       st.sampled_from(["a", "k", "ab", "", "k k"]))
def test_query_get_against_urllib(pairs, key):
    """url.ParseQuery vs urllib.parse.parse_qsl on queries without ';' and with well-formed escapes only (Go drops a pair
    with a malformed escape, urllib passes it through; Go rejects ';', older Pythons split on it)."""
    q = "&".join(urllib.parse.quote_plus(k, safe="") + "=" + urllib.parse.quote_plus(v, safe="") for k, v in pairs)
    want = ""
    for k, v in urllib.parse.parse_qsl(q, keep_blank_values=True):
        if k == key:
            want = v
            break
    assert O.query_get(q.encode(), key.encode()) == want.encode()


# ---- json.Unmarshal's syntax check (the scanner the Bind path restates) against Python's json module ----
from gofr_b200 import spec as S  # noqa: E402
from gofr_b200 import synth  # noqa: E402

_atoms = st.sampled_from([b"{", b"}", b"[", b"]", b":", b",", b'"a"', b'"id"', b'"\\u0041"', b'"\\x"', b'"\\ud83d\\ude00"', b'"\\ud83d"',
                          b"1", b"-1", b"01", b"1.5", b"1e5", b"1E+2", b"-", b"1.", b".5", b"true", b"false", b"null", b"nul", b"tru",
                          b" ", b"\n", b"\t", b'"', b"\\", b"\x01", b'"\x01"', b"0", b"-0", b"1e", b"{}", b"[]", b'"name":"x"', b"e", b"+1"])


def _python_accepts(body: bytes) -> bool:
    def no_constants(name):
        raise ValueError(name)  # NaN / Infinity are not JSON; Go rejects them
    try:
        json.loads(body.decode("utf-8"), parse_constant=no_constants)
        return True
    except (ValueError, RecursionError):
        return False


@settings(max_examples=600, deadline=None)
@given(st.lists(_atoms, min_size=0, max_size=10).map(b"".join))
def test_unmarshal_syntax_check_against_python_json(body):
    """Go's json.Unmarshal first checks the whole input for validity and reports a SyntaxError ("invalid character …",
    "unexpected end of JSON input") before storing anything; well-formed input can at most produce an
    UnmarshalTypeError.  Python's json.loads must agree on which inputs are well formed."""
    ot = O.OracleTable(synth.config3_spec())
    sid = synth.config3_spec().schemas[0].id
    ok, msg = ot.bind(sid, body)
    syntax_error = (not ok) and (msg.startswith(b"invalid character") or msg.startswith(b"unexpected end of JSON input"))
    assert syntax_error == (not _python_accepts(body)), (body, ok, msg)


# ---- the wire responses (FRAME_WIRE) through two independent HTTP/1.1 response parsers ----

def _h11_response(wire: bytes, head_request: bool):
    import h11
    c = h11.Connection(h11.CLIENT)
    c.send(h11.Request(method="HEAD" if head_request else "GET", target="/", headers=[("Host", "h")]))
    c.send(h11.EndOfMessage())
    c.receive_data(wire)
    ev = c.next_event()
    assert isinstance(ev, h11.Response), ev
    body = b""
    while True:
        e = c.next_event()
        if isinstance(e, h11.Data):
            body += bytes(e.data)
        elif isinstance(e, h11.EndOfMessage):
            break
        else:
            raise AssertionError(e)
    return ev.status_code, {k.decode(): v.decode() for k, v in ev.headers}, body


def _llhttp_response(wire: bytes):
    import httptools

    class P:
        def __init__(self):
            self.headers, self.body, self.done = {}, b"", False

        def on_header(self, k, v):
            self.headers[k.decode().lower()] = v.decode()

        def on_body(self, b):
            self.body += b

        def on_message_complete(self):
            self.done = True

    p = P()
    parser = httptools.HttpResponseParser(p)
    parser.feed_data(wire)
    return parser.get_status_code(), p.headers, p.body, p.done


def test_wire_responses_parse_as_http11():
    """every response of the mixed-traffic stream (200 / 301 / 404 / 405 / 500, HEAD and OPTIONS included) is one complete
    HTTP/1.1 message for h11 and for llhttp: status as reported in meta, Content-Length == body length, JSON bodies load"""
    from gofr_b200 import spec as S, synth
    bare = S.TableSpec(default_routes=False, routes=[S.Route(S.M_GET, "/only-get", S.H_STATIC_STRING, s0=b"x"),
                                                    S.Route(S.M_POST, "/only-post", S.H_NIL)])
    bare_batch = S.RequestBatch.pack([S.Req(S.M_POST, b"/only-get"), S.Req(S.M_HEAD, b"/only-post"), S.Req(S.M_GET, b"/nothing"),
                                      S.Req(S.M_GET, b"/only-get"), S.Req(S.M_POST, b"/only-post"), S.Req(S.M_GET, b"//only-get")])
    seen = set()
    for spec, batch in ((synth.config4_spec(), synth.config4_batch(1500)), (bare, bare_batch)):
        out, off, meta = O.OracleTable(spec).serve(batch, S.http_date(1_700_000_000))
        for i, wire in enumerate(O.responses(out, off)):
            status = int(meta[i]) & 0xFFFF
            if status == 0:
                continue                                  # host-only handler: no bytes
            head = int(batch.desc[i]["method"]) == S.M_HEAD
            st, hdr, body = _h11_response(wire, head)
            assert st == status, (i, wire[:40])
            if "content-length" in hdr and not head:
                assert int(hdr["content-length"]) == len(body)
            if hdr.get("content-type") == "application/json" and body:
                json.loads(body)
            assert hdr.get("date") == "Tue, 14 Nov 2023 22:13:20 GMT"
            if not head:
                st2, hdr2, body2, done = _llhttp_response(wire)
                assert st2 == status and body2 == body and done and hdr2 == hdr, (i, wire[:60])
            seen.add(status)
    assert {200, 301, 404, 405, 500} <= seen


def test_large_file_response_is_chunked_and_parses():
    """a File body of more than 2048 bytes is ONE Write past net/http's 2 KiB bufio.Writer: the reference answers with
    Transfer-Encoding: chunked (its own 13 149-byte favicon does).  Product (device code on the CPU) == oracle, and two
    independent HTTP parsers (h11, llhttp) read the chunked message back to exactly the blob."""
    from gofr_b200 import spec as S
    from gofr_b200.table import Table
    from tests.emu import emu
    date = S.http_date(1_700_000_000)
    for size in (1, 2048, 2049, 13149, 70000):
        blob = b"\x00\x00\x01\x00" + bytes((i * 131 + 7) & 0xFF for i in range(size - 4)) if size > 4 else b"x" * size
        spec = S.TableSpec(frame_mode=S.FRAME_WIRE, routes=[S.Route(S.M_GET, "/favicon.ico", S.H_FILE, s0=b"image/x-icon", blob=blob)])
        batch = S.RequestBatch.pack([S.Req(S.M_GET, b"/favicon.ico")])
        o1, f1, m1 = O.OracleTable(spec).serve(batch, date, out_cap=1 << 18)
        o2, f2, m2 = emu.serve(Table(spec).serialize(), batch, date, out_cap=1 << 18)
        wire = O.responses(o1, f1)[0]
        assert wire == O.responses(o2, f2)[0] and int(m1[0]) == int(m2[0])
        st, hdr, body = _h11_response(wire, False)
        st2, hdr2, body2, done = _llhttp_response(wire)
        assert st == st2 == 200 and body == body2 == blob and done
        if size > 2048:
            assert hdr.get("transfer-encoding") == "chunked" and "content-length" not in hdr
            assert wire.endswith(b"\r\n0\r\n\r\n") and (b"\r\n\r\n%x\r\n" % size) in wire
        else:
            assert int(hdr["content-length"]) == size and "transfer-encoding" not in hdr
        if size > 4:
            assert hdr["content-type"] == "image/x-icon"          # sniffed from the first 512 bytes, not the handler's value


def test_known_deviation_json_bodies_beyond_2k_are_chunked_by_the_reference():
    """json.Encoder.Encode hands the whole body to ONE Write; beyond 2048 bytes that Write bypasses response.w's buffer and the
    reference's response is chunked.  The product sends the same bytes with Content-Length (a stated deviation:
    the framing decision depends on the body length, which the kernels' header programs do not branch on).  This pins exactly
    what differs: the Content-Length line against a Transfer-Encoding line after Content-Type, and the chunk framing."""
    from gofr_b200 import spec as S
    date = S.http_date(1_700_000_000)
    spec = S.TableSpec(frame_mode=S.FRAME_WIRE, routes=[S.Route(S.M_GET, "/big", S.H_RESULT), S.Route(S.M_GET, "/small", S.H_RESULT)])
    big, small = b"y" * 2037, b"y" * 2036   # {"data":"…"}\n adds 12 bytes: 2049 (one past the buffer) and 2048 (fits exactly)
    batch = S.RequestBatch.pack([S.Req(S.M_GET, b"/big", data=S.result_record(S.RESULT_STRING, big)),
                                 S.Req(S.M_GET, b"/small", data=S.result_record(S.RESULT_STRING, small))])
    plain = O.responses(*O.OracleTable(spec).serve(batch, date)[:2])
    O.set_strict_chunking(True)
    try:
        strict = O.responses(*O.OracleTable(spec).serve(batch, date)[:2])
    finally:
        O.set_strict_chunking(False)
    assert plain[1] == strict[1]                                   # 2048 bytes: buffered, Content-Length either way
    hp, _, bp = plain[0].partition(b"\r\n\r\n")
    hs, _, bs = strict[0].partition(b"\r\n\r\n")
    assert len(bp) == 2049 and bs == b"%x\r\n" % len(bp) + bp + b"\r\n0\r\n\r\n"
    lp, ls = hp.split(b"\r\n"), hs.split(b"\r\n")
    assert [l for l in lp if l not in ls] == [b"Content-Length: 2049"] and [l for l in ls if l not in lp] == [b"Transfer-Encoding: chunked"]
    assert ls[-1] == b"Transfer-Encoding: chunked" and ls[-2].startswith(b"Content-Type: ")
    st, hdr, body = _h11_response(strict[0], False)
    assert st == 200 and body == bp and hdr.get("transfer-encoding") == "chunked"


def test_static_bodies_beyond_2k_are_chunked_like_the_reference():
    """handlers that return a constant: the body length is known when the table is sealed, so the product frames bodies of more
    than 2048 bytes the way net/http does (one chunk); 2048 bytes exactly still fit the buffer.  Product == oracle, h11 agrees."""
    from gofr_b200 import spec as S
    from gofr_b200.table import Table
    from tests.emu import emu
    date = S.http_date(1_700_000_000)
    for n, chunked in ((2036, False), (2037, True), (9000, True)):      # {"data":"…"}\n adds 12 bytes
        spec = S.TableSpec(frame_mode=S.FRAME_WIRE, routes=[S.Route(S.M_GET, "/s", S.H_STATIC_STRING, s0=b"q" * n)])
        b = S.RequestBatch.pack([S.Req(S.M_GET, b"/s")])
        o1, f1, m1 = O.OracleTable(spec).serve(b, date, out_cap=1 << 16)
        o2, f2, m2 = emu.serve(Table(spec).serialize(), b, date, out_cap=1 << 16)
        wire = O.responses(o1, f1)[0]
        assert wire == O.responses(o2, f2)[0]
        st, hdr, body = _h11_response(wire, False)
        assert st == 200 and body == b'{"data":"' + b"q" * n + b'"}\n'
        assert (hdr.get("transfer-encoding") == "chunked") == chunked and ("content-length" in hdr) == (not chunked)


def test_known_deviation_json_with_LP_at_offset_34_sniffs_as_a_font_in_the_reference():
    """http.DetectContentType's embedded-OpenType signature is 34 don't-care bytes followed by "LP": a JSON body that happens to
    carry "LP" at offset 34 leaves the reference with Content-Type: application/vnd.ms-fontobject.  The product's programs
    assume JSON text only ever sniffs as text/plain (a stated deviation); the oracle follows the product unless
    strict mode is on.  This pins the difference to that one header value."""
    from gofr_b200 import spec as S
    date = S.http_date(1_700_000_000)
    spec = S.TableSpec(frame_mode=S.FRAME_WIRE, routes=[S.Route(S.M_GET, "/s", S.H_RESULT)])
    text = b"x" * (34 - len(b'{"data":"')) + b"LP and so on"          # {"data":"xxxxxxxxxxxxxxxxxxxxxxxxxLP and so on"}
    batch = S.RequestBatch.pack([S.Req(S.M_GET, b"/s", data=S.result_record(S.RESULT_STRING, text)),
                                 S.Req(S.M_GET, b"/s", data=S.result_record(S.RESULT_STRING, b"y" + text))])
    plain = O.responses(*O.OracleTable(spec).serve(batch, date)[:2])
    O.set_strict_chunking(True)
    try:
        strict = O.responses(*O.OracleTable(spec).serve(batch, date)[:2])
    finally:
        O.set_strict_chunking(False)
    assert plain[0][34 + plain[0].index(b"\r\n\r\n") + 4:][:2] == b"LP"
    assert plain[1] == strict[1]                                   # one byte further along: nothing special
    assert plain[0].replace(b"Content-Type: text/plain; charset=utf-8", b"Content-Type: application/vnd.ms-fontobject") == strict[0]
    assert plain[0] != strict[0]

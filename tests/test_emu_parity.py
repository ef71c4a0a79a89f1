"""CPU parity: the serve kernel's per-request device code (run on the CPU by tests/emu) vs the oracle.

This is the logic half of the kernel — routing, sizing, the funnel-shift word writer, every response program — at
all four output alignments.  The tile machinery (TMA staging, block scan, look-back) only exists on the GPU and is
covered by tests/test_gpu_parity.py.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gofr_b200 import spec as S
from gofr_b200 import synth
from gofr_b200.table import Table
from tests import oracle as O
from tests.emu import emu

DATE = S.http_date(1789974595)


def _compare(spec, batch, misalign=0):
    ot = O.OracleTable(spec)
    img = Table(spec).serialize()
    o1, f1, m1 = ot.serve(batch, DATE)
    o2, f2, m2 = emu.serve(img, batch, DATE, misalign=misalign)
    r1, r2 = O.responses(o1, f1), O.responses(o2, f2)
    for i, (a, b) in enumerate(zip(r1, r2)):
        assert a == b, (i, batch.desc[i], a, b)
    assert (m1 == m2).all()
    return r1


@pytest.mark.parametrize("mis", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [S.FRAME_WIRE, S.FRAME_INTENDED, S.FRAME_BODY])
def test_config1(mis, mode):
    _compare(synth.config1_spec(mode), synth.config1_batch(300), mis)


@pytest.mark.parametrize("mode", [S.FRAME_WIRE, S.FRAME_INTENDED, S.FRAME_BODY])
def test_config2(mode):
    r = _compare(synth.config2_spec(mode), synth.config2_batch(2048), 3)
    want = {S.FRAME_WIRE: 521, S.FRAME_INTENDED: 512, S.FRAME_BODY: 256}[mode]
    assert {len(x) for x in r} == {want}


def test_config2_escape_heavy():
    _compare(synth.config2_spec(), synth.config2_batch(4096, escape_every=2), 1)


@pytest.mark.parametrize("mode", [S.FRAME_WIRE, S.FRAME_INTENDED, S.FRAME_BODY])
def test_config4_mixed(mode):
    _compare(synth.config4_spec(mode), synth.config4_batch(6000), 2)


def test_no_default_routes_mux_404_405():
    spec = S.TableSpec(default_routes=False, routes=[
        S.Route(S.M_GET, "/a", S.H_STATIC_STRING, s0=b"A"), S.Route(S.M_POST, "/b", S.H_NIL),
        S.Route(S.M_GET, "/u/{id:[0-9]+}", S.H_STATIC_STRING, s0=b"U"), S.Route(S.M_GET, "/f/{name}.json", S.H_NIL),
        S.Route(S.M_GET, "noslash", S.H_NIL), S.Route(S.M_GET, "/x/{a}-{b}/y", S.H_NIL),
        S.Route(S.M_GET, "/w/{rest:.*}", S.H_NIL), S.Route(S.M_GET, "/k/{v:[a-c]*}z", S.H_NIL)])
    reqs = []
    for m in (S.M_GET, S.M_POST, S.M_DELETE, S.M_HEAD, S.M_OPTIONS, S.M_OTHER):
        for p in (b"/a", b"/b", b"/zzz", b"/u/123", b"/u/12a", b"/f/x.json", b"/f/a.json.json", b"/f/.json", b"noslash",
                  b"/noslash", b"//a", b"/a/", b"/x/1-2-3/y", b"/x/-/y", b"/x/a-/y", b"/w/", b"/w/a/b/c", b"/k/abcz", b"/k/z",
                  b"/k/abdz", b"", b"/a/../a", b"/%zz", b"/u/\xc3\xa9"):
            reqs.append(S.Req(m, p, b"q=1" if len(reqs) % 3 == 0 else b""))
    for mode in (S.FRAME_WIRE, S.FRAME_BODY):
        spec.frame_mode = mode
        _compare(spec, S.RequestBatch.pack(reqs), 1)


def test_redirect_locations():
    spec = synth.config1_spec()
    paths = [b"//hello", b"/hello/../hello", b"/a b//c", b"/./", b"/a/b/../../..", b"hello", b"/x/./y/", b"/\xc3\xa9//", b"/a?b//c",
             b"/%41//", b"/a/./././b", b"/../../..", b"/a//", b"//", b"/..a/..", b"/a/.../..//"]
    reqs = [S.Req(S.M_GET, p, q, flags=f) for p in paths for q, f in ((b"", 0), (b"x=1&y=2", 0), (b"", S.REQ_FORCE_QUERY))]
    reqs += [S.Req(S.M_HEAD, b"//hello"), S.Req(S.M_OPTIONS, b"//hello")]
    for mis in range(4):
        _compare(spec, S.RequestBatch.pack(reqs), mis)


def test_param_values():
    spec = synth.config1_spec()
    qs = [b"", b"name=", b"name=gofr", b"name=a+b%20c", b"name=%zz&name=ok", b"name=a;b&name=c", b"&&name=x", b"name=%4",
          b"na%6De=v", b"x=1&name=%e2%82%ac&name=2", b"name=%ff%fe", b"name=%22%3C%3E%26%5C", b"name=%0a%09%0d%00%1f",
          b"name=%e2%80%a8%e2%80%a9", b"name=%c3", b"name=\xc3\xa9", b"name=\xff", b"name=a=b=c", b"Name=x", b"name",
          b"name=%F0%9F%98%80", b"name=%ed%a0%80", b"name=+", b"name=%2B", b"a=1&b=2&c=3&name=" + b"z" * 300]
    reqs = [S.Req(S.M_GET, b"/hello", q) for q in qs]
    for mis in range(4):
        _compare(spec, S.RequestBatch.pack(reqs), mis)


def test_bad_rows_answer_like_a_panic():
    spec = synth.config2_spec()
    good = synth.C2_SCHEMA.encode_row([1, "n", "e", True, 2])
    rows = [good, good[:10], good[:24], b"", good[:-1], good + b"extra"]
    _compare(spec, S.RequestBatch.pack([S.Req(S.M_GET, b"/api/v1/r03", b"", r) for r in rows]), 1)


def test_omitempty_and_ints():
    sc = S.Schema(5, "main.L", [S.Field("A", S.F_INT64, "a", True), S.Field("B", S.F_STRING, "b", True),
                                S.Field("C", S.F_BOOL, "c", True), S.Field("D", S.F_INT32, "d"), S.Field("E", S.F_INT, "<e>", True)])
    spec = S.TableSpec(schemas=[sc], routes=[S.Route(S.M_GET, "/l", S.H_ROW, schema_id=5)])
    vals = [0, 1, -1, 9, 10, 99, 100, 12345678, 123456789, 10 ** 15, 10 ** 16, 10 ** 16 - 1, -10 ** 16, 2 ** 63 - 1, -2 ** 63,
            99999999, 100000000, 9999999999999999, 10 ** 17 + 5]
    reqs = []
    for i, v in enumerate(vals):
        reqs.append(S.Req(S.M_GET, b"/l", b"", sc.encode_row([v, "x" * (i % 3), i % 2, (v % 2 ** 31) * (-1 if i % 2 else 1), -v if abs(v) < 2 ** 62 else 0])))
    reqs.append(S.Req(S.M_GET, b"/l", b"", sc.encode_row([0, "", False, 0, 0])))
    for mis in range(4):
        _compare(spec, S.RequestBatch.pack(reqs), mis)


_str = st.binary(min_size=0, max_size=40)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(_str, _str, st.integers(-2 ** 63, 2 ** 63 - 1), st.booleans()), min_size=1, max_size=12),
       st.integers(0, 3))
def test_random_strings_property(rows, mis):
    """Any byte strings through the struct encoder: escapes, invalid UTF-8, every source/destination alignment."""
    sc = synth.C2_SCHEMA
    spec = S.TableSpec(schemas=[sc], routes=[S.Route(S.M_GET, "/p", S.H_ROW, schema_id=1)])
    reqs = [S.Req(S.M_GET, b"/p", b"", sc.encode_row([i, a, b, f, i % 2 ** 31])) for a, b, i, f in rows]
    _compare(spec, S.RequestBatch.pack(reqs), mis)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.binary(min_size=0, max_size=30), min_size=1, max_size=10), st.integers(0, 3))
def test_random_queries_property(queries, mis):
    spec = synth.config1_spec()
    _compare(spec, S.RequestBatch.pack([S.Req(S.M_GET, b"/hello", b"name=" + q) for q in queries] +
                                       [S.Req(S.M_GET, b"/hello", q) for q in queries]), mis)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.text(alphabet="/.ab%? ", min_size=0, max_size=12), min_size=1, max_size=10))
def test_random_paths_property(paths):
    spec = S.TableSpec(routes=[S.Route(S.M_GET, "/a", S.H_NIL), S.Route(S.M_GET, "/a/{x}", S.H_NIL),
                               S.Route(S.M_GET, "/{y}/b/", S.H_NIL)])
    _compare(spec, S.RequestBatch.pack([S.Req(S.M_GET, p.encode()) for p in paths]), 1)


def test_path_params():
    """Request.PathParam (pkg/gofr/http/request.go:36-38): spans captured by mux's template regexp, leftmost-first."""
    spec = S.TableSpec(routes=[
        S.Route(S.M_GET, "/users/{id}", S.H_PATHPARAM_FORMAT, s0=b"id", s2=b"user ", s3=b"!"),
        S.Route(S.M_GET, "/f/{name}.json", S.H_PATHPARAM_FORMAT, s0=b"name", s2=b"", s3=b""),
        S.Route(S.M_GET, "/x/{a}-{b}/y", S.H_PATHPARAM_FORMAT, s0=b"b", s2=b"b=", s3=b""),
        S.Route(S.M_GET, "/x2/{a}-{b}/y", S.H_PATHPARAM_FORMAT, s0=b"a", s2=b"a=", s3=b""),
        S.Route(S.M_GET, "/w/{rest:.*}", S.H_PATHPARAM_FORMAT, s0=b"rest", s2=b"[", s3=b"]"),
        S.Route(S.M_GET, "/n/{id:[0-9]+}/p", S.H_PATHPARAM_FORMAT, s0=b"nosuch", s2=b"<", s3=b">"),
        S.Route(S.M_GET, "/d/{v}/{v}", S.H_PATHPARAM_FORMAT, s0=b"v", s2=b"", s3=b""),   # duplicated variable: dead route
    ])
    paths = [b"/users/42", b"/users/a\"b<c", b"/users/\xc3\xa9", b"/users/\xff", b"/users/", b"/f/x.json", b"/f/a.json.json",
             b"/x/1-2-3/y", b"/x2/1-2-3/y", b"/x/-/y", b"/w/", b"/w/a/b/c", b"/n/77/p", b"/d/1/1", b"/users/a b", b"/users/" + b"z" * 300]
    for mis in range(4):
        r = _compare(spec, S.RequestBatch.pack([S.Req(S.M_GET, p) for p in paths]), mis)
    assert r[0].endswith(b'{"data":"user 42!"}\n') and r[6].endswith(b'{"data":"a.json"}\n')
    assert r[7].endswith(b'{"data":"b=3"}\n') and r[8].endswith(b'{"data":"a=1-2"}\n')   # greedy first variable
    assert r[11].endswith(b'{"data":"[a/b/c]"}\n') and r[12].endswith(b'{"data":"\\u003c\\u003e"}\n')
    assert b" 404 " in r[13][:20]

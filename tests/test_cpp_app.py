"""The C++ stand-in of the reference's app API (include/gofr_b200.hpp): gofr.New / GET / PUT / POST / DELETE /
Context.Param / Context.PathParam / Run / ServeHTTP (pkg/gofr/gofr.go:49-73,152-177; pkg/gofr/http/request.go:28-38).

examples/cpp/server_routes.cpp is the reference's TestGofr_ServerRoutes (pkg/gofr/gofr_test.go:40-105) written against
it, with closures as handlers; this file builds it, runs it on the GPU and compares every response byte with the oracle
serving the same requests (the closures restated in Python below)."""
import os
import subprocess

import numpy as np
import pytest

from gofr_b200 import _build
from gofr_b200 import spec as S
from tests import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATE = S.http_date(1_700_000_000)


def _compile(src, exe, link=True):
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-o", exe]
    if link:
        cmd += [_build.LIB, "-Wl,-rpath," + os.path.dirname(_build.LIB)]
    subprocess.check_call(cmd)


def test_header_compiles_and_host_helpers_match_oracle(tmp_path):
    """no GPU: the example builds against the library; Param's query parsing agrees with the oracle's url.ParseQuery
    restatement on hostile queries; template variable names come out in mux's order"""
    from gofr_b200 import _abi
    _abi.lib()  # builds the .so if needed
    _compile(os.path.join(ROOT, "examples", "cpp", "server_routes.cpp"), str(tmp_path / "server_routes"))
    exe = str(tmp_path / "hpp_host_check")
    _compile(os.path.join(ROOT, "tests", "emu", "hpp_host_check.cpp"), exe)
    rng = np.random.default_rng(5)
    alphabet = [b"a", b"b", b"name", b"=", b"&", b"&", b";", b"+", b"%41", b"%zz", b"%", b"%4", b"x", b"=", b"%26", b"%3D", b"\xff", b" "]
    cases = [(b"name=Vikash", b"name"), (b"", b"name"), (b"name", b"name"), (b"a=1&name=&name=2", b"name"), (b"a=1;name=2&name=3", b"name"),
             (b"name=%zz&name=ok", b"name"), (b"n%61me=v", b"name"), (b"name=a+b%20c", b"name"), (b"=v", b""), (b"&&name=x&&", b"name"),
             (b"name=x=y", b"name"), (b"name=%", b"name"), (b"name=%4", b"name"), (b"name=a%", b"name")]
    for _ in range(3000):
        q = b"".join(alphabet[int(k)] for k in rng.integers(0, len(alphabet), int(rng.integers(0, 12))))
        cases.append((q, [b"name", b"a", b"b", b"x", b""][int(rng.integers(0, 5))]))
    lines = "".join("Q %s %s\n" % (q.hex() or "-", k.hex() or "-") for q, k in cases)
    tmpl = ["/a/{id}/b/{name:[a-z]+}", "/{x:[0-9]{2}}/{y}", "/plain", "/{a}-{b}.{c}", "/{v:.*}"]
    lines += "".join("T %s\n" % t.encode().hex() for t in tmpl)
    lines += "".join("R %d\n" % c for c in range(10))
    lines += "".join("V %d\n" % c for c in range(6))
    # request targets: path unescaping, the first '?' splits, a lone trailing '?' is ForceQuery, bad escapes are refused
    import re
    import urllib.parse
    tparts = [b"/", b"a", b"b c", b"%41", b"%2F", b"%2f", b"%zz", b"%4", b"%", b"?", b"?", b"x=1", b"&", b"%3F", b"+", b"\xc3\xa9", b"//", b".."]
    targets = [b"/", b"/a?", b"/a??", b"/a%3Fb?c", b"/%", b"/%4", b"/a%zzb", b"/x?y=%zz", b"a/b", b"", b"/p%41th/%2F?q=%41"]
    for _ in range(1500):
        targets.append(b"/" + b"".join(tparts[int(k)] for k in rng.integers(0, len(tparts), int(rng.integers(0, 8)))))
    lines += "".join("P %s\n" % (t.hex() or "-") for t in targets)
    out = subprocess.run([exe], input=lines, capture_output=True, text=True, check=True).stdout.splitlines()
    for (q, k), got in zip(cases, out):
        assert bytes.fromhex(got) == O.query_get(q, k), (q, k)
    assert out[len(cases):len(cases) + 5] == ["id,name", "x,y", "", "a,b,c", "v"]
    # the (data, err) → GOFR_H_RESULT record encoding against the Python packer the oracle-side tests use
    item = S.Schema(1, "main.Item", [S.Field("SKU", S.F_STRING, "sku"), S.Field("Qty", S.F_INT32, "qty"), S.Field("Big", S.F_INT64, "big"),
                                     S.Field("Ok", S.F_BOOL, "ok"), S.Field("N", S.F_INT, "n"), S.Field("Note", S.F_STRING, "note", omitempty=True)])
    bad = (0xFFFFFFFF).to_bytes(4, "little")
    want = [S.result_record(S.RESULT_STRING, b"Hello World!"), S.result_record(S.RESULT_ERROR, b"db: connection refused"),
            S.result_record(S.RESULT_NIL), S.result_record(S.RESULT_MISSING, b"http: no such file"),
            S.result_record(S.RESULT_DATA, item.encode_row(["A-1", -3, -5000000000, True, 1 << 40, "fragile"])),
            S.result_both(item, ["", 0, 0, False, 0, ""], b"partial"), bad, bad, bad, S.result_record(S.RESULT_ERROR, b"e")]
    got = [bytes.fromhex(l) for l in out[len(cases) + 5:len(cases) + 15]]
    assert got == want
    # rows of the wider data model (include/gofr_b200.h "Row format") against the Python packer
    addr = S.Schema(1, "main.Addr", [S.Field("City", S.F_STRING, "city"), S.Field("Zip", S.F_INT32, "zip", True),
                                     S.Field("Geo", S.F_FLOAT64, "geo", container=S.C_SLICE)])
    user = S.Schema(2, "main.User", [S.Field("Name", S.F_STRING, "name"), S.Field("Score", S.F_FLOAT64, "score"), S.Field("Home", S.F_STRUCT, "home", elem_schema=1),
                                     S.Field("Work", S.F_STRUCT, "work", True, S.C_PTR, 1), S.Field("Tags", S.F_STRING, "tags", False, S.C_SLICE),
                                     S.Field("Attrs", S.F_STRING, "attrs", True, S.C_MAP), S.Field("Hist", S.F_STRUCT, "hist", False, S.C_SLICE, 1),
                                     S.Field("N", S.F_INT64, "n", True, S.C_PTR), S.Field("Counts", S.F_INT64, "counts", False, S.C_MAP)])
    addrs = S.Schema(3, "[]main.Addr", [S.Field("", S.F_STRUCT, "", container=S.C_SLICE, elem_schema=1, flags=S.FIELD_BARE)])
    blob = S.Schema(4, "main.Blob", [S.Field("ID", S.F_UINT64, "id"), S.Field("Data", S.F_BYTES, "data"), S.Field("Sum", S.F_BYTES, "sum", True),
                                     S.Field("Ratio", S.F_FLOAT32, "ratio"), S.Field("Parts", S.F_BYTES, "parts", False, S.C_SLICE),
                                     S.Field("UM", S.F_UINT64, "um", False, S.C_MAP), S.Field("PF", S.F_FLOAT32, "pf", False, S.C_PTR),
                                     S.Field("At", S.F_TIME, "at"), S.Field("Seen", S.F_TIME, "seen", False, S.C_SLICE),
                                     S.Field("Kids", S.F_STRUCT, "kids", False, S.C_SLICE_PTR, 1), S.Field("PI", S.F_INT64, "pi", False, S.C_SLICE_PTR)])
    look = {1: addr, 2: user, 3: addrs, 4: blob}.__getitem__
    wantv = [S.result_record(S.RESULT_DATA, user.encode_row(["bo<b>", 1.5e-7, ["Paris", 0, [1.0, 2.5]], None, ["a", "b\n"], {"z": "1", "a": "2", "aa": "3"},
                                                             [["X", 7, None], ["Y", 0, []]], 5, None], look)),
             S.result_record(S.RESULT_DATA, user.encode_row(["", 0.0, ["", 0, None], ["W", 1, [-0.5]], None, None, None, None, {"k": -1}], look)),
             S.result_record(S.RESULT_DATA, addrs.encode_row([[["X", 7, None], ["Y", 2, [3.25]]]], look)),
             S.result_record(S.RESULT_DATA, addrs.encode_row([None], look)), bad,
             # uint64 / []byte (nil and not) / float32 members, by value, in a slice, a map and behind a pointer
             S.result_record(S.RESULT_DATA, blob.encode_row([2 ** 64 - 1, b"\x00\xff\x10", None, 0.1, [b"ab", None, b""], {"k": 2 ** 63, "j": 7}, 2.5,
                                                             (1709210096, 123456789, 19800), [(-62135596800, 0, 0), (0, 5, -3600)],
                                                             [None, ["K", 9, [0.5]], None], [4, None]], look))]
    assert [bytes.fromhex(l) for l in out[len(cases) + 15:len(cases) + 21]] == wantv
    for t, line in zip(targets, out[len(cases) + 21:]):
        path, sep, query = t.partition(b"?")
        bad = not t.startswith(b"/") or re.search(rb"%(?![0-9a-fA-F]{2})", path) is not None
        if bad:
            assert line == "ERR", t
        else:
            p_hex, q_hex, force = line.split(" ")
            assert bytes.fromhex(p_hex[1:]) == urllib.parse.unquote_to_bytes(path) and bytes.fromhex(q_hex[1:]) == query, t
            assert force == ("1" if sep and not query else "0"), t


def _records():
    """the closures of examples/cpp/server_routes.cpp, restated: (method, target, body) → result record"""
    hello = S.result_record(S.RESULT_STRING, b"Hello World!")
    person = S.Schema(1, "main.Person", [S.Field("ID", S.F_INT, "id"), S.Field("Name", S.F_STRING, "name"),
                                         S.Field("Admin", S.F_BOOL, "admin", omitempty=True)])
    cases = [("GET", "/hello", b"", hello), ("GET", "/hello2", b"", hello), ("PUT", "/hello", b"", hello), ("POST", "/hello", b"", hello),
             ("GET", "/params?name=Vikash", b"", S.result_record(S.RESULT_STRING, b"Hello Vikash!")),
             ("DELETE", "/delete", b"", S.result_record(S.RESULT_STRING, b"Success")),
             ("GET", "/greet", b"", hello),
             ("GET", "/greet?name=a%26b+c&name=second", b"", S.result_record(S.RESULT_STRING, b"Hello a&b c!")),
             ("GET", "/greet?name=%zz&x=1", b"", hello),
             ("GET", "/error", b"", S.result_record(S.RESULT_ERROR, b"some error occurred")),
             ("GET", "/users/42/posts/hello-world", b"", S.result_record(S.RESULT_STRING, b"user 42 post hello-world")),
             ("GET", "/users/4x2/posts/p", b"", b""),
             ("GET", "/person/root", b"", S.result_record(S.RESULT_DATA, person.encode_row([4, "root", True]))),
             ("GET", "/person/al%20ice", b"", S.result_record(S.RESULT_DATA, person.encode_row([6, "al ice", False]))),
             ("GET", "/person/nobody", b"", S.result_both(person, [0, "nobody", False], b"partial <result>")),
             ("GET", "/nil", b"", S.result_record(S.RESULT_NIL)),
             ("GET", "/file", b"", S.result_record(S.RESULT_MISSING, b"http: no such file")),
             ("GET", "/panic", b"", (0xFFFFFFFF).to_bytes(4, "little")),
             ("POST", "/echo", b'line1\n"quoted" <tag>', S.result_record(S.RESULT_STRING, b'line1\n"quoted" <tag>')),
             ("GET", "/hello/", b"", b""), ("GET", "/a/../hello", b"", b""), ("GET", "//hello", b"", b""), ("OPTIONS", "/hello", b"", b""), ("PATCH", "/hello", b"", b""),
             ("GET", "/.well-known/health", b"", b"")]
    # Context.Bind: the closure binds the body and returns the struct (or json.Unmarshal's error): the records the C++ side
    # builds from the GPU's rows must be the ones Go's json.Unmarshal (the oracle's restatement) leads to
    ot = O.OracleTable(S.TableSpec(schemas=[person], routes=[]))
    for body in (b'{"id":1,"name":"Bob"}', b'{"ID":7,"NAME":"caf\\u00e9 \\"q\\"","admin":true,"extra":[1,{"a":2}]}', b'{"id":"x"}', b"{bad", b""):
        ok, res = ot.bind(1, body)
        cases.append(("POST", "/people", body, S.result_record(S.RESULT_DATA if ok else S.RESULT_ERROR, res)))
    cases += [("GET", "/raw", b"", S.result_record(S.RESULT_RAW_STRING, b"just <text>")),
              ("GET", "/rawnil", b"", S.result_record(S.RESULT_RAW_NIL)),
              ("GET", "/rawperson", b"", S.result_record(S.RESULT_RAW_DATA, person.encode_row([9, "raw", True]))),
              ("GET", "/rawerr", b"", S.result_record(S.RESULT_RAW_STRING, b"x", S.RAW_ERR))]
    addr = S.Schema(2, "main.Addr", [S.Field("City", S.F_STRING, "city"), S.Field("Zip", S.F_INT32, "zip", True),
                                     S.Field("Geo", S.F_FLOAT64, "geo", container=S.C_SLICE)])
    user = S.Schema(3, "main.User", [S.Field("Name", S.F_STRING, "name"), S.Field("Score", S.F_FLOAT64, "score"), S.Field("Home", S.F_STRUCT, "home", elem_schema=2),
                                     S.Field("Work", S.F_STRUCT, "work", True, S.C_PTR, 2), S.Field("Tags", S.F_STRING, "tags", False, S.C_SLICE),
                                     S.Field("Attrs", S.F_STRING, "attrs", True, S.C_MAP)])
    addrs = S.Schema(4, "[]main.Addr", [S.Field("", S.F_STRUCT, "", container=S.C_SLICE, elem_schema=2, flags=S.FIELD_BARE)])
    look = {2: addr, 3: user, 4: addrs}.__getitem__
    cases += [("GET", "/user/bob", b"", S.result_record(S.RESULT_DATA, user.encode_row(["bob", 1.5e-7, ["Paris", 75001, [48.8566, 2.3522]], ["Lyon", 0, []],
                                                                                       ["a", "b<c>"], {"z": "1", "a": "2"}], look))),
              ("GET", "/user/nan", b"", S.result_record(S.RESULT_DATA, user.encode_row(["nan", float("nan"), ["", 0, None], None, None, None], look))),
              ("GET", "/addrs", b"", S.result_record(S.RESULT_DATA, addrs.encode_row([[["X", 7, None], ["Y", 0, [1e21, -0.0]]]], look))),
              ("GET", "/addrs?none=1", b"", S.result_record(S.RESULT_DATA, addrs.encode_row([None], look)))]
    routes = [("GET", "/hello", 0), ("GET", "/hello2", 0), ("PUT", "/hello", 0), ("POST", "/hello", 0), ("GET", "/params", 0),
              ("DELETE", "/delete", 0), ("GET", "/greet", 0), ("GET", "/error", 0), ("GET", "/users/{id:[0-9]+}/posts/{slug}", 0),
              ("GET", "/person/{name}", 1), ("GET", "/nil", 0), ("GET", "/file", 0), ("GET", "/panic", 0), ("POST", "/echo", 0),
              ("POST", "/people", 1), ("GET", "/raw", 0), ("GET", "/rawnil", 0), ("GET", "/rawperson", 1), ("GET", "/rawerr", 0),
              ("GET", "/user/{name}", 3), ("GET", "/addrs", 4)]
    spec = S.TableSpec(schemas=[person, addr, user, addrs], favicon=b"",
                       routes=[S.Route(S.method_code(m), p, S.H_RESULT, schema_id=sid) for m, p, sid in routes])
    return spec, cases


@pytest.mark.gpu
def test_reference_route_test_through_the_cpp_app(tmp_path):
    import urllib.parse
    from gofr_b200 import _abi
    _abi.lib()
    exe = str(tmp_path / "server_routes")
    _compile(os.path.join(ROOT, "examples", "cpp", "server_routes.cpp"), exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr          # the C++ side checks status and body against the reference's expectations
    got = [(int(l.split(" ")[0]), bytes.fromhex(l.split(" ")[1]) if len(l.split(" ")) > 1 else b"") for l in r.stdout.splitlines()]
    spec, cases = _records()
    reqs = []
    for m, target, _body, rec in cases:
        path, _, query = target.partition("?")
        reqs.append(S.Req(S.method_code(m), urllib.parse.unquote_to_bytes(path), query.encode(), data=rec))
    b = S.RequestBatch.pack(reqs)
    for i in range(b.n):
        b.trace_ids[i] = ((np.arange(16) + i * 16) % 256).astype(np.uint8)
    out, off, meta = O.OracleTable(spec).serve(b, DATE)
    want = O.responses(out, off)
    assert len(got) == len(want)
    for i, ((st, by), w) in enumerate(zip(got, want)):
        assert st == int(meta[i]) & 0xFFFF and by == w, (i, cases[i][:2])


@pytest.mark.gpu
def test_one_process_several_engines(tmp_path):
    """The non-Python multi-GPU story: one sealed table, one engine per device, one host thread per
    engine serving its contiguous shard; the concatenated shards equal the single-engine result.  Uses every visible GPU
    (two engines share the device when there is only one)."""
    import torch
    from gofr_b200 import _abi
    _abi.lib()
    exe = str(tmp_path / "two_engines")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "cpp", "two_engines.cpp"), _build.LIB, "-Wl,-rpath," + os.path.dirname(_build.LIB),
                    "-lpthread", "-o", exe], check=True)
    ndev = max(1, torch.cuda.device_count())
    r = subprocess.run([exe, str(max(2, ndev)), str(ndev)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 responses differ" in r.stdout, r.stdout

"""GOFR_H_RESULT — stage 2 of the split API: the closure ran on the host, Responder.Respond runs on its (data, err)
(pkg/gofr/http/responder.go:19-62; handler.ServeHTTP pkg/gofr/handler.go:32-36).  Stage 1 is gofr_route_device
(tests/test_route.py)."""
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

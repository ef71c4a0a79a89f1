"""gofr_proto_encode_device — proto3 message encoder + gRPC framing.

Reference behaviour: proto.Marshal of the message a unary handler returns, then grpc-go's 5-byte length prefix
(examples/grpc-server/grpc/hello_grpc.pb.go:73-89; protobuf-go v1.32.0, grpc-go v1.60.1).  Three layers:
  oracle (oracle/orc_proto.c)  vs  python google.protobuf with a descriptor built at run time (independent implementation)
  device code on the CPU (tests/emu) vs oracle, hypothesis-generated message types and values
  CUDA kernel vs oracle (-m gpu)"""
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gofr_b200 import spec as S
from tests import oracle as O
from tests.emu import emu

ALL_TYPES = [S.PB_DOUBLE, S.PB_FLOAT, S.PB_INT64, S.PB_UINT64, S.PB_INT32, S.PB_FIXED64, S.PB_FIXED32, S.PB_BOOL, S.PB_STRING,
             S.PB_BYTES, S.PB_UINT32, S.PB_SFIXED32, S.PB_SFIXED64, S.PB_SINT32, S.PB_SINT64]   # PB_ENUM encodes like INT32


def _py_class(fields):
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto(name="t.proto", package="t", syntax="proto3")
    m = fdp.message_type.add(name="M")
    for f in fields:
        m.field.add(name="f%d" % f.number, number=f.number, type=f.type, label=1)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("t.M"))


def _py_serialize(fields, values):
    cls = _py_class(fields)
    out = []
    for msg in values:
        kw = {}
        for f, v in zip(fields, msg):
            if f.type == S.PB_STRING:
                v = v.decode("utf-8") if isinstance(v, (bytes, bytearray)) else v
            kw["f%d" % f.number] = v
        body = cls(**kw).SerializeToString()
        out.append(b"\x00" + len(body).to_bytes(4, "big") + body)
    return out


def _frames(out, off):
    return [out[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]


def _hello_fields():
    return [S.ProtoField(1, S.PB_STRING)]      # HelloResponse{message = 1} — examples/grpc-server/grpc/hello.proto:8-10


def test_hello_response_matches_the_hello_path():
    """the fused Hello kernel's response is this encoder applied to HelloResponse{message: "Hello <name>!"}"""
    f = _hello_fields()
    msgs = [["Hello World!"], ["Hello gofr!"], ["Hello " + "n" * 200 + "!"], [""]]
    rows, off = S.pack_proto_rows(f, msgs)
    out, o, meta = O.proto_encode(f, rows, off)
    fr = _frames(out, o)
    assert fr[0] == b"\x00\x00\x00\x00\x0e\x0a\x0cHello World!"
    assert fr[3] == b"\x00\x00\x00\x00\x00"                       # zero value: empty message, still a frame
    assert fr == _py_serialize(f, msgs) and not meta.any()
    # the Hello oracle on the matching requests produces the same frames
    req = b"".join(b"\x00" + (2 + len(n)).to_bytes(4, "big") + b"\x0a" + bytes([len(n)]) + n for n in (b"", b"gofr"))
    roff = np.array([0, 7, 7 + 7 + 4], dtype=np.uint32)
    h_out, h_off, _ = O.grpc_hello(np.frombuffer(req, dtype=np.uint8), roff)
    assert _frames(h_out, h_off) == fr[:2]


def test_oracle_matches_python_protobuf_on_edge_values():
    f = [S.ProtoField(1, S.PB_INT32), S.ProtoField(2, S.PB_INT64), S.ProtoField(3, S.PB_UINT32), S.ProtoField(4, S.PB_UINT64),
         S.ProtoField(5, S.PB_SINT32), S.ProtoField(6, S.PB_SINT64), S.ProtoField(7, S.PB_BOOL), S.ProtoField(8, S.PB_FIXED32),
         S.ProtoField(9, S.PB_FIXED64), S.ProtoField(10, S.PB_SFIXED32), S.ProtoField(11, S.PB_SFIXED64), S.ProtoField(12, S.PB_FLOAT),
         S.ProtoField(13, S.PB_DOUBLE), S.ProtoField(14, S.PB_STRING), S.ProtoField(15, S.PB_BYTES), S.ProtoField(16, S.PB_STRING),
         S.ProtoField(2047, S.PB_INT32), S.ProtoField(2048, S.PB_BOOL), S.ProtoField(536870911, S.PB_UINT32)]
    msgs = [
        [0] * 6 + [False, 0, 0, 0, 0, 0.0, 0.0, "", b"", "", 0, False, 0],
        [-1, -1, 2**32 - 1, 2**64 - 1, -1, -1, True, 2**32 - 1, 2**64 - 1, -1, -1, -0.0, -0.0, "é", b"\xff\x00", "x", -2**31, True, 1],
        [2**31 - 1, 2**63 - 1, 127, 128, 2**31 - 1, 2**63 - 1, True, 1, 1, 1, 1, 1.5, 1e300, "a" * 127, b"b" * 128, "c" * 300, 1, False, 300],
        [-2**31, -2**63, 16383, 16384, -2**31, -2**63, False, 0, 0, -2**31, -2**63, float("inf"), float("nan"), " ", b"", "", 0, False, 0],
        [1, 1, 1, 1, 1, 1, True, 0, 0, 0, 0, 1e-45, 5e-324, "", b"\x00", "", 0, False, 0],
    ]
    rows, off = S.pack_proto_rows(f, msgs)
    out, o, meta = O.proto_encode(f, rows, off)
    assert not meta.any()
    got, want = _frames(out, o), _py_serialize(f, msgs)
    for i in range(len(msgs)):
        assert got[i] == want[i], i


def test_invalid_utf8_and_malformed_rows():
    f = [S.ProtoField(1, S.PB_STRING), S.ProtoField(2, S.PB_BYTES), S.ProtoField(3, S.PB_INT64)]
    rows, off = S.pack_proto_rows(f, [["ok", b"\xff", 1], [b"\xff", b"", 1], [b"\xed\xa0\x80", b"", 0], [b"\xf4\x90\x80\x80", b"x", 0],
                                      [b"\xc3", b"", 0], ["fine é€\U0001F600", b"", 2]])
    bad = bytearray(rows.tobytes())
    o2 = off.copy()
    # two more rows, hand made: a string longer than the row, a row shorter than its fixed part
    extra1 = (100).to_bytes(4, "little") + (0).to_bytes(4, "little") + (0).to_bytes(8, "little") + b"abcd"
    extra2 = b"\x01\x00\x00\x00"
    body = bytes(bad[:int(off[-1])]) + extra1 + extra2 + b"\0" * 8
    o2 = np.concatenate([off, np.array([int(off[-1]) + len(extra1), int(off[-1]) + len(extra1) + len(extra2)], dtype=np.uint32)])
    rows2 = np.frombuffer(body, dtype=np.uint8).copy()
    out, o, meta = O.proto_encode(f, rows2, o2)
    assert list(meta) == [S.GRPC_OK, S.GRPC_BAD_UTF8, S.GRPC_BAD_UTF8, S.GRPC_BAD_UTF8, S.GRPC_BAD_UTF8, S.GRPC_OK, S.GRPC_BAD_ROW,
                          S.GRPC_BAD_ROW]
    fr = _frames(out, o)
    assert all(fr[i] == b"" for i in (1, 2, 3, 4, 6, 7))
    for mis in (0, 5):
        e_out, e_off, e_meta = emu.proto_encode(f, rows2, o2, mis)
        assert np.array_equal(e_meta, meta) and np.array_equal(e_off, o + mis)
        assert e_out[mis:int(e_off[-1])].tobytes() == out[:int(o[-1])].tobytes()


_value = {
    S.PB_DOUBLE: st.floats(allow_nan=False, width=64) | st.sampled_from([0.0, -0.0, float("inf")]),
    S.PB_FLOAT: st.floats(allow_nan=False, width=32) | st.sampled_from([0.0, -0.0]),
    S.PB_INT64: st.integers(-2**63, 2**63 - 1) | st.sampled_from([0, 1, -1]), S.PB_UINT64: st.integers(0, 2**64 - 1) | st.just(0),
    S.PB_INT32: st.integers(-2**31, 2**31 - 1) | st.sampled_from([0, -1]), S.PB_FIXED64: st.integers(0, 2**64 - 1) | st.just(0),
    S.PB_FIXED32: st.integers(0, 2**32 - 1) | st.just(0), S.PB_BOOL: st.booleans(),
    S.PB_STRING: st.text(max_size=40) | st.text(alphabet="ab", min_size=100, max_size=300),
    S.PB_BYTES: st.binary(max_size=40), S.PB_UINT32: st.integers(0, 2**32 - 1) | st.just(0),
    S.PB_SFIXED32: st.integers(-2**31, 2**31 - 1), S.PB_SFIXED64: st.integers(-2**63, 2**63 - 1),
    S.PB_SINT32: st.integers(-2**31, 2**31 - 1) | st.just(0), S.PB_SINT64: st.integers(-2**63, 2**63 - 1) | st.just(0),
}


@st.composite
def _typed_messages(draw):
    nf = draw(st.integers(1, 10))
    numbers = sorted(draw(st.sets(st.integers(1, 40) | st.integers(2040, 2060) | st.integers(2**29 - 3, 2**29 - 1), min_size=nf, max_size=nf)))
    fields = [S.ProtoField(n, draw(st.sampled_from(ALL_TYPES))) for n in numbers]
    n_msgs = draw(st.integers(1, 6))
    msgs = [[draw(_value[f.type]) for f in fields] for _ in range(n_msgs)]
    return fields, msgs


@settings(max_examples=150, deadline=None)
@given(_typed_messages(), st.integers(0, 15))
def test_random_message_types_three_way(tm, mis):
    """python protobuf == oracle == device code (CPU emulation), for random flat proto3 message types and values"""
    fields, msgs = tm
    msgs = [[v.encode("utf-8", "surrogatepass").decode("utf-8", "replace") if isinstance(v, str) else v for v in m] for m in msgs]
    rows, off = S.pack_proto_rows(fields, msgs)
    out, o, meta = O.proto_encode(fields, rows, off)
    assert not meta.any()
    assert _frames(out, o) == _py_serialize(fields, msgs)
    e_out, e_off, e_meta = emu.proto_encode(fields, rows, off, mis)
    assert np.array_equal(e_meta, meta) and np.array_equal(e_off, o + mis)
    assert e_out[mis:int(e_off[-1])].tobytes() == out[:int(o[-1])].tobytes()


def _bulk(n, seed=7):
    rng = np.random.default_rng(seed)
    fields = [S.ProtoField(1, S.PB_STRING), S.ProtoField(2, S.PB_INT64), S.ProtoField(3, S.PB_SINT32), S.ProtoField(4, S.PB_BOOL),
              S.ProtoField(5, S.PB_DOUBLE), S.ProtoField(7, S.PB_BYTES), S.ProtoField(9, S.PB_FIXED32), S.ProtoField(300, S.PB_INT32),
              S.ProtoField(301, S.PB_STRING)]
    names = [b"", b"a", b"caf\xc3\xa9", b"x" * 40, b"\xe2\x82\xac" * 5, b"y" * 200]
    msgs = []
    for k in range(n):
        r = rng.integers(0, 1 << 62, 6)
        msgs.append([names[int(r[0]) % 6], int(r[1]) - (1 << 61) if k % 3 else 0, int(r[2] % 2001) - 1000, bool(r[3] & 1),
                     float(int(r[4]) % 1000) / 8 if k % 4 else 0.0, bytes(int(b) & 0xFF for b in r[:int(r[5]) % 5]),
                     int(r[5]) & 0xFFFFFFFF if k % 5 else 0, -int(r[0] % 50000) if k % 2 else 7,
                     b"\xff" if k % 97 == 0 else names[int(r[1]) % 6]])
    return fields, msgs


def test_emu_bulk_matches_oracle():
    fields, msgs = _bulk(3000)
    rows, off = S.pack_proto_rows(fields, msgs)
    out, o, meta = O.proto_encode(fields, rows, off)
    assert (meta == S.GRPC_BAD_UTF8).sum() == len(range(0, 3000, 97))
    e_out, e_off, e_meta = emu.proto_encode(fields, rows, off, 3)
    assert np.array_equal(e_meta, meta) and np.array_equal(e_off, o + 3)
    assert e_out[3:int(e_off[-1])].tobytes() == out[:int(o[-1])].tobytes()


@pytest.mark.gpu
def test_gpu_matches_oracle():
    from gofr_b200 import synth
    from gofr_b200.engine import Engine
    from gofr_b200.table import Table
    eng = Engine(Table(synth.config1_spec()), 0)
    for fields, msgs in (_bulk(50000), ([S.ProtoField(1, S.PB_STRING)], [["Hello %d!" % k] for k in range(10000)]),
                         ([S.ProtoField(5, S.PB_UINT64)], [[k * 977] for k in range(1000)]), (_bulk(1)[0], [])):
        rows, off = S.pack_proto_rows(fields, msgs)
        out, o, meta = O.proto_encode(fields, rows, off)
        d_out, d_off, d_meta = eng.proto_encode_device(fields, rows, off)
        g_off = d_off.cpu().numpy().view(np.uint32)
        assert np.array_equal(g_off, o)
        if len(msgs):
            assert np.array_equal(d_meta.cpu().numpy().view(np.uint32), meta)
        assert d_out[:int(o[-1])].cpu().numpy().tobytes() == out[:int(o[-1])].tobytes()
    # argument checks: descending field numbers, unknown type, too many fields
    from gofr_b200 import _abi
    rows, off = S.pack_proto_rows([S.ProtoField(1, S.PB_BOOL)], [[True]])
    for bad in ([S.ProtoField(2, S.PB_BOOL), S.ProtoField(1, S.PB_BOOL)], [S.ProtoField(1, 11)], [S.ProtoField(0, S.PB_BOOL)],
                [S.ProtoField(k + 1, S.PB_BOOL) for k in range(33)]):
        with pytest.raises(_abi.GofrError):
            eng.proto_encode_device(bad, rows, off)
    eng.close()

"""Small instances of every C-ABI entry point, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gofr_b200 import spec as S, synth, _abi
from gofr_b200.engine import Engine, pin_batch, pinned_array
from gofr_b200.table import Table

date = S.http_date(1_700_000_000)
for name, spec, batch in (("config2", synth.config2_spec(), synth.config2_batch(700, escape_every=9)),
                          ("config3", synth.config3_spec(), synth.config3_batch(500)),
                          ("config4", synth.config4_spec(), synth.config4_batch(900))):
    eng = Engine(Table(spec), 0)
    db = eng.upload(batch)
    resp = eng.alloc_responses(batch.n, batch.n * 1200 + 4096)
    eng.serve_device(db, date, resp)
    eng.serve_device_slots(db, date, 1024)
    eng.route_device(db)
    eng.set_chunk(256)
    hb = pin_batch(batch)
    out = pinned_array(batch.n * 1200 + 4096)
    off = pinned_array(4 * (batch.n + 1), np.uint32)
    meta = pinned_array(4 * batch.n, np.uint32)
    got = eng.serve_host(hb, date, out, off, meta)
    out2 = np.zeros(batch.n * 1200 + 4096, dtype=np.uint8)   # pageable buffers: the cudaMemcpy pipeline
    off2 = np.zeros(batch.n + 1, dtype=np.uint32)
    meta2 = np.zeros(batch.n, dtype=np.uint32)
    got2 = eng.serve_host(batch, date, out2, off2, meta2)
    torch.cuda.synchronize()
    assert got == got2 and np.array_equal(out[:got], out2[:got])
    print(name, "ok", got)
    eng.close()
eng = Engine(Table(synth.config1_spec()), 0)
eng.request_log_device(synth.reqlog_batch(600, hostile_every=3, rpc_every=4))
frames, off = synth.config5_frames(1000)
d_in = torch.from_numpy(np.concatenate([frames, np.zeros(64, np.uint8)])).cuda()
d_off = torch.from_numpy(off.view(np.int32)).cuda()
cap = int(frames.size) + 40 * 1000
d_out = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
d_ooff = torch.zeros(1001, dtype=torch.int32, device="cuda")
d_meta = torch.zeros(1000, dtype=torch.int32, device="cuda")
_abi.check(_abi.lib().gofr_grpc_hello_device(eng._e, d_in.data_ptr(), d_off.data_ptr(), 1000, d_out.data_ptr(), cap,
                                             d_ooff.data_ptr(), d_meta.data_ptr(), torch.cuda.current_stream().cuda_stream), "grpc")
torch.cuda.synchronize()
raw, roff = synth.http_messages(900, seed_msgs=[b"GET /a HTTP/1.0\r\nHost: h\r\n\r\n", b"", b"POST /e HTTP/1.1\r\nHost: h\r\nContent-Length: 3\r\n\r\nabc"])
eng.http_parse_device(raw, roff)
torch.cuda.synchronize()
print("reqlog + grpc + http ok")
# additions of the end of round 1: proto3 encoder, host-buffer routing, slot host path, string outcomes, front-end
fields = [S.ProtoField(1, S.PB_STRING), S.ProtoField(2, S.PB_INT64), S.ProtoField(3, S.PB_SINT32), S.ProtoField(5, S.PB_DOUBLE),
          S.ProtoField(7, S.PB_BYTES), S.ProtoField(300, S.PB_INT32)]
msgs = [["n" * (k % 70), k * 977 - 5000, (k % 41) - 20, float(k % 7), bytes([k & 0xFF]) * (k % 9), -k if k % 3 else 0] for k in range(700)]
msgs[13][0] = b"\xff"
rows, roff2 = S.pack_proto_rows(fields, msgs)
d_o, d_oo, _ = eng.proto_encode_device(fields, rows, roff2)
oo = d_oo.cpu().numpy().view(np.uint32)
eng.proto_decode_device(fields, d_o[:int(oo[-1])].cpu().numpy(), oo)
torch.cuda.synchronize()
spec = S.TableSpec(routes=[S.Route(S.M_GET, "/hello", S.H_RESULT), S.Route(S.M_GET, "/u/{id}", S.H_RESULT)])
reqs = [S.Req(S.M_GET, b"/hello" if k % 2 else b"/u/%d" % k, data=S.result_record(S.RESULT_STRING if k % 3 else S.RESULT_ERROR, b"<v %d>" % k * (k % 30)))
        for k in range(500)]
b2 = S.RequestBatch.pack(reqs)
e2 = Engine(Table(spec), 0)
e2.set_chunk(128)
e2.route_host(b2)
so = np.zeros((b2.n, 1024), dtype=np.uint8); sl = np.zeros(b2.n, dtype=np.uint32); sm = np.zeros(b2.n, dtype=np.uint32)
e2.serve_host_slots(b2, date, 1024, so.reshape(-1), sl, sm)
from gofr_b200.frontend import Frontend
fe = Frontend(e2, max_batch=16, max_wait_us=100, slot_bytes=1024, max_request_bytes=2048)
for k in range(40):
    r = reqs[k]
    fe.serve(r.method, r.path, r.query, r.data)
fe.close()
e2.close()
torch.cuda.synchronize()
print("proto + route_host + slots host + front-end ok")

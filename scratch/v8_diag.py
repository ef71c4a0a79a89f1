"""Diagnostic for the 256-bit store variant: mismatch runs of the packed config-4 output against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gofr_b200 import spec as S, synth
from gofr_b200.engine import Engine
from gofr_b200.table import Table
from tests import oracle as O
DATE = S.http_date(1789974595)
spec, batch = synth.config4_spec(), synth.config4_batch(5000)
eng = Engine(Table(spec), 0)
o1, f1, m1 = O.OracleTable(spec).serve(batch, DATE)
total = int(f1[-1])
resp = eng.alloc_responses(batch.n, total + 1024)
base = resp.out.data_ptr()
eng.serve_device(eng.upload(batch), DATE, resp)
torch.cuda.synchronize()
out, off, meta = resp.to_host()
bad = np.nonzero(out[:total] != o1[:total])[0]
print("out base mod 128:", base % 128, "mismatching bytes:", bad.size)
runs = np.split(bad, np.nonzero(np.diff(bad) != 1)[0] + 1) if bad.size else []
for r in runs[:40]:
    a, b = int(r[0]), int(r[-1]) + 1
    i = int(np.searchsorted(f1, a, side="right") - 1)
    print(f"run [{a},{b}) len {b-a} sector_off {(base+a)%32} resp {i} start {int(f1[i])} (mod32 {(base+int(f1[i]))%32}) len {int(f1[i+1]-f1[i])} off_in_resp {a-int(f1[i])} got {bytes(out[a:a+8])!r} meta {int(meta[i])&0xffff} tile_lane {i%128}")

"""Per-GPU engine: thin wrapper over gofr_engine_* (include/gofr_b200.h).

torch is plumbing only: it owns device buffers and the CUDA stream whose raw handle is passed through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _abi
from . import spec as S
from .table import Table


@dataclass
class DeviceBatch:
    """A request batch resident in HBM (torch uint8 tensors)."""
    desc: "torch.Tensor"
    trace_ids: "torch.Tensor"
    arena: "torch.Tensor"
    n: int
    input_bytes: int


@dataclass
class DeviceResponses:
    out: "torch.Tensor"       # uint8[out_cap]
    out_off: "torch.Tensor"   # int32 view of uint32[n+1]
    meta: "torch.Tensor"      # int32 view of uint32[n]
    n: int

    def to_host(self):
        import torch
        off = self.out_off.cpu().numpy().view(np.uint32)
        meta = self.meta.cpu().numpy().view(np.uint32)
        total = int(off[self.n])
        out = self.out[:total].cpu().numpy()
        return out, off, meta


class Engine:
    def __init__(self, table: Table, device: int = 0):
        L = _abi.lib()
        self._e = C.c_void_p()
        self.table = table
        self.device = device
        _abi.check(L.gofr_engine_create(C.byref(self._e), table.handle, device), "gofr_engine_create")

    def close(self):
        if self._e:
            _abi.lib().gofr_engine_destroy(self._e)
            self._e = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ----
    def set_tile(self, in_bytes_per_req: int):
        _abi.check(_abi.lib().gofr_engine_set_tile(self._e, in_bytes_per_req), "gofr_engine_set_tile")

    def set_chunk(self, n: int):
        _abi.check(_abi.lib().gofr_engine_set_chunk(self._e, n), "gofr_engine_set_chunk")

    def set_timing(self, on: bool):
        _abi.check(_abi.lib().gofr_engine_set_timing(self._e, 1 if on else 0), "gofr_engine_set_timing")

    def slot_ctas(self, force: int = 0) -> int:
        """CTAs/SM of the slot-layout serve kernel in effect (4: wide instance, 5); force=4/5 overrides the engine's choice."""
        v = C.c_int(0)
        _abi.check(_abi.lib().gofr_engine_slot_ctas(self._e, int(force), C.byref(v)), "gofr_engine_slot_ctas")
        return v.value

    def geometry(self):
        g, b, s, m = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _abi.check(_abi.lib().gofr_engine_geometry(self._e, C.byref(g), C.byref(b), C.byref(s), C.byref(m)),
                   "gofr_engine_geometry")
        return {"grid": g.value, "blocks_per_sm": b.value, "smem_bytes": s.value, "sm_count": m.value}

    def launch_count(self) -> int:
        return int(_abi.lib().gofr_engine_launch_count(self._e))

    def kernel_time_ms(self, reset: bool = False):
        ms, n = C.c_double(), C.c_uint64()
        _abi.check(_abi.lib().gofr_engine_kernel_time_ms(self._e, C.byref(ms), C.byref(n), 1 if reset else 0),
                   "gofr_engine_kernel_time_ms")
        return ms.value, n.value

    def overflowed(self, reset: bool = True) -> bool:
        f = C.c_int()
        _abi.check(_abi.lib().gofr_engine_overflowed(self._e, C.byref(f), 1 if reset else 0), "gofr_engine_overflowed")
        return bool(f.value)

    # ---- device-resident path ----
    def upload(self, batch: S.RequestBatch) -> DeviceBatch:
        import torch
        dev = torch.device("cuda", self.device)
        desc = torch.from_numpy(batch.desc.view(np.uint8).reshape(-1).copy()).to(dev)
        ids = torch.from_numpy(batch.trace_ids.reshape(-1).copy()).to(dev)
        pad = (-batch.arena.size) % 16 + 16
        arena = torch.from_numpy(np.concatenate([batch.arena, np.zeros(pad, dtype=np.uint8)])).to(dev)
        return DeviceBatch(desc, ids, arena, batch.n, batch.input_bytes())

    def alloc_responses(self, n: int, out_cap: int) -> DeviceResponses:
        import torch
        dev = torch.device("cuda", self.device)
        out = torch.empty(out_cap + 64, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        meta = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        return DeviceResponses(out, off, meta, n)

    def serve_device(self, b: DeviceBatch, date: bytes, resp: DeviceResponses, stream=None) -> None:
        """One fused launch on torch's current stream (or `stream`).  Asynchronous."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        assert len(date) == 29
        _abi.check(_abi.lib().gofr_serve_device(self._e, b.desc.data_ptr(), b.trace_ids.data_ptr(), b.arena.data_ptr(),
                                                b.n, date, resp.out.data_ptr(), resp.out.numel() - 64,
                                                resp.out_off.data_ptr(), resp.meta.data_ptr(), st.cuda_stream),
                   "gofr_serve_device")

    # ---- routing only: stage 1 of the split API for closures that stay on the host ----
    def route_device(self, b: DeviceBatch, stream=None):
        """gofr_route_device → (meta int32[n] = status | route << 16, vars int32[n, 8] = off | len << 16)."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        meta = torch.zeros(max(b.n, 1), dtype=torch.int32, device=dev)
        vars_ = torch.zeros((max(b.n, 1), 8), dtype=torch.int32, device=dev)
        _abi.check(_abi.lib().gofr_route_device(self._e, b.desc.data_ptr(), b.arena.data_ptr(), b.n, meta.data_ptr(),
                                                vars_.data_ptr(), st.cuda_stream), "gofr_route_device")
        return meta[:b.n], vars_[:b.n]

    def proto_encode_device(self, fields, rows: np.ndarray, row_off: np.ndarray, out_cap: Optional[int] = None, stream=None):
        """gofr_proto_encode_device: rows (spec.pack_proto_rows) → packed gRPC frames.  Returns (out uint8, out_off int32
        view of uint32[n+1], meta int32 view of uint32[n]) on the device."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        n = len(row_off) - 1
        ft = (_abi.ProtoField * max(len(fields), 1))(*[_abi.ProtoField(f.number, f.type) for f in fields])
        d_rows = torch.from_numpy(np.concatenate([np.ascontiguousarray(rows), np.zeros(16, np.uint8)])).to(dev)
        d_off = torch.from_numpy(row_off.view(np.int32)).to(dev)
        cap = out_cap if out_cap is not None else int(rows.size) * 3 + 32 * n * max(len(fields), 1) + 64
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        meta = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        _abi.check(_abi.lib().gofr_proto_encode_device(self._e, ft, len(fields), d_rows.data_ptr(), d_off.data_ptr(), n,
                                                       out.data_ptr(), cap, off.data_ptr(), meta.data_ptr(), st.cuda_stream),
                   "gofr_proto_encode_device")
        return out, off, meta[:n]

    def proto_encode_nested_device(self, msgs, root: int, rows: np.ndarray, row_off: np.ndarray, out_cap: Optional[int] = None, stream=None):
        """gofr_proto_encode_nested_device: rows (spec.pack_proto_nested_rows) of message types with nested / repeated fields →
        packed gRPC frames.  Returns (out, out_off, meta) on the device."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        n = len(row_off) - 1
        nm, nf, k = S.proto_nested_tables(msgs)
        d_rows = torch.from_numpy(np.concatenate([np.ascontiguousarray(rows), np.zeros(16, np.uint8)])).to(dev)
        d_off = torch.from_numpy(row_off.view(np.int32)).to(dev)
        cap = out_cap if out_cap is not None else int(rows.size) * 4 + 64 * n + 64
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        meta = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        _abi.check(_abi.lib().gofr_proto_encode_nested_device(self._e, nm.ctypes.data, len(msgs), nf.ctypes.data, k, root, d_rows.data_ptr(),
                                                              d_off.data_ptr(), n, out.data_ptr(), cap, off.data_ptr(), meta.data_ptr(),
                                                              st.cuda_stream), "gofr_proto_encode_nested_device")
        return out, off, meta[:n]

    def proto_decode_nested_device(self, msgs, root: int, frames: np.ndarray, in_off: np.ndarray, rows_cap: Optional[int] = None, stream=None):
        """gofr_proto_decode_nested_device: packed gRPC frames of message types with nested / repeated fields → rows.
        Returns (rows uint8, row_off, meta) on the device."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        n = len(in_off) - 1
        nm, nf, k = S.proto_nested_tables(msgs)
        d_in = torch.from_numpy(np.concatenate([np.ascontiguousarray(frames), np.zeros(16, np.uint8)])).to(dev)
        d_off = torch.from_numpy(in_off.view(np.int32)).to(dev)
        cap = rows_cap if rows_cap is not None else int(frames.size) * 10 + 4096 * max(n, 1)
        rows = torch.empty(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        meta = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        _abi.check(_abi.lib().gofr_proto_decode_nested_device(self._e, nm.ctypes.data, len(msgs), nf.ctypes.data, k, root, d_in.data_ptr(),
                                                              d_off.data_ptr(), n, rows.data_ptr(), cap, off.data_ptr(), meta.data_ptr(),
                                                              st.cuda_stream), "gofr_proto_decode_nested_device")
        return rows, off, meta[:n]

    def proto_decode_device(self, fields, frames: np.ndarray, in_off: np.ndarray, rows_cap: Optional[int] = None, stream=None):
        """gofr_proto_decode_device: packed gRPC frames → rows.  Returns (rows uint8, row_off, meta) on the device."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        n = len(in_off) - 1
        ft = (_abi.ProtoField * max(len(fields), 1))(*[_abi.ProtoField(f.number, f.type) for f in fields])
        d_in = torch.from_numpy(np.concatenate([np.ascontiguousarray(frames), np.zeros(16, np.uint8)])).to(dev)
        d_off = torch.from_numpy(in_off.view(np.int32)).to(dev)
        cap = rows_cap if rows_cap is not None else int(frames.size) + (8 * len(fields) + 8) * n + 64
        rows = torch.empty(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        meta = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        _abi.check(_abi.lib().gofr_proto_decode_device(self._e, ft, len(fields), d_in.data_ptr(), d_off.data_ptr(), n,
                                                       rows.data_ptr(), cap, off.data_ptr(), meta.data_ptr(), st.cuda_stream),
                   "gofr_proto_decode_device")
        return rows, off, meta[:n]

    def route_host(self, batch: S.RequestBatch):
        """gofr_batch_route: the same for a batch in host memory → (meta uint32[n], vars uint32[n, 8])."""
        rb = _abi.ReqBatch(desc=batch.desc.ctypes.data, trace_ids=batch.trace_ids.ctypes.data, arena=batch.arena.ctypes.data,
                           arena_bytes=batch.arena.size, n=batch.n, date=b"")
        meta = np.zeros(max(batch.n, 1), dtype=np.uint32)
        vars_ = np.zeros((max(batch.n, 1), 8), dtype=np.uint32)
        _abi.check(_abi.lib().gofr_batch_route(self._e, C.byref(rb), meta.ctypes.data, vars_.ctypes.data), "gofr_batch_route")
        return meta[:batch.n], vars_[:batch.n]

    # ---- HTTP/1.1 request heads → request descriptors (the step in front of the router) ----
    def http_parse_device(self, raw: np.ndarray, raw_off: np.ndarray, stream=None):
        """gofr_http_parse_device on host arrays (uploaded here).  Returns torch tensors (desc uint8[n*16], arena uint8,
        status int32[n], spans int64[n, 6]) — desc/arena can be fed to the serve calls as they are."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        n = len(raw_off) - 1
        d_raw = torch.from_numpy(np.concatenate([raw, np.zeros(32, dtype=np.uint8)])).to(dev)
        d_off = torch.from_numpy(raw_off.astype(np.uint32).view(np.int32)).to(dev)
        desc = torch.zeros(max(n, 1) * 16, dtype=torch.uint8, device=dev)
        arena = torch.zeros(int(raw.size) + 48, dtype=torch.uint8, device=dev)
        status = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        spans = torch.zeros((max(n, 1), 6), dtype=torch.int64, device=dev)
        _abi.check(_abi.lib().gofr_http_parse_device(self._e, d_raw.data_ptr(), d_off.data_ptr(), n, desc.data_ptr(),
                                                     arena.data_ptr(), status.data_ptr(), spans.data_ptr(), st.cuda_stream),
                   "gofr_http_parse_device")
        return desc[:n * 16], arena, status[:n], spans[:n]

    # ---- RequestLog lines (middleware.Logging → logger.Log), device resident ----
    def request_log_device(self, batch: S.LogBatch, out_cap: Optional[int] = None, stream=None):
        """Uploads a LogBatch, runs gofr_requestlog_device, returns (out, out_off) torch tensors (uint8, int32 view)."""
        import torch
        dev = torch.device("cuda", self.device)
        n = batch.n
        d_desc = torch.from_numpy(batch.desc.view(np.uint8).reshape(-1).copy()).to(dev)
        d_ids = torch.from_numpy(batch.trace_ids.reshape(-1).copy()).to(dev)
        d_arena = torch.from_numpy(np.concatenate([batch.arena, np.zeros(48, dtype=np.uint8)])).to(dev)
        if out_cap is None:
            out_cap = 400 * n + 6 * int(batch.arena.size) + 64
        d_out = torch.empty(out_cap + 64, dtype=torch.uint8, device=dev)
        d_off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        self.request_log_resident(d_desc, d_ids, d_arena, n, d_out, out_cap, d_off, stream)
        return d_out, d_off

    def request_log_resident(self, d_desc, d_ids, d_arena, n: int, d_out, out_cap: int, d_off, stream=None) -> None:
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _abi.check(_abi.lib().gofr_requestlog_device(self._e, d_desc.data_ptr(), d_ids.data_ptr(), d_arena.data_ptr(), n,
                                                     d_out.data_ptr(), out_cap, d_off.data_ptr(), st.cuda_stream),
                   "gofr_requestlog_device")

    def serve_device_slots(self, b: DeviceBatch, date: bytes, slot_bytes: int, out=None, out_len=None, meta=None, stream=None):
        """gofr_serve_device_slots: response i at out[i * slot_bytes:], its length in out_len[i].  Returns
        (out uint8[n * slot_bytes], out_len int32[n], meta int32[n]); pass the buffers to reuse / pre-fill them."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        if out is None:
            out = torch.empty(max(b.n, 1) * slot_bytes, dtype=torch.uint8, device=dev)
        if out_len is None:
            out_len = torch.zeros(max(b.n, 1), dtype=torch.int32, device=dev)
        if meta is None:
            meta = torch.zeros(max(b.n, 1), dtype=torch.int32, device=dev)
        assert len(date) == 29
        _abi.check(_abi.lib().gofr_serve_device_slots(self._e, b.desc.data_ptr(), b.trace_ids.data_ptr(), b.arena.data_ptr(),
                                                      b.n, date, out.data_ptr(), slot_bytes, out_len.data_ptr(),
                                                      meta.data_ptr(), st.cuda_stream), "gofr_serve_device_slots")
        return out, out_len[:b.n], meta[:b.n]

    def bind_device(self, b: DeviceBatch, schema_id: int, slot_bytes: int, stream=None):
        """gofr_bind_device: Context.Bind for host closures.  Returns (rows uint8[n * slot_bytes], len int32[n], status int32[n])."""
        import torch
        dev = torch.device("cuda", self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        rows = torch.zeros(max(b.n, 1) * slot_bytes, dtype=torch.uint8, device=dev)
        ln = torch.zeros(max(b.n, 1), dtype=torch.int32, device=dev)
        status = torch.zeros(max(b.n, 1), dtype=torch.int32, device=dev)
        _abi.check(_abi.lib().gofr_bind_device(self._e, schema_id, b.desc.data_ptr(), b.arena.data_ptr(), b.n, rows.data_ptr(),
                                               slot_bytes, ln.data_ptr(), status.data_ptr(), st.cuda_stream), "gofr_bind_device")
        return rows, ln[:b.n], status[:b.n]

    # ---- host path (the call a user makes): host buffers in, host buffers out ----
    def serve_host(self, batch: S.RequestBatch, date: bytes, out: np.ndarray, out_off: np.ndarray, meta: np.ndarray) -> int:
        L = _abi.lib()
        rb = _abi.ReqBatch(desc=batch.desc.ctypes.data, trace_ids=batch.trace_ids.ctypes.data,
                           arena=batch.arena.ctypes.data, arena_bytes=batch.arena.size, n=batch.n, date=date)
        ob = _abi.RespBatch(out=out.ctypes.data, out_cap=out.size, out_off=out_off.ctypes.data, meta=meta.ctypes.data,
                            out_bytes=0)
        t = C.c_uint64()
        _abi.check(L.gofr_batch_submit(self._e, C.byref(rb), C.byref(ob), C.byref(t)), "gofr_batch_submit")
        _abi.check(L.gofr_batch_wait(self._e, t.value), "gofr_batch_wait")
        return int(ob.out_bytes)


def _serve_host_slots(self, batch: S.RequestBatch, date: bytes, slot_bytes: int, out: np.ndarray, out_len: np.ndarray,
                      meta: np.ndarray) -> None:
    """gofr_batch_submit_slots + wait: host buffers in, response i in out[i * slot_bytes:], its length in out_len[i]."""
    L = _abi.lib()
    rb = _abi.ReqBatch(desc=batch.desc.ctypes.data, trace_ids=batch.trace_ids.ctypes.data, arena=batch.arena.ctypes.data,
                       arena_bytes=batch.arena.size, n=batch.n, date=date)
    sb = _abi.SlotBatch(out=out.ctypes.data, slot_bytes=slot_bytes, reserved=0, out_len=out_len.ctypes.data, meta=meta.ctypes.data)
    t = C.c_uint64()
    _abi.check(L.gofr_batch_submit_slots(self._e, C.byref(rb), C.byref(sb), C.byref(t)), "gofr_batch_submit_slots")
    _abi.check(L.gofr_batch_wait(self._e, t.value), "gofr_batch_wait")


Engine.serve_host_slots = _serve_host_slots


def pinned_array(nbytes: int, dtype=np.uint8) -> np.ndarray:
    """numpy view over cudaMallocHost memory (gofr_alloc_pinned).  The memory lives until process exit."""
    L = _abi.lib()
    p = L.gofr_alloc_pinned(max(nbytes, 1))
    if not p:
        raise MemoryError(L.gofr_last_error().decode())
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
    a = np.frombuffer(buf, dtype=np.uint8)[:nbytes]
    return a.view(dtype)


def pin_batch(batch: S.RequestBatch) -> S.RequestBatch:
    d = pinned_array(batch.desc.nbytes).view(S.DESC_DTYPE)
    d[:] = batch.desc
    i = pinned_array(batch.trace_ids.nbytes).reshape(-1, 16)
    i[:] = batch.trace_ids
    a = pinned_array(batch.arena.nbytes)
    a[:] = batch.arena
    return S.RequestBatch(d, i, a)

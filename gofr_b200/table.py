"""Route table: registration → sealed image, through the C ABI (mirrors Router.Add, pkg/gofr/http/router.go:30-33)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from . import spec as S


class Table:
    def __init__(self, spec: S.TableSpec | None = None, image: bytes | None = None):
        L = _abi.lib()
        self._t = C.c_void_p()
        self.spec = spec
        self.route_ids = []
        if image is not None:
            buf = np.frombuffer(image, dtype=np.uint8)
            _abi.check(L.gofr_table_deserialize(C.byref(self._t), buf.ctypes.data, len(image)), "gofr_table_deserialize")
            return
        assert spec is not None
        _abi.check(L.gofr_table_create(C.byref(self._t), spec.frame_mode), "gofr_table_create")
        for sc in spec.schemas:
            arr = (_abi.FieldDesc * len(sc.fields))()
            keep = []
            for i, f in enumerate(sc.fields):
                g, j = f.go_name.encode(), f.json_name.encode()
                keep += [g, j]
                arr[i].go_name, arr[i].json_name, arr[i].kind, arr[i].omitempty = g, j, f.kind, 1 if f.omitempty else 0
                arr[i].container, arr[i].flags, arr[i].elem_schema = f.container, f.flags, f.elem_schema
            _abi.check(L.gofr_table_add_schema(self._t, sc.id, sc.go_type.encode(), arr, len(sc.fields)),
                       "gofr_table_add_schema")
        for r in spec.routes:
            h = _abi.HandlerDesc(kind=r.kind, schema_id=r.schema_id, s0=r.s0, s0_len=len(r.s0), s1=r.s1, s1_len=len(r.s1),
                                 s2=r.s2, s2_len=len(r.s2), s3=r.s3, s3_len=len(r.s3), blob=r.blob, blob_len=len(r.blob))
            rid = C.c_uint32()
            p = r.pattern.encode()
            _abi.check(L.gofr_table_add_route(self._t, r.method, p, len(p), C.byref(h), C.byref(rid)),
                       f"gofr_table_add_route({r.pattern})")
            self.route_ids.append(rid.value)
        if spec.default_routes:
            _abi.check(L.gofr_table_add_default_routes(self._t, spec.favicon, len(spec.favicon)),
                       "gofr_table_add_default_routes")
        _abi.check(L.gofr_table_seal(self._t), "gofr_table_seal")

    @property
    def handle(self):
        return self._t

    def serialize(self) -> bytes:
        L = _abi.lib()
        n = C.c_uint64(0)
        _abi.check(L.gofr_table_serialize(self._t, None, C.byref(n)), "gofr_table_serialize")
        buf = np.zeros(n.value, dtype=np.uint8)
        _abi.check(L.gofr_table_serialize(self._t, buf.ctypes.data, C.byref(n)), "gofr_table_serialize")
        return buf.tobytes()

    def slot_ctas(self) -> int:
        """CTAs/SM of the slot-layout serve kernel an engine will pick for this table (4: the 128-register instance)"""
        v = C.c_int(0)
        _abi.check(_abi.lib().gofr_table_slot_ctas(self._t, C.byref(v)), "gofr_table_slot_ctas")
        return v.value

    def route_count(self) -> int:
        return _abi.lib().gofr_table_route_count(self._t)

    def __del__(self):
        try:
            if self._t:
                _abi.lib().gofr_table_destroy(self._t)
        except Exception:
            pass

// grpc_device.cuh — per-frame device logic of the gRPC message path: the unary Hello of BASELINE config 5 (first part)
// and the proto3 encoder / decoder for flat message types (second part).
//
// Replaces, for a batch of length-prefixed messages, what grpc-go + the generated code do per RPC in the reference:
//   dec(in) in _Hello_SayHello_Handler (examples/grpc-server/grpc/hello_grpc.pb.go:73-89): strip the 5-byte gRPC
//   header (compressed flag + big-endian u32 length), proto.Unmarshal HelloRequest{name = 1} (hello.proto:4-6)
//   → Server.SayHello (examples/grpc-server/grpc/server.go:12-21): "Hello " + (name or "World") + "!"
//   → proto.Marshal HelloResponse{message = 1} (hello.proto:8-10) + 5-byte header.
// Wire rules follow protobuf-go v1.32.0 / grpc-go v1.60.1 (go.mod:11,23): varints ≤ 10 bytes, field numbers
// 1..2^29-1, unknown fields of every wire type skipped (groups balanced), the last occurrence of field 1 wins, proto3
// strings must be valid UTF-8.  __host__ __device__ so tests/emu runs the same code on the CPU.
#pragma once
#include "engine_internal.h"
#include "serve_device.cuh"

namespace gofr {

struct HelloReq {
    uint32_t status;      // GOFR_GRPC_*
    uint32_t name_off;    // offset of the name inside the frame
    uint32_t name_len;
    uint32_t out_len;     // bytes of the response frame (0 on error)
};

// protowire.ConsumeVarint
GOFR_HD int grpc_varint(const uint8_t* p, uint32_t n, uint64_t* v) {
    uint64_t x = 0;
    for (uint32_t i = 0; i < 10; i++) {
        if (i >= n) return -1;
        uint32_t b = p[i];
        if (i == 9 && b > 1) return -1;
        x |= (uint64_t)(b & 0x7F) << (7 * i);
        if (b < 0x80) { *v = x; return (int)i + 1; }
    }
    return -1;
}

GOFR_HD bool grpc_utf8_ok(const uint8_t* s, uint32_t n) {
    uint32_t i = 0;
    while (i < n) {
        if (s[i] < 0x80) { i++; continue; }
        uint32_t L = utf8_len_at(s + i, n - i);
        if (!L) return false;
        i += L;
    }
    return true;
}

// the same with ASCII runs taken four bytes at a time once the address is word aligned (message strings are longer than
// the Hello path's names, where the extra test costs more than it saves: 0.070 → 0.073 ms)
GOFR_HD bool proto_utf8_ok(const uint8_t* s, uint32_t n) {
    uint32_t i = 0;
    while (i < n) {
        if (s[i] < 0x80) {
            i++;
            if ((((uintptr_t)(s + i)) & 3u) == 0)
                while (i + 4 <= n && (*(const uint32_t*)(s + i) & 0x80808080u) == 0) i += 4;
            continue;
        }
        uint32_t L = utf8_len_at(s + i, n - i);
        if (!L) return false;
        i += L;
    }
    return true;
}

GOFR_HD uint32_t varint_len(uint32_t v) { return v < 0x80 ? 1u : v < 0x4000 ? 2u : v < 0x200000 ? 3u : v < 0x10000000 ? 4u : 5u; }

constexpr int kMaxGroupDepth = 16;  // deeper unknown-group nesting is reported as BAD_PROTO (upstream allows more)

// parse + size one frame [f, f+fn)
GOFR_HD HelloReq hello_parse(const uint8_t* f, uint32_t fn) {
    HelloReq r = {GOFR_GRPC_OK, 0, 0, 0};
    if (fn < 5) { r.status = GOFR_GRPC_BAD_LENGTH; return r; }
    if (f[0] == 1) { r.status = GOFR_GRPC_COMPRESSED; return r; }  // no compressor is registered (pkg/gofr/grpc.go:23-26)
    if (f[0] != 0) { r.status = GOFR_GRPC_BAD_LENGTH; return r; }
    uint32_t L = (uint32_t)f[1] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[3] << 8 | f[4];
    if (L != fn - 5) { r.status = GOFR_GRPC_BAD_LENGTH; return r; }
    const uint8_t* p = f + 5;
    uint32_t i = 0, depth = 0;
    uint32_t stack[kMaxGroupDepth];
    while (i < L) {
        uint64_t tag, v;
        int k = grpc_varint(p + i, L - i, &tag);
        if (k < 0) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
        i += (uint32_t)k;
        uint64_t num = tag >> 3;
        uint32_t wt = (uint32_t)(tag & 7);
        if (num == 0 || num > 0x1FFFFFFFull) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
        if (wt == 0) {
            k = grpc_varint(p + i, L - i, &v);
            if (k < 0) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            i += (uint32_t)k;
        } else if (wt == 1) {
            if (L - i < 8) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            i += 8;
        } else if (wt == 5) {
            if (L - i < 4) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            i += 4;
        } else if (wt == 2) {
            k = grpc_varint(p + i, L - i, &v);
            if (k < 0) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            i += (uint32_t)k;
            if (v > L - i) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            if (num == 1 && depth == 0) {
                if (!grpc_utf8_ok(p + i, (uint32_t)v)) { r.status = GOFR_GRPC_BAD_UTF8; return r; }
                r.name_off = 5 + i;
                r.name_len = (uint32_t)v;
            }
            i += (uint32_t)v;
        } else if (wt == 3) {
            if (depth == kMaxGroupDepth) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            stack[depth++] = (uint32_t)num;
        } else if (wt == 4) {
            if (depth == 0 || stack[depth - 1] != (uint32_t)num) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
            depth--;
        } else { r.status = GOFR_GRPC_BAD_PROTO; return r; }
    }
    if (depth != 0) { r.status = GOFR_GRPC_BAD_PROTO; return r; }
    uint32_t nl = r.name_len ? r.name_len : 5;  // "World"
    uint32_t ml = 6 + nl + 1;                   // fmt.Sprintf("Hello %s!", name)
    r.out_len = 5 + 1 + varint_len(ml) + ml;
    return r;
}

// write the response frame at dst (arbitrary alignment inside the packed output)
GOFR_HD void hello_emit(const uint8_t* f, const HelloReq r, uint8_t* dst, uint32_t* stage_col) {
    if (!r.out_len) return;
    Writer w;
    w.init(dst, stage_col);
    uint32_t nl = r.name_len ? r.name_len : 5;
    uint32_t ml = 7 + nl, vl = varint_len(ml), plen = 1 + vl + ml;
    w.put4(0u | (plen >> 24) << 8 | ((plen >> 16) & 0xFF) << 16 | ((plen >> 8) & 0xFF) << 24);  // 00, be32[0..2]
    w.putk((plen & 0xFF) | 0x0Au << 8, 2);                                                      // be32[3], tag
    for (uint32_t v = ml;;) {  // varint(len(message))
        uint32_t b = v & 0x7F;
        v >>= 7;
        w.putc(v ? b | 0x80 : b);  // at most 21 bytes precede the name: the staging buffer cannot fill up
        if (!v) break;
    }
    w.put4('H' | 'e' << 8 | 'l' << 16 | 'l' << 24);
    w.putk('o' | ' ' << 8, 2);
    if (r.name_len) w.copy<false>(f + r.name_off, r.name_len);
    else { w.put4('W' | 'o' << 8 | 'r' << 16 | 'l' << 24); w.putc('d'); }
    w.reserve(2);
    w.putc('!');
    w.finish();
}

// ---------------------------------------------------------------------------------------------------------------
// proto3 message encoder (gofr_proto_encode_device): proto.Marshal of the message a unary handler
// returns (examples/grpc-server/grpc/hello_grpc.pb.go:73-89 → grpc-go's proto codec, protobuf-go v1.32.0) plus the
// 5-byte length-prefixed-message header, for flat messages with scalar fields.  Fields are emitted in the order given
// (= field-number order), a field holding its zero value is skipped (proto3 implicit presence), strings must be valid
// UTF-8.  Rows use the GOFR_H_ROW layout (fixed words, then string bytes).
// ---------------------------------------------------------------------------------------------------------------
static_assert(GOFR_PROTO_MAX_FIELDS == 32, "ProtoSchema (engine_internal.h) is sized for 32 fields");
static_assert(GOFR_PB_DOUBLE == 1 && GOFR_PB_FLOAT == 2 && GOFR_PB_INT64 == 3 && GOFR_PB_UINT64 == 4 && GOFR_PB_INT32 == 5 &&
              GOFR_PB_FIXED64 == 6 && GOFR_PB_FIXED32 == 7 && GOFR_PB_BOOL == 8 && GOFR_PB_STRING == 9 && GOFR_PB_BYTES == 12 &&
              GOFR_PB_ENUM == 14 && GOFR_PB_SFIXED32 == 15 && GOFR_PB_SFIXED64 == 16 && GOFR_PB_SINT32 == 17 && GOFR_PB_SINT64 == 18,
              "proto_class (engine_internal.h) spells the FieldDescriptorProto.Type numbers out");

struct ProtoMsg {
    uint32_t status;   // GOFR_GRPC_OK / GOFR_GRPC_BAD_UTF8 / GOFR_GRPC_BAD_ROW
    uint32_t out_len;  // 5 + message bytes (0 on error)
};

GOFR_HD uint32_t varint_len64(uint64_t v) {
    // ceil(bits / 7), bits = position of the highest set bit (1 for v == 0)
#if defined(__CUDA_ARCH__)
    const uint32_t bits = 64u - (uint32_t)__clzll((long long)(v | 1));
#else
    const uint32_t bits = 64u - (uint32_t)__builtin_clzll(v | 1);
#endif
    return (bits * 9u + 64u) >> 6;  // == (bits + 6) / 7 for bits in 1..64
}
// the varint payload of a varint-typed field (what follows the tag), from the row's words
GOFR_HD uint64_t proto_varint_value(uint32_t cls, uint32_t w0, uint32_t w1) {
    const uint64_t v64 = (uint64_t)w0 | (uint64_t)w1 << 32;
    if (cls & PC_ZIGZAG) return (cls & PC_64) ? (v64 << 1) ^ (uint64_t)((int64_t)v64 >> 63) : (uint32_t)((w0 << 1) ^ (uint32_t)((int32_t)w0 >> 31));
    if (cls & PC_SIGNEXT) return (uint64_t)(int64_t)(int32_t)w0;  // negative: ten bytes
    if (cls & PC_BOOL) return w0 ? 1u : 0u;
    return v64;  // int64 / uint64 / uint32 (w1 == 0)
}

// size pass: validates the row and returns the exact frame length
GOFR_HD ProtoMsg proto_size(const ProtoSchema& S, const uint8_t* row, uint32_t rn, bool aligned) {
    ProtoMsg m = {GOFR_GRPC_OK, 0};
    if (!aligned || rn < S.fixed_bytes) { m.status = GOFR_GRPC_BAD_ROW; return m; }
    const uint32_t* w = (const uint32_t*)row;
    uint32_t wi = 0, spos = S.fixed_bytes, len = 0;
    for (uint32_t k = 0; k < S.n_fields; k++) {
        const uint32_t cls = S.cls[k], tl = varint_len(S.tag[k]);
        const uint32_t w0 = w[wi], w1 = (cls & PC_64) ? w[wi + 1] : 0u;
        wi += (cls & PC_64) ? 2u : 1u;
        const uint32_t wire = cls & PC_WIRE;
        if (wire == 2) {
            if (w0 > rn - spos) { m.status = GOFR_GRPC_BAD_ROW; return m; }
            if ((cls & PC_UTF8) && !proto_utf8_ok(row + spos, w0)) { m.status = GOFR_GRPC_BAD_UTF8; return m; }
            if (w0) len += tl + varint_len(w0) + w0;
            spos += w0;
        } else if (w0 | w1) {
            len += tl + (wire == 1 ? 8u : wire == 5 ? 4u : varint_len64(proto_varint_value(cls, w0, w1)));
        }
    }
    m.out_len = 5 + len;
    return m;
}

GOFR_HD void proto_put_varint(Writer& w, uint64_t v) {
    for (;;) {
        const uint32_t b = (uint32_t)v & 0x7Fu;
        v >>= 7;
        w.putc(v ? b | 0x80u : b);
        if (!v) break;
    }
}

// emit pass: the frame at dst (arbitrary alignment inside the packed output)
GOFR_HD void proto_emit(const ProtoSchema& S, const uint8_t* row, const ProtoMsg m, uint8_t* dst, uint32_t* stage_col) {
    if (!m.out_len) return;
    Writer w;
    w.init(dst, stage_col);
    const uint32_t plen = m.out_len - 5;
    w.put4(0u | (plen >> 24) << 8 | ((plen >> 16) & 0xFF) << 16 | ((plen >> 8) & 0xFF) << 24);  // 00, be32[0..2]
    w.putc(plen & 0xFF);
    const uint32_t* rw = (const uint32_t*)row;
    uint32_t wi = 0, spos = S.fixed_bytes;
    for (uint32_t k = 0; k < S.n_fields; k++) {
        const uint32_t cls = S.cls[k];
        const uint32_t w0 = rw[wi], w1 = (cls & PC_64) ? rw[wi + 1] : 0u;
        wi += (cls & PC_64) ? 2u : 1u;
        const uint32_t wire = cls & PC_WIRE;
        if (wire == 2) {
            if (w0) {
                w.reserve(4);  // tag (<= 5 bytes) + length (<= 5) + carried bytes
                proto_put_varint(w, S.tag[k]);
                proto_put_varint(w, w0);
                w.copy<false>(row + spos, w0);
            }
            spos += w0;
        } else if (w0 | w1) {
            w.reserve(6);  // tag (<= 5 bytes) + payload (<= 10) + carried bytes: at most 5 new words
            proto_put_varint(w, S.tag[k]);
            if (wire == 1) { w.put4(w0); w.put4(w1); }
            else if (wire == 5) w.put4(w0);
            else proto_put_varint(w, proto_varint_value(cls, w0, w1));
        }
    }
    w.finish();
}

// ---------------------------------------------------------------------------------------------------------------
// proto3 message decoder (gofr_proto_decode_device): dec(in) of a unary handler (hello_grpc.pb.go:73-89) for any flat
// message type — length-prefixed frame → proto.Unmarshal → the field values as a row (the layout the encoder reads).
// Same wire rules as hello_parse (varints <= 10 bytes, numbers 1..2^29-1, unknown fields and balanced groups skipped);
// a known field carrying a foreign wire type is an unknown field; scalars: last occurrence wins, 32-bit kinds keep the
// low 32 bits of the varint, bool is v != 0, sint kinds are zigzag decoded; strings must be valid UTF-8.
// ---------------------------------------------------------------------------------------------------------------
struct ProtoRow {
    uint32_t status;   // GOFR_GRPC_*
    uint32_t out_len;  // bytes of the row (multiple of 4; 0 on error)
    // per field: scalar words (lo, hi) — or, for string / bytes, the value's offset in the frame and its length
    uint32_t a[GOFR_PROTO_MAX_FIELDS];
    uint32_t b[GOFR_PROTO_MAX_FIELDS];
};

GOFR_HD void proto_decode_scan(const ProtoSchema& S, const uint8_t* f, uint32_t fn, ProtoRow& r) {
    r.status = GOFR_GRPC_OK;
    r.out_len = 0;
    for (uint32_t k = 0; k < S.n_fields; k++) { r.a[k] = 0; r.b[k] = 0; }
    if (fn < 5) { r.status = GOFR_GRPC_BAD_LENGTH; return; }
    if (f[0] == 1) { r.status = GOFR_GRPC_COMPRESSED; return; }
    if (f[0] != 0) { r.status = GOFR_GRPC_BAD_LENGTH; return; }
    const uint32_t L = (uint32_t)f[1] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[3] << 8 | f[4];
    if (L != fn - 5) { r.status = GOFR_GRPC_BAD_LENGTH; return; }
    const uint8_t* p = f + 5;
    uint32_t i = 0, depth = 0;
    uint32_t stack[kMaxGroupDepth];
    while (i < L) {
        uint64_t tag, v = 0;
        int k = grpc_varint(p + i, L - i, &tag);
        if (k < 0) { r.status = GOFR_GRPC_BAD_PROTO; return; }
        i += (uint32_t)k;
        const uint64_t num = tag >> 3;
        const uint32_t wt = (uint32_t)(tag & 7);
        if (num == 0 || num > 0x1FFFFFFFull) { r.status = GOFR_GRPC_BAD_PROTO; return; }
        // the schema field this tag addresses: same number AND same wire type, outside any unknown group
        uint32_t fld = 0xFFFFFFFFu;
        if (depth == 0) {
            const uint32_t want = (uint32_t)tag;  // number << 3 | wire, exactly what S.tag holds
            for (uint32_t q = 0; q < S.n_fields; q++)
                if (S.tag[q] == want) fld = q;
        }
        uint32_t poff = 0, plen = 0;
        if (wt == 0) {
            k = grpc_varint(p + i, L - i, &v);
            if (k < 0) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            i += (uint32_t)k;
        } else if (wt == 1) {
            if (L - i < 8) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            for (int q = 7; q >= 0; q--) v = v << 8 | p[i + (uint32_t)q];
            i += 8;
        } else if (wt == 5) {
            if (L - i < 4) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            for (int q = 3; q >= 0; q--) v = v << 8 | p[i + (uint32_t)q];
            i += 4;
        } else if (wt == 2) {
            k = grpc_varint(p + i, L - i, &v);
            if (k < 0) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            i += (uint32_t)k;
            if (v > L - i) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            poff = 5 + i;
            plen = (uint32_t)v;
            i += plen;
        } else if (wt == 3) {
            if (depth == kMaxGroupDepth) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            stack[depth++] = (uint32_t)num;
            continue;
        } else if (wt == 4) {
            if (depth == 0 || stack[depth - 1] != (uint32_t)num) { r.status = GOFR_GRPC_BAD_PROTO; return; }
            depth--;
            continue;
        } else { r.status = GOFR_GRPC_BAD_PROTO; return; }
        if (fld == 0xFFFFFFFFu) continue;
        const uint32_t cls = S.cls[fld];
        if (wt == 2) {
            if ((cls & PC_UTF8) && !proto_utf8_ok(f + poff, plen)) { r.status = GOFR_GRPC_BAD_UTF8; return; }
            r.a[fld] = poff;
            r.b[fld] = plen;
        } else if (wt != 0 || ((cls & PC_64) && !(cls & PC_ZIGZAG))) {
            r.a[fld] = (uint32_t)v;            // fixed32 / fixed64 / int64 / uint64: the bits
            r.b[fld] = (uint32_t)(v >> 32);
        } else if (cls & PC_ZIGZAG) {
            if (cls & PC_64) {
                const uint64_t z = (v >> 1) ^ (uint64_t)-(int64_t)(v & 1);
                r.a[fld] = (uint32_t)z;
                r.b[fld] = (uint32_t)(z >> 32);
            } else {
                const uint32_t x = (uint32_t)v;
                r.a[fld] = (x >> 1) ^ (uint32_t)-(int32_t)(x & 1);
            }
        } else if (cls & PC_BOOL) {
            r.a[fld] = v != 0 ? 1u : 0u;
        } else {
            r.a[fld] = (uint32_t)v;            // int32 / uint32 / enum: the low 32 bits
        }
    }
    if (depth != 0) { r.status = GOFR_GRPC_BAD_PROTO; return; }
    uint32_t need = S.fixed_bytes;
    for (uint32_t k = 0; k < S.n_fields; k++)
        if ((S.cls[k] & PC_WIRE) == 2) need += r.b[k];
    r.out_len = (need + 3u) & ~3u;
}

GOFR_HD void proto_decode_emit(const ProtoSchema& S, const uint8_t* f, const ProtoRow& r, uint8_t* dst, uint32_t* stage_col) {
    if (!r.out_len) return;
    Writer w;
    w.init(dst, stage_col);
    uint32_t produced = S.fixed_bytes;
    for (uint32_t k = 0; k < S.n_fields; k++) {
        const uint32_t cls = S.cls[k];
        w.reserve(3);
        if ((cls & PC_WIRE) == 2) w.put4(r.b[k]);
        else { w.put4(r.a[k]); if (cls & PC_64) w.put4(r.b[k]); }
    }
    for (uint32_t k = 0; k < S.n_fields; k++)
        if ((S.cls[k] & PC_WIRE) == 2 && r.b[k]) { w.copy<false>(f + r.a[k], r.b[k]); produced += r.b[k]; }
    w.reserve(2);
    for (; produced & 3u; produced++) w.putc(0);
    w.finish();
}

}  // namespace gofr

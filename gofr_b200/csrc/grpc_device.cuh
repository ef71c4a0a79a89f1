// grpc_device.cuh — per-frame device logic of the gRPC unary Hello path (BASELINE config 5).
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

}  // namespace gofr

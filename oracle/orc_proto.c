/*
 * orc_proto.c — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h; parity unpinned: no Go toolchain in this image).
 *
 * What the reference's gRPC server does with the message a unary handler returns (examples/grpc-server/grpc/
 * hello_grpc.pb.go:73-89 hands the handler's response to grpc-go): the proto codec calls proto.Marshal (protobuf-go
 * v1.32.0, go.mod:23) and grpc-go v1.60.1 (go.mod:11) prepends the 5-byte length-prefixed-message header.  Restated
 * for flat proto3 messages with scalar fields ("general proto3 schema encoder — multi-field,
 * varint / zigzag / fixed"):
 *   - fields are emitted in field-number order (protobuf-go orders a generated message's coders by number);
 *   - implicit presence: a field holding its zero value is not emitted (ints 0, false, "", +0.0 — a float whose bits
 *     are not all zero, e.g. -0.0, IS emitted);
 *   - int32 / enum: negative values are sign-extended to 64 bits (10-byte varint); sint32 / sint64: zigzag;
 *     fixed / float / double: little-endian raw bits; string / bytes: length-delimited, strings must be valid UTF-8
 *     (proto.Marshal fails with "string field contains invalid UTF-8" → the RPC fails, no frame is produced);
 *   - the frame: 0x00 (not compressed: no compressor is registered, pkg/gofr/grpc.go:23-26), big-endian u32 length.
 * Independent check: tests/test_proto.py builds the same message types with python google.protobuf (descriptor built at
 * run time) and compares SerializeToString() byte for byte.
 *
 * Row format (the one GOFR_H_ROW rows use): per field, in the order given, 64-bit kinds two LE words (lo, hi), 32-bit
 * kinds and bool one word, string / bytes one word holding the byte length; then the bytes of all string / bytes fields
 * concatenated.  Row i is rows[row_off[i] .. row_off[i+1]); offsets are multiples of 4.
 */
#include "gofr_oracle.h"
#include "orc_internal.h"

enum { PB_DOUBLE = 1, PB_FLOAT = 2, PB_INT64 = 3, PB_UINT64 = 4, PB_INT32 = 5, PB_FIXED64 = 6, PB_FIXED32 = 7, PB_BOOL = 8,
       PB_STRING = 9, PB_BYTES = 12, PB_UINT32 = 13, PB_ENUM = 14, PB_SFIXED32 = 15, PB_SFIXED64 = 16, PB_SINT32 = 17,
       PB_SINT64 = 18 };
enum { ST_OK = 0, ST_BAD_UTF8 = 4, ST_BAD_ROW = 5 };

static int kind_words(uint32_t t) {
    switch (t) {
        case PB_DOUBLE: case PB_INT64: case PB_UINT64: case PB_FIXED64: case PB_SFIXED64: case PB_SINT64: return 2;
        case PB_FLOAT: case PB_INT32: case PB_FIXED32: case PB_BOOL: case PB_STRING: case PB_BYTES: case PB_UINT32:
        case PB_ENUM: case PB_SFIXED32: case PB_SINT32: return 1;
        default: return 0;
    }
}

typedef struct { uint8_t* p; size_t n, cap; } buf;
static void b_byte(buf* b, uint8_t c) {
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 256; b->p = (uint8_t*)realloc(b->p, b->cap); }
    b->p[b->n++] = c;
}
static void b_varint(buf* b, uint64_t v) {
    while (v >= 0x80) { b_byte(b, (uint8_t)(v | 0x80)); v >>= 7; }
    b_byte(b, (uint8_t)v);
}
static void b_tag(buf* b, uint32_t number, uint32_t wire) { b_varint(b, (uint64_t)number << 3 | wire); }
static void b_le(buf* b, uint64_t v, int bytes) { for (int k = 0; k < bytes; k++) b_byte(b, (uint8_t)(v >> (8 * k))); }

static int str_utf8_valid(const uint8_t* s, size_t n) {  /* utf8.Valid */
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) { i++; continue; }
        size_t need;
        uint8_t lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) need = 1;
        else if (c >= 0xE0 && c <= 0xEF) { need = 2; if (c == 0xE0) lo = 0xA0; if (c == 0xED) hi = 0x9F; }
        else if (c >= 0xF0 && c <= 0xF4) { need = 3; if (c == 0xF0) lo = 0x90; if (c == 0xF4) hi = 0x8F; }
        else return 0;
        if (n - i <= need) return 0;
        if (s[i + 1] < lo || s[i + 1] > hi) return 0;
        for (size_t k = 2; k <= need; k++) if ((s[i + k] & 0xC0) != 0x80) return 0;
        i += need + 1;
    }
    return 1;
}

/* proto.Marshal of one row; returns ST_* */
static int marshal_row(const uint32_t* ftab, uint32_t nf, const uint8_t* row, size_t rn, buf* out) {
    size_t fixed = 0;
    for (uint32_t k = 0; k < nf; k++) fixed += 4 * (size_t)kind_words(ftab[2 * k + 1]);
    if (rn < fixed) return ST_BAD_ROW;
    size_t wpos = 0, spos = fixed;
    for (uint32_t k = 0; k < nf; k++) {
        const uint32_t number = ftab[2 * k], type = ftab[2 * k + 1];
        uint32_t w0, w1 = 0;
        memcpy(&w0, row + wpos, 4);
        if (kind_words(type) == 2) memcpy(&w1, row + wpos + 4, 4);
        wpos += 4 * (size_t)kind_words(type);
        const uint64_t v64 = (uint64_t)w0 | (uint64_t)w1 << 32;
        switch (type) {
            case PB_INT64: case PB_UINT64:
                if (v64) { b_tag(out, number, 0); b_varint(out, v64); }
                break;
            case PB_SINT64:
                if (v64) { b_tag(out, number, 0); b_varint(out, (v64 << 1) ^ (uint64_t)((int64_t)v64 >> 63)); }
                break;
            case PB_INT32: case PB_ENUM:
                if (w0) { b_tag(out, number, 0); b_varint(out, (uint64_t)(int64_t)(int32_t)w0); }
                break;
            case PB_UINT32:
                if (w0) { b_tag(out, number, 0); b_varint(out, w0); }
                break;
            case PB_SINT32:
                if (w0) { b_tag(out, number, 0); b_varint(out, (uint32_t)((w0 << 1) ^ (uint32_t)((int32_t)w0 >> 31))); }
                break;
            case PB_BOOL:
                if (w0) { b_tag(out, number, 0); b_byte(out, 1); }
                break;
            case PB_FIXED64: case PB_SFIXED64: case PB_DOUBLE:
                if (v64) { b_tag(out, number, 1); b_le(out, v64, 8); }
                break;
            case PB_FIXED32: case PB_SFIXED32: case PB_FLOAT:
                if (w0) { b_tag(out, number, 5); b_le(out, w0, 4); }
                break;
            case PB_STRING: case PB_BYTES: {
                if ((size_t)w0 > rn - spos) return ST_BAD_ROW;
                if (type == PB_STRING && !str_utf8_valid(row + spos, w0)) return ST_BAD_UTF8;
                if (w0) {
                    b_tag(out, number, 2);
                    b_varint(out, w0);
                    for (uint32_t j = 0; j < w0; j++) b_byte(out, row[spos + j]);
                }
                spos += w0;
                break;
            }
            default: return ST_BAD_ROW;
        }
    }
    return ST_OK;
}

int orc_proto_encode(const uint32_t* fields, uint32_t n_fields, const uint8_t* rows, const uint32_t* row_off, uint32_t n,
                     uint8_t* out, uint64_t out_cap, uint32_t* out_off, uint32_t* meta) {
    uint64_t pos = 0;
    buf b = {NULL, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
        out_off[i] = (uint32_t)pos;
        b.n = 0;
        int st = (row_off[i] & 3u) ? ST_BAD_ROW : marshal_row(fields, n_fields, rows + row_off[i], row_off[i + 1] - row_off[i], &b);
        meta[i] = (uint32_t)st;
        if (st != ST_OK) continue;
        if (pos + 5 + b.n > out_cap) { free(b.p); return -1; }
        uint8_t* o = out + pos;
        o[0] = 0;
        o[1] = (uint8_t)(b.n >> 24); o[2] = (uint8_t)(b.n >> 16); o[3] = (uint8_t)(b.n >> 8); o[4] = (uint8_t)b.n;
        if (b.n) memcpy(o + 5, b.p, b.n);
        pos += 5 + b.n;
    }
    out_off[n] = (uint32_t)pos;
    free(b.p);
    return 0;
}

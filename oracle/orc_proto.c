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

/* ---------------------------------------------------------------------------------------------------------------
 * The other direction: what dec(in) does for the request of a unary RPC (examples/grpc-server/grpc/
 * hello_grpc.pb.go:73-89 calls dec, grpc-go has stripped nothing yet at the level modelled here): check the 5-byte
 * length-prefixed-message header, proto.Unmarshal into a flat proto3 message type, hand the field values on — here as a
 * row in the layout the encoder takes.  protobuf-go v1.32.0 rules restated:
 *   - tags: field number 1 .. 2^29-1, wire types 0,1,2,5 and groups 3/4 (only ever unknown here, skipped balanced);
 *   - a known field arriving with another wire type than its own is an unknown field (skipped);
 *   - scalars: the last occurrence wins; int32 / uint32 / enum keep the low 32 bits of the varint, bool is v != 0,
 *     sint32 zigzag-decodes the low 32 bits, sint64 all 64;
 *   - proto3 strings must be valid UTF-8 (every occurrence, also one that is overwritten later);
 *   - absent fields hold their zero value.
 * meta: 0 ok, 1 compressed flag set (no decompressor registered, pkg/gofr/grpc.go:23-26), 2 bad length prefix,
 * 3 malformed message, 4 invalid UTF-8.  A failed frame yields an empty row.
 * Independent check: tests/test_proto.py parses the same frames with python google.protobuf and packs rows from the
 * field values it reports.
 * --------------------------------------------------------------------------------------------------------------- */
enum { ST_COMPRESSED = 1, ST_BAD_LENGTH = 2, ST_BAD_PROTO = 3 };
#define DEC_MAX_FIELDS 32
#define DEC_MAX_GROUP_DEPTH 16

typedef struct { uint64_t scalar; const uint8_t* str; size_t str_len; } dec_value;

static int dec_varint(const uint8_t* p, size_t n, uint64_t* v) { /* protowire.ConsumeVarint */
    uint64_t x = 0;
    for (int i = 0; i < 10; i++) {
        if ((size_t)i >= n) return -1;
        uint8_t b = p[i];
        if (i == 9 && b > 1) return -1;
        x |= (uint64_t)(b & 0x7F) << (7 * i);
        if (b < 0x80) { *v = x; return i + 1; }
    }
    return -1;
}

static int wire_of(uint32_t t) {
    switch (t) {
        case PB_STRING: case PB_BYTES: return 2;
        case PB_DOUBLE: case PB_FIXED64: case PB_SFIXED64: return 1;
        case PB_FLOAT: case PB_FIXED32: case PB_SFIXED32: return 5;
        default: return 0;
    }
}

static int unmarshal(const uint32_t* ftab, uint32_t nf, const uint8_t* p, size_t n, dec_value* val) {
    size_t i = 0;
    uint32_t group_stack[DEC_MAX_GROUP_DEPTH];
    int depth = 0;
    while (i < n) {
        uint64_t tag, v = 0;
        int k = dec_varint(p + i, n - i, &tag);
        if (k < 0) return ST_BAD_PROTO;
        i += (size_t)k;
        const uint64_t num = tag >> 3;
        const int wt = (int)(tag & 7);
        if (num == 0 || num > 0x1FFFFFFF) return ST_BAD_PROTO;
        int field = -1;
        if (depth == 0)
            for (uint32_t f = 0; f < nf; f++)
                if (ftab[2 * f] == num && wire_of(ftab[2 * f + 1]) == wt) field = (int)f;
        const uint8_t* payload = p + i;
        size_t plen = 0;
        switch (wt) {
            case 0:
                k = dec_varint(p + i, n - i, &v);
                if (k < 0) return ST_BAD_PROTO;
                i += (size_t)k;
                break;
            case 1:
                if (n - i < 8) return ST_BAD_PROTO;
                for (int b = 7; b >= 0; b--) v = v << 8 | p[i + (size_t)b];
                i += 8;
                break;
            case 5:
                if (n - i < 4) return ST_BAD_PROTO;
                for (int b = 3; b >= 0; b--) v = v << 8 | p[i + (size_t)b];
                i += 4;
                break;
            case 2:
                k = dec_varint(p + i, n - i, &v);
                if (k < 0) return ST_BAD_PROTO;
                i += (size_t)k;
                if (v > n - i) return ST_BAD_PROTO;
                payload = p + i;
                plen = (size_t)v;
                i += plen;
                break;
            case 3:
                if (depth == DEC_MAX_GROUP_DEPTH) return ST_BAD_PROTO;
                group_stack[depth++] = (uint32_t)num;
                continue;
            case 4:
                if (depth == 0 || group_stack[depth - 1] != (uint32_t)num) return ST_BAD_PROTO;
                depth--;
                continue;
            default: return ST_BAD_PROTO;
        }
        if (field < 0) continue;
        const uint32_t t = ftab[2 * field + 1];
        switch (t) {
            case PB_STRING:
                if (!str_utf8_valid(payload, plen)) return ST_BAD_UTF8;
                /* fall through */
            case PB_BYTES: val[field].str = payload; val[field].str_len = plen; break;
            case PB_INT32: case PB_UINT32: case PB_ENUM: val[field].scalar = (uint32_t)v; break;
            case PB_SINT32: { uint32_t x = (uint32_t)v; val[field].scalar = (uint32_t)((x >> 1) ^ (uint32_t)-(int32_t)(x & 1)); break; }
            case PB_SINT64: val[field].scalar = (v >> 1) ^ (uint64_t)-(int64_t)(v & 1); break;
            case PB_BOOL: val[field].scalar = v != 0; break;
            default: val[field].scalar = v; break; /* int64, uint64, fixed*, float, double: the bits */
        }
    }
    return depth == 0 ? ST_OK : ST_BAD_PROTO;
}

int orc_proto_decode(const uint32_t* fields, uint32_t n_fields, const uint8_t* in, const uint32_t* in_off, uint32_t n,
                     uint8_t* rows, uint64_t rows_cap, uint32_t* row_off, uint32_t* meta) {
    if (n_fields > DEC_MAX_FIELDS) return -2;
    uint64_t pos = 0;
    for (uint32_t i = 0; i < n; i++) {
        row_off[i] = (uint32_t)pos;
        const uint8_t* f = in + in_off[i];
        const size_t fn = in_off[i + 1] - in_off[i];
        dec_value val[DEC_MAX_FIELDS];
        memset(val, 0, sizeof val);
        int st;
        if (fn < 5) st = ST_BAD_LENGTH;
        else if (f[0] == 1) st = ST_COMPRESSED;
        else if (f[0] != 0) st = ST_BAD_LENGTH;
        else {
            const uint32_t L = (uint32_t)f[1] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[3] << 8 | f[4];
            st = (size_t)L != fn - 5 ? ST_BAD_LENGTH : unmarshal(fields, n_fields, f + 5, L, val);
        }
        meta[i] = (uint32_t)st;
        if (st != ST_OK) continue;
        size_t need = 0;
        for (uint32_t k = 0; k < n_fields; k++) need += 4 * (size_t)kind_words(fields[2 * k + 1]) + val[k].str_len;
        need = (need + 3) & ~(size_t)3;
        if (pos + need > rows_cap) return -1;
        uint8_t* o = rows + pos;
        size_t w = 0;
        for (uint32_t k = 0; k < n_fields; k++) {
            const uint32_t t = fields[2 * k + 1];
            const uint64_t s = (t == PB_STRING || t == PB_BYTES) ? (uint64_t)val[k].str_len : val[k].scalar;
            for (int b = 0; b < 4 * kind_words(t); b++) o[w++] = (uint8_t)(s >> (8 * b));
        }
        for (uint32_t k = 0; k < n_fields; k++)
            if (val[k].str_len) { memcpy(o + w, val[k].str, val[k].str_len); w += val[k].str_len; }
        while (w < need) o[w++] = 0;
        pos += need;
    }
    row_off[n] = (uint32_t)pos;
    return 0;
}

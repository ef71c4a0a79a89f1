/*
 * orc_proto_nested.c — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h; parity unpinned: no Go toolchain in this image).
 *
 * proto.Marshal (protobuf-go v1.32.0, go.mod:23) + grpc-go's 5-byte length prefix (v1.60.1, go.mod:11) for proto3 message
 * types with NESTED and REPEATED fields — what the reference's gRPC server would put on the wire for such a response
 * (examples/grpc-server/grpc/hello_grpc.pb.go:73-89 hands whatever message the handler returns to grpc-go).  Rules
 * restated (impl/codec_field.go, codec_gen.go, encode.go):
 *   - a message's fields in ascending field-number order;
 *   - singular scalars: skipped at their zero value (bits all zero); strings must be valid UTF-8;
 *   - singular message: written when set (tag, length, content), an empty message included;
 *   - repeated numeric scalars (everything but string / bytes): packed — one tag with wire type 2, the payload length, the
 *     values; no bytes at all for an empty list;
 *   - repeated string / bytes / message: tag + length + payload for every element, empty ones included.
 * Written recursively (a message is marshalled into its own buffer, then copied behind its length) — the device code walks
 * with an explicit stack and sizes nested messages with a separate pass.  Independent check: tests/test_proto_nested.py
 * builds the same types with python google.protobuf at run time and compares SerializeToString() byte for byte.
 *
 * Description: msgs = n_msgs pairs (first_field, n_fields); fields = n_fields quadruples (number, type, repeated, msg).
 * Rows: include/gofr_b200.h "Row format" as gofr_proto_encode_nested_device documents it.
 */
#include "gofr_oracle.h"
#include "orc_internal.h"

enum { T_DOUBLE = 1, T_FLOAT = 2, T_INT64 = 3, T_UINT64 = 4, T_INT32 = 5, T_FIXED64 = 6, T_FIXED32 = 7, T_BOOL = 8, T_STRING = 9,
       T_MESSAGE = 11, T_BYTES = 12, T_UINT32 = 13, T_ENUM = 14, T_SFIXED32 = 15, T_SFIXED64 = 16, T_SINT32 = 17, T_SINT64 = 18 };
enum { N_OK = 0, N_BAD_UTF8 = 4, N_BAD_ROW = 5 };

typedef struct {
    const uint32_t* msgs;
    const uint32_t* fields;
    uint32_t n_msgs;
    const uint8_t* var; /* cursor in the row's variable part */
    const uint8_t* end;
    int status;
    int depth;
} nctx;

static uint32_t ld32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint64_t ld64(const uint8_t* p) { return (uint64_t)ld32(p) | (uint64_t)ld32(p + 4) << 32; }
static int is64(uint32_t t) { return t == T_DOUBLE || t == T_INT64 || t == T_UINT64 || t == T_FIXED64 || t == T_SFIXED64 || t == T_SINT64; }

static void put_varint(obuf* b, uint64_t v) {
    while (v >= 0x80) { ob_putc(b, (uint8_t)(v | 0x80)); v >>= 7; }
    ob_putc(b, (uint8_t)v);
}

/* words a message type owns in the fixed part of a row (singular messages inline behind a presence word) */
static uint32_t fixed_words(const nctx* c, uint32_t m, int guard) {
    if (guard > 16) return 0;
    uint32_t w = 0;
    const uint32_t first = c->msgs[2 * m], nf = c->msgs[2 * m + 1];
    for (uint32_t k = 0; k < nf; k++) {
        const uint32_t* f = c->fields + 4 * (first + k);
        if (f[2]) w += 1;
        else if (f[1] == T_MESSAGE) w += 1 + fixed_words(c, f[3], guard + 1);
        else w += is64(f[1]) ? 2 : 1;
    }
    return w;
}

static int utf8_ok(const uint8_t* s, size_t n) { /* utf8.Valid */
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
        if (i + need >= n) return 0; /* truncated sequence */
        if (s[i + 1] < lo || s[i + 1] > hi) return 0;
        for (size_t k = 2; k <= need; k++) if (s[i + k] < 0x80 || s[i + k] > 0xBF) return 0;
        i += need + 1;
    }
    return 1;
}

static const uint8_t* take(nctx* c, size_t n) {
    if ((size_t)(c->end - c->var) < n) { c->status = N_BAD_ROW; return NULL; }
    const uint8_t* p = c->var;
    c->var += n;
    return p;
}

/* the payload of one scalar value (what follows the tag; for packed lists: one element) */
static void scalar_payload(obuf* b, uint32_t type, const uint8_t* p) {
    const uint32_t w0 = ld32(p);
    const uint64_t v64 = is64(type) ? ld64(p) : w0;
    switch (type) {
        case T_INT64: case T_UINT64: put_varint(b, v64); break;
        case T_SINT64: put_varint(b, (v64 << 1) ^ (uint64_t)((int64_t)v64 >> 63)); break;
        case T_INT32: case T_ENUM: put_varint(b, (uint64_t)(int64_t)(int32_t)w0); break;
        case T_UINT32: put_varint(b, w0); break;
        case T_SINT32: put_varint(b, (uint32_t)((w0 << 1) ^ (uint32_t)((int32_t)w0 >> 31))); break;
        case T_BOOL: ob_putc(b, w0 ? 1 : 0); break;
        case T_FIXED64: case T_SFIXED64: case T_DOUBLE: for (int k = 0; k < 8; k++) ob_putc(b, (uint8_t)(v64 >> (8 * k))); break;
        default: for (int k = 0; k < 4; k++) ob_putc(b, (uint8_t)(w0 >> (8 * k))); break; /* FIXED32, SFIXED32, FLOAT */
    }
}
static uint32_t scalar_wire(uint32_t t) {
    if (t == T_DOUBLE || t == T_FIXED64 || t == T_SFIXED64) return 1;
    if (t == T_FLOAT || t == T_FIXED32 || t == T_SFIXED32) return 5;
    return 0;
}

static void marshal_msg(nctx* c, uint32_t m, const uint8_t* fixed, obuf* out);

/* tag + length + content of the message of type m whose fixed part is at fx */
static void marshal_field_msg(nctx* c, uint32_t number, uint32_t m, const uint8_t* fx, obuf* out) {
    obuf sub;
    ob_init(&sub);
    marshal_msg(c, m, fx, &sub);
    put_varint(out, (uint64_t)number << 3 | 2);
    put_varint(out, sub.n);
    ob_put(out, sub.p, sub.n);
    ob_free(&sub);
}

static void marshal_msg(nctx* c, uint32_t m, const uint8_t* fixed, obuf* out) {
    if (++c->depth > 16) { c->status = N_BAD_ROW; return; }
    const uint32_t first = c->msgs[2 * m], nf = c->msgs[2 * m + 1];
    for (uint32_t k = 0; k < nf && c->status == N_OK; k++) {
        const uint32_t* f = c->fields + 4 * (first + k);
        const uint32_t number = f[0], type = f[1], repeated = f[2], sub = f[3];
        const uint8_t* p = fixed;
        if (repeated) fixed += 4;
        else if (type == T_MESSAGE) fixed += 4 + 4 * (size_t)fixed_words(c, sub, 0);
        else fixed += is64(type) ? 8 : 4;
        if (!repeated) {
            if (type == T_MESSAGE) {
                if (ld32(p)) marshal_field_msg(c, number, sub, p + 4, out);
            } else if (type == T_STRING || type == T_BYTES) {
                const uint32_t len = ld32(p);
                const uint8_t* s = take(c, len);
                if (!s) break;
                if (type == T_STRING && !utf8_ok(s, len)) { c->status = N_BAD_UTF8; break; }
                if (len) { put_varint(out, (uint64_t)number << 3 | 2); put_varint(out, len); ob_put(out, s, len); }
            } else if (is64(type) ? ld64(p) != 0 : ld32(p) != 0) {
                put_varint(out, (uint64_t)number << 3 | scalar_wire(type));
                scalar_payload(out, type, p);
            }
            continue;
        }
        const uint32_t n = ld32(p);
        if ((size_t)(c->end - c->var) / 4 < n) { c->status = N_BAD_ROW; break; } /* every element owns at least a word */
        if (type == T_MESSAGE) {
            const size_t fb = 4 * (size_t)fixed_words(c, sub, 0);
            for (uint32_t i = 0; i < n && c->status == N_OK; i++) {
                const uint8_t* fx = take(c, fb);
                if (fx) marshal_field_msg(c, number, sub, fx, out);
            }
        } else if (type == T_STRING || type == T_BYTES) {
            for (uint32_t i = 0; i < n && c->status == N_OK; i++) {
                const uint8_t* lp = take(c, 4);
                if (!lp) break;
                const uint32_t len = ld32(lp);
                const uint8_t* s = take(c, len);
                if (!s) break;
                if (type == T_STRING && !utf8_ok(s, len)) { c->status = N_BAD_UTF8; break; }
                put_varint(out, (uint64_t)number << 3 | 2);
                put_varint(out, len);
                ob_put(out, s, len);
            }
        } else if (n) {
            obuf pk;
            ob_init(&pk);
            const size_t eb = is64(type) ? 8 : 4;
            for (uint32_t i = 0; i < n; i++) {
                const uint8_t* e = take(c, eb);
                if (!e) break;
                scalar_payload(&pk, type, e);
            }
            if (c->status == N_OK) { put_varint(out, (uint64_t)number << 3 | 2); put_varint(out, pk.n); ob_put(out, pk.p, pk.n); }
            ob_free(&pk);
        }
    }
    c->depth--;
}

int orc_proto_encode_nested(const uint32_t* msgs, uint32_t n_msgs, const uint32_t* fields, uint32_t n_fields, uint32_t root,
                            const uint8_t* rows, const uint32_t* row_off, uint32_t n, uint8_t* out, uint64_t out_cap,
                            uint32_t* out_off, uint32_t* meta) {
    (void)n_fields;
    uint64_t pos = 0;
    obuf b;
    ob_init(&b);
    for (uint32_t i = 0; i < n; i++) {
        out_off[i] = (uint32_t)pos;
        b.n = 0;
        nctx c = {msgs, fields, n_msgs, NULL, NULL, N_OK, 0};
        const uint8_t* row = rows + row_off[i];
        const size_t rn = row_off[i + 1] - row_off[i], fb = 4 * (size_t)fixed_words(&c, root, 0);
        if (rn < fb) c.status = N_BAD_ROW;
        else {
            c.var = row + fb;
            c.end = row + rn;
            marshal_msg(&c, root, row, &b);
        }
        meta[i] = (uint32_t)c.status;
        if (c.status != N_OK) continue;
        if (pos + 5 + b.n > out_cap) { ob_free(&b); return -1; }
        uint8_t* o = out + pos;
        o[0] = 0;
        o[1] = (uint8_t)(b.n >> 24); o[2] = (uint8_t)(b.n >> 16); o[3] = (uint8_t)(b.n >> 8); o[4] = (uint8_t)b.n;
        if (b.n) memcpy(o + 5, b.p, b.n);
        pos += 5 + b.n;
    }
    out_off[n] = (uint32_t)pos;
    ob_free(&b);
    return 0;
}

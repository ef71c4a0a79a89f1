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

/* ---------------------------------------------------------------------------------------------------------------
 * The other direction (gofr_proto_decode_nested_device): 5-byte header check + proto.Unmarshal into a message type with
 * nested / repeated fields, the values out as a row of the layout above.  protobuf-go v1.32.0 rules restated: unknown
 * fields and balanced groups skipped; a known number with a foreign wire type is unknown, except that repeated numeric
 * scalars take both the packed and the unpacked form; singular scalars / strings: last occurrence wins, every string
 * occurrence must be valid UTF-8; repeated string / bytes / message: one element per occurrence in wire order.  A singular
 * message field that occurs more than once would be merged by protobuf-go: reported as 6 (DEFER), like the device code.
 * Status precedence: 3 malformed > 4 invalid UTF-8 > 6 DEFER.  Written list-first (collect every field's occurrences of a
 * message, then lay the row out, recursing per nested message) — the device code scans the bytes once per field instead.
 * --------------------------------------------------------------------------------------------------------------- */
enum { D_COMPRESSED = 1, D_BAD_LENGTH = 2, D_BAD_PROTO = 3, D_BAD_UTF8 = 4, D_DEFER = 6 };

typedef struct { uint32_t wt; uint64_t v; const uint8_t* p; size_t n; } occ;
typedef struct { occ* o; size_t n, cap; } occ_list;
typedef struct { const uint32_t* msgs; const uint32_t* fields; int bad_utf8, defer; } dctx;

static int d_varint(const uint8_t* p, size_t n, uint64_t* v) { /* protowire.ConsumeVarint: <= 10 bytes, the tenth <= 1 */
    uint64_t x = 0;
    for (size_t i = 0; i < n && i < 10; i++) {
        const uint8_t b = p[i];
        if (i == 9 && b > 1) return -1;
        x |= (uint64_t)(b & 0x7F) << (7 * i);
        if (b < 0x80) { *v = x; return (int)i + 1; }
    }
    return -1;
}
static void occ_add(occ_list* l, occ o) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 4; l->o = (occ*)realloc(l->o, l->cap * sizeof(occ)); }
    l->o[l->n++] = o;
}
static uint32_t own_wire(const uint32_t* f) { return f[1] == T_MESSAGE || f[1] == T_STRING || f[1] == T_BYTES ? 2 : scalar_wire(f[1]); }

static void scalar_words(uint32_t type, uint32_t wt, uint64_t v, obuf* out) {
    uint32_t w0 = (uint32_t)v, w1 = (uint32_t)(v >> 32);
    if (wt == 0) {
        if (type == T_SINT64) { const uint64_t z = (v >> 1) ^ (uint64_t)-(int64_t)(v & 1); w0 = (uint32_t)z; w1 = (uint32_t)(z >> 32); }
        else if (type == T_SINT32) { const uint32_t x = (uint32_t)v; w0 = (x >> 1) ^ (uint32_t)-(int32_t)(x & 1); }
        else if (type == T_BOOL) w0 = v != 0;
    }
    ob_put(out, &w0, 4);
    if (is64(type)) ob_put(out, &w1, 4);
}

/* one message of type m in [p, p + n): its fixed part appended to `fixed`, its variable part to `var`.  0 or D_BAD_PROTO. */
static int dec_msg(dctx* c, uint32_t m, const uint8_t* p, size_t n, obuf* fixed, obuf* var, int depth) {
    if (depth > 16) return D_BAD_PROTO;
    const uint32_t first = c->msgs[2 * m], nf = c->msgs[2 * m + 1];
    occ_list* L = (occ_list*)calloc(nf, sizeof(occ_list));
    int rc = 0;
    /* ---- split into fields, wire order ---- */
    size_t i = 0;
    uint32_t gstack[16];
    int g = 0;
    while (i < n && !rc) {
        uint64_t tag, v = 0;
        int k = d_varint(p + i, n - i, &tag);
        if (k < 0) { rc = D_BAD_PROTO; break; }
        i += (size_t)k;
        const uint64_t num = tag >> 3;
        const uint32_t wt = (uint32_t)(tag & 7);
        if (num == 0 || num > 0x1FFFFFFFull) { rc = D_BAD_PROTO; break; }
        occ o = {wt, 0, NULL, 0};
        if (wt == 0) { k = d_varint(p + i, n - i, &v); if (k < 0) { rc = D_BAD_PROTO; break; } i += (size_t)k; }
        else if (wt == 1) { if (n - i < 8) { rc = D_BAD_PROTO; break; } v = ld64(p + i); i += 8; }
        else if (wt == 5) { if (n - i < 4) { rc = D_BAD_PROTO; break; } v = ld32(p + i); i += 4; }
        else if (wt == 2) {
            k = d_varint(p + i, n - i, &v);
            if (k < 0) { rc = D_BAD_PROTO; break; }
            i += (size_t)k;
            if (v > n - i) { rc = D_BAD_PROTO; break; }
            o.p = p + i; o.n = (size_t)v;
            i += (size_t)v;
        } else if (wt == 3) { if (g == 16) { rc = D_BAD_PROTO; break; } gstack[g++] = (uint32_t)num; continue; }
        else if (wt == 4) { if (g == 0 || gstack[g - 1] != (uint32_t)num) { rc = D_BAD_PROTO; break; } g--; continue; }
        else { rc = D_BAD_PROTO; break; }
        o.v = v;
        if (g) continue; /* inside an unknown group */
        for (uint32_t q = 0; q < nf; q++) {
            const uint32_t* f = c->fields + 4 * (first + q);
            if (f[0] != (uint32_t)num) continue;
            const uint32_t own = own_wire(f);
            const int packed_ok = f[2] && own != 2 && wt == 2;
            if (wt == own || packed_ok) occ_add(&L[q], o);
            break;
        }
    }
    if (!rc && g) rc = D_BAD_PROTO;
    /* ---- lay the row out, field order ---- */
    for (uint32_t q = 0; q < nf && !rc; q++) {
        const uint32_t* f = c->fields + 4 * (first + q);
        const uint32_t type = f[1], repeated = f[2], sub = f[3];
        const occ_list* l = &L[q];
        if (!repeated) {
            if (type == T_MESSAGE) {
                const uint32_t present = l->n ? 1 : 0;
                ob_put(fixed, &present, 4);
                if (l->n > 1) c->defer = 1;
                if (!l->n) { nctx t = {c->msgs, c->fields, 0, NULL, NULL, 0, 0}; for (uint32_t k = 0; k < 4 * fixed_words(&t, sub, 0); k++) ob_putc(fixed, 0); }
                for (size_t j = 0; j < l->n && !rc; j++) { /* every occurrence is parsed (validation); the first one is kept */
                    obuf fx, vr;
                    ob_init(&fx); ob_init(&vr);
                    rc = dec_msg(c, sub, l->o[j].p, l->o[j].n, &fx, &vr, depth + 1);
                    if (!rc && j == 0) { ob_put(fixed, fx.p, fx.n); ob_put(var, vr.p, vr.n); }
                    ob_free(&fx); ob_free(&vr);
                }
            } else if (type == T_STRING || type == T_BYTES) {
                for (size_t j = 0; j < l->n; j++)
                    if (type == T_STRING && !utf8_ok(l->o[j].p, l->o[j].n)) c->bad_utf8 = 1;
                const uint32_t len = l->n ? (uint32_t)l->o[l->n - 1].n : 0;
                ob_put(fixed, &len, 4);
                if (len) ob_put(var, l->o[l->n - 1].p, len);
            } else {
                if (l->n) scalar_words(type, l->o[l->n - 1].wt, l->o[l->n - 1].v, fixed);
                else { const uint64_t z = 0; ob_put(fixed, &z, is64(type) ? 8 : 4); }
            }
            continue;
        }
        if (type == T_MESSAGE) {
            const uint32_t cnt = (uint32_t)l->n;
            ob_put(fixed, &cnt, 4);
            for (size_t j = 0; j < l->n && !rc; j++) {
                obuf fx, vr;
                ob_init(&fx); ob_init(&vr);
                rc = dec_msg(c, sub, l->o[j].p, l->o[j].n, &fx, &vr, depth + 1);
                if (!rc) { ob_put(var, fx.p, fx.n); ob_put(var, vr.p, vr.n); }
                ob_free(&fx); ob_free(&vr);
            }
        } else if (type == T_STRING || type == T_BYTES) {
            const uint32_t cnt = (uint32_t)l->n;
            ob_put(fixed, &cnt, 4);
            for (size_t j = 0; j < l->n; j++) {
                if (type == T_STRING && !utf8_ok(l->o[j].p, l->o[j].n)) c->bad_utf8 = 1;
                const uint32_t len = (uint32_t)l->o[j].n;
                ob_put(var, &len, 4);
                ob_put(var, l->o[j].p, len);
            }
        } else {
            obuf el;
            ob_init(&el);
            uint32_t cnt = 0;
            const uint32_t ewt = scalar_wire(type);
            for (size_t j = 0; j < l->n && !rc; j++) {
                if (l->o[j].wt != 2) { scalar_words(type, l->o[j].wt, l->o[j].v, &el); cnt++; continue; }
                const uint8_t* q2 = l->o[j].p;
                size_t left = l->o[j].n;
                while (left && !rc) { /* the packed payload: elements back to back */
                    uint64_t v = 0;
                    if (ewt == 0) { int k = d_varint(q2, left, &v); if (k < 0) { rc = D_BAD_PROTO; break; } q2 += k; left -= (size_t)k; }
                    else if (ewt == 1) { if (left < 8) { rc = D_BAD_PROTO; break; } v = ld64(q2); q2 += 8; left -= 8; }
                    else { if (left < 4) { rc = D_BAD_PROTO; break; } v = ld32(q2); q2 += 4; left -= 4; }
                    scalar_words(type, ewt, v, &el);
                    cnt++;
                }
            }
            ob_put(fixed, &cnt, 4);
            ob_put(var, el.p, el.n);
            ob_free(&el);
        }
    }
    for (uint32_t q = 0; q < nf; q++) free(L[q].o);
    free(L);
    return rc;
}

int orc_proto_decode_nested(const uint32_t* msgs, uint32_t n_msgs, const uint32_t* fields, uint32_t n_fields, uint32_t root,
                            const uint8_t* in, const uint32_t* in_off, uint32_t n, uint8_t* rows, uint64_t rows_cap,
                            uint32_t* row_off, uint32_t* meta) {
    (void)n_msgs; (void)n_fields;
    uint64_t pos = 0;
    for (uint32_t i = 0; i < n; i++) {
        row_off[i] = (uint32_t)pos;
        const uint8_t* f = in + in_off[i];
        const size_t fn = in_off[i + 1] - in_off[i];
        int st = 0;
        if (fn < 5) st = D_BAD_LENGTH;
        else if (f[0] == 1) st = D_COMPRESSED;
        else if (f[0] != 0) st = D_BAD_LENGTH;
        else if ((((size_t)f[1] << 24) | ((size_t)f[2] << 16) | ((size_t)f[3] << 8) | f[4]) != fn - 5) st = D_BAD_LENGTH;
        obuf fx, vr;
        ob_init(&fx); ob_init(&vr);
        if (!st) {
            dctx c = {msgs, fields, 0, 0};
            st = dec_msg(&c, root, f + 5, fn - 5, &fx, &vr, 0);
            if (!st) st = c.bad_utf8 ? D_BAD_UTF8 : c.defer ? D_DEFER : 0;
        }
        meta[i] = (uint32_t)st;
        if (!st) {
            const size_t len = (fx.n + vr.n + 3) & ~(size_t)3;
            if (pos + len > rows_cap) { ob_free(&fx); ob_free(&vr); return -1; }
            memset(rows + pos, 0, len);
            memcpy(rows + pos, fx.p, fx.n);
            if (vr.n) memcpy(rows + pos + fx.n, vr.p, vr.n);
            pos += len;
        }
        ob_free(&fx); ob_free(&vr);
    }
    row_off[n] = (uint32_t)pos;
    return 0;
}

/*
 * orc_grpc.c — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h).
 *
 * Config 5: the unary Hello RPC of examples/grpc-server, message level only (HTTP/2 + HPACK live in grpc-go and are
 * out of scope).  Restates:
 *   _Hello_SayHello_Handler   examples/grpc-server/grpc/hello_grpc.pb.go:73-89   dec(in): strip the 5-byte gRPC
 *                                                                                 length-prefixed-message header,
 *                                                                                 proto.Unmarshal HelloRequest
 *   Server.SayHello           examples/grpc-server/grpc/server.go:12-21          "Hello " + (name or "World") + "!"
 *   HelloRequest/Response     examples/grpc-server/grpc/hello.proto:4-10         field 1, string
 * Wire arithmetic: protobuf-go v1.32.0 (protowire varint/tag/length rules, proto3 string UTF-8 validation, unknown
 * fields skipped incl. groups) and grpc-go v1.60.1 (1-byte compressed flag + big-endian u32 length).
 * Independent check: tests cross-check against python google.protobuf using the descriptor bytes restated from
 * hello.proto.
 */
#include "gofr_oracle.h"
#include "orc_internal.h"

#include <pthread.h>

enum { GRPC_OK = 0, GRPC_COMPRESSED = 1, GRPC_BAD_LENGTH = 2, GRPC_BAD_PROTO = 3, GRPC_BAD_UTF8 = 4 };

/* protowire.ConsumeVarint: ≤10 bytes, 10th byte ≤ 1.  returns bytes consumed or -1 */
static int consume_varint(const uint8_t* p, size_t n, uint64_t* v) {
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

static int utf8_valid(const uint8_t* s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t b0 = s[i];
        if (b0 < 0x80) { i++; continue; }
        if (b0 < 0xC2 || b0 > 0xF4) return 0;
        int need = b0 < 0xE0 ? 1 : b0 < 0xF0 ? 2 : 3;
        if (i + (size_t)need >= n) return 0; /* truncated sequence */
        uint8_t lo = 0x80, hi = 0xBF;
        if (b0 == 0xE0) lo = 0xA0;
        if (b0 == 0xED) hi = 0x9F;
        if (b0 == 0xF0) lo = 0x90;
        if (b0 == 0xF4) hi = 0x8F;
        if (s[i + 1] < lo || s[i + 1] > hi) return 0;
        for (int k = 2; k <= need; k++)
            if ((s[i + (size_t)k] & 0xC0) != 0x80) return 0;
        i += (size_t)need + 1;
    }
    return 1;
}

#define MAX_GROUP_DEPTH 16 /* upstream recursion limit is far deeper; deeper nesting is reported as BAD_PROTO here */

/* proto.Unmarshal(HelloRequest): returns GRPC_* and the name span */
static int parse_hello_request(const uint8_t* p, size_t n, const uint8_t** name, size_t* name_len) {
    *name = NULL;
    *name_len = 0;
    size_t i = 0;
    uint32_t group_stack[MAX_GROUP_DEPTH];
    int depth = 0;
    while (i < n) {
        uint64_t tag;
        int k = consume_varint(p + i, n - i, &tag);
        if (k < 0) return GRPC_BAD_PROTO;
        i += (size_t)k;
        uint64_t num = tag >> 3;
        int wt = (int)(tag & 7);
        if (num == 0 || num > 0x1FFFFFFF) return GRPC_BAD_PROTO; /* protowire: invalid field number */
        uint64_t v;
        switch (wt) {
            case 0:
                k = consume_varint(p + i, n - i, &v);
                if (k < 0) return GRPC_BAD_PROTO;
                i += (size_t)k;
                break;
            case 1:
                if (n - i < 8) return GRPC_BAD_PROTO;
                i += 8;
                break;
            case 5:
                if (n - i < 4) return GRPC_BAD_PROTO;
                i += 4;
                break;
            case 2:
                k = consume_varint(p + i, n - i, &v);
                if (k < 0) return GRPC_BAD_PROTO;
                i += (size_t)k;
                if (v > n - i) return GRPC_BAD_PROTO;
                if (num == 1 && depth == 0) {
                    if (!utf8_valid(p + i, (size_t)v)) return GRPC_BAD_UTF8; /* proto3 string */
                    *name = p + i; /* last occurrence wins */
                    *name_len = (size_t)v;
                }
                i += (size_t)v;
                break;
            case 3: /* start group (only ever unknown here): skip to the matching end group */
                if (depth == MAX_GROUP_DEPTH) return GRPC_BAD_PROTO;
                group_stack[depth++] = (uint32_t)num;
                break;
            case 4:
                if (depth == 0 || group_stack[depth - 1] != (uint32_t)num) return GRPC_BAD_PROTO;
                depth--;
                break;
            default: return GRPC_BAD_PROTO;
        }
    }
    if (depth != 0) return GRPC_BAD_PROTO;
    return GRPC_OK;
}

static size_t put_varint(uint8_t* out, uint64_t v) {
    size_t k = 0;
    while (v >= 0x80) { out[k++] = (uint8_t)(v | 0x80); v >>= 7; }
    out[k++] = (uint8_t)v;
    return k;
}

typedef struct {
    const uint8_t* in;
    const uint32_t* in_off;
    uint32_t lo, hi;
    uint8_t* out;
    uint64_t base, cap, end;
    uint32_t* out_off;
    uint32_t* meta;
    int rc;
} gshard;

static void* gshard_run(void* arg) {
    gshard* s = (gshard*)arg;
    uint64_t pos = s->base;
    for (uint32_t i = s->lo; i < s->hi; i++) {
        const uint8_t* f = s->in + s->in_off[i];
        size_t fn = s->in_off[i + 1] - s->in_off[i];
        s->out_off[i] = (uint32_t)pos;
        int st = GRPC_OK;
        const uint8_t* name = NULL;
        size_t nl = 0;
        if (fn < 5) st = GRPC_BAD_LENGTH;
        else if (f[0] == 1) st = GRPC_COMPRESSED; /* no compressor registered (pkg/gofr/grpc.go:23-26) */
        else if (f[0] != 0) st = GRPC_BAD_LENGTH;
        else {
            uint32_t L = (uint32_t)f[1] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[3] << 8 | f[4];
            if ((size_t)L != fn - 5) st = GRPC_BAD_LENGTH;
            else st = parse_hello_request(f + 5, L, &name, &nl);
        }
        s->meta[i] = (uint32_t)st;
        if (st != GRPC_OK) continue;
        if (nl == 0) { name = (const uint8_t*)"World"; nl = 5; } /* server.go:14-17 */
        size_t ml = 6 + nl + 1;                                   /* fmt.Sprintf("Hello %s!", name) */
        uint8_t vb[10];
        size_t vn = put_varint(vb, ml);
        size_t total = 5 + 1 + vn + ml;
        if (pos + total > s->base + s->cap) { s->rc = -1; break; }
        uint8_t* o = s->out + pos;
        uint32_t plen = (uint32_t)(1 + vn + ml);
        o[0] = 0; o[1] = (uint8_t)(plen >> 24); o[2] = (uint8_t)(plen >> 16); o[3] = (uint8_t)(plen >> 8); o[4] = (uint8_t)plen;
        o[5] = 0x0A; /* field 1, wire type 2 */
        memcpy(o + 6, vb, vn);
        memcpy(o + 6 + vn, "Hello ", 6);
        memcpy(o + 12 + vn, name, nl);
        o[12 + vn + nl] = '!';
        pos += total;
    }
    s->end = pos;
    return NULL;
}

int orc_grpc_hello(const uint8_t* in, const uint32_t* in_off, uint32_t n, uint8_t* out, uint64_t out_cap,
                   uint32_t* out_off, uint32_t* meta, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n && n > 0) nthreads = (int)n;
    gshard* sh = (gshard*)calloc((size_t)nthreads, sizeof(gshard));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    uint64_t slice = out_cap / (uint64_t)nthreads;
    for (int k = 0; k < nthreads; k++) {
        sh[k] = (gshard){in, in_off, (uint32_t)((uint64_t)n * k / nthreads), (uint32_t)((uint64_t)n * (k + 1) / nthreads),
                         out, slice * (uint64_t)k, slice, 0, out_off, meta, 0};
        if (nthreads == 1) gshard_run(&sh[k]);
        else pthread_create(&th[k], NULL, gshard_run, &sh[k]);
    }
    int rc = 0;
    for (int k = 0; k < nthreads; k++) {
        if (nthreads > 1) pthread_join(th[k], NULL);
        if (sh[k].rc) rc = -1;
    }
    out_off[n] = (uint32_t)sh[nthreads - 1].end;
    free(sh);
    free(th);
    return rc;
}

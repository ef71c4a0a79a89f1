/*
 * gofr_b200.h — C ABI of the B200-native GoFr request hot path.
 *
 * This is the drop-in boundary.  The reference (JigarJoshi04/gofr) has no FFI of its own: its seams
 * are in-process Go interfaces.  Each entry point below names the reference interface it stands in for, so a Go shim
 * (go/gofrb200/) binds them one-to-one over cgo:
 *
 *   reference interface (file:line)                              replaced by
 *   ------------------------------------------------------------ ---------------------------------------------
 *   http.NewRouter                pkg/gofr/http/router.go:17-28    gofr_table_create
 *   Router.Add / App.add          pkg/gofr/http/router.go:30-33,   gofr_table_add_route
 *                                 pkg/gofr/gofr.go:171-177
 *   App.Run default routes        pkg/gofr/gofr.go:102-107         gofr_table_add_default_routes
 *   (route table frozen at Run)   pkg/gofr/gofr.go:90-126          gofr_table_seal / _serialize / _deserialize
 *   mux.Router.ServeHTTP → middleware chain → handler.ServeHTTP → Responder.Respond → net/http framing
 *                                 pkg/gofr/http/router.go:14,
 *                                 pkg/gofr/http/middleware/{tracer,logger,cors}.go,
 *                                 pkg/gofr/handler.go:32-36,
 *                                 pkg/gofr/http/responder.go:19-41 gofr_serve_device[_slots] / gofr_batch_submit+wait
 *   Request.Param / Bind          pkg/gofr/http/request.go:28-47   (fused into the serve kernel per handler kind)
 *   _Hello_SayHello_Handler       examples/grpc-server/grpc/
 *                                 hello_grpc.pb.go:73-89           gofr_grpc_hello_device
 *   dec(in) / proto.Marshal of any flat proto3 message type (the codec grpc-go calls around a unary handler)
 *                                 hello_grpc.pb.go:73-89           gofr_proto_decode_device / gofr_proto_encode_device
 *   Router.Match + mux.Vars (routing only, closures on the host)
 *                                 pkg/gofr/http/router.go:14,30-33 gofr_route_device / gofr_batch_route
 *   router.ServeHTTP per request under net/http's conn goroutine
 *                                 pkg/gofr/httpServer.go:29-33     gofr_frontend_serve (batches the concurrent callers)
 *   net/http readRequest + url.ParseRequestURI (stdlib, reached from pkg/gofr/httpServer.go:29-33)
 *                                                                  gofr_http_parse_device
 *   middleware.Logging's RequestLog line → logger.Log
 *                                 pkg/gofr/http/middleware/logger.go:24-33,41-84,
 *                                 pkg/gofr/logging/logger.go:37-74 gofr_requestlog_device
 *
 * Conventions (mirroring the reference's "log and continue", responder.go:29,40): plain C types only, integer status
 * codes, never abort; per-request outcomes are reported in the response meta column; the table is immutable after
 * seal (the reference never mutates routes after Run()); no callbacks into the caller from CUDA threads (cgo-safe).
 * Values that are non-deterministic in the reference are INPUTS here: the 16-byte OTel trace id per request
 * (middleware/logger.go:46-47) and the IMF-fixdate `Date` string per batch (net/http).
 * Device-resident INPUT buffers of variable-length bytes (d_arena, d_in of frames, d_rows, d_raw, the strings of log
 * records): 16-byte aligned, and the allocation extends at least 16 bytes past the last byte an offset names — tiles are
 * staged in 16-byte units and the copy loops read whole aligned words; the bytes there are never looked at.  The host-buffer
 * entry points (gofr_batch_*) stage their inputs themselves and ask nothing of the kind.
 */
#ifndef GOFR_B200_H
#define GOFR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOFR_ABI_VERSION 1

/* ---- status codes (returned by every entry point) ---- */
enum {
    GOFR_OK = 0,
    GOFR_ERR_INVALID = 1,      /* bad argument */
    GOFR_ERR_UNSUPPORTED = 2,  /* e.g. a {var:regexp} outside the supported class, a float field */
    GOFR_ERR_NOMEM = 3,
    GOFR_ERR_CUDA = 4,         /* a CUDA runtime call failed; gofr_last_error() has the text */
    GOFR_ERR_SEALED = 5,       /* mutation after seal */
    GOFR_ERR_NOT_SEALED = 6,
    GOFR_ERR_CAPACITY = 7,     /* batch / arena / output larger than the engine was created for */
    GOFR_ERR_NO_DEVICE = 8     /* no CUDA device: the product path has no CPU fallback, by design */
};

/* ---- request methods.  mux's methodMatcher compares exact upper-case strings; the shim maps the exact strings
 *      below and sends everything else (including lower-case spellings) as GOFR_M_OTHER, which matches no
 *      method-restricted route. ---- */
enum {
    GOFR_M_GET = 0, GOFR_M_HEAD = 1, GOFR_M_POST = 2, GOFR_M_PUT = 3, GOFR_M_PATCH = 4, GOFR_M_DELETE = 5,
    GOFR_M_CONNECT = 6, GOFR_M_OPTIONS = 7, GOFR_M_TRACE = 8, GOFR_M_OTHER = 15,
    GOFR_M_ANY = 255 /* registration only: no method matcher (PathPrefix("/") catch-all, gofr.go:104) */
};

/* ---- framing modes ---- */
enum {
    GOFR_FRAME_WIRE = 0,     /* what net/http 1.21 puts on the socket: Responder sets Content-type AFTER WriteHeader
                                (responder.go:21 vs :39) so the type is sniffed: text/plain; charset=utf-8 */
    GOFR_FRAME_INTENDED = 1, /* what the reference's tests observe through the live recorder map
                                (responder_test.go:28, gofr_test.go:104): Content-Type: application/json */
    GOFR_FRAME_BODY = 2      /* response body only (unambiguous in both modes) */
};

/* ---- declarative handler kinds.  User handlers are arbitrary Go closures (handler.go:12) and cannot run on the
 *      GPU; the kinds below are the closures the reference's examples/tests use, expressed as data so the whole
 *      request can be served by one fused kernel.  GOFR_H_HOST is the general case: the GPU routes, the host runs
 *      the closure, the GPU encodes (two launches). ---- */
enum {
    GOFR_H_HOST = 0,          /* arbitrary closure on the host; serve emits nothing and reports route_id */
    GOFR_H_STATIC_STRING = 1, /* return s0, nil                       gofr_test.go:62-64 */
    GOFR_H_STATIC_ERROR = 2,  /* return nil, errors.New(s0)           examples/http-server/main.go:41-43 */
    GOFR_H_NIL = 3,           /* return nil, nil                      handler_test.go:29 */
    GOFR_H_PARAM_FORMAT = 4,  /* v := c.Param(s0); if v == "" { v = s1 }; return s2 + v + s3, nil
                                 examples/http-server/main.go:31-39, gofr_test.go:80-82 */
    GOFR_H_ROW = 5,           /* return <struct of schema>, nil; field values arrive in the request's data section */
    GOFR_H_BIND_ECHO = 6,     /* var v T; if err := c.Bind(&v); err != nil { return nil, err }; return v, nil
                                 T: a flat struct of int / int32 / int64 / float64 / bool / string fields.  A float64 member
                                 gets strconv.ParseFloat's correctly rounded value through ParseFloat's own fast steps (the
                                 exact cases, then Eisel-Lemire; sure overflows → the UnmarshalTypeError Go reports, sure
                                 underflows → 0); a request with a literal Go needs its big-decimal slow path for (an exact
                                 half-way case, a subnormal, 1e308-ish) comes back with status 0 ("run on the host"), like a
                                 body nested deeper than 64 levels */
    GOFR_H_HEALTH = 7,        /* healthHandler with no datasources: map{} handler.go:38-40, container.go:26-38 */
    GOFR_H_MISSING_FILE = 8,  /* catchAllHandler: return nil, http.ErrMissingFile      handler.go:51-53 */
    GOFR_H_FILE = 9,          /* return response.File{Content: blob, ContentType: s0}   handler.go:42-49 */
    GOFR_H_PANIC = 10,        /* handler panics; middleware.panicRecovery answers       middleware/logger.go:91-114 */
    GOFR_H_PATHPARAM_FORMAT = 11, /* v := c.PathParam(s0); return s2 + v + s3, nil         pkg/gofr/http/request.go:36-38 */
    GOFR_H_RESULT = 12        /* stage 2 of the split API: the closure already ran on the host (after gofr_route_device);
                                 its (data, err) arrives in the request's data section and Responder.Respond runs here
                                 (pkg/gofr/http/responder.go:19-62).  Data section: one LE u32 outcome word, then
                                   GOFR_RESULT_DATA     a row of the route's schema (as for GOFR_H_ROW)      → 200 {"data":{…}}
                                   GOFR_RESULT_ERROR    u32 length + err.Error() bytes                      → 500 {"error":{"message":…}}
                                   GOFR_RESULT_NIL      nothing                                             → 200 {}
                                   GOFR_RESULT_MISSING  u32 length + message; errors.Is(err, http.ErrMissingFile) → 404
                                   GOFR_RESULT_BOTH     (data, err) both non-nil: one row whose first field is the message
                                                        string followed by the schema's fields (length word + fixed
                                                        words, then message bytes + string bytes) → 500 {"error":…,"data":…}
                                   GOFR_RESULT_STRING   data is a Go string (what most of the reference's example handlers
                                                        return, examples/http-server/main.go:29-41): u32 length + bytes
                                                        → 200 {"data":"…"}  (the route needs no schema for this outcome)
                                 The closure returned a response.Raw (pkg/gofr/http/response/raw.go:3-5): Respond encodes
                                 Raw.Data bare, without the {"error":…,"data":…} envelope (responder.go:24-26); the error, if
                                 any, only picks the status code (written at responder.go:21 before the type switch):
                                   GOFR_RESULT_RAW_DATA    a row of the route's schema                      → {…}\n
                                   GOFR_RESULT_RAW_STRING  u32 length + bytes                               → "…"\n
                                   GOFR_RESULT_RAW_NIL     nothing (Raw{}: Data is a nil interface)         → null\n
                                 with the error in bits 8..15 of the outcome word: GOFR_RESULT_RAW_OK (err == nil, 200),
                                 GOFR_RESULT_RAW_ERR (500), GOFR_RESULT_RAW_MISSING (errors.Is(err, http.ErrMissingFile), 404);
                                 e.g. GOFR_RESULT_RAW_STRING | GOFR_RESULT_RAW_ERR << 8. */
};
enum { GOFR_RESULT_DATA = 0, GOFR_RESULT_ERROR = 1, GOFR_RESULT_NIL = 2, GOFR_RESULT_MISSING = 3, GOFR_RESULT_BOTH = 4,
       GOFR_RESULT_STRING = 5, GOFR_RESULT_RAW_DATA = 6, GOFR_RESULT_RAW_STRING = 7, GOFR_RESULT_RAW_NIL = 8 };
enum { GOFR_RESULT_RAW_OK = 0, GOFR_RESULT_RAW_ERR = 1, GOFR_RESULT_RAW_MISSING = 2 };

/* ---- struct field kinds for response / Bind schemas ---- */
enum {
    GOFR_F_INT64 = 1,
    GOFR_F_INT32 = 2,
    GOFR_F_BOOL = 3,
    GOFR_F_STRING = 4,
    GOFR_F_INT = 5,    /* Go int (64-bit); only the type name in Bind error text differs from INT64 */
    GOFR_F_FLOAT64 = 6, /* encoding/json floatEncoder: shortest digits that round-trip, 'e' form below 1e-6 and from 1e21,
                           "e-0X" written "e-X"; NaN / ±Inf make Encode fail: the response keeps its status and headers
                           and has no body (responder.go:40 drops the error).  Bind schemas: see GOFR_H_BIND_ECHO */
    GOFR_F_STRUCT = 7,  /* a struct value of the schema `elem_schema`                            response schemas only */
    GOFR_F_UINT64 = 8,  /* uint64 / uint / uintptr: two words, unsigned decimal (narrower unsigned types fit INT64, int8 /
                           int16 fit INT32: the text is the same)                               response schemas only */
    GOFR_F_BYTES = 9,   /* []byte: encoding/json writes base64.StdEncoding with padding in quotes, null when nil.  Row:
                           one fixed word = the byte length (GOFR_NIL_COUNT: nil), the bytes in the variable part;
                           as an element E([]byte) = u32 length (GOFR_NIL_COUNT: nil) + bytes     response schemas only */
    GOFR_F_FLOAT32 = 10, /* float32: one word (IEEE-754 bits), the shortest digits that identify the FLOAT32
                           (strconv.AppendFloat(…, -1, 32)), same format rules and NaN / Inf behaviour as FLOAT64
                                                                                                response schemas only */
    GOFR_F_TIME = 11    /* time.Time: Time.MarshalJSON = the quoted RFC 3339 text with nanoseconds, trailing zeros of the
                           fraction trimmed, "Z" for offset 0 else ±hh:mm ("2006-01-02T15:04:05.999999999Z07:00").  Four
                           words: Unix seconds (lo, hi; the zero Time is -62135596800), nanoseconds (0 .. 999999999), zone
                           offset in seconds east of UTC (int32).  A year outside [0, 9999] or a zone hour outside [0, 23]
                           makes MarshalJSON — and with it Encode — fail: the response keeps its status and headers and
                           has no body, as for NaN (a nanoseconds word >= 1e9, which no Time holds, is answered the same way).  Never "empty" for omitempty (a struct).     response schemas only */
};
/* What the field holds of its kind T (response schemas only; Bind schemas take GOFR_C_VALUE of kinds 1..6):
 *   GOFR_C_VALUE  T          GOFR_C_PTR  *T (nil → null)      GOFR_C_SLICE  []T (nil → null, empty → [])
 *   GOFR_C_MAP    map[string]T, T not a struct (nil → null; keys sorted bytewise like encoding/json does)
 * One container level per field; deeper types nest through GOFR_F_STRUCT (a struct field may again be a pointer, slice
 * or map).  interface{} values, [][]T (other than [][]byte) and maps of structs are not modelled: GOFR_ERR_UNSUPPORTED. */
enum { GOFR_C_VALUE = 0, GOFR_C_PTR = 1, GOFR_C_SLICE = 2, GOFR_C_MAP = 3,
       GOFR_C_SLICE_PTR = 4 /* []*T (what ORMs hand back): like []T, each element preceded by a presence word (0: nil → null) */ };
#define GOFR_FIELD_BARE 0x01u /* the schema has this ONE field and stands for the field's own type: the handler returns a
                                 []T, map[string]T, *T or float64 rather than a struct; no {"name":…} around the value */
#define GOFR_NIL_COUNT 0xFFFFFFFFu /* count word of a nil slice / map */

typedef struct gofr_field_desc {
    const char* go_name;   /* Go field name (case-insensitive Bind fallback, error text) */
    const char* json_name; /* `json:"..."` name; NULL/"" = go_name */
    uint8_t kind;          /* GOFR_F_* */
    uint8_t omitempty;     /* isEmptyValue: false, 0, 0.0 (either sign), "", nil pointer, nil or empty slice / map;
                              a struct value is never empty */
    uint8_t container;     /* GOFR_C_* */
    uint8_t flags;         /* GOFR_FIELD_* */
    uint16_t elem_schema;  /* GOFR_F_STRUCT: schema id of the struct type (added before this schema: no cycles) */
    uint8_t reserved[2];
} gofr_field_desc;

typedef struct gofr_handler_desc {
    uint32_t kind;        /* GOFR_H_* */
    uint32_t schema_id;   /* GOFR_H_ROW, GOFR_H_BIND_ECHO */
    const char* s0; uint32_t s0_len;
    const char* s1; uint32_t s1_len;
    const char* s2; uint32_t s2_len;
    const char* s3; uint32_t s3_len;
    const uint8_t* blob; uint32_t blob_len; /* GOFR_H_FILE */
} gofr_handler_desc;

/* ---- request batch (host or device memory, SoA): desc[n], trace_id[n][16], arena[arena_bytes] ----
 * A request's variable bytes are contiguous in the arena:
 *   [ URL.Path (decoded) | URL.RawQuery | pad to 4 | data section ]
 * arena_off must be a multiple of 4.  The data section is the request body for GOFR_H_BIND_ECHO routes and the
 * handler-result row for GOFR_H_ROW routes.  Row format of a struct S: its FIXED part, then its VARIABLE part.
 *   fixed part, little-endian 32-bit words, fields in schema order:
 *     INT64 / INT / UINT64 / FLOAT64  two words (lo, hi; FLOAT64: the IEEE-754 bits)   INT32 / FLOAT32  one   BOOL  one, 0/1
 *     BYTES                  one word, the byte length, GOFR_NIL_COUNT when nil
 *     TIME                   four words: seconds lo, hi, nanoseconds, zone offset
 *     STRING                 one word, the byte length                            STRUCT  the struct's fixed part, inline
 *     *T                     one word 0 (nil) / 1, then T's fixed words (ignored, and T's variable part ABSENT, when nil)
 *     []T, []*T, map[string]T  one word, the element count, GOFR_NIL_COUNT when nil
 *   variable part, bytes, fields in schema order, nothing aligned:
 *     STRING / BYTES  the bytes   STRUCT / *STRUCT  the struct's variable part   other scalars  nothing
 *     []T            count elements E(T)            map[string]T   count entries: u32 key length, key bytes, E(T)
 *     []*T           count elements: u32 presence (0 = nil, nothing follows; else 1), E(T)
 *     E(T): scalars — their fixed words; STRING / BYTES — u32 length + bytes; STRUCT — its fixed part + its variable part
 * A flat struct of scalars and strings therefore is: one word per field (INT64: two) followed by the string bytes of all
 * STRING fields concatenated in schema order. */
typedef struct gofr_req_desc {
    uint32_t arena_off;
    uint16_t path_len;
    uint16_t query_len;
    uint32_t data_len;
    uint8_t method; /* GOFR_M_* */
    uint8_t flags;  /* GOFR_REQ_* */
    uint16_t aux;
} gofr_req_desc;

#define GOFR_REQ_FORCE_QUERY 0x01u /* url.URL.ForceQuery: target ended in a bare '?' (only visible in a 301 Location) */

/* ---- response batch: out[out_off[i] .. out_off[i+1]) are request i's bytes, packed in request order;
 *      meta[i] = status (low 16 bits) | route_id << 16 (0xFFFF = no user route; status 0 = GOFR_H_HOST pending) ---- */
#define GOFR_META_STATUS(m) ((uint32_t)(m) & 0xFFFFu)
#define GOFR_META_ROUTE(m) ((uint32_t)(m) >> 16)
#define GOFR_ROUTE_NONE 0xFFFFu

typedef struct gofr_table gofr_table;
typedef struct gofr_engine gofr_engine;
typedef uint64_t gofr_ticket;

/* ================= route table (startup, single thread) ================= */
int gofr_table_create(gofr_table** out, uint32_t frame_mode);
void gofr_table_destroy(gofr_table*);
int gofr_table_add_schema(gofr_table*, uint32_t schema_id, const char* go_type_name, const gofr_field_desc* fields,
                          uint32_t n_fields);
/* Call order is match priority (mux scans routes in registration order).  Patterns: literals, {name} (= [^/]+) and
 * {name:regexp} where the regexp is a concatenation of quantified units: a unit is a bracket class, \d, \w, '.', an escaped
 * punctuation character or a literal character; a quantifier is + * ? {n} {n,} {n,m} (n, m <= 250) — e.g. [0-9]+,
 * \d{4}-\d{2}, v[0-9]+, [a-z]+\.[a-z]{2,4}.  Alternation, groups, anchors, lazy quantifiers, \s / \p{..} / POSIX classes and
 * counted repetition of classes with non-ASCII members ('.', negated classes: Go counts runes) → GOFR_ERR_UNSUPPORTED. */
int gofr_table_add_route(gofr_table*, uint32_t method, const char* pattern, uint32_t pattern_len,
                         const gofr_handler_desc* handler, uint32_t* route_id_out);
/* GET /.well-known/health, GET /favicon.ico, PathPrefix("/") catch-all — appended after the user routes, as
 * App.Run does.  Without this call the table behaves like the router before Run() (gofr_test.go:85-95): mux's own
 * plain-text 404 / empty 405 answer unmatched requests and no middleware runs for them. */
int gofr_table_add_default_routes(gofr_table*, const uint8_t* favicon, uint32_t favicon_len);
int gofr_table_seal(gofr_table*);
/* The sealed image is position independent: rank 0 seals and serializes, the bytes are broadcast (NCCL), every
 * other rank deserializes.  *len_inout: capacity in, size out (call with buf NULL to query). */
int gofr_table_serialize(const gofr_table*, uint8_t* buf, uint64_t* len_inout);
int gofr_table_deserialize(gofr_table** out, const uint8_t* buf, uint64_t len);
uint32_t gofr_table_route_count(const gofr_table*);
uint32_t gofr_table_max_response_bytes(const gofr_table*, uint32_t max_data_len);
/* Upper bound of ONE response given that request's sizes (what a host batcher sums up to size its output buffer). */
uint32_t gofr_table_response_bound(const gofr_table*, uint32_t path_len, uint32_t query_len, uint32_t data_len);

/* ================= engine (one per GPU / process) ================= */
int gofr_engine_create(gofr_engine** out, const gofr_table* sealed, int device);
void gofr_engine_destroy(gofr_engine*);

/* Device-resident batch: every pointer is device memory, `stream` is a cudaStream_t (NULL = default stream).
 * One fused launch: route match + middleware predicates + handler kind + JSON encode + HTTP framing.
 * date is the 29-byte IMF-fixdate for this batch (host memory).  out_off[n] receives the packed size.
 * d_arena must be 16-byte aligned and its allocation must extend at least 16 bytes past the last request byte; d_out must be
 * 16-byte aligned.  Launches on one engine must be stream-ordered with respect to each other (they share scratch). */
int gofr_serve_device(gofr_engine*, const gofr_req_desc* d_desc, const uint8_t* d_trace_ids, const uint8_t* d_arena,
                      uint32_t n, const char* date29, uint8_t* d_out, uint64_t out_cap, uint32_t* d_out_off,
                      uint32_t* d_meta, void* stream);

/* The same launch with a SLOT layout instead of packed offsets: response i is written at d_out + i * slot_bytes
 * (slot_bytes a multiple of 16, d_out 16-byte aligned) and d_out_len[i] receives its length; the bytes between the end
 * of a response and the next 16-byte boundary are zeroed, the rest of the slot is left untouched.  A response longer
 * than the slot is not written: d_out_len[i] > slot_bytes tells the caller to serve that request through the packed
 * call.  Slots remove the only dependence between requests (the running output offset): no scan, no look-back between
 * tiles, and every response starts aligned — this is the layout a ring of pinned response buffers wants anyway. */
int gofr_serve_device_slots(gofr_engine*, const gofr_req_desc* d_desc, const uint8_t* d_trace_ids, const uint8_t* d_arena,
                            uint32_t n, const char* date29, uint8_t* d_out, uint32_t slot_bytes, uint32_t* d_out_len,
                            uint32_t* d_meta, void* stream);

/* Host batch: pinned (gofr_alloc_pinned) or pageable caller-owned buffers; the engine pipelines
 * H2D → kernel → D2H in chunks over its own streams.  submit is thread-safe; wait blocks the caller. */
typedef struct gofr_req_batch {
    const gofr_req_desc* desc;
    const uint8_t* trace_ids;
    const uint8_t* arena;
    uint64_t arena_bytes;
    uint32_t n;
    char date[29];
    uint8_t pad[3];
} gofr_req_batch;

typedef struct gofr_resp_batch {
    uint8_t* out;       /* capacity out_cap */
    uint64_t out_cap;
    uint32_t* out_off;  /* n+1 */
    uint32_t* meta;     /* n */
    uint64_t out_bytes; /* filled by wait */
} gofr_resp_batch;

/* submit is thread-safe (submits are serialised inside the engine) and currently runs the batch to completion before it
 * returns; wait(ticket) reports that batch's status exactly once.  With pinned caller buffers egress is device driven
 * (no host stall per chunk); pageable buffers fall back to an exact-size cudaMemcpy pipeline. */
int gofr_batch_submit(gofr_engine*, const gofr_req_batch* in, gofr_resp_batch* out, gofr_ticket* ticket);
int gofr_batch_wait(gofr_engine*, gofr_ticket ticket);

/* Host batch with the slot layout: response i lands in out + i * slot_bytes (a ring of fixed-size response buffers),
 * its length in out_len[i] (> slot_bytes: not written, serve that request through the packed call).  Chunk sizes are
 * known on the host, so the results leave with plain asynchronous copies — no device-side packing, no egress kernel.
 * slot_bytes: a positive multiple of 16; out: 16-byte aligned, n * slot_bytes bytes.  Whole slots are copied back: the
 * bytes of a slot behind its (zero-padded) response are unspecified. */
typedef struct gofr_slot_batch {
    uint8_t* out;
    uint32_t slot_bytes;
    uint32_t reserved;
    uint32_t* out_len; /* n */
    uint32_t* meta;    /* n */
} gofr_slot_batch;
int gofr_batch_submit_slots(gofr_engine*, const gofr_req_batch* in, gofr_slot_batch* out, gofr_ticket* ticket);
int gofr_engine_set_chunk(gofr_engine*, uint32_t requests_per_chunk);
/* Tile geometry: shared-memory staging budget for request bytes, per request.  Tiles whose request bytes exceed the
 * budget are still served correctly, straight from HBM. */
int gofr_engine_set_tile(gofr_engine*, uint32_t in_bytes_per_req);
int gofr_engine_geometry(const gofr_engine*, uint32_t* grid, uint32_t* blocks_per_sm, uint32_t* smem_bytes,
                         uint32_t* sm_count);
/* Slot layout: which instance of the serve kernel answers — 5 CTAs/SM with 96 registers (the default, the faster one on every
 * table measured) or 4 CTAs/SM with 128 registers.  ctas_per_sm 0 = query, 4 or 5 = force; *in_effect (may be NULL)
 * receives the value in effect.  Tuning knob, no effect on results. */
int gofr_engine_slot_ctas(gofr_engine*, int ctas_per_sm, int* in_effect);
/* the engine's default for a sealed table (5), computable without a GPU */
int gofr_table_slot_ctas(const gofr_table*, int* ctas_per_sm);
/* gofr_serve_device reports an undersized d_out through this flag (the launch itself is asynchronous). */
int gofr_engine_overflowed(gofr_engine*, int* flag_out, int reset);
int gofr_engine_set_timing(gofr_engine*, int on);

/* Binds the CALLING thread to the CPUs of the NUMA node the GPU hangs off (sysfs: /sys/bus/pci/devices/<bus id>/numa_node)
 * and makes that node the preferred one for the pages the thread touches from now on — call it before gofr_alloc_pinned
 * and before the threads that fill the batches are started.  With eight GPUs on two sockets, pinned buffers that all sit
 * on one socket make half of the H2D/D2H traffic cross the inter-socket link: the reference has nothing like it (its
 * goroutines are scheduled by the Go runtime, pkg/gofr/httpServer.go:29-33); a host that feeds GPUs needs it.
 * numa_node_out (may be NULL) receives the node, -1 if unknown.  GOFR_ERR_UNSUPPORTED when the topology cannot be read. */
int gofr_bind_host_thread(int device, int* numa_node_out);
void* gofr_alloc_pinned(size_t bytes);
void gofr_free_pinned(void*);

/* gRPC unary Hello (config 5): length-prefixed HelloRequest frames in, length-prefixed HelloResponse frames out,
 * packed; frame i of the input starts at in_off[i].  meta[i] = 0 ok, else a GOFR_GRPC_* error (frame emitted empty). */
enum { GOFR_GRPC_OK = 0, GOFR_GRPC_COMPRESSED = 1, GOFR_GRPC_BAD_LENGTH = 2, GOFR_GRPC_BAD_PROTO = 3, GOFR_GRPC_BAD_UTF8 = 4,
       GOFR_GRPC_BAD_ROW = 5,
       GOFR_GRPC_DEFER = 6 /* gofr_proto_decode_nested_device: a valid frame outside its subset (a singular message field sent
                              more than once, which protobuf-go merges): decode it on the host */ };
int gofr_grpc_hello_device(gofr_engine*, const uint8_t* d_in, const uint32_t* d_in_off, uint32_t n, uint8_t* d_out,
                           uint64_t out_cap, uint32_t* d_out_off, uint32_t* d_meta, void* stream);

/* proto3 message encoder: proto.Marshal of the message a unary gRPC handler returns + the
 * 5-byte length-prefixed-message header grpc-go puts in front (examples/grpc-server/grpc/hello_grpc.pb.go:73-89 hands
 * the response to grpc-go; protobuf-go v1.32.0, grpc-go v1.60.1), for a batch of flat messages with scalar fields.
 *   fields     the message type: (number, GOFR_PB_* type) in ascending field-number order — the order proto.Marshal
 *              emits them; at most GOFR_PROTO_MAX_FIELDS
 *   d_rows     row i = d_rows[d_row_off[i] .. d_row_off[i+1]), offsets multiples of 4, GOFR_H_ROW layout: per field
 *              (same order) 64-bit types two LE words (lo, hi), the others one (float / double: IEEE bits; string /
 *              bytes: byte length), then the string bytes in field order.  The buffer must stay readable for 8 bytes
 *              past the last row.
 *   output     packed frames in row order, d_out_off[n+1]; d_meta[i] = GOFR_GRPC_OK, GOFR_GRPC_BAD_UTF8 (a string
 *              field is not valid UTF-8: proto.Marshal fails, the RPC fails — no frame), GOFR_GRPC_BAD_ROW (malformed
 *              row from the host shim — no frame).
 * Zero values are not emitted (proto3 implicit presence; -0.0 is not a zero value); negative int32 / enum values take
 * ten bytes; sint32 / sint64 are zigzag encoded. */
enum {
    GOFR_PB_DOUBLE = 1, GOFR_PB_FLOAT = 2, GOFR_PB_INT64 = 3, GOFR_PB_UINT64 = 4, GOFR_PB_INT32 = 5, GOFR_PB_FIXED64 = 6,
    GOFR_PB_FIXED32 = 7, GOFR_PB_BOOL = 8, GOFR_PB_STRING = 9, GOFR_PB_BYTES = 12, GOFR_PB_UINT32 = 13, GOFR_PB_ENUM = 14,
    GOFR_PB_SFIXED32 = 15, GOFR_PB_SFIXED64 = 16, GOFR_PB_SINT32 = 17, GOFR_PB_SINT64 = 18
}; /* google.protobuf.FieldDescriptorProto.Type */
#define GOFR_PROTO_MAX_FIELDS 32
typedef struct gofr_proto_field {
    uint32_t number; /* 1 .. 2^29-1 */
    uint32_t type;   /* GOFR_PB_* */
} gofr_proto_field;
int gofr_proto_encode_device(gofr_engine*, const gofr_proto_field* fields, uint32_t n_fields, const uint8_t* d_rows,
                             const uint32_t* d_row_off, uint32_t n, uint8_t* d_out, uint64_t out_cap, uint32_t* d_out_off,
                             uint32_t* d_meta, void* stream);
/* The other direction — dec(in) of the same handler: length-prefixed frames in (frame i = d_in[d_in_off[i] ..
 * d_in_off[i+1]), as for gofr_grpc_hello_device), proto.Unmarshal into the message type, the field values out as rows in
 * the layout above (row sizes are multiples of 4; absent fields hold their zero value), packed, d_row_off[n+1].
 * d_meta[i]: GOFR_GRPC_OK or COMPRESSED / BAD_LENGTH / BAD_PROTO / BAD_UTF8 as for the Hello call (a failed frame gives an
 * empty row).  Unknown fields — and known fields arriving with a foreign wire type — are skipped; the last occurrence of
 * a field wins; int32 / uint32 / enum keep the low 32 bits of the varint, bool is v != 0.  d_in must stay readable for 8
 * bytes past the last frame. */
int gofr_proto_decode_device(gofr_engine*, const gofr_proto_field* fields, uint32_t n_fields, const uint8_t* d_in,
                             const uint32_t* d_in_off, uint32_t n, uint8_t* d_rows, uint64_t rows_cap, uint32_t* d_row_off,
                             uint32_t* d_meta, void* stream);

/* The same encoder for message types with NESTED and REPEATED fields (proto.Marshal as protobuf-go v1.32.0 does it: fields
 * in field-number order, singular zero values skipped, a set message field written even when empty, repeated numeric
 * scalars packed, repeated strings / bytes / messages one tag per element).  msgs[m] names fields[first_field ..
 * first_field + n_fields) (ascending numbers); a field of type GOFR_PB_MESSAGE names its message type in `msg`; `root` is the
 * type of the rows.  At most 8 message types with 48 fields in all, nested at most 4 levels deep, no recursive types;
 * map fields are not taken (protobuf-go writes them in random order: there is no byte sequence to match).
 * Rows: the layout of "Row format" above — fixed words of a message in field order (64-bit kinds two words, string /
 * bytes their length, a singular message a presence word 0 / 1 followed by its own fixed part, a repeated field its
 * element count), then the variable part in field order (string bytes; a set message's variable part; the elements of a
 * repeated field: scalars as their words, strings as u32 length + bytes, messages as fixed part + variable part).
 * Output and d_meta as for gofr_proto_encode_device. */
#define GOFR_PB_MESSAGE 11u
typedef struct gofr_proto_nfield {
    uint32_t number;
    uint8_t type;     /* GOFR_PB_* or GOFR_PB_MESSAGE */
    uint8_t repeated; /* 0 / 1 */
    uint16_t msg;     /* GOFR_PB_MESSAGE: index into msgs */
} gofr_proto_nfield;
typedef struct gofr_proto_nmsg {
    uint16_t first_field, n_fields;
} gofr_proto_nmsg;
int gofr_proto_encode_nested_device(gofr_engine*, const gofr_proto_nmsg* msgs, uint32_t n_msgs, const gofr_proto_nfield* fields,
                                    uint32_t n_fields, uint32_t root, const uint8_t* d_rows, const uint32_t* d_row_off, uint32_t n,
                                    uint8_t* d_out, uint64_t out_cap, uint32_t* d_out_off, uint32_t* d_meta, void* stream);
/* The other direction for the same message types: length-prefixed frames -> rows (proto.Unmarshal as protobuf-go does it:
 * unknown fields and balanced groups skipped at every level; a known number with a foreign wire type is unknown, except that
 * repeated numeric scalars are accepted packed and unpacked; singular scalars and strings: the last occurrence wins;
 * repeated strings / bytes / messages: one element per occurrence in wire order; strings must be valid UTF-8).  d_meta[i]:
 * GOFR_GRPC_OK or COMPRESSED / BAD_LENGTH / BAD_PROTO / BAD_UTF8 / DEFER (a failed frame gives an empty row); rows are
 * packed, 4-byte aligned, zero padded to a multiple of 4, d_row_off[n + 1]. */
int gofr_proto_decode_nested_device(gofr_engine*, const gofr_proto_nmsg* msgs, uint32_t n_msgs, const gofr_proto_nfield* fields,
                                    uint32_t n_fields, uint32_t root, const uint8_t* d_in, const uint32_t* d_in_off, uint32_t n,
                                    uint8_t* d_rows, uint64_t rows_cap, uint32_t* d_row_off, uint32_t* d_meta, void* stream);
/* validation of such a description without a GPU; desc_out (>= 512 bytes, may be NULL with desc_cap 0 to validate only...
 * then GOFR_ERR_CAPACITY means "valid") receives the kernel's internal descriptor (test infrastructure reads it) */
int gofr_proto_nested_describe(const gofr_proto_nmsg* msgs, uint32_t n_msgs, const gofr_proto_nfield* fields, uint32_t n_fields,
                               uint32_t root, void* desc_out, uint32_t desc_cap);

/* Bind as a stage of the split API: Context.Bind(&v) for closures that stay on the host (pkg/gofr/context.go:52-54 ->
 * pkg/gofr/http/request.go:40-47 -> json.Unmarshal(body, &v), v a struct of a registered schema).  Body i is the data
 * section of request i (same descriptors / arena as the serve calls); result i lies in its own slot d_rows + i * slot_bytes
 * (slot_bytes a multiple of 16, d_rows 16-byte aligned), zero padded to the next 16-byte boundary:
 *   d_status[i] = GOFR_BIND_OK     the struct as a row in the GOFR_H_ROW layout (one LE u32 word per field in schema order,
 *                                  two for INT64 / INT / FLOAT64, STRING = byte length; then the DECODED string bytes in schema order:
 *                                  escapes resolved, invalid UTF-8 replaced by U+FFFD, as json.Unmarshal stores them).
 *                                  Keys are matched exactly, then case-insensitively; unknown keys are skipped; later
 *                                  duplicates win; null leaves the zero value.
 *   d_status[i] = GOFR_BIND_ERROR  err.Error() of json.Unmarshal: a syntax error ("invalid character 'x' after object key",
 *                                  "unexpected end of JSON input"), or the FIRST UnmarshalTypeError ("json: cannot unmarshal
 *                                  string into Go struct field T.f of type int64"); the partially filled struct Go also
 *                                  leaves behind in that case is not reported
 *   d_status[i] = GOFR_BIND_HOST   not decided on the device (nesting deeper than 64, or a number for a float64 member that
 *                                  is one of the rare literals left to Go's slow path — see GOFR_H_BIND_ECHO): run encoding/json on the host
 * d_len[i] = bytes of the result; a value above slot_bytes means it did not fit and nothing was written.
 * Pins: pkg/gofr/http/request_test.go:17-30, pkg/gofr/context_test.go:23-49. */
enum { GOFR_BIND_OK = 0, GOFR_BIND_ERROR = 1, GOFR_BIND_HOST = 2 };
int gofr_bind_device(gofr_engine* e, uint32_t schema_id, const gofr_req_desc* d_desc, const uint8_t* d_arena, uint32_t n,
                     uint8_t* d_rows, uint32_t slot_bytes, uint32_t* d_len, uint32_t* d_status, void* stream);
/* The same stage with host buffers (what a host shim calls between gofr_batch_route and its closures); synchronous. */
int gofr_batch_bind(gofr_engine* e, uint32_t schema_id, const gofr_req_batch* in, uint8_t* rows, uint32_t slot_bytes,
                    uint32_t* len, uint32_t* status);

/* Stage 1 of the split API for routes whose closure stays on the host (GOFR_H_HOST): everything mux.Router.ServeHTTP
 * and the middleware chain decide before handler.ServeHTTP runs (pkg/gofr/http/router.go:14,30-33;
 * middleware/cors.go:10-13; pkg/gofr/handler.go:32-36), for a whole batch resident in HBM.
 *   d_meta[i] = status | route_id << 16, status: 301 cleanPath redirect, 404 / 405 mux's own handlers (route_id =
 *               GOFR_ROUTE_NONE), 200 OPTIONS answered by the CORS middleware, 0 = the handler of route_id runs;
 *   d_vars[i * GOFR_MAX_PATH_VARS + k] = offset | length << 16 of the k-th path variable inside the request's path
 *               (mux.Vars in template order — what Request.PathParam reads, pkg/gofr/http/request.go:36-38);
 *               0xFFFFFFFF for unused slots.
 * The host then runs the closures and hands their results to gofr_serve_device / gofr_batch_submit (same table: the
 * route is registered with the handler kind that describes its result, e.g. GOFR_H_ROW). */
#define GOFR_MAX_PATH_VARS 8
int gofr_route_device(gofr_engine*, const gofr_req_desc* d_desc, const uint8_t* d_arena, uint32_t n, uint32_t* d_meta,
                      uint32_t* d_vars, void* stream);
/* The same for a batch in host memory (synchronous: H2D of descriptors and paths, the route kernel, D2H of meta and vars
 * — n and n * GOFR_MAX_PATH_VARS words).  What the C++ stand-in of the app API (include/gofr_b200.hpp) calls before it
 * runs the closures. */
int gofr_batch_route(gofr_engine*, const gofr_req_batch* in, uint32_t* meta, uint32_t* vars);

/* The JSON line middleware.Logging hands to logger.Log after every request (non-terminal writer:
 * json.NewEncoder(out).Encode(logEntry{Level: INFO, Time: time.Now(), Message: RequestLog{...}})):
 *   {"Level":"INFO","time":"<RFC 3339 nano>","message":{"id":"<32 hex>","start_time":"...","response_time":<µs>,
 *    "method":"..","user_agent":"..","ip":"..","uri":"..","response":<status>}}\n
 * with `omitempty` on every message field (logger.go:24-33).  What the reference reads from the clock and from the
 * request is an INPUT per record: the three clock readings, the local zone offset, and five byte strings laid back to
 * back in the arena in this order: req.Method | req.UserAgent() | first X-Forwarded-For header value | req.RemoteAddr |
 * req.RequestURI.  getIPAddress (first comma-separated element, else RemoteAddr, strings.TrimSpace) runs on the device.
 * Lines are packed: line i = d_out[d_out_off[i] .. d_out_off[i+1]).  The arena needs 32 readable bytes after the last
 * record byte.  trace ids: the same 16 bytes per request gofr_serve_device takes. */
/* kind GOFR_LOG_RPC is the gRPC LoggingInterceptor's line instead (pkg/gofr/grpc/log.go:15-50): logger.Infof("%s",
 * RPCLog{id, startTime, responseTime, method}) — the message is the STRING json.Marshal(RPCLog), escaped once more by
 * the outer encoder; `method` carries info.FullMethod, the other four strings are ignored. */
enum { GOFR_LOG_REQUEST = 0, GOFR_LOG_RPC = 1 };
typedef struct gofr_log_desc { /* 48 bytes */
    int64_t start_unix_ns;   /* start := time.Now()                        logger.go:44 */
    int64_t elapsed_ns;      /* time.Since(start)                          logger.go:53 */
    int64_t log_unix_ns;     /* time.Now() in logger.logf                  logging/logger.go:55 */
    uint32_t arena_off;
    uint16_t method_len, ua_len, xff_len, remote_len, uri_len;
    uint16_t status;         /* StatusResponseWriter.status (0: WriteHeader never ran → field omitted) */
    int32_t tz_offset_s;     /* offset of time.Local at that instant, seconds east of UTC */
    uint32_t kind;           /* GOFR_LOG_REQUEST or GOFR_LOG_RPC */
} gofr_log_desc;
int gofr_requestlog_device(gofr_engine*, const gofr_log_desc* d_desc, const uint8_t* d_trace_ids, const uint8_t* d_arena,
                           uint32_t n, uint8_t* d_out, uint64_t out_cap, uint32_t* d_out_off, void* stream);

/* Batching front-end: the reference serves one request per goroutine (net/http conn.serve →
 * router.ServeHTTP, pkg/gofr/httpServer.go:29-33); the GPU wants batches.  Any number of threads call
 * gofr_frontend_serve with ONE request each and block until its response is ready; a dispatcher thread closes a batch
 * when it holds max_batch requests or its oldest request has waited max_wait_us, serves it through
 * gofr_batch_submit_slots and wakes the callers.  A small ring of pinned batches rotates (one fills while another is in
 * flight; a batch whose callers are slow to pick up their responses is skipped).  The Date header of a batch is the
 * wall clock at dispatch (gofr_frontend_set_clock pins it for tests).
 * A response longer than slot_bytes is served again on its own through the packed call (same Date), into resp; serve
 * returns GOFR_ERR_CAPACITY (with *resp_len set) only when the response does not fit resp_cap.
 * destroy may only be called when no thread is inside gofr_frontend_serve. */
typedef struct gofr_frontend gofr_frontend;
int gofr_frontend_create(gofr_frontend** out, gofr_engine*, uint32_t max_batch, uint32_t max_wait_us, uint32_t slot_bytes,
                         uint32_t max_request_bytes);
void gofr_frontend_destroy(gofr_frontend*);
int gofr_frontend_set_clock(gofr_frontend*, int64_t unix_seconds);
int gofr_frontend_stats(gofr_frontend*, uint64_t* batches, uint64_t* requests);
int gofr_frontend_serve(gofr_frontend*, uint8_t method, const uint8_t* path, uint16_t path_len, const uint8_t* query,
                        uint16_t query_len, uint8_t flags, const uint8_t* data, uint32_t data_len, const uint8_t trace_id[16],
                        uint8_t* resp, uint32_t resp_cap, uint32_t* resp_len, uint32_t* meta);

/* HTTP/1.1 request heads: what net/http's readRequest + url.ParseRequestURI hand to
 * Router.ServeHTTP, for a batch of raw request messages resident in HBM, in the layout the serve calls consume.
 *   d_raw / d_raw_off: message i = d_raw[d_raw_off[i] .. d_raw_off[i+1]) — exactly one request (head + body) as framed
 *                      by the connection reader; d_raw needs 16 readable bytes after the last message.
 *   d_status[i]      : GOFR_HTTP_OK, or GOFR_HTTP_DEFER = "not in the subset below: parse it on the host" (nothing is
 *                      claimed about such a message — it may be perfectly valid).  Never a different answer.
 *   d_desc[i], d_arena: for OK messages the gofr_req_desc and its bytes (decoded URL.Path | URL.RawQuery | pad4 | body) at
 *                      arena_off = d_raw_off[i] rounded up to 4 (the parsed form is never longer than the message, so
 *                      d_arena is a buffer as large as d_raw + 16 and no offsets have to be computed across requests).
 *   d_spans[i][k]    : offset into d_raw | length << 32 of METHOD, TARGET (= RequestURI), USER_AGENT, X_FORWARDED_FOR,
 *                      HOST, BODY — the strings a gofr_log_desc needs come straight from here.
 * The subset (anything else → DEFER): request line `METHOD SP target SP HTTP/1.1 CRLF`; METHOD one of GET HEAD POST PUT
 * PATCH DELETE OPTIONS; target in origin form (`/…`, not `//…`), bytes 0x21–0x7E, no '#', every '%' in the path part
 * followed by two hex digits; header lines `name: value CRLF` with RFC 7230 token names, no leading whitespace (no
 * obs-fold), value bytes HTAB / 0x20–0x7E / ≥ 0x80, CRLF line ends only, head ≤ 16 KiB; exactly one Host header with a
 * non-empty value of [A-Za-z0-9.:_-] and brackets; at most one Content-Length (1–9 digits) and then exactly that many body
 * bytes; or one `Transfer-Encoding: chunked` (no Content-Length beside it) and a chunk stream `hex-size CRLF data CRLF …
 * 0 CRLF CRLF` with sizes of 1–8 hex digits, no chunk extensions, no trailer fields and nothing after it — the arena then
 * holds the de-chunked body, as the handler's io.ReadAll(r.Body) would (BODY span: the chunk stream as received); otherwise
 * no bytes after the head; no Expect, Upgrade or Trailer header; a Connection header only with the value keep-alive.  For such a message URL.Path is the percent-decoded path, URL.RawQuery what follows
 * the first '?', ForceQuery a trailing '?' with nothing after it, header values are trimmed of spaces and tabs. */
enum { GOFR_HTTP_OK = 0, GOFR_HTTP_DEFER = 1 };
enum { GOFR_HTTP_SPAN_METHOD = 0, GOFR_HTTP_SPAN_TARGET = 1, GOFR_HTTP_SPAN_USER_AGENT = 2, GOFR_HTTP_SPAN_XFF = 3,
       GOFR_HTTP_SPAN_HOST = 4, GOFR_HTTP_SPAN_BODY = 5, GOFR_HTTP_SPANS = 6 };
int gofr_http_parse_device(gofr_engine*, const uint8_t* d_raw, const uint32_t* d_raw_off, uint32_t n, gofr_req_desc* d_desc,
                           uint8_t* d_arena, uint32_t* d_status, uint64_t* d_spans, void* stream);

/* number of kernels launched by this engine so far (bench.py reports it as gpu_launches) */
uint64_t gofr_engine_launch_count(const gofr_engine*);
/* duration in ms of the kernels launched since the last reset, measured with CUDA events on the launch stream.
 * Timing is OFF by default (each timed launch costs two event objects): gofr_engine_set_timing(e, 1) first. */
int gofr_engine_kernel_time_ms(gofr_engine*, double* total_ms, uint64_t* launches, int reset);

const char* gofr_last_error(void);
uint32_t gofr_abi_version(void);
void gofr_format_http_date(int64_t unix_seconds, char out29[29]);

#ifdef __cplusplus
}
#endif
#endif /* GOFR_B200_H */

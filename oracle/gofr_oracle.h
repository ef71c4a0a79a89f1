/*
 * gofr_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C11, scalar, one request at a time) of the reference's request hot path, used as the parity
 * oracle and as the timed CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  The product (gofr_b200/csrc) never links, imports or calls it.
 *
 * PARITY PIN STATUS: the reference cannot be run here (no Go toolchain) and its own tests pin this
 * path only semantically (status codes, decoded data values, header values) plus ONE byte-exact encoding/json
 * string (pkg/gofr/grpc/log_test.go:28).  The oracle is checked against every one of those pins
 * (tests/test_oracle_golden.py, tests/golden/reference_pins.json) and, for config 5, against an independent
 * implementation (python google.protobuf).  Byte-level parity of full HTTP responses against real Go output is
 * therefore **parity unpinned**: the arithmetic follows the published behaviour of the pinned dependencies
 * (gorilla/mux v1.8.1, Go 1.21 encoding/json, net/http, net/url, path; protobuf-go v1.32.0, grpc-go v1.60.1).
 */
#ifndef GOFR_ORACLE_H
#define GOFR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_table orc_table;

/* frame modes / methods / handler kinds / field kinds use the same numeric values as include/gofr_b200.h so the
 * tests can drive both with one description; the code is independent. */

orc_table* orc_table_new(int frame_mode);
void orc_table_free(orc_table*);
int orc_add_schema(orc_table*, int schema_id, const char* go_type_name, int n_fields, const char* const* go_names,
                   const char* const* json_names, const int* kinds, const int* omitempty);
/* the wider data model of the schema added last: per field GOFR_C_* container, GOFR_FIELD_* flags, element schema id */
int orc_schema_extend(orc_table*, const int* containers, const int* flags, const int* elem_schemas);
/* method: 0..15 or 255 (= PathPrefix, no method matcher).  Returns route id (>=0) or <0 on error. */
int orc_add_route(orc_table*, int method, const char* pattern, int pattern_len, int hkind, int schema_id,
                  const char* s0, int s0_len, const char* s1, int s1_len, const char* s2, int s2_len, const char* s3,
                  int s3_len, const uint8_t* blob, int blob_len);
int orc_add_default_routes(orc_table*, const uint8_t* favicon, int favicon_len);

/* Serve n requests laid out exactly as include/gofr_b200.h describes (desc[n] 16 B each, ids[n][16], arena).
 * Output packed in request order, out_off[n+1], meta[n].  nthreads > 1 shards contiguously over pthreads; each
 * thread then packs into its own slice of `out` (out_off stays absolute, slices are not adjacent).
 * Returns 0, or -1 if out_cap is too small. */
int orc_serve(const orc_table*, const void* desc, const uint8_t* ids, const uint8_t* arena, uint32_t n,
              const char* date29, uint8_t* out, uint64_t out_cap, uint32_t* out_off, uint32_t* meta, int nthreads);

/* gRPC Hello (config 5). */
/* proto.Marshal of flat proto3 messages + the gRPC length prefix (orc_proto.c); fields = n_fields pairs (number, type) */
int orc_proto_encode(const uint32_t* fields, uint32_t n_fields, const uint8_t* rows, const uint32_t* row_off, uint32_t n,
                     uint8_t* out, uint64_t out_cap, uint32_t* out_off, uint32_t* meta);
/* the other direction: length-prefixed frames → rows (proto.Unmarshal of a flat proto3 message type) */
int orc_proto_decode(const uint32_t* fields, uint32_t n_fields, const uint8_t* in, const uint32_t* in_off, uint32_t n,
                     uint8_t* rows, uint64_t rows_cap, uint32_t* row_off, uint32_t* meta);
/* the same for message types with nested / repeated fields (orc_proto_nested.c): msgs = n_msgs pairs (first_field, n_fields),
 * fields = n_fields quadruples (number, type, repeated, message type index) */
int orc_proto_encode_nested(const uint32_t* msgs, uint32_t n_msgs, const uint32_t* fields, uint32_t n_fields, uint32_t root,
                            const uint8_t* rows, const uint32_t* row_off, uint32_t n, uint8_t* out, uint64_t out_cap,
                            uint32_t* out_off, uint32_t* meta);
int orc_proto_decode_nested(const uint32_t* msgs, uint32_t n_msgs, const uint32_t* fields, uint32_t n_fields, uint32_t root,
                            const uint8_t* in, const uint32_t* in_off, uint32_t n, uint8_t* rows, uint64_t rows_cap,
                            uint32_t* row_off, uint32_t* meta);
int orc_grpc_hello(const uint8_t* in, const uint32_t* in_off, uint32_t n, uint8_t* out, uint64_t out_cap,
                   uint32_t* out_off, uint32_t* meta, int nthreads);

/* ---- unit-level entry points for the golden-vector tests; each returns the number of bytes written to out
 *      (or a negative error) ---- */
int orc_json_string(const uint8_t* s, int n, uint8_t* out, int cap);      /* encoding/json string, HTML-safe */
int orc_json_int(int64_t v, uint8_t* out, int cap);
int orc_json_float64(double x, uint8_t* out, int cap);                     /* encoding/json floatEncoder; 0 = not encodable */
int orc_json_float32(float x, uint8_t* out, int cap);
void orc_set_strict_chunking(int on); /* JSON bodies > 2048 B framed as the reference does (chunked); default: as the product does */                      /* the same for a float32 */
int orc_encode_row_json(const orc_table*, int schema_id, const uint8_t* row, int n, uint8_t* out, int cap);                        /* strconv.AppendInt base 10 */
int orc_clean_path(const uint8_t* p, int n, uint8_t* out, int cap);        /* mux cleanPath */
int orc_query_get(const uint8_t* q, int qn, const uint8_t* key, int kn, uint8_t* out, int cap); /* URL.Query().Get */
int orc_escape_path(const uint8_t* p, int n, uint8_t* out, int cap);       /* url.escape(p, encodePath) */
/* Route match only: returns route id, or -1 (mux 404), -2 (mux 405), -3 (301 redirect). */
int orc_match(const orc_table*, int method, const uint8_t* path, int path_len);
/* json.Unmarshal(body, &struct) for a schema.  Writes a row (same format as the request data-section row) into
 * row_out on success and returns its length; on error returns -(length of message) and writes the message. */
int orc_bind(const orc_table*, int schema_id, const uint8_t* body, int n, uint8_t* row_out, int cap);
/* RPCLog.String() (pkg/gofr/grpc/log.go:15-25): the reference's only byte-exact encoding/json golden. */
int orc_rpclog_string(const char* id, const char* start_time, int64_t response_time, const char* method, uint8_t* out,
                      int cap);
/* Router.Match + mux.Vars for a batch (the oracle of gofr_route_device); desc = gofr_req_desc[n]. */
int orc_route_batch(const orc_table*, const void* desc, const uint8_t* arena, uint32_t n, uint32_t* meta, uint32_t* vars,
              int max_vars);
/* HTTP/1.1 request heads (the oracle of gofr_http_parse_device, orc_http.c): status 0 = parsed, 1 = deferred. */
int orc_http_parse(const uint8_t* raw, const uint32_t* raw_off, uint32_t n, void* desc, uint8_t* arena, uint32_t* status,
                   uint64_t* spans);
/* The JSON line middleware.Logging → logger.Log writes per request (orc_reqlog.c; logger.go:41-70, logging/logger.go:37-74).
 * desc: n records of 48 bytes in the layout of gofr_log_desc (include/gofr_b200.h); lines are packed back to back. */
int orc_request_log(const void* desc, const uint8_t* ids, const uint8_t* arena, uint32_t n, uint8_t* out, uint64_t out_cap,
                    uint32_t* out_off);
void orc_format_http_date(int64_t unix_seconds, char out29[29]);

#ifdef __cplusplus
}
#endif
#endif

// Package gofrb200 binds libgofr_b200.so (include/gofr_b200.h) over cgo.
//
// UNVERIFIED SOURCE: there is no Go toolchain in the build image or on the GPU box, so this file has never been
// compiled.  It shows the exact binding a GoFr maintainer would add; every C call is one declared in
// include/gofr_b200.h.  The C++/Python harnesses in this repository exercise the same ABI.
//
// How it plugs into GoFr (reference paths):
//   - App.add (pkg/gofr/gofr.go:171-177) additionally calls Table.AddRoute with the declarative handler kind (or
//     HHost for an arbitrary closure);
//   - App.Run (pkg/gofr/gofr.go:90-126) calls Table.AddDefaultRoutes, Table.Seal and NewEngine instead of
//     httpServer.Run, and starts a Batcher that collects parsed requests from the connection goroutines;
//   - the Batcher hands batches to Engine.Serve and fans the response bytes back to the connections.
package gofrb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gofr_b200 -lgofr_b200 -Wl,-rpath,${SRCDIR}/../../gofr_b200
#include <stdlib.h>
#include "gofr_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"time"
	"unsafe"
)

// Method codes: mux compares exact upper-case method strings, everything else is MOther.
const (
	MGet = iota
	MHead
	MPost
	MPut
	MPatch
	MDelete
	MConnect
	MOptions
	MTrace
	MOther = 15
)

var methodCodes = map[string]uint8{"GET": MGet, "HEAD": MHead, "POST": MPost, "PUT": MPut, "PATCH": MPatch,
	"DELETE": MDelete, "CONNECT": MConnect, "OPTIONS": MOptions, "TRACE": MTrace}

func MethodCode(m string) uint8 {
	if c, ok := methodCodes[m]; ok {
		return c
	}
	return MOther
}

func check(rc C.int, where string) error {
	if rc == 0 {
		return nil
	}
	return fmt.Errorf("%s: gofr error %d: %s", where, int(rc), C.GoString(C.gofr_last_error()))
}

// Table mirrors the route slice of mux.Router: routes in registration order, frozen by Seal.
type Table struct{ t *C.gofr_table }

func NewTable(frameMode uint32) (*Table, error) {
	var t *C.gofr_table
	if err := check(C.gofr_table_create(&t, C.uint32_t(frameMode)), "gofr_table_create"); err != nil {
		return nil, err
	}
	return &Table{t}, nil
}

// Handler is the declarative description of a gofr.Handler closure (GOFR_H_* in the header).
type Handler struct {
	Kind     uint32
	SchemaID uint32
	S0, S1   string
	S2, S3   string
	Blob     []byte
}

func (t *Table) AddRoute(method string, pattern string, h Handler) (uint32, error) {
	var d C.gofr_handler_desc
	d.kind = C.uint32_t(h.Kind)
	d.schema_id = C.uint32_t(h.SchemaID)
	cs := func(s string) (*C.char, C.uint32_t) { return C.CString(s), C.uint32_t(len(s)) }
	d.s0, d.s0_len = cs(h.S0)
	d.s1, d.s1_len = cs(h.S1)
	d.s2, d.s2_len = cs(h.S2)
	d.s3, d.s3_len = cs(h.S3)
	defer func() {
		C.free(unsafe.Pointer(d.s0)); C.free(unsafe.Pointer(d.s1)); C.free(unsafe.Pointer(d.s2)); C.free(unsafe.Pointer(d.s3))
	}()
	if len(h.Blob) > 0 {
		d.blob = (*C.uint8_t)(unsafe.Pointer(&h.Blob[0]))
		d.blob_len = C.uint32_t(len(h.Blob))
	}
	p := C.CString(pattern)
	defer C.free(unsafe.Pointer(p))
	var id C.uint32_t
	m := C.uint32_t(MethodCode(method))
	err := check(C.gofr_table_add_route(t.t, m, p, C.uint32_t(len(pattern)), &d, &id), "gofr_table_add_route")
	return uint32(id), err
}

func (t *Table) AddDefaultRoutes(favicon []byte) error {
	var p *C.uint8_t
	if len(favicon) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&favicon[0]))
	}
	return check(C.gofr_table_add_default_routes(t.t, p, C.uint32_t(len(favicon))), "gofr_table_add_default_routes")
}

func (t *Table) Seal() error { return check(C.gofr_table_seal(t.t), "gofr_table_seal") }

// Engine is one GPU.
type Engine struct{ e *C.gofr_engine }

func NewEngine(t *Table, device int) (*Engine, error) {
	var e *C.gofr_engine
	if err := check(C.gofr_engine_create(&e, t.t, C.int(device)), "gofr_engine_create"); err != nil {
		return nil, err
	}
	return &Engine{e}, nil
}

// Batch is the pinned SoA staging area one batch of parsed requests is appended to by the connection goroutines.
type Batch struct {
	Desc     []C.gofr_req_desc // n
	TraceIDs []byte            // n*16
	Arena    []byte
}

// Serve runs one batch through the GPU.  out/off/meta are caller-owned (ideally from gofr_alloc_pinned).
func (e *Engine) Serve(b *Batch, now time.Time, out []byte, off, meta []uint32) (int, error) {
	n := len(b.Desc)
	if n == 0 {
		return 0, nil
	}
	if len(off) < n+1 || len(meta) < n {
		return 0, errors.New("gofrb200: off/meta too small")
	}
	var in C.gofr_req_batch
	in.desc = &b.Desc[0]
	in.trace_ids = (*C.uint8_t)(unsafe.Pointer(&b.TraceIDs[0]))
	in.arena = (*C.uint8_t)(unsafe.Pointer(&b.Arena[0]))
	in.arena_bytes = C.uint64_t(len(b.Arena))
	in.n = C.uint32_t(n)
	C.gofr_format_http_date(C.int64_t(now.Unix()), &in.date[0])
	var o C.gofr_resp_batch
	o.out = (*C.uint8_t)(unsafe.Pointer(&out[0]))
	o.out_cap = C.uint64_t(len(out))
	o.out_off = (*C.uint32_t)(unsafe.Pointer(&off[0]))
	o.meta = (*C.uint32_t)(unsafe.Pointer(&meta[0]))
	var ticket C.gofr_ticket
	if err := check(C.gofr_batch_submit(e.e, &in, &o, &ticket), "gofr_batch_submit"); err != nil {
		return 0, err
	}
	if err := check(C.gofr_batch_wait(e.e, ticket), "gofr_batch_wait"); err != nil {
		return 0, err
	}
	return int(o.out_bytes), nil
}

// LogRecord is what middleware.Logging's deferred function reads from the clock and the request
// (pkg/gofr/http/middleware/logger.go:49-63); Append packs it as one gofr_log_desc plus its strings.
type LogBatch struct {
	Desc     []C.gofr_log_desc
	TraceIDs []byte // 16 bytes per record
	Arena    []byte // method | user_agent | x_forwarded_for | remote_addr | request_uri per record
}

func (b *LogBatch) Append(start time.Time, elapsed time.Duration, now time.Time, traceID [16]byte, method, userAgent, xff,
	remoteAddr, requestURI string, status int) {
	_, off := start.Zone()
	var d C.gofr_log_desc
	d.start_unix_ns = C.int64_t(start.UnixNano())
	d.elapsed_ns = C.int64_t(elapsed.Nanoseconds())
	d.log_unix_ns = C.int64_t(now.UnixNano())
	d.arena_off = C.uint32_t(len(b.Arena))
	d.method_len, d.ua_len, d.xff_len = C.uint16_t(len(method)), C.uint16_t(len(userAgent)), C.uint16_t(len(xff))
	d.remote_len, d.uri_len = C.uint16_t(len(remoteAddr)), C.uint16_t(len(requestURI))
	d.status = C.uint16_t(status)
	d.tz_offset_s = C.int32_t(off)
	d.kind = C.GOFR_LOG_REQUEST
	b.Desc = append(b.Desc, d)
	b.TraceIDs = append(b.TraceIDs, traceID[:]...)
	b.Arena = append(append(append(append(append(b.Arena, method...), userAgent...), xff...), remoteAddr...), requestURI...)
}

// RequestLogDevice runs gofr_requestlog_device on buffers that already live in device memory (the batcher uploads the
// LogBatch next to the request batch); lines come back packed, line i = out[off[i]:off[i+1]].
func (e *Engine) RequestLogDevice(dDesc, dTraceIDs, dArena unsafe.Pointer, n int, dOut unsafe.Pointer, outCap uint64,
	dOff unsafe.Pointer, stream unsafe.Pointer) error {
	return check(C.gofr_requestlog_device(e.e, (*C.gofr_log_desc)(dDesc), (*C.uint8_t)(dTraceIDs), (*C.uint8_t)(dArena),
		C.uint32_t(n), (*C.uint8_t)(dOut), C.uint64_t(outCap), (*C.uint32_t)(dOff), stream), "gofr_requestlog_device")
}

// RouteDevice is stage 1 of the split API for closures that stay in Go: routing decisions and mux.Vars spans for a
// batch resident in device memory (gofr_route_device).  meta[i] = status | route << 16; status 0 = run handler `route`.
func (e *Engine) RouteDevice(dDesc, dArena unsafe.Pointer, n int, dMeta, dVars unsafe.Pointer, stream unsafe.Pointer) error {
	return check(C.gofr_route_device(e.e, (*C.gofr_req_desc)(dDesc), (*C.uint8_t)(dArena), C.uint32_t(n), (*C.uint32_t)(dMeta),
		(*C.uint32_t)(dVars), stream), "gofr_route_device")
}

// BindDevice is Context.Bind as a GPU stage for closures that stay in Go (gofr_bind_device): body i (the data section of
// request i) is decoded with encoding/json's rules into a row of schema schemaID (status 0), or into err.Error() (status 1),
// in slot i of dRows; status 2 = decide on the host.  pkg/gofr/context.go:52-54, pkg/gofr/http/request.go:40-47.
func (e *Engine) BindDevice(schemaID uint32, dDesc, dArena unsafe.Pointer, n int, dRows unsafe.Pointer, slotBytes uint32,
	dLen, dStatus unsafe.Pointer, stream unsafe.Pointer) error {
	return check(C.gofr_bind_device(e.e, C.uint32_t(schemaID), (*C.gofr_req_desc)(dDesc), (*C.uint8_t)(dArena), C.uint32_t(n),
		(*C.uint8_t)(dRows), C.uint32_t(slotBytes), (*C.uint32_t)(dLen), (*C.uint32_t)(dStatus), stream), "gofr_bind_device")
}

// BindHostThread pins the calling OS thread (runtime.LockOSThread first) to the CPUs and memory of the NUMA node the GPU
// hangs off; call it before AllocPinned and before the goroutines that fill batches are started (gofr_bind_host_thread).
func BindHostThread(device int) (numaNode int, err error) {
	var node C.int
	err = check(C.gofr_bind_host_thread(C.int(device), &node), "gofr_bind_host_thread")
	return int(node), err
}

// ServeDeviceSlots writes response i into its own slotBytes-sized, 16-byte aligned slot of dOut and its length into
// dOutLen[i] (gofr_serve_device_slots): the layout a ring of fixed-size response buffers maps onto directly.
func (e *Engine) ServeDeviceSlots(dDesc, dTraceIDs, dArena unsafe.Pointer, n int, now time.Time, dOut unsafe.Pointer,
	slotBytes uint32, dOutLen, dMeta unsafe.Pointer, stream unsafe.Pointer) error {
	var date [29]C.char
	C.gofr_format_http_date(C.int64_t(now.Unix()), &date[0])
	return check(C.gofr_serve_device_slots(e.e, (*C.gofr_req_desc)(dDesc), (*C.uint8_t)(dTraceIDs), (*C.uint8_t)(dArena),
		C.uint32_t(n), &date[0], (*C.uint8_t)(dOut), C.uint32_t(slotBytes), (*C.uint32_t)(dOutLen), (*C.uint32_t)(dMeta), stream),
		"gofr_serve_device_slots")
}

// HTTPParseDevice turns raw HTTP/1.1 request messages (back to back in dRaw, message i = dRaw[off[i]:off[i+1]]) into
// request descriptors + arena for the serve calls (gofr_http_parse_device).  status[i] != 0 (GOFR_HTTP_DEFER): hand
// that connection's bytes to net/http as before.
func (e *Engine) HTTPParseDevice(dRaw, dRawOff unsafe.Pointer, n int, dDesc, dArena, dStatus, dSpans unsafe.Pointer,
	stream unsafe.Pointer) error {
	return check(C.gofr_http_parse_device(e.e, (*C.uint8_t)(dRaw), (*C.uint32_t)(dRawOff), C.uint32_t(n), (*C.gofr_req_desc)(dDesc),
		(*C.uint8_t)(dArena), (*C.uint32_t)(dStatus), (*C.uint64_t)(dSpans), stream), "gofr_http_parse_device")
}

// Frontend gathers single requests from many goroutines into batches (gofr_frontend_*): the per-request call that stands
// where router.ServeHTTP stands in the reference (pkg/gofr/httpServer.go:29-33).
type Frontend struct{ f *C.gofr_frontend }

func NewFrontend(e *Engine, maxBatch, maxWaitMicros, slotBytes, maxRequestBytes int) (*Frontend, error) {
	var f *C.gofr_frontend
	if err := check(C.gofr_frontend_create(&f, e.e, C.uint32_t(maxBatch), C.uint32_t(maxWaitMicros), C.uint32_t(slotBytes),
		C.uint32_t(maxRequestBytes)), "gofr_frontend_create"); err != nil {
		return nil, err
	}
	return &Frontend{f}, nil
}

// Close must not run while a Serve call is in progress.
func (f *Frontend) Close() { C.gofr_frontend_destroy(f.f); f.f = nil }

// Serve blocks until the batch this request joined has been served; resp[:n] is the response (wire bytes in FRAME_WIRE
// mode), meta = status | route id << 16.  A response that does not fit resp or the slot returns an error with n set.
func (f *Frontend) Serve(method uint8, path, query []byte, flags uint8, body []byte, traceID *[16]byte, resp []byte) (n int, meta uint32, err error) {
	var cn, cm C.uint32_t
	rc := C.gofr_frontend_serve(f.f, C.uint8_t(method), bytesPtr(path), C.uint16_t(len(path)), bytesPtr(query), C.uint16_t(len(query)),
		C.uint8_t(flags), bytesPtr(body), C.uint32_t(len(body)), (*C.uint8_t)(unsafe.Pointer(&traceID[0])), bytesPtr(resp),
		C.uint32_t(len(resp)), &cn, &cm)
	return int(cn), uint32(cm), check(rc, "gofr_frontend_serve")
}

func bytesPtr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0]))
}

// ProtoField is one field of a flat proto3 message type: (number, FieldDescriptorProto.Type).
type ProtoField struct{ Number, Type uint32 }

// ProtoEncodeDevice marshals n rows (GOFR_H_ROW layout, resident in HBM) as messages of the given type and frames them
// for gRPC (gofr_proto_encode_device): what proto.Marshal + grpc-go's msgHeader do for a unary handler's response.
func (e *Engine) ProtoEncodeDevice(fields []ProtoField, dRows, dRowOff unsafe.Pointer, n int, dOut unsafe.Pointer, outCap uint64,
	dOutOff, dMeta unsafe.Pointer, stream unsafe.Pointer) error {
	var fp *C.gofr_proto_field
	if len(fields) > 0 {
		fp = (*C.gofr_proto_field)(unsafe.Pointer(&fields[0]))
	}
	return check(C.gofr_proto_encode_device(e.e, fp, C.uint32_t(len(fields)), (*C.uint8_t)(dRows), (*C.uint32_t)(dRowOff), C.uint32_t(n),
		(*C.uint8_t)(dOut), C.uint64_t(outCap), (*C.uint32_t)(dOutOff), (*C.uint32_t)(dMeta), stream), "gofr_proto_encode_device")
}

// ProtoDecodeDevice is the other direction (gofr_proto_decode_device): length-prefixed request frames → rows.
func (e *Engine) ProtoDecodeDevice(fields []ProtoField, dIn, dInOff unsafe.Pointer, n int, dRows unsafe.Pointer, rowsCap uint64,
	dRowOff, dMeta unsafe.Pointer, stream unsafe.Pointer) error {
	var fp *C.gofr_proto_field
	if len(fields) > 0 {
		fp = (*C.gofr_proto_field)(unsafe.Pointer(&fields[0]))
	}
	return check(C.gofr_proto_decode_device(e.e, fp, C.uint32_t(len(fields)), (*C.uint8_t)(dIn), (*C.uint32_t)(dInOff), C.uint32_t(n),
		(*C.uint8_t)(dRows), C.uint64_t(rowsCap), (*C.uint32_t)(dRowOff), (*C.uint32_t)(dMeta), stream), "gofr_proto_decode_device")
}

// ProtoNField / ProtoNMsg describe message types with nested and repeated fields (gofr_proto_nfield / gofr_proto_nmsg):
// Msgs[m] owns Fields[FirstField : FirstField+NFields] in ascending field-number order; a field of Type 11
// (TYPE_MESSAGE) names its message type in Msg.  protoreflect gives all of it: for a generated message,
// walk md.Fields() (fd.Number(), fd.Kind(), fd.IsList(), fd.Message()).
type ProtoNField struct {
	Number   uint32
	Type     uint8
	Repeated uint8
	Msg      uint16
}
type ProtoNMsg struct{ FirstField, NFields uint16 }

// ProtoEncodeNestedDevice marshals n rows as messages of type msgs[root] and frames them for gRPC
// (gofr_proto_encode_nested_device): proto.Marshal for responses that are not flat — nested messages (a presence word
// and the message's fixed part inline in the row), repeated fields (a count, the elements in the row's variable part).
func (e *Engine) ProtoEncodeNestedDevice(msgs []ProtoNMsg, fields []ProtoNField, root uint32, dRows, dRowOff unsafe.Pointer, n int,
	dOut unsafe.Pointer, outCap uint64, dOutOff, dMeta unsafe.Pointer, stream unsafe.Pointer) error {
	if len(msgs) == 0 || len(fields) == 0 {
		return errors.New("gofrb200: empty message description")
	}
	return check(C.gofr_proto_encode_nested_device(e.e, (*C.gofr_proto_nmsg)(unsafe.Pointer(&msgs[0])), C.uint32_t(len(msgs)),
		(*C.gofr_proto_nfield)(unsafe.Pointer(&fields[0])), C.uint32_t(len(fields)), C.uint32_t(root), (*C.uint8_t)(dRows),
		(*C.uint32_t)(dRowOff), C.uint32_t(n), (*C.uint8_t)(dOut), C.uint64_t(outCap), (*C.uint32_t)(dOutOff), (*C.uint32_t)(dMeta), stream),
		"gofr_proto_encode_nested_device")
}

// SlotCTAs queries (ctas == 0) or forces (4 / 5) the instance of the slot-layout serve kernel (gofr_engine_slot_ctas).
func (e *Engine) SlotCTAs(ctas int) (int, error) {
	var v C.int
	err := check(C.gofr_engine_slot_ctas(e.e, C.int(ctas), &v), "gofr_engine_slot_ctas")
	return int(v), err
}

// ProtoDecodeNestedDevice is the other direction (gofr_proto_decode_nested_device): length-prefixed request frames → rows of
// msgs[root].  dMeta[i] == 6 (GOFR_GRPC_DEFER): a valid frame this decoder leaves to proto.Unmarshal on the host.
func (e *Engine) ProtoDecodeNestedDevice(msgs []ProtoNMsg, fields []ProtoNField, root uint32, dIn, dInOff unsafe.Pointer, n int,
	dRows unsafe.Pointer, rowsCap uint64, dRowOff, dMeta unsafe.Pointer, stream unsafe.Pointer) error {
	if len(msgs) == 0 || len(fields) == 0 {
		return errors.New("gofrb200: empty message description")
	}
	return check(C.gofr_proto_decode_nested_device(e.e, (*C.gofr_proto_nmsg)(unsafe.Pointer(&msgs[0])), C.uint32_t(len(msgs)),
		(*C.gofr_proto_nfield)(unsafe.Pointer(&fields[0])), C.uint32_t(len(fields)), C.uint32_t(root), (*C.uint8_t)(dIn),
		(*C.uint32_t)(dInOff), C.uint32_t(n), (*C.uint8_t)(dRows), C.uint64_t(rowsCap), (*C.uint32_t)(dRowOff), (*C.uint32_t)(dMeta), stream),
		"gofr_proto_decode_nested_device")
}

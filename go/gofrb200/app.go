// UNVERIFIED SOURCE (see gofrb200.go: no Go toolchain exists where this repository is built and tested).
//
// The app-level shim: gofr.New / GET / PUT / POST / DELETE / Context.Param / PathParam / Bind / Run with the reference's
// signatures (pkg/gofr/gofr.go:49-73,152-177; pkg/gofr/context.go:12-54; pkg/gofr/handler.go:12), handlers unchanged —
// arbitrary closures — served through the split API:
//
//	Engine.RouteHost   gofr_batch_route   mux match + middleware decisions + mux.Vars spans      (GPU)
//	the closures       Handler(c)         only for requests whose handler the reference would call (host, this file)
//	Engine.Serve       gofr_batch_submit  Responder.Respond + net/http framing                    (GPU)
//
// include/gofr_b200.hpp is the same flow in C++, compiled and tested (examples/cpp/server_routes.cpp is the reference's
// TestGofr_ServerRoutes written against it); keep the two in step.
package gofrb200

/*
#include <stdlib.h>
#include "gofr_b200.h"
*/
import "C"

import (
	"encoding/binary"
	"encoding/json"
	"errors"
	"io"
	"math"
	"net/http"
	"reflect"
	"strings"
	"time"
	"unsafe"
)

const (
	resultData    = 0
	resultError   = 1
	resultNil     = 2
	resultMissing = 3
	resultBoth    = 4
	resultString  = 5
	// response.Raw outcomes; the error selector goes into bits 8..15 (include/gofr_b200.h)
	resultRawData   = 6
	resultRawString = 7
	resultRawNil    = 8
	rawOK           = 0
	rawErr          = 1
	rawMissing      = 2
	hResult         = 12 // GOFR_H_RESULT
)

// Raw is the reference's response.Raw (pkg/gofr/http/response/raw.go:3-5): a handler that returns Raw{Data: v} gets v
// encoded bare, without the {"error":…,"data":…} envelope (pkg/gofr/http/responder.go:24-26).
type Raw struct {
	Data interface{}
}

// Context is what a handler sees (pkg/gofr/context.go:12-27): the request behind the reference's Request interface.
type Context struct {
	req        *http.Request
	body       []byte
	pathParams map[string]string
}

func (c *Context) Param(key string) string     { return c.req.URL.Query().Get(key) } // http/request.go:28-30
func (c *Context) PathParam(key string) string { return c.pathParams[key] }          // http/request.go:36-38
func (c *Context) Bind(i interface{}) error    { return json.Unmarshal(c.body, &i) } // http/request.go:40-47

// HandlerFunc is the reference's gofr.Handler (pkg/gofr/handler.go:12); `Handler` in this package is the declarative
// description of gofrb200.go.
type HandlerFunc func(c *Context) (interface{}, error)

type route struct {
	method, pattern string
	fn              HandlerFunc
	schemaID        uint32
	rtype           reflect.Type
	vars            []string
}

// App mirrors gofr.App for the HTTP path.
type App struct {
	table   *Table
	engine  *Engine
	routes  []route
	schemas map[reflect.Type]uint32
	order   []reflect.Type // registration order: a struct type comes after the struct types its fields use
	regErr  error          // first type the GPU encoder does not model (reported by Run)
}

// New is gofr.New() (pkg/gofr/gofr.go:49-73) without config, container and servers other than the HTTP path.
func New() (*App, error) {
	t, err := NewTable(0) // GOFR_FRAME_WIRE
	if err != nil {
		return nil, err
	}
	return &App{table: t, schemas: map[reflect.Type]uint32{}}, nil
}

// GET / PUT / POST / DELETE (pkg/gofr/gofr.go:152-169).  returns: optionally one sample value of the struct type the
// handler returns, so that its fields can be registered before the table is sealed (Go needs no such hint because
// encoding/json reflects at run time; the GPU encoder wants the field list up front).
func (a *App) GET(pattern string, h HandlerFunc, returns ...interface{})    { a.add("GET", pattern, h, returns...) }
func (a *App) PUT(pattern string, h HandlerFunc, returns ...interface{})    { a.add("PUT", pattern, h, returns...) }
func (a *App) POST(pattern string, h HandlerFunc, returns ...interface{})   { a.add("POST", pattern, h, returns...) }
func (a *App) DELETE(pattern string, h HandlerFunc, returns ...interface{}) { a.add("DELETE", pattern, h, returns...) }

// add is App.add (pkg/gofr/gofr.go:171-177): registration order is match priority.
func (a *App) add(method, pattern string, h HandlerFunc, returns ...interface{}) {
	r := route{method: method, pattern: pattern, fn: h, vars: templateVars(pattern)}
	if len(returns) > 0 {
		r.rtype = reflect.TypeOf(returns[0])
		r.schemaID = a.register(r.rtype, map[reflect.Type]bool{})
	}
	a.routes = append(a.routes, r)
}

// templateVars lists the variable names of a mux template in order ("{name}" or "{name:pattern}", nested braces allowed).
func templateVars(pattern string) []string {
	var names []string
	level, start := 0, 0
	for i := 0; i < len(pattern); i++ {
		switch pattern[i] {
		case '{':
			if level == 0 {
				start = i + 1
			}
			level++
		case '}':
			if level > 0 {
				level--
				if level == 0 {
					inner := pattern[start:i]
					if k := strings.IndexByte(inner, ':'); k >= 0 {
						inner = inner[:k]
					}
					names = append(names, inner)
				}
			}
		}
	}
	return names
}

const (
	cValue = 0 // GOFR_C_*: T, *T, []T, map[string]T
	cPtr   = 1
	cSlice = 2
	cMap   = 3
	cSlicePtr = 4 // []*T
	fStruct   = 7
	fieldBare = 1
	nilCount  = 0xFFFFFFFF
)

var timeType = reflect.TypeOf(time.Time{})

// timeWords: Unix seconds, nanoseconds, zone offset — what Time.MarshalJSON's text is made of.
func timeWords(b []byte, t time.Time) []byte {
	_, off := t.Zone()
	b = binary.LittleEndian.AppendUint64(b, uint64(t.Unix()))
	return u32(u32(b, uint32(t.Nanosecond())), uint32(int32(off)))
}

// isBytes: []byte (any []T with T of kind uint8 without its own marshaller): base64 in encoding/json, kind GOFR_F_BYTES here.
func isBytes(t reflect.Type) bool { return t.Kind() == reflect.Slice && t.Elem().Kind() == reflect.Uint8 }

// typeDesc resolves a Go type to (GOFR_F_* kind, GOFR_C_* container, struct type of a GOFR_F_STRUCT).  ok = false: a type the
// GPU encoder does not model (interface{}, [][]T other than [][]byte, maps of structs or with non-string keys, arrays,
// channels ...): such a route stays on the host path.  Integers narrower than the row's words are widened (int8 / int16 →
// INT32, uint8 / uint16 / uint32 → INT64): the JSON text is the same.
func typeDesc(t reflect.Type) (kind, container uint8, elem reflect.Type, ok bool) {
	switch {
	case isBytes(t): // a []byte by value: kind 9 below
	case t.Kind() == reflect.Ptr:
		container, t = cPtr, t.Elem()
	case t.Kind() == reflect.Slice && t.Elem().Kind() == reflect.Ptr: // []*T: what ORMs hand back
		container, t = cSlicePtr, t.Elem().Elem()
	case t.Kind() == reflect.Slice:
		container, t = cSlice, t.Elem()
	case t.Kind() == reflect.Map:
		if t.Key().Kind() != reflect.String {
			return 0, 0, nil, false
		}
		container, t = cMap, t.Elem()
	}
	if isBytes(t) {
		return 9, container, nil, true
	}
	if t == timeType { // time.Time: Time.MarshalJSON's RFC 3339 text (kind GOFR_F_TIME), not its unexported fields
		return 11, container, nil, true
	}
	switch t.Kind() {
	case reflect.Int64, reflect.Uint8, reflect.Uint16, reflect.Uint32:
		kind = 1
	case reflect.Int32, reflect.Int8, reflect.Int16:
		kind = 2
	case reflect.Uint64, reflect.Uint, reflect.Uintptr:
		kind = 8
	case reflect.Float32:
		kind = 10
	case reflect.Bool:
		kind = 3
	case reflect.String:
		kind = 4
	case reflect.Int:
		kind = 5
	case reflect.Float64:
		kind = 6
	case reflect.Struct:
		if container == cMap {
			return 0, 0, nil, false
		}
		kind, elem = fStruct, t
	default:
		return 0, 0, nil, false
	}
	return kind, container, elem, true
}

// register gives t (a struct type, or a non-struct type a handler returns: a "bare" schema) a schema id, registering the
// struct types below it first.  Recursive types (a struct reaching itself through a pointer or a slice) are refused: the
// device walker has a fixed frame stack.
func (a *App) register(t reflect.Type, visiting map[reflect.Type]bool) uint32 {
	if id, ok := a.schemas[t]; ok {
		return id
	}
	fail := func(msg string) uint32 {
		if a.regErr == nil {
			a.regErr = errors.New("gofrb200: " + t.String() + ": " + msg)
		}
		return 0
	}
	if visiting[t] {
		return fail("recursive type")
	}
	visiting[t] = true
	defer delete(visiting, t)
	if t.Kind() == reflect.Struct && t != timeType { // a bare time.Time is a value, not a struct to walk
		for i := 0; i < t.NumField(); i++ {
			f := t.Field(i)
			if name, _, _ := strings.Cut(f.Tag.Get("json"), ","); name == "-" {
				continue
			}
			if f.PkgPath != "" || f.Anonymous {
				return fail("field " + f.Name + ": unexported and embedded fields are not modelled")
			}
			_, _, elem, ok := typeDesc(f.Type)
			if !ok {
				return fail("field " + f.Name + ": type not supported by the GPU encoder")
			}
			if elem != nil && a.register(elem, visiting) == 0 {
				return 0
			}
		}
	} else {
		_, _, elem, ok := typeDesc(t)
		if !ok {
			return fail("type not supported by the GPU encoder")
		}
		if elem != nil && a.register(elem, visiting) == 0 {
			return 0
		}
	}
	id := uint32(len(a.schemas) + 1)
	a.schemas[t] = id
	a.order = append(a.order, t)
	return id
}

// Run is App.Run (pkg/gofr/gofr.go:90-126) minus the listener: struct types, routes, default routes, seal, engine.
func (a *App) Run(device int, favicon []byte) error {
	if a.regErr != nil {
		return a.regErr
	}
	for _, t := range a.order {
		id := a.schemas[t]
		var fields []C.gofr_field_desc
		var keep []unsafe.Pointer
		desc := func(goName, jsonName string, ft reflect.Type, omitempty bool, flags uint8) {
			kind, container, elem, _ := typeDesc(ft)
			var d C.gofr_field_desc
			d.go_name = C.CString(goName)
			d.json_name = C.CString(jsonName)
			keep = append(keep, unsafe.Pointer(d.go_name), unsafe.Pointer(d.json_name))
			d.kind = C.uint8_t(kind)
			d.container = C.uint8_t(container)
			d.flags = C.uint8_t(flags)
			if elem != nil {
				d.elem_schema = C.uint16_t(a.schemas[elem])
			}
			if omitempty {
				d.omitempty = 1
			}
			fields = append(fields, d)
		}
		if t.Kind() == reflect.Struct && t != timeType {
			for i := 0; i < t.NumField(); i++ {
				f := t.Field(i)
				name, opts, _ := strings.Cut(f.Tag.Get("json"), ",")
				if name == "-" {
					continue
				}
				desc(f.Name, name, f.Type, strings.Contains(","+opts+",", ",omitempty,"), 0)
			}
		} else {
			desc("", "", t, false, fieldBare)
		}
		tn := C.CString(t.String())
		keep = append(keep, unsafe.Pointer(tn))
		var fp *C.gofr_field_desc
		if len(fields) > 0 {
			fp = &fields[0]
		}
		err := check(C.gofr_table_add_schema(a.table.t, C.uint32_t(id), tn, fp, C.uint32_t(len(fields))), "gofr_table_add_schema")
		for _, p := range keep {
			C.free(p)
		}
		if err != nil {
			return err
		}
	}
	for _, r := range a.routes {
		if _, err := a.table.AddRoute(r.method, r.pattern, Handler{Kind: hResult, SchemaID: r.schemaID}); err != nil {
			return err
		}
	}
	if err := a.table.AddDefaultRoutes(favicon); err != nil {
		return err
	}
	if err := a.table.Seal(); err != nil {
		return err
	}
	e, err := NewEngine(a.table, device)
	a.engine = e
	return err
}

// RouteHost is gofr_batch_route: stage 1 for a batch in host memory.
func (e *Engine) RouteHost(b *Batch, meta, vars []uint32) error {
	var in C.gofr_req_batch
	in.desc = &b.Desc[0]
	in.trace_ids = (*C.uint8_t)(unsafe.Pointer(&b.TraceIDs[0]))
	in.arena = (*C.uint8_t)(unsafe.Pointer(&b.Arena[0]))
	in.arena_bytes = C.uint64_t(len(b.Arena))
	in.n = C.uint32_t(len(b.Desc))
	return check(C.gofr_batch_route(e.e, &in, (*C.uint32_t)(unsafe.Pointer(&meta[0])), (*C.uint32_t)(unsafe.Pointer(&vars[0]))),
		"gofr_batch_route")
}

func pad4(b []byte) []byte {
	for len(b)%4 != 0 {
		b = append(b, 0)
	}
	return b
}

func u32(b []byte, v uint32) []byte { return binary.LittleEndian.AppendUint32(b, v) }

// scalarWords appends the fixed words of a scalar (include/gofr_b200.h "Row format").
func scalarWords(b []byte, v reflect.Value) []byte {
	if v.Type() == timeType {
		return timeWords(b, v.Interface().(time.Time))
	}
	switch v.Kind() {
	case reflect.Int64, reflect.Int:
		return binary.LittleEndian.AppendUint64(b, uint64(v.Int()))
	case reflect.Uint64, reflect.Uint, reflect.Uintptr, reflect.Uint8, reflect.Uint16, reflect.Uint32: // the narrow ones ride as INT64
		return binary.LittleEndian.AppendUint64(b, v.Uint())
	case reflect.Int32, reflect.Int8, reflect.Int16:
		return u32(b, uint32(int32(v.Int())))
	case reflect.Float32:
		return u32(b, math.Float32bits(float32(v.Float())))
	case reflect.Bool:
		if v.Bool() {
			return u32(b, 1)
		}
		return u32(b, 0)
	case reflect.Float64:
		return binary.LittleEndian.AppendUint64(b, math.Float64bits(v.Float()))
	}
	return b
}

// fixedBytes is the size of the fixed words a field of type t owns.
func fixedBytes(t reflect.Type) int {
	switch t.Kind() {
	case reflect.Slice, reflect.Map:
		return 4
	case reflect.Ptr:
		return 4 + fixedBytes(t.Elem())
	case reflect.Struct:
		if t == timeType {
			return 16
		}
		n := 0
		for i := 0; i < t.NumField(); i++ {
			if name, _, _ := strings.Cut(t.Field(i).Tag.Get("json"), ","); name != "-" {
				n += fixedBytes(t.Field(i).Type)
			}
		}
		return n
	case reflect.Int64, reflect.Int, reflect.Float64, reflect.Uint64, reflect.Uint, reflect.Uintptr, reflect.Uint8, reflect.Uint16, reflect.Uint32:
		return 8
	}
	return 4
}

// encodePlain: a T by value — its fixed words to fixed, its variable part to vars.
func encodePlain(v reflect.Value, fixed, vars []byte) ([]byte, []byte) {
	if isBytes(v.Type()) { // length word (nilCount: the nil slice), the bytes in the variable part
		if v.IsNil() {
			return u32(fixed, nilCount), vars
		}
		return u32(fixed, uint32(v.Len())), append(vars, v.Bytes()...)
	}
	if v.Type() == timeType {
		return scalarWords(fixed, v), vars
	}
	switch v.Kind() {
	case reflect.String:
		return u32(fixed, uint32(v.Len())), append(vars, v.String()...)
	case reflect.Struct:
		return encodeRow(v, fixed, vars)
	}
	return scalarWords(fixed, v), vars
}

// encodeElement: E(T), an element of a slice or map, entirely in the variable part.
func encodeElement(v reflect.Value, vars []byte) []byte {
	if isBytes(v.Type()) {
		if v.IsNil() {
			return u32(vars, nilCount)
		}
		return append(u32(vars, uint32(v.Len())), v.Bytes()...)
	}
	if v.Type() == timeType {
		return scalarWords(vars, v)
	}
	switch v.Kind() {
	case reflect.String:
		return append(u32(vars, uint32(v.Len())), v.String()...)
	case reflect.Struct:
		fx, vr := encodeRow(v, nil, nil)
		return append(append(vars, fx...), vr...)
	}
	return scalarWords(vars, v)
}

// encodeField: one field (or the bare value of a non-struct type).
func encodeField(f reflect.Value, fixed, vars []byte) ([]byte, []byte) {
	if isBytes(f.Type()) {
		return encodePlain(f, fixed, vars)
	}
	switch f.Kind() {
	case reflect.Ptr:
		if f.IsNil() {
			return append(fixed, make([]byte, fixedBytes(f.Type()))...), vars
		}
		return encodePlain(f.Elem(), u32(fixed, 1), vars)
	case reflect.Slice:
		if f.IsNil() {
			return u32(fixed, nilCount), vars
		}
		fixed = u32(fixed, uint32(f.Len()))
		for i := 0; i < f.Len(); i++ {
			e := f.Index(i)
			if e.Kind() == reflect.Ptr { // []*T: a presence word, then the pointee
				if e.IsNil() {
					vars = u32(vars, 0)
					continue
				}
				vars, e = u32(vars, 1), e.Elem()
			}
			vars = encodeElement(e, vars)
		}
		return fixed, vars
	case reflect.Map:
		if f.IsNil() {
			return u32(fixed, nilCount), vars
		}
		fixed = u32(fixed, uint32(f.Len()))
		it := f.MapRange() // any order: the device sorts the keys like encoding/json does
		for it.Next() {
			k := it.Key().String()
			vars = encodeElement(it.Value(), append(u32(vars, uint32(len(k))), k...))
		}
		return fixed, vars
	}
	return encodePlain(f, fixed, vars)
}

// encodeRow lays a struct value out as a handler-result row: fixed words in field order, then the variable part.
func encodeRow(v reflect.Value, fixed, vars []byte) ([]byte, []byte) {
	t := v.Type()
	for i := 0; i < v.NumField(); i++ {
		if name, _, _ := strings.Cut(t.Field(i).Tag.Get("json"), ","); name == "-" {
			continue
		}
		fixed, vars = encodeField(v.Field(i), fixed, vars)
	}
	return fixed, vars
}

// encodeValue: a value of a registered type — a struct, or the bare value of a non-struct type.
func encodeValue(v reflect.Value) (fixed, vars []byte) {
	if v.Kind() == reflect.Struct && v.Type() != timeType {
		return encodeRow(v, nil, nil)
	}
	return encodeField(v, nil, nil)
}

// resultRecord describes (data, err) for Responder.Respond (pkg/gofr/http/responder.go:19-62) as a GOFR_H_RESULT record.
func (a *App) resultRecord(r *route, data interface{}, err error) []byte {
	var rec []byte
	if raw, ok := data.(Raw); ok {
		// response.Raw (pkg/gofr/http/response/raw.go:3-5): Respond encodes raw.Data bare; the error only picks the status
		es := uint32(rawOK)
		if err != nil {
			es = rawErr
			if errors.Is(err, http.ErrMissingFile) {
				es = rawMissing
			}
		}
		rv := reflect.ValueOf(raw.Data)
		switch {
		case raw.Data == nil:
			rec = u32(rec, resultRawNil|es<<8)
		case r.rtype != nil && rv.Type() == r.rtype:
			fixed, strs := encodeValue(rv)
			rec = append(append(u32(rec, resultRawData|es<<8), fixed...), strs...)
		default:
			if s, ok := raw.Data.(string); ok {
				rec = append(u32(u32(rec, resultRawString|es<<8), uint32(len(s))), s...)
			} else {
				rec = u32(rec, 0xFFFFFFFF)
			}
		}
		return rec
	}
	rv := reflect.ValueOf(data)
	isStruct := data != nil && r.rtype != nil && rv.Type() == r.rtype
	switch {
	case isStruct && err != nil:
		fixed, strs := encodeValue(rv)
		msg := err.Error()
		rec = u32(u32(rec, resultBoth), uint32(len(msg)))
		rec = append(append(append(rec, fixed...), msg...), strs...)
	case isStruct:
		fixed, strs := encodeValue(rv)
		rec = append(append(u32(rec, resultData), fixed...), strs...)
	case err != nil:
		kind := uint32(resultError)
		if errors.Is(err, http.ErrMissingFile) {
			kind = resultMissing
		}
		msg := err.Error()
		rec = append(u32(u32(rec, kind), uint32(len(msg))), msg...)
	case data == nil:
		rec = u32(rec, resultNil)
	default:
		if s, ok := data.(string); ok {
			rec = append(u32(u32(rec, resultString), uint32(len(s))), s...)
		} else {
			rec = u32(rec, 0xFFFFFFFF) // a type the GPU encoder was not told about: answered like a panic; serve it on the host instead
		}
	}
	return rec
}

// ServeBatch is router.ServeHTTP (pkg/gofr/httpServer.go:29-33) for a batch of parsed requests; response i is the wire
// bytes the reference's server would have written for reqs[i].  traceIDs: 16 bytes per request (the tracer's span ids).
func (a *App) ServeBatch(reqs []*http.Request, traceIDs []byte, now time.Time) ([][]byte, error) {
	n := len(reqs)
	if n == 0 {
		return nil, nil
	}
	b := &Batch{Desc: make([]C.gofr_req_desc, n), TraceIDs: traceIDs}
	bodies := make([][]byte, n)
	for i, r := range reqs {
		d := &b.Desc[i]
		d.arena_off = C.uint32_t(len(b.Arena))
		d.path_len = C.uint16_t(len(r.URL.Path))
		d.query_len = C.uint16_t(len(r.URL.RawQuery))
		d.method = C.uint8_t(MethodCode(r.Method))
		if r.URL.ForceQuery {
			d.flags = 1 // GOFR_REQ_FORCE_QUERY
		}
		b.Arena = pad4(append(append(b.Arena, r.URL.Path...), r.URL.RawQuery...))
		if r.Body != nil {
			bodies[i], _ = io.ReadAll(r.Body)
		}
	}
	b.Arena = append(b.Arena, make([]byte, 64)...)
	meta := make([]uint32, n)
	vars := make([]uint32, n*8) // GOFR_MAX_PATH_VARS
	if err := a.engine.RouteHost(b, meta, vars); err != nil {
		return nil, err
	}
	// the closures, only where the reference would have called one
	arena := make([]byte, 0, len(b.Arena))
	var bound uint64 = 4096
	for i, r := range reqs {
		var rec []byte
		status, rid := meta[i]&0xFFFF, int(meta[i]>>16)
		if status == 0 && rid < len(a.routes) {
			rt := &a.routes[rid]
			c := &Context{req: r, body: bodies[i], pathParams: map[string]string{}}
			for k, name := range rt.vars {
				if k >= 8 { // GOFR_MAX_PATH_VARS
					break
				}
				if v := vars[i*8+k]; v != 0xFFFFFFFF {
					c.pathParams[name] = r.URL.Path[(v & 0xFFFF) : (v&0xFFFF)+(v>>16)]
				}
			}
			rec = a.call(rt, c)
		}
		d := &b.Desc[i]
		d.arena_off = C.uint32_t(len(arena))
		d.data_len = C.uint32_t(len(rec))
		arena = pad4(append(append(arena, r.URL.Path...), r.URL.RawQuery...))
		arena = pad4(append(arena, rec...))
		bound += uint64(C.gofr_table_response_bound(a.table.t, C.uint32_t(d.path_len), C.uint32_t(d.query_len), d.data_len))
	}
	b.Arena = append(arena, make([]byte, 64)...)
	out := make([]byte, bound)
	off := make([]uint32, n+1)
	if _, err := a.engine.Serve(b, now, out, off, meta); err != nil {
		return nil, err
	}
	resp := make([][]byte, n)
	for i := range resp {
		resp[i] = out[off[i]:off[i+1]]
	}
	return resp, nil
}

// call runs one closure; a panic is answered like the reference's panicRecovery (middleware/logger.go:91-114).
func (a *App) call(rt *route, c *Context) (rec []byte) {
	defer func() {
		if recover() != nil {
			rec = u32(nil, 0xFFFFFFFF)
		}
	}()
	data, err := rt.fn(c)
	return a.resultRecord(rt, data, err)
}

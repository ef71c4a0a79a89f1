// value_device.cuh — the generic value encoder of the serve path: encoding/json for the fields whose SHAPE depends on the
// row (pointers, slices, map[string]T, nested structs with omitempty members, bare schemas).
//
// Replaces, for those fields, what Responder.Respond reaches through json.NewEncoder(w).Encode(resp)
// (pkg/gofr/http/responder.go:32-40; Go 1.21 encoding/json encode.go: structEncoder, ptrEncoder, sliceEncoder /
// arrayEncoder, mapEncoder, floatEncoder).  Flat fields never come here: the table builder turns them into straight-line
// op programs (table_build.cpp struct_level_ops); one OP_VALUE op per shape-dependent field calls value_encode, which walks
// the field's share of the row (include/gofr_b200.h "Row format") with an explicit frame stack — device code does not
// recurse — of at most kMaxValueDepth struct levels (checked when the schema is added).
//
// Out of line and byte-at-a-time on purpose: it is the general path, the hot programs are literals and scalars.
// included by serve_device.cuh
#pragma once
#include "float_device.cuh"

namespace gofr {

enum ValueStatus : uint32_t {
    VAL_OK = 0,
    VAL_MALFORMED = 1,    // the row ends before the walk does: answered like a handler panic (as for flat rows)
    VAL_UNENCODABLE = 2,  // a float64 is NaN / ±Inf: json.Encoder.Encode fails, the response keeps headers and loses its body
};

GOFR_HD uint32_t ld32u(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
GOFR_HD uint64_t ld64u(const uint8_t* p) { return (uint64_t)ld32u(p) | (uint64_t)ld32u(p + 4) << 32; }

// float64 → text through the Writer; returns the length, 0 for NaN / ±Inf
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t emit_f64(Writer* w, uint64_t bits) {
    uint8_t buf[32];
    const uint32_t n = json_float64_text(bits, buf);
    if (EMIT)
        for (uint32_t k = 0; k < n; k++) w->put1(buf[k]);
    return n;
}

// float32 → text (floatEncoder with bits == 32: the shortest digits that identify the float32); 0 for NaN / ±Inf
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t emit_f32(Writer* w, uint32_t bits) {
    uint8_t buf[32];
    const uint32_t n = json_float32_text(bits, buf);
    if (EMIT)
        for (uint32_t k = 0; k < n; k++) w->put1(buf[k]);
    return n;
}
// uint64 → decimal (uintEncoder: strconv.AppendUint)
template <bool EMIT>
GOFR_HD uint32_t emit_u64_slow(Writer* w, uint64_t v) {
    uint8_t d[20];
    uint32_t n = 0;
    do { d[n++] = (uint8_t)('0' + (uint32_t)(v % 10)); v /= 10; } while (v);
    if (EMIT)
        for (uint32_t k = n; k-- > 0;) w->put1(d[k]);
    return n;
}
// []byte → base64.StdEncoding with padding (encodeByteSlice), without the quotes
template <bool EMIT>
GOFR_HD uint32_t emit_base64(Writer* w, const uint8_t* p, uint32_t n) {
    if (EMIT) {
        auto ch = [](uint32_t x) -> uint32_t { return x < 26 ? 'A' + x : x < 52 ? 'a' + (x - 26) : x < 62 ? '0' + (x - 52) : x == 62 ? '+' : '/'; };
        uint32_t i = 0;
        for (; i + 3 <= n; i += 3) {
            const uint32_t t = (uint32_t)p[i] << 16 | (uint32_t)p[i + 1] << 8 | p[i + 2];
            w->put1(ch(t >> 18)); w->put1(ch((t >> 12) & 63)); w->put1(ch((t >> 6) & 63)); w->put1(ch(t & 63));
        }
        if (n - i == 1) {
            const uint32_t t = (uint32_t)p[i] << 16;
            w->put1(ch(t >> 18)); w->put1(ch((t >> 12) & 63)); w->put1('='); w->put1('=');
        } else if (n - i == 2) {
            const uint32_t t = (uint32_t)p[i] << 16 | (uint32_t)p[i + 1] << 8;
            w->put1(ch(t >> 18)); w->put1(ch((t >> 12) & 63)); w->put1(ch((t >> 6) & 63)); w->put1('=');
        }
    }
    return (n + 2) / 3 * 4;
}
// time.Time → Time.MarshalJSON's text, quotes included (time/format_rfc3339.go appendFormatRFC3339 + appendStrictRFC3339,
// Go 1.21): "2006-01-02T15:04:05[.fraction]" + "Z" | ±hh:mm.  sec = Unix seconds, off = zone offset in seconds.
// Returns the length; 0 when MarshalJSON fails (year outside [0, 9999], zone hour outside [0, 23]) — and for words no Time
// can hold (nanoseconds >= 1e9): the CONTENT of a scalar never makes a row "malformed", only its framing does, so that
// "what ends the walk first" does not depend on the order values are looked at (map values are written in key order).
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t emit_time_json(Writer* w, int64_t sec, uint32_t nsec, int32_t off) {
    if (nsec >= 1000000000u) return 0;
    if (sec < -70000000000ll || sec > 300000000000ll) return 0;  // far outside years [0, 9999] (and sec + off cannot overflow)
    const int64_t local = sec + off;
    // 0000-01-01T00:00:00 .. 9999-12-31T23:59:59 in the zone's wall clock; the zone itself: |offset| < 24 h
    if (local < -62167219200ll || local >= 253402300800ll) return 0;
    int32_t zone = off / 60;  // minutes, truncated toward zero like Go
    const bool zneg = zone < 0;
    if (zneg) zone = -zone;
    if (zone / 60 >= 24) return 0;
    int64_t days = local / 86400;
    int32_t sod = (int32_t)(local - days * 86400);
    if (sod < 0) { sod += 86400; days -= 1; }
    // days since 1970-01-01 → proleptic Gregorian date (the usual era / day-of-era arithmetic)
    const int32_t z = (int32_t)days + 719468;
    const int32_t era = (z >= 0 ? z : z - 146096) / 146097;
    const uint32_t doe = (uint32_t)(z - era * 146097);
    const uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const uint32_t mp = (5 * doy + 2) / 153;
    const uint32_t day = doy - (153 * mp + 2) / 5 + 1, month = mp < 10 ? mp + 3 : mp - 9;
    const uint32_t year = (uint32_t)((int32_t)yoe + era * 400) + (month <= 2 ? 1u : 0u);
    uint32_t fd = 0;  // digits of the fraction that survive the trailing-zero trim
    if (nsec) { fd = 9; for (uint32_t t = nsec; t % 10 == 0; t /= 10) fd--; }
    const uint32_t total = 2 + 19 + (fd ? 1 + fd : 0) + (off == 0 ? 1 : 6);
    if (!EMIT) return total;
    auto two = [&](uint32_t x) { w->put1('0' + x / 10); w->put1('0' + x % 10); };
    w->put1('"');
    two(year / 100); two(year % 100); w->put1('-'); two(month); w->put1('-'); two(day); w->put1('T');
    two((uint32_t)sod / 3600); w->put1(':'); two((uint32_t)sod / 60 % 60); w->put1(':'); two((uint32_t)sod % 60);
    if (fd) {
        w->put1('.');
        uint32_t div = 100000000u;
        for (uint32_t k = 0; k < fd; k++, div /= 10) w->put1('0' + nsec / div % 10);
    }
    if (off == 0) w->put1('Z');
    else { w->put1(zneg ? '-' : '+'); two((uint32_t)zone / 60); w->put1(':'); two((uint32_t)zone % 60); }
    w->put1('"');
    return total;
}
// bytes a scalar kind owns (fixed words, or as an element of a slice / map)
GOFR_HD uint32_t value_scalar_bytes(uint32_t kind) {
    return kind == GOFR_F_TIME ? 16u : (kind == GOFR_F_INT64 || kind == GOFR_F_INT || kind == GOFR_F_FLOAT64 || kind == GOFR_F_UINT64) ? 8u : 4u;
}

// isEmptyValue (encode.go) of a field whose fixed words are at p: false, 0, 0.0 of either sign, "", nil pointer, nil or
// empty slice / map; a struct value is never empty
GOFR_HD bool value_field_empty(uint32_t kind, uint32_t container, const uint8_t* p) {
    if (container == GOFR_C_PTR) return ld32u(p) == 0;
    if (container == GOFR_C_SLICE || container == GOFR_C_MAP || container == GOFR_C_SLICE_PTR) { const uint32_t n = ld32u(p); return n == 0 || n == GOFR_NIL_COUNT; }
    if (kind == GOFR_F_INT64 || kind == GOFR_F_INT || kind == GOFR_F_UINT64) return ld64u(p) == 0;
    if (kind == GOFR_F_FLOAT64) return (ld64u(p) << 1) == 0;
    if (kind == GOFR_F_FLOAT32) return (ld32u(p) << 1) == 0;
    if (kind == GOFR_F_BYTES) { const uint32_t n = ld32u(p); return n == 0 || n == GOFR_NIL_COUNT; }  // len(v) == 0: nil or empty
    if (kind == GOFR_F_STRUCT || kind == GOFR_F_TIME) return false;  // isEmptyValue knows no empty struct (time.Time is one)
    return ld32u(p) == 0;  // INT32, BOOL, STRING (its length)
}

struct ValueFrame {
    const uint8_t* fixed;  // fixed part of the struct this frame walks
    uint32_t slice_left;   // elements still to come of the slice of structs the frame is in the middle of
    uint16_t schema;
    uint16_t slice_elem;   // schema of those elements
    uint8_t next_field;
    uint8_t first;         // no member written yet (comma logic)
    uint8_t in_slice;      // 1: first element pending, 2: later elements
    uint8_t single;        // root frame: exactly one field, no braces, no key (the OP_VALUE field itself)
    uint8_t slice_ptr;     // the slice is a []*T: a presence word before every element
};

// Encodes field `fidx` of schema `sidx` whose fixed words are at `fixed`; its variable bytes start at `var` (avail bytes
// left in the data section).  EMIT=false only counts.  Returns the bytes produced; *consumed = variable bytes walked,
// *status = ValueStatus.
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t value_encode(Writer* w, const TableView tv, uint32_t sidx, uint32_t fidx, const uint8_t* fixed,
                                       const uint8_t* var, uint32_t avail, uint32_t* consumed, uint32_t* status) {
    ValueFrame st[kMaxValueDepth + 1];
    int depth = 0;
    uint32_t out = 0, err = VAL_OK;
    const uint8_t* const var0 = var;
    const uint8_t* const end = var + avail;
    const SchemaRec* const schemas = tv.schemas();

    auto fields_of = [&](uint32_t s) { return (const FieldRec*)(tv.base + schemas[s].fields_off); };
    auto take = [&](uint32_t n) -> const uint8_t* {
        if ((uint32_t)(end - var) < n) { err = VAL_MALFORMED; return nullptr; }
        const uint8_t* p = var;
        var += n;
        return p;
    };
    auto put_c = [&](uint32_t c) { if (EMIT) w->put1(c); out += 1; };
    auto put_bytes = [&](const uint8_t* p, uint32_t n) {
        if (EMIT) for (uint32_t k = 0; k < n; k++) w->put1(p[k]);
        out += n;
    };
    auto put_null = [&]() { if (EMIT) { w->reserve_out(1); w->put4('n' | 'u' << 8 | 'l' << 16 | 'l' << 24); } out += 4; };
    auto string_val = [&](const uint8_t* s, uint32_t len) {
        put_c('"');
        out += json_escape_slow<EMIT>(w, s, len);
        put_c('"');
    };
    auto scalar = [&](uint32_t kind, const uint8_t* p) {
        if (kind == GOFR_F_INT64 || kind == GOFR_F_INT || kind == GOFR_F_INT32) {
            const int64_t v = kind == GOFR_F_INT32 ? (int64_t)(int32_t)ld32u(p) : (int64_t)ld64u(p);
            if (EMIT) w->reserve_out(6);
            out += emit_i64<EMIT>(w, v);
        } else if (kind == GOFR_F_BOOL) {
            const bool t = ld32u(p) != 0;
            if (EMIT) { w->reserve_out(2); if (t) w->put4('t' | 'r' << 8 | 'u' << 16 | 'e' << 24); else { w->put4('f' | 'a' << 8 | 'l' << 16 | 's' << 24); w->putc('e'); } }
            out += t ? 4u : 5u;
        } else if (kind == GOFR_F_FLOAT64) {
            const uint32_t n = emit_f64<EMIT>(w, ld64u(p));
            if (!n) err = VAL_UNENCODABLE;
            out += n;
        } else if (kind == GOFR_F_FLOAT32) {
            const uint32_t n = emit_f32<EMIT>(w, ld32u(p));
            if (!n) err = VAL_UNENCODABLE;
            out += n;
        } else if (kind == GOFR_F_UINT64) {
            out += emit_u64_slow<EMIT>(w, ld64u(p));
        } else if (kind == GOFR_F_TIME) {
            const uint32_t n = emit_time_json<EMIT>(w, (int64_t)ld64u(p), ld32u(p + 8), (int32_t)ld32u(p + 12));
            if (!n) err = VAL_UNENCODABLE;
            out += n;
        } else err = VAL_MALFORMED;
    };
    // a []byte whose length word (GOFR_NIL_COUNT: nil) has been read: null, or the base64 text in quotes
    auto bytes_val = [&](uint32_t len) {
        if (len == GOFR_NIL_COUNT) { put_null(); return; }
        const uint8_t* s = take(len);
        if (!s) return;
        put_c('"');
        out += emit_base64<EMIT>(w, s, len);
        put_c('"');
    };
    auto push_struct = [&](uint32_t schema, const uint8_t* fx) {
        if (depth > kMaxValueDepth) { err = VAL_MALFORMED; return; }
        ValueFrame& f = st[depth++];
        f.fixed = fx; f.slice_left = 0; f.schema = (uint16_t)schema; f.slice_elem = 0;
        f.next_field = 0; f.first = 1; f.in_slice = 0; f.single = 0; f.slice_ptr = 0;
        put_c('{');
    };
    // E(T) of a STRING / scalar element, taken from the variable part (struct elements go through push_struct)
    auto leaf_element = [&](uint32_t kind) {
        if (kind == GOFR_F_STRING) {
            const uint8_t* lp = take(4);
            if (!lp) return;
            const uint32_t len = ld32u(lp);
            const uint8_t* s = take(len);
            if (s) string_val(s, len);
        } else if (kind == GOFR_F_BYTES) {
            const uint8_t* lp = take(4);
            if (lp) bytes_val(ld32u(lp));
        } else {
            const uint8_t* p = take(value_scalar_bytes(kind));
            if (p) scalar(kind, p);
        }
    };
    // T whose fixed words are at p
    auto plain = [&](const FieldRec& F, const uint8_t* p) {
        if (F.kind == GOFR_F_STRING) {
            const uint32_t len = ld32u(p);
            const uint8_t* s = take(len);
            if (s) string_val(s, len);
        } else if (F.kind == GOFR_F_BYTES) bytes_val(ld32u(p));
        else if (F.kind == GOFR_F_STRUCT) push_struct(F.elem, p);
        else scalar(F.kind, p);
    };
    // map[string]T, T a string or a scalar: entries (u32 key length, key, E(T)) in the row's order; encoding/json sorts
    // the keys bytewise.  No scratch memory: n selection passes over the entries, each picking the smallest key above
    // the one written before (ties by position, so a malformed row with a repeated key still matches the oracle).
    auto map_value = [&](uint32_t kind, uint32_t n) {
        const uint8_t* const first_entry = var;
        auto skip_entry = [&](const uint8_t*& q, const uint8_t*& key, uint32_t& klen, const uint8_t*& val) -> bool {
            if ((uint32_t)(end - q) < 4u) return false;
            klen = ld32u(q); q += 4;
            if ((uint32_t)(end - q) < klen) return false;
            key = q; q += klen;
            val = q;
            if (kind == GOFR_F_STRING || kind == GOFR_F_BYTES) {
                if ((uint32_t)(end - q) < 4u) return false;
                uint32_t vlen = ld32u(q); q += 4;
                if (kind == GOFR_F_BYTES && vlen == GOFR_NIL_COUNT) vlen = 0;  // a nil []byte owns no bytes
                if ((uint32_t)(end - q) < vlen) return false;
                q += vlen;
            } else {
                const uint32_t vb = value_scalar_bytes(kind);
                if ((uint32_t)(end - q) < vb) return false;
                q += vb;
            }
            return true;
        };
        {   // validate once and find where the map ends
            const uint8_t* q = first_entry;
            for (uint32_t i = 0; i < n; i++) {
                const uint8_t *k, *v;
                uint32_t kl;
                if (!skip_entry(q, k, kl, v)) { err = VAL_MALFORMED; return; }
            }
            var = q;
        }
        auto less = [&](const uint8_t* a, uint32_t an, uint32_t ai, const uint8_t* b, uint32_t bn, uint32_t bi) -> bool {
            const uint32_t m = an < bn ? an : bn;
            for (uint32_t k = 0; k < m; k++)
                if (a[k] != b[k]) return a[k] < b[k];
            if (an != bn) return an < bn;
            return ai < bi;
        };
        put_c('{');
        const uint8_t* last_key = nullptr;
        uint32_t last_len = 0, last_idx = 0;
        for (uint32_t done = 0; done < n && err == VAL_OK; done++) {
            const uint8_t *best_key = nullptr, *best_val = nullptr;
            uint32_t best_len = 0, best_idx = 0;
            const uint8_t* q = first_entry;
            for (uint32_t i = 0; i < n; i++) {
                const uint8_t *k, *v;
                uint32_t kl;
                skip_entry(q, k, kl, v);
                if (last_key && !less(last_key, last_len, last_idx, k, kl, i)) continue;  // written already
                if (!best_key || less(k, kl, i, best_key, best_len, best_idx)) { best_key = k; best_len = kl; best_idx = i; best_val = v; }
            }
            if (done) put_c(',');
            string_val(best_key, best_len);
            put_c(':');
            const uint8_t* save = var;
            var = best_val;  // leaf_element reads at the cursor; the map's extent was validated above
            leaf_element(kind);
            var = save;
            last_key = best_key; last_len = best_len; last_idx = best_idx;
        }
        put_c('}');
    };
    // the value of field F of frame fi (fixed words at p); may push a struct frame or start a slice of structs
    auto field_value = [&](const FieldRec& F, const uint8_t* p, int fi) {
        if (F.container == GOFR_C_VALUE) { plain(F, p); return; }
        if (F.container == GOFR_C_PTR) {
            if (!ld32u(p)) put_null(); else plain(F, p + 4);
            return;
        }
        const uint32_t n = ld32u(p);
        if (n == GOFR_NIL_COUNT) { put_null(); return; }
        if (F.container == GOFR_C_MAP) {
            if (F.kind == GOFR_F_STRUCT) { err = VAL_MALFORMED; return; }
            map_value(F.kind, n);
            return;
        }
        const bool eptr = F.container == GOFR_C_SLICE_PTR;
        put_c('[');
        if (F.kind == GOFR_F_STRUCT) {
            if (n == 0) { put_c(']'); return; }
            st[fi].slice_left = n; st[fi].slice_elem = F.elem; st[fi].in_slice = 1; st[fi].slice_ptr = eptr;
            return;
        }
        if ((uint32_t)(end - var) / 4u < n) { err = VAL_MALFORMED; return; }  // every element owns at least a word
        for (uint32_t i = 0; i < n && err == VAL_OK; i++) {
            if (i) put_c(',');
            if (eptr) {  // []*T: nil elements are null
                const uint8_t* pw = take(4);
                if (!pw) break;
                if (!ld32u(pw)) { put_null(); continue; }
            }
            leaf_element(F.kind);
        }
        put_c(']');
    };

    {   // root frame: the one field the op names
        const FieldRec& F = fields_of(sidx)[fidx];
        ValueFrame& r = st[depth++];
        r.fixed = fixed - (size_t)F.word * 4; r.slice_left = 0; r.schema = (uint16_t)sidx; r.slice_elem = 0;
        r.next_field = (uint8_t)fidx; r.first = 1; r.in_slice = 0; r.single = 1; r.slice_ptr = 0;
    }
    while (depth > 0 && err == VAL_OK) {
        const int fi = depth - 1;
        if (st[fi].slice_left) {  // next element of a slice of structs
            if (st[fi].in_slice == 2) put_c(',');
            st[fi].in_slice = 2;
            st[fi].slice_left--;
            if (st[fi].slice_ptr) {  // []*T: a nil element is null and owns nothing else
                const uint8_t* pw = take(4);
                if (!pw) continue;
                if (!ld32u(pw)) { put_null(); continue; }
            }
            const uint32_t es = st[fi].slice_elem;
            const uint8_t* fx = take((uint32_t)schemas[es].fixed_words * 4u);
            if (fx) push_struct(es, fx);
            continue;
        }
        if (st[fi].in_slice) { put_c(']'); st[fi].in_slice = 0; }
        const SchemaRec& S = schemas[st[fi].schema];
        if (st[fi].single ? st[fi].first == 0 : st[fi].next_field >= S.n_fields) {
            if (!st[fi].single) put_c('}');
            depth--;
            continue;
        }
        const FieldRec& F = fields_of(st[fi].schema)[st[fi].next_field];
        const uint8_t* p = st[fi].fixed + (size_t)F.word * 4;
        st[fi].next_field++;
        if (st[fi].single) st[fi].first = 0;
        else {
            // an empty value owns no bytes of the variable part (zero-length string, nil pointer, no elements)
            if (F.omitempty && value_field_empty(F.kind, F.container, p)) continue;
            if (!st[fi].first) put_c(',');
            st[fi].first = 0;
            put_bytes(tv.lit_bytes(F.key_off), F.key_len);
        }
        field_value(F, p, fi);
    }
    *consumed = (uint32_t)(var - var0);
    *status = err;
    return out;
}

// OP_KEY's emptiness test for the kinds of the wider data model (kind | container << 4), out of line: the interpreter's
// own code stays what it was for flat structs
GOFR_HD_NOINLINE bool value_key_empty(uint32_t okind, const uint8_t* p) { return value_field_empty(okind & 15u, okind >> 4, p); }

// OP_F64 / OP_VALUE of the interpreter (serve_device.cuh run_prog) behind ONE out-of-line call.  raw = the op; row = the
// row's fixed part; [data, data + data_len) the data section, `used` bytes of it consumed so far (fixed part + variable
// cursor).  Returns the bytes produced; *consumed = bytes of the variable part walked; *status = ValueStatus.
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t value_op(Writer* w, const TableView tv, const uint4 raw, const uint32_t* row, const uint8_t* data,
                                   uint32_t data_len, uint32_t used, uint32_t* consumed, uint32_t* status) {
    *consumed = 0;
    *status = VAL_OK;
    if ((raw.x & 0xFFu) == OP_F64) {
        const uint32_t n = emit_f64<EMIT>(w, (uint64_t)row[raw.z] | (uint64_t)row[raw.z + 1] << 32);
        if (!n) *status = VAL_UNENCODABLE;
        return n;
    }
    if (used > data_len) { *status = VAL_MALFORMED; return 0; }
    return value_encode<EMIT>(w, tv, raw.w, (raw.x >> 8) & 0xFFu, (const uint8_t*)(row + raw.z), data + used, data_len - used, consumed, status);
}

}  // namespace gofr

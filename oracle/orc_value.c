/* orc_value.c — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h).
 *
 * encoding/json (Go 1.21, go.mod:3) for the values Responder.Respond hands to json.NewEncoder(w).Encode
 * (pkg/gofr/http/responder.go:32-40) beyond flat structs: float64, nested structs, pointers, slices and
 * map[string]T — walked straight off the handler-result row (include/gofr_b200.h "Row format").  Follows
 *   encoding/json/encode.go: structEncoder (declaration order, omitempty via isEmptyValue), ptrEncoder (nil → null),
 *   sliceEncoder / arrayEncoder (nil → null, else [a,b]), mapEncoder (nil → null, keys sorted with sort.Slice on the
 *   key STRINGS, i.e. bytewise), floatEncoder (below), boolEncoder, intEncoder, stringEncoder (orc_enc_string).
 * The float text does NOT share the product's algorithm: the product computes shortest digits with Ryu
 * (gofr_b200/csrc/float_device.cuh); here they are found by trying precisions with the C library's correctly rounded
 * printf / strtod, and both are checked against Python's repr() in tests/test_values.py.
 */
#include <math.h>
#include <stdio.h>
#include <time.h>

#include "orc_internal.h"

typedef struct orc_table orc_table;

/* strconv.FormatFloat(f, 'g'-like switch, -1, 64) as encoding/json's floatEncoder uses it (encode.go floatEncoder.encode):
 *   fmt = 'f'; if abs != 0 && (abs < 1e-6 || abs >= 1e21) fmt = 'e'; AppendFloat(b, f, fmt, -1, 64);
 *   then for 'e': "e-0X" → "e-X" (n >= 4 && b[n-4]=='e' && b[n-3]=='-' && b[n-2]=='0').
 * Returns the length, 0 for NaN / ±Inf (UnsupportedValueError). */
/* f32: ax holds a float32 value exactly and "the same float" means the same float32 (bitSize 32 of AppendFloat) */
static int round_trips_w(const char* s, double x, int f32) { return f32 ? strtof(s, NULL) == (float)x : strtod(s, NULL) == x; }
#define round_trips(s, x) round_trips_w(s, x, f32)

/* shortest decimal digits (no dot, no leading zeros beyond a single 0) and the decimal exponent of the first digit */
static void shortest_digits(double ax, char* digits, int* nd, int* exp10, int f32) {
    char buf[64];
    const int maxp = f32 ? 9 : 17;
    for (int p = 1; p <= maxp; p++) {
        snprintf(buf, sizeof buf, "%.*e", p - 1, ax);
        int ok = round_trips(buf, ax);
        if (!ok && p < maxp) {
            /* the interval of decimals that parse back to ax is not symmetric around it at a power of two: a
             * p-digit decimal other than the correctly rounded one may still lie inside.  Try its two neighbours. */
            char m[32];
            int e = 0, k = 0;
            for (const char* q = buf; *q && *q != 'e'; q++) if (*q != '.') m[k++] = *q;
            m[k] = 0;
            sscanf(strchr(buf, 'e') + 1, "%d", &e);
            for (int dir = -1; dir <= 1 && !ok; dir += 2) {
                char c[32];
                memcpy(c, m, (size_t)k + 1);
                int ce = e, i = k - 1;
                if (dir > 0) {
                    while (i >= 0 && c[i] == '9') c[i--] = '0';
                    if (i < 0) { memmove(c + 1, c, (size_t)k); c[0] = '1'; c[k] = 0; ce++; } else c[i]++;
                } else {
                    while (i >= 0 && c[i] == '0') c[i--] = '9';
                    if (i < 0) continue;
                    c[i]--;
                    if (c[0] == '0') continue; /* lost a digit: not a p-digit decimal any more */
                }
                char t[64];
                snprintf(t, sizeof t, "%c.%se%d", c[0], c + 1, ce);
                if (round_trips(t, ax)) { ok = 1; snprintf(buf, sizeof buf, "%s", t); }
            }
        }
        if (ok) break;
    }
    /* buf = d[.ddd]e[+-]XX */
    int k = 0, e = 0;
    for (const char* q = buf; *q && *q != 'e'; q++) if (*q != '.') digits[k++] = *q;
    sscanf(strchr(buf, 'e') + 1, "%d", &e);
    while (k > 1 && digits[k - 1] == '0') k--; /* %e pads to the precision asked for */
    digits[k] = 0;
    *nd = k;
    *exp10 = e;
}

static int float_text_w(double x, char* out, int f32);
int orc_float_text(double x, char* out) { return float_text_w(x, out, 0); }
/* floatEncoder with bits == 32: the cutoffs are compared as float32 ("Must use float32 comparisons for underlying float32
 * value to get precise cutoffs right") */
int orc_float32_text(float x, char* out) { return float_text_w((double)x, out, 1); }
static int float_text_w(double x, char* out, int f32) {
    if (isnan(x) || isinf(x)) return 0;
    int n = 0;
    if (signbit(x)) out[n++] = '-';
    double ax = fabs(x);
    if (ax == 0) { out[n++] = '0'; return n; }
    char d[32];
    int nd, e;
    shortest_digits(ax, d, &nd, &e, f32);
    if (f32 ? ((float)ax < 1e-6f || (float)ax >= 1e21f) : (ax < 1e-6 || ax >= 1e21)) {
        /* %e: d.ddde±XX, at least two exponent digits */
        out[n++] = d[0];
        if (nd > 1) { out[n++] = '.'; memcpy(out + n, d + 1, (size_t)nd - 1); n += nd - 1; }
        out[n++] = 'e';
        out[n++] = e < 0 ? '-' : '+';
        int ae = e < 0 ? -e : e;
        char eb[8];
        int en = snprintf(eb, sizeof eb, "%02d", ae);
        memcpy(out + n, eb, (size_t)en);
        n += en;
        /* clean up e-09 to e-9 */
        if (n >= 4 && out[n - 4] == 'e' && out[n - 3] == '-' && out[n - 2] == '0') { out[n - 2] = out[n - 1]; n--; }
        return n;
    }
    /* %f with the shortest digits */
    if (e < 0) {
        out[n++] = '0'; out[n++] = '.';
        for (int k = 0; k < -e - 1; k++) out[n++] = '0';
        memcpy(out + n, d, (size_t)nd); n += nd;
        return n;
    }
    for (int k = 0; k <= e; k++) out[n++] = k < nd ? d[k] : '0';
    if (nd > e + 1) { out[n++] = '.'; memcpy(out + n, d + e + 1, (size_t)(nd - e - 1)); n += nd - e - 1; }
    return n;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* row walker                                                                                                    */
/* ------------------------------------------------------------------------------------------------------------ */

typedef struct {
    const orc_table* t;
    const uint8_t* var; /* cursor in the variable part */
    const uint8_t* end; /* end of the data section */
    int failed;         /* UnsupportedValueError somewhere: Encode writes nothing */
    int malformed;
} walk;
/* The walk stops at the FIRST thing that ends it — a value encoding/json cannot write, or a row that ends too early.  (Go
 * stops at the first UnsupportedValueError as well; for malformed rows there is no reference behaviour, and what matters
 * is that the product and this file answer alike: what comes first in field order wins.  Inside ONE map the entries'
 * framing is checked as a whole before any value is looked at, because that is how the device code has to do it.) */
#define WALK_OVER(w) ((w)->malformed || (w)->failed)

static uint32_t rd32u(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint64_t rd64u(const uint8_t* p) { return (uint64_t)rd32u(p) | (uint64_t)rd32u(p + 4) << 32; }

static int kind_words(int kind) { return kind == F_TIME ? 4 : (kind == F_INT64 || kind == F_INT || kind == F_FLOAT64 || kind == F_UINT64) ? 2 : 1; }

/* encodeByteSlice: base64.StdEncoding (padding) in quotes; a nil slice is null */
static void enc_bytes(obuf* b, walk* w, uint32_t len) {
    static const char A[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    if (len == 0xFFFFFFFFu) { ob_puts(b, "null"); return; }
    if ((size_t)(w->end - w->var) < len) { w->malformed = 1; return; }
    const uint8_t* s = w->var;
    w->var += len;
    ob_putc(b, '"');
    uint32_t i = 0;
    for (; i + 3 <= len; i += 3) {
        uint32_t t = (uint32_t)s[i] << 16 | (uint32_t)s[i + 1] << 8 | s[i + 2];
        ob_putc(b, (uint8_t)A[t >> 18]); ob_putc(b, (uint8_t)A[(t >> 12) & 63]); ob_putc(b, (uint8_t)A[(t >> 6) & 63]); ob_putc(b, (uint8_t)A[t & 63]);
    }
    if (len - i) {
        uint32_t t = (uint32_t)s[i] << 16 | (len - i == 2 ? (uint32_t)s[i + 1] << 8 : 0u);
        ob_putc(b, (uint8_t)A[t >> 18]); ob_putc(b, (uint8_t)A[(t >> 12) & 63]);
        ob_putc(b, len - i == 2 ? (uint8_t)A[(t >> 6) & 63] : (uint8_t)'=');
        ob_putc(b, '=');
    }
    ob_putc(b, '"');
}

/* words of the fixed part of a struct / of one field */
int orc_schema_fixed_words(const orc_table* t, const orc_schema* sc);
static int field_fixed_words(const orc_table* t, const orc_field* f) {
    if (f->container == C_SLICE || f->container == C_MAP || f->container == C_SLICE_PTR) return 1;
    int w = 0;
    if (f->kind == F_STRUCT) {
        const orc_schema* es = orc_find_schema(t, f->elem_schema);
        w = es ? orc_schema_fixed_words(t, es) : 0;
    } else w = kind_words(f->kind);
    return w + (f->container == C_PTR ? 1 : 0);
}
int orc_schema_fixed_words(const orc_table* t, const orc_schema* sc) {
    int w = 0;
    for (int i = 0; i < sc->n_fields; i++) w += field_fixed_words(t, &sc->f[i]);
    return w;
}

static void enc_struct(obuf* b, walk* w, const orc_schema* sc, const uint8_t* fixed);

/* take n bytes of the variable part */
static const uint8_t* take(walk* w, size_t n) {
    if ((size_t)(w->end - w->var) < n) { w->malformed = 1; return NULL; }
    const uint8_t* p = w->var;
    w->var += n;
    return p;
}

/* a scalar whose words sit at p (fixed part, or taken from the variable part for elements) */
static void enc_scalar(obuf* b, walk* w, int kind, const uint8_t* p) {
    switch (kind) {
        case F_INT64: case F_INT: orc_enc_int(b, (int64_t)rd64u(p)); break;
        case F_INT32: orc_enc_int(b, (int32_t)rd32u(p)); break;
        case F_BOOL: ob_puts(b, rd32u(p) ? "true" : "false"); break;
        case F_FLOAT64: {
            uint64_t bits = rd64u(p);
            double x;
            memcpy(&x, &bits, 8);
            char tmp[40];
            int n = orc_float_text(x, tmp);
            if (!n) w->failed = 1;
            ob_put(b, tmp, (size_t)n);
            break;
        }
        case F_FLOAT32: {
            uint32_t bits = rd32u(p);
            float x;
            memcpy(&x, &bits, 4);
            char tmp[40];
            int n = orc_float32_text(x, tmp);
            if (!n) w->failed = 1;
            ob_put(b, tmp, (size_t)n);
            break;
        }
        case F_TIME: { /* Time.MarshalJSON (time/time.go, format_rfc3339.go; Go 1.21): appendFormatRFC3339(b, true) between
                        * quotes, then appendStrictRFC3339's checks: year exactly four digits wide, zone hour < 24.  The civil
                        * fields come from the C library's gmtime_r of the zone's wall clock. */
            int64_t sec = (int64_t)rd64u(p);
            uint32_t nsec = rd32u(p + 8);
            int32_t off = (int32_t)rd32u(p + 12);
            if (nsec >= 1000000000u) { w->failed = 1; break; } /* no Time holds this: answered like a Time that cannot be marshalled
                                                                 * (a scalar's content never makes the ROW malformed: see WALK_OVER) */
            if (sec < -70000000000ll || sec > 300000000000ll) { w->failed = 1; break; } /* far outside [0, 9999]; no overflow below */
            time_t local = (time_t)(sec + off);
            struct tm g;
            if (!gmtime_r(&local, &g)) { w->failed = 1; break; }
            long year = (long)g.tm_year + 1900;
            int zone = off / 60, zneg = zone < 0;
            if (zneg) zone = -zone;
            if (year < 0 || year > 9999 || zone / 60 >= 24) { w->failed = 1; break; }
            char tmp[64];
            int n = snprintf(tmp, sizeof tmp, "\"%04ld-%02d-%02dT%02d:%02d:%02d", year, g.tm_mon + 1, g.tm_mday, g.tm_hour, g.tm_min, g.tm_sec);
            if (nsec) {
                char fr[16];
                snprintf(fr, sizeof fr, "%09u", nsec);
                int k = 9;
                while (fr[k - 1] == '0') k--;
                tmp[n++] = '.';
                memcpy(tmp + n, fr, (size_t)k);
                n += k;
            }
            if (off == 0) tmp[n++] = 'Z';
            else n += snprintf(tmp + n, sizeof tmp - (size_t)n, "%c%02d:%02d", zneg ? '-' : '+', zone / 60, zone % 60);
            tmp[n++] = '"';
            ob_put(b, tmp, (size_t)n);
            break;
        }
        case F_UINT64: { /* uintEncoder: strconv.AppendUint(b, v.Uint(), 10) */
            char tmp[24];
            int n = snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)rd64u(p));
            ob_put(b, tmp, (size_t)n);
            break;
        }
        default: w->malformed = 1;
    }
}

/* E(T): an element of a slice / map, entirely in the variable part */
static void enc_element(obuf* b, walk* w, const orc_field* f) {
    if (f->kind == F_STRING) {
        const uint8_t* lp = take(w, 4);
        if (!lp) return;
        uint32_t len = rd32u(lp);
        const uint8_t* s = take(w, len);
        if (!s) return;
        orc_enc_string(b, s, len);
    } else if (f->kind == F_BYTES) {
        const uint8_t* lp = take(w, 4);
        if (lp) enc_bytes(b, w, rd32u(lp));
    } else if (f->kind == F_STRUCT) {
        const orc_schema* es = orc_find_schema(w->t, f->elem_schema);
        if (!es) { w->malformed = 1; return; }
        const uint8_t* fx = take(w, (size_t)orc_schema_fixed_words(w->t, es) * 4);
        if (!fx) return;
        enc_struct(b, w, es, fx);
    } else {
        const uint8_t* p = take(w, (size_t)kind_words(f->kind) * 4);
        if (p) enc_scalar(b, w, f->kind, p);
    }
}

/* T whose fixed words sit at p */
static void enc_plain(obuf* b, walk* w, const orc_field* f, const uint8_t* p) {
    if (f->kind == F_STRING) {
        uint32_t len = rd32u(p);
        const uint8_t* s = take(w, len);
        if (s) orc_enc_string(b, s, len);
    } else if (f->kind == F_BYTES) {
        enc_bytes(b, w, rd32u(p));
    } else if (f->kind == F_STRUCT) {
        const orc_schema* es = orc_find_schema(w->t, f->elem_schema);
        if (!es) { w->malformed = 1; return; }
        enc_struct(b, w, es, p);
    } else enc_scalar(b, w, f->kind, p);
}

static int key_less(const uint8_t* a, uint32_t an, const uint8_t* b, uint32_t bn) {
    uint32_t m = an < bn ? an : bn;
    int c = m ? memcmp(a, b, m) : 0;
    return c < 0 || (c == 0 && an < bn);
}

/* the value of field f (fixed words at p) */
static void enc_field_value(obuf* b, walk* w, const orc_field* f, const uint8_t* p) {
    switch (f->container) {
        case C_VALUE: enc_plain(b, w, f, p); break;
        case C_PTR:
            if (!rd32u(p)) ob_puts(b, "null"); else enc_plain(b, w, f, p + 4);
            break;
        case C_SLICE_PTR: /* []*T: ptrEncoder per element — nil → null */
        case C_SLICE: {
            uint32_t n = rd32u(p);
            if (n == 0xFFFFFFFFu) { ob_puts(b, "null"); break; }
            /* a count that cannot fit (every scalar or string element owns at least a word) is malformed before any element
             * is looked at; slices of structs are simply walked */
            if (f->kind != F_STRUCT && (size_t)(w->end - w->var) / 4 < n) { w->malformed = 1; break; }
            ob_putc(b, '[');
            for (uint32_t i = 0; i < n && !WALK_OVER(w); i++) {
                if (i) ob_putc(b, ',');
                if (f->container == C_SLICE_PTR) {
                    const uint8_t* pw = take(w, 4);
                    if (!pw) break;
                    if (!rd32u(pw)) { ob_puts(b, "null"); continue; }
                }
                enc_element(b, w, f);
            }
            ob_putc(b, ']');
            break;
        }
        case C_MAP: {
            uint32_t n = rd32u(p);
            if (n == 0xFFFFFFFFu) { ob_puts(b, "null"); break; }
            if (f->kind == F_STRUCT || (size_t)(w->end - w->var) / 4 < n) { w->malformed = 1; break; }
            /* collect the entries, encode each value, then emit in key order (mapEncoder: sort.Slice by key string) */
            typedef struct { const uint8_t* k; uint32_t kn; obuf v; } ent;
            ent* es = (ent*)calloc(n ? n : 1, sizeof(ent));
            uint32_t got = 0;
            for (; got < n && !w->malformed; got++) {
                const uint8_t* lp = take(w, 4);
                if (!lp) break;
                es[got].kn = rd32u(lp);
                es[got].k = take(w, es[got].kn);
                if (!es[got].k) break;
                ob_init(&es[got].v);
                enc_element(&es[got].v, w, f);
            }
            if (!w->malformed) {
                for (uint32_t i = 1; i < n; i++) { /* insertion sort: stable, maps are small */
                    ent x = es[i];
                    uint32_t j = i;
                    while (j > 0 && key_less(x.k, x.kn, es[j - 1].k, es[j - 1].kn)) { es[j] = es[j - 1]; j--; }
                    es[j] = x;
                }
                ob_putc(b, '{');
                for (uint32_t i = 0; i < n; i++) {
                    if (i) ob_putc(b, ',');
                    orc_enc_string(b, es[i].k, es[i].kn);
                    ob_putc(b, ':');
                    ob_put(b, es[i].v.p, es[i].v.n);
                }
                ob_putc(b, '}');
            }
            for (uint32_t i = 0; i < n; i++) ob_free(&es[i].v);
            free(es);
            break;
        }
        default: w->malformed = 1;
    }
}

/* isEmptyValue (encode.go): false, 0, 0.0, "", nil pointer, len 0 slice / map; structs never */
static int field_empty(const orc_field* f, const uint8_t* p) {
    if (f->container == C_PTR) return rd32u(p) == 0;
    if (f->container == C_SLICE || f->container == C_MAP || f->container == C_SLICE_PTR) { uint32_t n = rd32u(p); return n == 0 || n == 0xFFFFFFFFu; }
    switch (f->kind) {
        case F_INT64: case F_INT: case F_UINT64: return rd64u(p) == 0;
        case F_FLOAT64: return (rd64u(p) << 1) == 0; /* +0 and -0 */
        case F_FLOAT32: return (uint32_t)(rd32u(p) << 1) == 0;
        case F_BYTES: { uint32_t n = rd32u(p); return n == 0 || n == 0xFFFFFFFFu; } /* len(v) == 0 */
        case F_STRUCT: case F_TIME: return 0; /* a struct is never empty */
        default: return rd32u(p) == 0; /* INT32, BOOL, STRING (length) */
    }
}

static void enc_struct(obuf* b, walk* w, const orc_schema* sc, const uint8_t* fixed) {
    if (sc->n_fields == 1 && (sc->f[0].flags & FIELD_BARE)) { enc_field_value(b, w, &sc->f[0], fixed); return; }
    ob_putc(b, '{');
    int first = 1;
    for (int i = 0; i < sc->n_fields && !WALK_OVER(w); i++) {
        const orc_field* f = &sc->f[i];
        const uint8_t* p = fixed;
        fixed += (size_t)field_fixed_words(w->t, f) * 4;
        /* an empty value owns no bytes of the variable part (zero-length string, nil pointer, no elements), so
         * skipping it needs no cursor movement */
        if (f->omitempty && field_empty(f, p)) continue;
        if (!first) ob_putc(b, ',');
        first = 0;
        orc_enc_string(b, (const uint8_t*)f->json_name, strlen(f->json_name));
        ob_putc(b, ':');
        enc_field_value(b, w, f, p);
    }
    ob_putc(b, '}');
}

/* The JSON of a row of schema sc whose fixed part is at `fixed` (fixed_avail bytes readable) and whose variable part is
 * [var, end).  Returns 0 (b holds the text), -1 malformed row, -2 the value cannot be encoded (NaN / ±Inf). */
int orc_enc_row(obuf* b, const orc_table* t, const orc_schema* sc, const uint8_t* fixed, size_t fixed_avail,
                const uint8_t* var, const uint8_t* end) {
    if ((size_t)orc_schema_fixed_words(t, sc) * 4 > fixed_avail) return -1;
    walk w = {t, var, end, 0, 0};
    enc_struct(b, &w, sc, fixed);
    if (w.malformed) return -1;
    return w.failed ? -2 : 0;
}

int orc_schema_is_flat(const orc_schema* sc) {
    for (int i = 0; i < sc->n_fields; i++)
        if (sc->f[i].container != C_VALUE || sc->f[i].kind > F_INT || (sc->f[i].flags & FIELD_BARE)) return 0;
    return 1;
}

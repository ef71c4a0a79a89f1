/* orc_reqlog.c — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h; parity unpinned: no Go toolchain in this image).
 *
 * CPU restatement of the per-request log line the reference writes after every HTTP request:
 *   middleware.Logging            pkg/gofr/http/middleware/logger.go:41-70   (builds RequestLog, calls logger.Log)
 *   RequestLog                    pkg/gofr/http/middleware/logger.go:24-33   (field order, json tags, omitempty)
 *   getIPAddress                  pkg/gofr/http/middleware/logger.go:72-84   (first X-Forwarded-For element, else RemoteAddr; TrimSpace)
 *   logger.logf / logEntry        pkg/gofr/logging/logger.go:37-41,43-74     (non-terminal: json.NewEncoder(out).Encode(entry))
 *   Level.MarshalJSON             pkg/gofr/logging/level.go:64-70            ("INFO" for logger.Log)
 * kind 1 is the gRPC interceptor's line instead:
 *   RPCLog / String / LoggingInterceptor   pkg/gofr/grpc/log.go:15-25,27-50   logger.Infof("%s", l): the message is the
 *   STRING json.Marshal(RPCLog) — so the inner document is JSON-escaped once more by the outer encoder.
 * and of the standard-library pieces those call: time.Time.MarshalJSON (RFC 3339 with nanoseconds, trailing zeros
 * of the fraction removed, "Z" for a zero offset), Time.Format("2006-01-02T15:04:05.999999999-07:00") (same fraction
 * rule, always a numeric offset), strings.Split / strings.TrimSpace (Unicode White_Space), encoding/json string
 * escaping with HTML escaping on, `omitempty` on strings and integers.
 */
#include "gofr_oracle.h"
#include "orc_internal.h"

typedef struct {
    int64_t start_unix_ns;  /* start := time.Now() */
    int64_t elapsed_ns;     /* time.Since(start) */
    int64_t log_unix_ns;    /* time.Now() inside logger.logf */
    uint32_t arena_off;     /* method | user_agent | x_forwarded_for | remote_addr | request_uri, back to back */
    uint16_t method_len, ua_len, xff_len, remote_len, uri_len;
    uint16_t status;        /* StatusResponseWriter.status (0 = WriteHeader never called) */
    int32_t tz_offset_s;    /* offset of time.Local at that instant, seconds east of UTC */
    uint32_t kind;          /* 0: RequestLog (HTTP middleware), 1: RPCLog (gRPC LoggingInterceptor) */
} orc_log_desc;

/* days since 1970-01-01 → proleptic Gregorian civil date (what time.Time.Date computes) */
static void civil_from_days(int64_t z, int* y, int* m, int* d) {
    z += 719468;
    int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    int64_t doe = z - era * 146097;
    int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    int64_t mp = (5 * doy + 2) / 153;
    *d = (int)(doy - (153 * mp + 2) / 5 + 1);
    *m = (int)(mp < 10 ? mp + 3 : mp - 9);
    *y = (int)(yoe + era * 400 + (*m <= 2));
}

static void put2(obuf* b, int v) { ob_putc(b, (uint8_t)('0' + v / 10)); ob_putc(b, (uint8_t)('0' + v % 10)); }

/* time.Time.AppendFormat for the two layouts on this path; zulu: "Z07:00" (true) or "-07:00" (false) */
static void fmt_time(obuf* b, int64_t unix_ns, int32_t off, int zulu) {
    int64_t sec = unix_ns / 1000000000;
    int64_t ns = unix_ns % 1000000000;
    if (ns < 0) { ns += 1000000000; sec -= 1; }
    int64_t local = sec + off;
    int64_t days = local / 86400, sod = local % 86400;
    if (sod < 0) { sod += 86400; days -= 1; }
    int y, m, d;
    civil_from_days(days, &y, &m, &d);
    /* "2006": zero padded to 4 (years of an int64 nanosecond clock are 1677..2262) */
    ob_putc(b, (uint8_t)('0' + y / 1000 % 10)); ob_putc(b, (uint8_t)('0' + y / 100 % 10));
    ob_putc(b, (uint8_t)('0' + y / 10 % 10)); ob_putc(b, (uint8_t)('0' + y % 10));
    ob_putc(b, '-'); put2(b, m); ob_putc(b, '-'); put2(b, d); ob_putc(b, 'T');
    put2(b, (int)(sod / 3600)); ob_putc(b, ':'); put2(b, (int)(sod / 60 % 60)); ob_putc(b, ':'); put2(b, (int)(sod % 60));
    if (ns) { /* ".999999999": trailing zeros trimmed, nothing at all for a whole second */
        char f[9];
        int64_t v = ns;
        for (int k = 8; k >= 0; k--) { f[k] = (char)('0' + v % 10); v /= 10; }
        int n = 9;
        while (n > 0 && f[n - 1] == '0') n--;
        ob_putc(b, '.');
        ob_put(b, f, (size_t)n);
    }
    if (zulu && off == 0) { ob_putc(b, 'Z'); return; }
    int zone = off / 60; /* truncates toward zero like Go */
    if (zone < 0) { ob_putc(b, '-'); zone = -zone; } else ob_putc(b, '+');
    put2(b, zone / 60); ob_putc(b, ':'); put2(b, zone % 60);
}

/* length of a Unicode White_Space rune starting at p (0 if none): unicode.IsSpace */
static int space_at(const uint8_t* p, size_t n) {
    if (n == 0) return 0;
    uint8_t c = p[0];
    if (c == ' ' || (c >= '\t' && c <= '\r')) return 1;
    if (c == 0xC2 && n >= 2 && (p[1] == 0x85 || p[1] == 0xA0)) return 2;
    if (n >= 3) {
        if (c == 0xE1 && p[1] == 0x9A && p[2] == 0x80) return 3;                 /* U+1680 */
        if (c == 0xE2 && p[1] == 0x80 && ((p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF)) return 3;
        if (c == 0xE2 && p[1] == 0x81 && p[2] == 0x9F) return 3;                 /* U+205F */
        if (c == 0xE3 && p[1] == 0x80 && p[2] == 0x80) return 3;                 /* U+3000 */
    }
    return 0;
}
/* length of a White_Space rune ENDING at p+n (utf8.DecodeLastRune + unicode.IsSpace) */
static int space_before(const uint8_t* p, size_t n) {
    if (n == 0) return 0;
    uint8_t c = p[n - 1];
    if (c < 0x80) return (c == ' ' || (c >= '\t' && c <= '\r')) ? 1 : 0;
    if (n >= 2 && space_at(p + n - 2, 2) == 2) return 2;
    if (n >= 3 && space_at(p + n - 3, 3) == 3) return 3;
    return 0;
}

static void trim_space(const uint8_t** s, size_t* n) {
    int k;
    while ((k = space_at(*s, *n)) != 0) { *s += k; *n -= (size_t)k; }
    while ((k = space_before(*s, *n)) != 0) *n -= (size_t)k;
}

static void key_str(obuf* b, int* first, const char* key, const uint8_t* s, size_t n) {
    if (n == 0) return; /* omitempty */
    if (!*first) ob_putc(b, ',');
    *first = 0;
    ob_putc(b, '"'); ob_puts(b, key); ob_puts(b, "\":");
    orc_enc_string(b, s, n);
}
static void key_int(obuf* b, int* first, const char* key, int64_t v) {
    if (v == 0) return; /* omitempty */
    if (!*first) ob_putc(b, ',');
    *first = 0;
    ob_putc(b, '"'); ob_puts(b, key); ob_puts(b, "\":");
    orc_enc_int(b, v);
}

int orc_request_log(const void* desc_v, const uint8_t* ids, const uint8_t* arena, uint32_t n, uint8_t* out, uint64_t out_cap,
                    uint32_t* out_off) {
    const orc_log_desc* desc = (const orc_log_desc*)desc_v;
    obuf b;
    ob_init(&b);
    uint64_t pos = 0;
    for (uint32_t i = 0; i < n; i++) {
        const orc_log_desc* d = &desc[i];
        const uint8_t* method = arena + d->arena_off;
        const uint8_t* ua = method + d->method_len;
        const uint8_t* xff = ua + d->ua_len;
        const uint8_t* remote = xff + d->xff_len;
        const uint8_t* uri = remote + d->remote_len;
        b.n = 0;
        ob_puts(&b, "{\"Level\":\"INFO\",\"time\":\"");
        fmt_time(&b, d->log_unix_ns, d->tz_offset_s, 1);
        /* reqID := ...TraceID().String(): 32 lower-case hex digits, never empty */
        char hex[32];
        for (int k = 0; k < 16; k++) {
            hex[2 * k] = "0123456789abcdef"[ids[(size_t)i * 16 + k] >> 4];
            hex[2 * k + 1] = "0123456789abcdef"[ids[(size_t)i * 16 + k] & 15];
        }
        if (d->kind == 1) {
            /* grpc/log.go:35-44: l.String() = json.Marshal(RPCLog) (no omitempty), logged with Infof("%s", l):
             * entry.Message is that string, encoded as a JSON string by the outer json.Encoder */
            obuf in, st;
            ob_init(&in);
            ob_init(&st);
            fmt_time(&st, d->start_unix_ns, d->tz_offset_s, 0);
            ob_puts(&in, "{\"id\":");
            orc_enc_string(&in, (const uint8_t*)hex, 32);
            ob_puts(&in, ",\"startTime\":");
            orc_enc_string(&in, st.p, st.n);
            ob_puts(&in, ",\"responseTime\":");
            orc_enc_int(&in, d->elapsed_ns / 1000); /* time.Since(start).Microseconds() */
            ob_puts(&in, ",\"method\":");
            orc_enc_string(&in, method, d->method_len); /* info.FullMethod */
            ob_putc(&in, '}');
            ob_puts(&b, "\",\"message\":");
            orc_enc_string(&b, in.p, in.n);
            ob_puts(&b, "}\n");
            ob_free(&in);
            ob_free(&st);
            out_off[i] = (uint32_t)pos;
            if (pos + b.n > out_cap) { ob_free(&b); return -1; }
            memcpy(out + pos, b.p, b.n);
            pos += b.n;
            continue;
        }
        ob_puts(&b, "\",\"message\":{");
        int first = 1;
        key_str(&b, &first, "id", (const uint8_t*)hex, 32);
        obuf st;
        ob_init(&st);
        fmt_time(&st, d->start_unix_ns, d->tz_offset_s, 0);
        key_str(&b, &first, "start_time", st.p, st.n);
        ob_free(&st);
        key_int(&b, &first, "response_time", d->elapsed_ns / 1000); /* Nanoseconds() / 1000, truncating */
        key_str(&b, &first, "method", method, d->method_len);
        key_str(&b, &first, "user_agent", ua, d->ua_len);
        /* getIPAddress: ips := strings.Split(xff, ","); ip := ips[0]; if ip == "" { ip = RemoteAddr }; TrimSpace(ip) */
        const uint8_t* ip = xff;
        size_t ipn = 0;
        while (ipn < d->xff_len && xff[ipn] != ',') ipn++;
        if (ipn == 0) { ip = remote; ipn = d->remote_len; }
        trim_space(&ip, &ipn);
        key_str(&b, &first, "ip", ip, ipn);
        key_str(&b, &first, "uri", uri, d->uri_len);
        key_int(&b, &first, "response", d->status);
        ob_puts(&b, "}}\n");
        out_off[i] = (uint32_t)pos;
        if (pos + b.n > out_cap) { ob_free(&b); return -1; }
        memcpy(out + pos, b.p, b.n);
        pos += b.n;
    }
    out_off[n] = (uint32_t)pos;
    ob_free(&b);
    return 0;
}

// reqlog_device.cuh — the per-request log line of middleware.Logging, one thread per record.
//
// Reference: pkg/gofr/http/middleware/logger.go:24-33 (RequestLog: field order, json tags, omitempty), :41-70 (what
// each field is), :72-84 (getIPAddress); pkg/gofr/logging/logger.go:37-41,43-74 (logEntry{Level,time,message} encoded
// with json.NewEncoder on a non-terminal writer); pkg/gofr/logging/level.go:64-70 (Level.MarshalJSON → "INFO").
// Standard-library behaviour restated here: time.Time.MarshalJSON (RFC 3339, fraction with trailing zeros removed,
// "Z" for offset 0), Time.Format("2006-01-02T15:04:05.999999999-07:00"), strings.Split(..)[0], strings.TrimSpace
// (Unicode White_Space), encoding/json strings (HTML escaping on) — the last through serve_device.cuh.
//
// The code is __host__ __device__ like serve_device.cuh: tests/emu runs it on the CPU against the oracle.
#pragma once
#include "serve_device.cuh"

namespace gofr {

struct LogDesc {  // = gofr_log_desc (include/gofr_b200.h), 48 bytes
    int64_t start_ns, elapsed_ns, log_ns;
    uint32_t arena_off;
    uint16_t method_len, ua_len, xff_len, remote_len, uri_len, status;
    int32_t tz_offset_s;
    uint32_t kind;  // 0: RequestLog (HTTP middleware), 1: RPCLog (gRPC LoggingInterceptor, pkg/gofr/grpc/log.go:15-50)
};
static_assert(sizeof(LogDesc) == 48, "LogDesc layout");

// what the size pass learns about one record
struct LogCtx {
    uint32_t ip_off, ip_len;  // getIPAddress result, relative to the record's first arena byte
    uint32_t esc_mask;        // bit0 method, bit1 user_agent, bit2 ip, bit3 uri: needs the slow escape path
    uint32_t total_len;
};

// ---- time.Time → "2006-01-02T15:04:05[.fraction]" + zone ---------------------------------------------------------
struct CivilTime {
    uint32_t year, month, day, hour, minute, second, ns;
};

GOFR_HD CivilTime civil_time(int64_t unix_ns, int32_t off) {
    int64_t sec = unix_ns / 1000000000ll;
    int64_t ns = unix_ns - sec * 1000000000ll;
    if (ns < 0) { ns += 1000000000ll; sec -= 1; }
    const int64_t local = sec + off;
    int64_t days = local / 86400;
    int32_t sod = (int32_t)(local - days * 86400);
    if (sod < 0) { sod += 86400; days -= 1; }
    // days since 1970-01-01 → civil date (proleptic Gregorian, as time.Time.Date); |days| < 2^18 on this clock
    const int32_t z = (int32_t)days + 719468;
    const int32_t era = (z >= 0 ? z : z - 146096) / 146097;
    const uint32_t doe = (uint32_t)(z - era * 146097);
    const uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const uint32_t mp = (5 * doy + 2) / 153;
    CivilTime t;
    t.day = doy - (153 * mp + 2) / 5 + 1;
    t.month = mp < 10 ? mp + 3 : mp - 9;
    t.year = (uint32_t)((int32_t)yoe + era * 400) + (t.month <= 2 ? 1u : 0u);
    t.hour = (uint32_t)sod / 3600;
    t.minute = (uint32_t)sod / 60 % 60;
    t.second = (uint32_t)sod % 60;
    t.ns = (uint32_t)ns;
    return t;
}

// digits of the ".999999999" fraction that survive the trailing-zero trim (0 for a whole second)
GOFR_HD uint32_t frac_digits(uint32_t ns) {
    if (!ns) return 0;
    uint32_t n = 9;
    while (ns % 10 == 0) { ns /= 10; n--; }
    return n;
}

GOFR_HD uint32_t two_digits(uint32_t v) { return ('0' + v / 10) | ('0' + v % 10) << 8; }

// appends the time; zulu = layout "Z07:00" (MarshalJSON), else "-07:00" (start_time).  At most 35 bytes = 9 words.
template <bool EMIT>
GOFR_HD uint32_t emit_time(Writer* w, int64_t unix_ns, int32_t off, bool zulu) {
    const CivilTime t = civil_time(unix_ns, off);
    const uint32_t fd = frac_digits(t.ns);
    const bool z = zulu && off == 0;
    const uint32_t total = 19 + (fd ? 1 + fd : 0) + (z ? 1 : 6);
    if (!EMIT) return total;
    const uint32_t mo = two_digits(t.month), da = two_digits(t.day), ho = two_digits(t.hour), mi = two_digits(t.minute),
                   se = two_digits(t.second);
    w->put4(ascii4(t.year));                                  // "2006"
    w->put4('-' | mo << 8 | '-' << 24);                       // "-01-"
    w->put4(da | 'T' << 16 | (ho & 0xFF) << 24);              // "02T1"
    w->put4((ho >> 8) | ':' << 8 | mi << 16);                 // "5:04"
    w->putk(':' | se << 8, 3);                                // ":05"
    if (fd) {
        // nine zero-padded digits; the first fd of them follow the dot
        const uint32_t hi = t.ns / 100000u, lo = t.ns - hi * 100000u;  // hi < 10000: digits 1-4; lo: digits 5-9
        const uint32_t d5 = lo / 10000u, l4 = lo - d5 * 10000u;
        const uint32_t W0 = ascii4(hi), W1 = ('0' + d5) | (ascii4(l4) << 8), W2 = ascii4(l4) >> 24;
        w->putc('.');
        if (fd >= 4) w->put4(W0); else w->putk(W0, fd);
        if (fd >= 8) w->put4(W1); else if (fd > 4) w->putk(W1, fd - 4);
        if (fd == 9) w->putc(W2);
    }
    if (z) { w->putc('Z'); return total; }
    int32_t zone = off / 60;  // truncates toward zero, like Go
    uint32_t sign = '+';
    if (zone < 0) { sign = '-'; zone = -zone; }
    const uint32_t zh = two_digits((uint32_t)zone / 60), zm = two_digits((uint32_t)zone % 60);
    w->put4(sign | zh << 8 | ':' << 24);                      // "+07:"
    w->putk(zm, 2);                                           // "00"
    return total;
}

// ---- strings.TrimSpace ---------------------------------------------------------------------------------------------
// length of a Unicode White_Space rune starting at p (0 if none)
GOFR_HD uint32_t space_at(const uint8_t* p, uint32_t n) {
    if (!n) return 0;
    const uint32_t c = p[0];
    if (c == ' ' || (c >= '\t' && c <= '\r')) return 1;
    if (c == 0xC2 && n >= 2 && (p[1] == 0x85 || p[1] == 0xA0)) return 2;
    if (n >= 3) {
        const uint32_t b = p[1], d = p[2];
        if (c == 0xE1 && b == 0x9A && d == 0x80) return 3;
        if (c == 0xE2 && b == 0x80 && ((d >= 0x80 && d <= 0x8A) || d == 0xA8 || d == 0xA9 || d == 0xAF)) return 3;
        if (c == 0xE2 && b == 0x81 && d == 0x9F) return 3;
        if (c == 0xE3 && b == 0x80 && d == 0x80) return 3;
    }
    return 0;
}
// length of a White_Space rune ending at p + n (utf8.DecodeLastRune finds the same start: the nearest lead byte)
GOFR_HD uint32_t space_before(const uint8_t* p, uint32_t n) {
    if (!n) return 0;
    const uint32_t c = p[n - 1];
    if (c < 0x80) return (c == ' ' || (c >= '\t' && c <= '\r')) ? 1u : 0u;
    if (n >= 2 && space_at(p + n - 2, 2) == 2) return 2;
    if (n >= 3 && space_at(p + n - 3, 3) == 3) return 3;
    return 0;
}

// getIPAddress: ips := strings.Split(xff, ","); ip := ips[0]; if ip == "" { ip = RemoteAddr }; strings.TrimSpace(ip)
GOFR_HD_NOINLINE void reqlog_ip(const uint8_t* rec, const LogDesc d, uint32_t* off_out, uint32_t* len_out) {
    uint32_t off = (uint32_t)d.method_len + d.ua_len, n = 0;
    while (n < d.xff_len && rec[off + n] != ',') n++;
    if (n == 0) { off += d.xff_len; n = d.remote_len; }
    for (uint32_t k; (k = space_at(rec + off, n)) != 0;) { off += k; n -= k; }
    for (uint32_t k; (k = space_before(rec + off, n)) != 0;) n -= k;
    *off_out = off;
    *len_out = n;
}

// ---- literals: compile-time strings leave as immediate words --------------------------------------------------------
template <uint32_t N>
GOFR_HD void put_lit(Writer* w, const char (&s)[N]) {
    constexpr uint32_t L = N - 1;
#pragma unroll
    for (uint32_t k = 0; k + 4 <= L; k += 4)
        w->put4((uint32_t)(uint8_t)s[k] | (uint32_t)(uint8_t)s[k + 1] << 8 | (uint32_t)(uint8_t)s[k + 2] << 16 |
                (uint32_t)(uint8_t)s[k + 3] << 24);
    constexpr uint32_t r = L & 3u, b = L & ~3u;
    if (r == 1) w->putk((uint32_t)(uint8_t)s[b], 1);
    if (r == 2) w->putk((uint32_t)(uint8_t)s[b] | (uint32_t)(uint8_t)s[b + (r > 1 ? 1 : 0)] << 8, 2);
    if (r == 3)
        w->putk((uint32_t)(uint8_t)s[b] | (uint32_t)(uint8_t)s[b + (r > 1 ? 1 : 0)] << 8 |
                    (uint32_t)(uint8_t)s[b + (r > 2 ? 2 : 0)] << 16,
                3);
}

// `,"key":"` + JSON string contents + `"` — or nothing for an empty string (omitempty)
template <bool EMIT, uint32_t N>
GOFR_HD uint32_t log_string(Writer* w, const char (&key)[N], const uint8_t* s, uint32_t n, bool staged, uint32_t bit,
                            uint32_t& esc_mask) {
    if (!n) return 0;
    uint32_t body;
    if (!EMIT) {
        const bool esc = staged ? json_needs_escape<true>(SrcMem<true>::from(s), n) : json_needs_escape<false>(s, n);
        if (esc) { esc_mask |= bit; body = json_escape_slow<false>(nullptr, s, n); }
        else body = n;
    } else {
        w->reserve(8);
        put_lit(w, key);
        if (esc_mask & bit) GOFR_SLOW_CALL(w, json_escape_slow<true>(tw, s, n));
        else if (staged) w->copy<true>(SrcMem<true>::from(s), n);
        else GOFR_SLOW_CALL(w, emit_bytes(*tw, s, n));
        w->reserve(8);
        w->putc('"');
        body = 0;
    }
    return (N - 1) + body + 1;
}

// info.FullMethod inside RPCLog inside the log entry: escaped by json.Marshal(RPCLog), then once more because the
// marshalled document is logged as a string (grpc/log.go:22-25,43: logger.Infof("%s", l)).  Rune by rune; rare.
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t rpc_method_slow(Writer* w, const uint8_t* p, uint32_t len) {
    uint32_t out = 0;
    for (uint32_t i = 0; i < len;) {
        if (EMIT) w->reserve_out(4);
        const uint32_t c = p[i];
        uint32_t L = 1, u = 0;  // u != 0: the rune leaves as \\uXXXX with these four hex digits (as a word)
        if (c < 0x80) {
            if (c == '"' || c == '\\') {  // escaped twice: three backslashes and a quote, or four backslashes
                if (EMIT) w->put4('\\' | '\\' << 8 | '\\' << 16 | c << 24);
                out += 4; i++; continue;
            }
            if (c == '\n' || c == '\r' || c == '\t') {
                if (EMIT) w->putk('\\' | '\\' << 8 | (c == '\n' ? 'n' : c == '\r' ? 'r' : 't') << 16, 3);
                out += 3; i++; continue;
            }
            if (c >= 0x20 && c != '<' && c != '>' && c != '&') {
                if (EMIT) w->put1(c);
                out += 1; i++; continue;
            }
            u = '0' | '0' << 8 | hex_lc(c >> 4) << 16 | hex_lc(c & 15) << 24;
        } else {
            L = utf8_len_at(p + i, len - i);
            if (L == 0) { L = 1; u = 'f' | 'f' << 8 | 'f' << 16 | 'd' << 24; }
            else if (L == 3 && c == 0xE2 && p[i + 1] == 0x80 && (p[i + 2] == 0xA8 || p[i + 2] == 0xA9))
                u = '2' | '0' << 8 | '2' << 16 | (p[i + 2] == 0xA8 ? '8' : '9') << 24;
        }
        if (u) {
            if (EMIT) { w->putk('\\' | '\\' << 8 | 'u' << 16, 3); w->put4(u); }
            out += 7;
        } else {
            if (EMIT) for (uint32_t k = 0; k < L; k++) w->put1(p[i + k]);
            out += L;
        }
        i += L;
    }
    return out;
}

// The gRPC interceptor's line: {"Level":"INFO","time":"…","message":"{\\"id\\":\\"…\\",\\"startTime\\":\\"…\\",
// \\"responseTime\\":N,\\"method\\":\\"…\\"}"}\n  (RPCLog has no omitempty; every inner quote is escaped by the outer encoder)
// a literal of at most 32 bytes: emitted (EMIT) or just counted
template <bool EMIT, uint32_t N>
GOFR_HD uint32_t lit(Writer* w, const char (&s)[N]) {
    static_assert(N - 1 <= 32, "one reserve(8) covers the literal");
    if (EMIT) { w->reserve(8); put_lit(w, s); }
    return N - 1;
}

template <bool EMIT>
GOFR_HD void rpclog_run(const LogDesc& d, const uint8_t* rec, bool staged, const uint32_t id[4], LogCtx& c, Writer* w) {
    uint32_t len = 0;
    if (!EMIT) { c.esc_mask = 0; c.ip_off = c.ip_len = 0; }
    len += lit<EMIT>(w, "{\"Level\":\"INFO\",\"time\":\"");
    if (EMIT) w->reserve(9);
    len += emit_time<EMIT>(w, d.log_ns, d.tz_offset_s, true);
    len += lit<EMIT>(w, "\",\"message\":\"{\\\"id\\\":\\\"");
    if (EMIT) {
        w->reserve(8);
#pragma unroll
        for (int k = 0; k < 4; k++) { uint32_t a, b; hex8(id[k], a, b); w->put4(a); w->put4(b); }
    }
    len += 32;
    len += lit<EMIT>(w, "\\\",\\\"startTime\\\":\\\"");
    if (EMIT) w->reserve(9);
    len += emit_time<EMIT>(w, d.start_ns, d.tz_offset_s, false);
    len += lit<EMIT>(w, "\\\",\\\"responseTime\\\":");
    if (EMIT) w->reserve(8);
    len += emit_i64<EMIT>(w, d.elapsed_ns / 1000);  // time.Since(start).Microseconds()
    len += lit<EMIT>(w, ",\\\"method\\\":\\\"");
    const uint32_t n = d.method_len;
    if (n) {
        if (!EMIT) {
            const bool esc = staged ? json_needs_escape<true>(SrcMem<true>::from(rec), n) : json_needs_escape<false>(rec, n);
            if (esc) { c.esc_mask |= 1u; len += rpc_method_slow<false>(nullptr, rec, n); }
            else len += n;
        } else if (c.esc_mask & 1u) GOFR_SLOW_CALL(w, rpc_method_slow<true>(tw, rec, n));
        else if (staged) w->copy<true>(SrcMem<true>::from(rec), n);
        else GOFR_SLOW_CALL(w, emit_bytes(*tw, rec, n));
    }
    len += lit<EMIT>(w, "\\\"}\"}\n");
    if (!EMIT) c.total_len = len;
}

// One record.  EMIT=false: fills c (ip span, escape bits, total_len).  EMIT=true: writes the line through w.
// rec = first arena byte of the record (method | user_agent | x_forwarded_for | remote_addr | request_uri).
template <bool EMIT>
GOFR_HD void reqlog_run(const LogDesc& d, const uint8_t* rec, bool staged, const uint32_t id[4], LogCtx& c, Writer* w) {
    if (d.kind == 1) { rpclog_run<EMIT>(d, rec, staged, id, c, w); return; }
    uint32_t len = 0;
    if (!EMIT) { c.esc_mask = 0; reqlog_ip(rec, d, &c.ip_off, &c.ip_len); }
    if (EMIT) { w->reserve(8); put_lit(w, "{\"Level\":\"INFO\",\"time\":\""); w->reserve(9); }
    len += 24 + emit_time<EMIT>(w, d.log_ns, d.tz_offset_s, true);
    if (EMIT) {
        w->reserve(8);
        put_lit(w, "\",\"message\":{\"id\":\"");
        w->reserve(8);
#pragma unroll
        for (int k = 0; k < 4; k++) { uint32_t a, b; hex8(id[k], a, b); w->put4(a); w->put4(b); }
        w->reserve(8);
        put_lit(w, "\",\"start_time\":\"");
        w->reserve(9);
    }
    len += 19 + 32 + 16 + emit_time<EMIT>(w, d.start_ns, d.tz_offset_s, false) + 1;
    if (EMIT) { w->reserve(8); w->putc('"'); }
    const int64_t rt = d.elapsed_ns / 1000;  // time.Since(start).Nanoseconds() / 1000
    if (rt != 0) {
        if (EMIT) { put_lit(w, ",\"response_time\":"); w->reserve(8); }
        len += 17 + emit_i64<EMIT>(w, rt);
    }
    const uint8_t* p = rec;
    len += log_string<EMIT>(w, ",\"method\":\"", p, d.method_len, staged, 1u, c.esc_mask);
    p += d.method_len;
    len += log_string<EMIT>(w, ",\"user_agent\":\"", p, d.ua_len, staged, 2u, c.esc_mask);
    len += log_string<EMIT>(w, ",\"ip\":\"", rec + c.ip_off, c.ip_len, staged, 4u, c.esc_mask);
    p += (uint32_t)d.ua_len + d.xff_len + d.remote_len;
    len += log_string<EMIT>(w, ",\"uri\":\"", p, d.uri_len, staged, 8u, c.esc_mask);
    if (d.status != 0) {
        if (EMIT) { w->reserve(8); put_lit(w, ",\"response\":"); }
        len += 12 + emit_u32<EMIT>(w, d.status);
    }
    if (EMIT) { w->reserve(8); put_lit(w, "}}\n"); }
    len += 3;
    if (!EMIT) c.total_len = len;
}

GOFR_HD void reqlog_size(const LogDesc& d, const uint8_t* rec, bool staged, LogCtx& c) {
    const uint32_t none[4] = {0, 0, 0, 0};
    reqlog_run<false>(d, rec, staged, none, c, nullptr);
}

GOFR_HD void reqlog_emit(const LogDesc& d, const uint8_t* rec, bool staged, const uint32_t id[4], LogCtx& c, uint8_t* dst,
                         uint32_t* ring_col) {
    Writer w;
    w.init(dst, ring_col);
    reqlog_run<true>(d, rec, staged, id, c, &w);
    w.finish();
}

}  // namespace gofr

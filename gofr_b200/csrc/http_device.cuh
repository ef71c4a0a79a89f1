// http_device.cuh — one HTTP/1.1 request head, one thread (gofr_http_parse_device).
//
// Restates, for the conservative subset spelled out in include/gofr_b200.h, what Go 1.21's net/http readRequest
// (request line via parseRequestLine, headers via textproto.Reader.ReadMIMEHeader, Host / Content-Length handling) and
// net/url.ParseRequestURI (origin form: Path = unescape(path part), RawQuery = after the first '?', ForceQuery) hand
// to Router.ServeHTTP — the step before pkg/gofr/http/router.go:14 on the reference's path (reached from
// pkg/gofr/httpServer.go:29-33).  Anything outside the subset is DEFERRED to the host, never answered differently.
// __host__ __device__ like serve_device.cuh: tests/emu runs it on the CPU against the oracle.
#pragma once
#include "serve_device.cuh"

namespace gofr {

struct HttpOut {
    uint32_t status;  // GOFR_HTTP_OK / GOFR_HTTP_DEFER
    uint32_t path_len, query_len, data_len, method, flags;
    uint64_t spans[GOFR_HTTP_SPANS];  // offset into the message | length << 32
};

GOFR_HD bool http_tchar(uint32_t c) {  // RFC 7230 token character: digits, letters and !#$%&'*+-.^_`|~ (a 128-bit set)
    const uint32_t w = c < 32 ? 0u : c < 64 ? 0x03FF6CFAu : c < 96 ? 0xC7FFFFFEu : c < 128 ? 0x57FFFFFFu : 0u;
    return (w >> (c & 31u)) & 1u;
}
GOFR_HD int http_hex(uint32_t c) {
    if (c - '0' < 10u) return (int)(c - '0');
    c |= 0x20u;
    return c - 'a' < 6u ? (int)(c - 'a' + 10) : -1;
}
// case-insensitive compare of a header name with a lower-case literal of the same length
GOFR_HD bool http_name_is(const uint8_t* p, uint32_t n, const char* lit, uint32_t ln) {
    if (n != ln) return false;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t c = p[k];
        if (c - 'A' < 26u) c |= 0x20u;
        if (c != (uint32_t)(uint8_t)lit[k]) return false;
    }
    return true;
}

// Parses message m[0..n); on success writes Path | RawQuery | pad4 | body at dst (4-byte aligned) and fills o.
// Returns with o.status = GOFR_HTTP_DEFER as soon as anything leaves the subset.
GOFR_HD_NOINLINE void http_parse(const uint8_t* m, uint32_t n, uint8_t* dst, HttpOut* out) {
    HttpOut o;
    o.status = GOFR_HTTP_DEFER;
    o.path_len = o.query_len = o.data_len = o.method = o.flags = 0;
    for (int k = 0; k < GOFR_HTTP_SPANS; k++) o.spans[k] = 0;
    *out = o;
    // ---- request line: METHOD SP target SP HTTP/1.1 CRLF ----
    uint32_t ml = 0;
    while (ml < n && ml < 8 && m[ml] != ' ') ml++;
    if (ml == n || m[ml] != ' ') return;
    uint32_t method;
    if (ml == 3 && m[0] == 'G' && m[1] == 'E' && m[2] == 'T') method = GOFR_M_GET;
    else if (ml == 4 && m[0] == 'H' && m[1] == 'E' && m[2] == 'A' && m[3] == 'D') method = GOFR_M_HEAD;
    else if (ml == 4 && m[0] == 'P' && m[1] == 'O' && m[2] == 'S' && m[3] == 'T') method = GOFR_M_POST;
    else if (ml == 3 && m[0] == 'P' && m[1] == 'U' && m[2] == 'T') method = GOFR_M_PUT;
    else if (ml == 5 && m[0] == 'P' && m[1] == 'A' && m[2] == 'T' && m[3] == 'C' && m[4] == 'H') method = GOFR_M_PATCH;
    else if (ml == 6 && m[0] == 'D' && m[1] == 'E' && m[2] == 'L' && m[3] == 'E' && m[4] == 'T' && m[5] == 'E') method = GOFR_M_DELETE;
    else if (ml == 7 && m[0] == 'O' && m[1] == 'P' && m[2] == 'T' && m[3] == 'I' && m[4] == 'O' && m[5] == 'N' && m[6] == 'S') method = GOFR_M_OPTIONS;
    else return;
    const uint32_t t0 = ml + 1;
    uint32_t t1 = t0, qmark = 0xFFFFFFFFu;
    while (t1 < n && m[t1] != ' ') {
        const uint32_t c = m[t1];
        if (c < 0x21 || c > 0x7E || c == '#') return;
        if (c == '?' && qmark == 0xFFFFFFFFu) qmark = t1;
        t1++;
    }
    if (t1 == t0 || t1 >= n || m[t0] != '/' || (t1 - t0 >= 2 && m[t0 + 1] == '/')) return;
    if (t1 - t0 > 8192) return;
    if (t1 + 11 > n) return;
    {
        const char v[] = " HTTP/1.1\r\n";
        for (uint32_t k = 0; k < 11; k++) if (m[t1 + k] != (uint8_t)v[k]) return;
    }
    const uint32_t path_end = qmark == 0xFFFFFFFFu ? t1 : qmark;
    for (uint32_t k = t0; k < path_end; k++)
        if (m[k] == '%' && (k + 2 >= path_end || http_hex(m[k + 1]) < 0 || http_hex(m[k + 2]) < 0)) return;

    // ---- header lines ----
    uint32_t pos = t1 + 11, hosts = 0, cls = 0, chunked = 0;
    uint64_t content_length = 0;
    uint64_t sp_ua = 0, sp_xff = 0, sp_host = 0;
    bool have_ua = false, have_xff = false;
    for (;;) {
        if (pos + 2 > n) return;                                    // no room for the blank line
        if (m[pos] == '\r') { if (m[pos + 1] != '\n') return; pos += 2; break; }
        if (pos > 16384) return;
        const uint32_t n0 = pos;
        while (pos < n && http_tchar(m[pos])) pos++;
        if (pos == n0 || pos >= n || m[pos] != ':') return;         // empty name, space before the colon, obs-fold, …
        const uint32_t nl = pos - n0;
        pos++;
        while (pos < n && (m[pos] == ' ' || m[pos] == '\t')) pos++;
        const uint32_t v0 = pos;
        for (; pos < n; pos++) {  // one compare per byte on the common path (printable or obs-text)
            const uint32_t c = m[pos];
            if (c >= 0x20 && c != 0x7F) continue;
            if (c == '\t') continue;
            if (c == '\r') break;
            return;
        }
        if (pos + 1 >= n || m[pos + 1] != '\n') return;
        uint32_t v1 = pos;
        while (v1 > v0 && (m[v1 - 1] == ' ' || m[v1 - 1] == '\t')) v1--;
        pos += 2;
        const uint8_t* nm = m + n0;
        const uint64_t span = (uint64_t)v0 | (uint64_t)(v1 - v0) << 32;
        if (http_name_is(nm, nl, "host", 4)) { hosts++; sp_host = span; }
        else if (http_name_is(nm, nl, "user-agent", 10)) { if (!have_ua) { sp_ua = span; have_ua = true; } }
        else if (http_name_is(nm, nl, "x-forwarded-for", 15)) { if (!have_xff) { sp_xff = span; have_xff = true; } }
        else if (http_name_is(nm, nl, "content-length", 14)) {
            cls++;
            if (v1 == v0 || v1 - v0 > 9) return;
            content_length = 0;
            for (uint32_t k = v0; k < v1; k++) {
                if ((uint32_t)m[k] - (uint32_t)'0' >= 10u) return;
                content_length = content_length * 10 + (m[k] - '0');
            }
        } else if (http_name_is(nm, nl, "connection", 10)) {
            if (!http_name_is(m + v0, v1 - v0, "keep-alive", 10)) return;
        } else if (http_name_is(nm, nl, "transfer-encoding", 17)) {
            // exactly one coding, "chunked" (net/http readTransfer: any other coding is "unsupported transfer encoding")
            if (chunked || !http_name_is(m + v0, v1 - v0, "chunked", 7)) return;
            chunked = 1;
        } else if (http_name_is(nm, nl, "expect", 6) || http_name_is(nm, nl, "upgrade", 7) || http_name_is(nm, nl, "trailer", 7)) return;
    }
    if (hosts != 1 || cls > 1) return;
    if (chunked && cls) return;  // both: net/http drops the Content-Length of a chunked request — left to it
    {
        const uint32_t h0 = (uint32_t)sp_host, hl = (uint32_t)(sp_host >> 32);
        if (hl == 0) return;
        for (uint32_t k = 0; k < hl; k++) {
            const uint32_t c = m[h0 + k];
            if (!(c - '0' < 10u || (c | 0x20u) - 'a' < 26u || c == '.' || c == ':' || c == '-' || c == '_' || c == '[' || c == ']')) return;
        }
    }
    // ---- body: Content-Length bytes, or a chunk stream (net/http internal chunkedReader: `hex-size CRLF data CRLF` ... `0 CRLF
    //      CRLF`); the subset has no chunk extensions, no whitespace in the size line, no trailer fields, sizes of at most
    //      8 hex digits, and nothing after the last chunk.  Validated as a whole before a byte is written ----
    uint32_t body_len = n - pos;
    if (chunked) {
        uint32_t q = pos, total = 0;
        for (;;) {
            uint32_t size = 0, nd = 0;
            while (q < n && nd < 9 && http_hex(m[q]) >= 0) { size = size << 4 | (uint32_t)http_hex(m[q]); q++; nd++; }
            if (nd == 0 || nd > 8 || q + 2 > n || m[q] != '\r' || m[q + 1] != '\n') return;
            q += 2;
            if (size == 0) {
                if (q + 2 != n || m[q] != '\r' || m[q + 1] != '\n') return;  // trailer fields or bytes after the message
                break;
            }
            if (size > n - q || n - q - size < 2 || m[q + size] != '\r' || m[q + size + 1] != '\n') return;
            q += size + 2;
            total += size;
        }
        body_len = total;
    } else if (cls ? (uint64_t)body_len != content_length : body_len != 0) return;

    // ---- URL.Path | URL.RawQuery | pad4 | body ----
    uint32_t w = 0;
    for (uint32_t k = t0; k < path_end; k++) {
        uint32_t c = m[k];
        if (c == '%') { c = (uint32_t)(http_hex(m[k + 1]) << 4 | http_hex(m[k + 2])); k += 2; }
        dst[w++] = (uint8_t)c;
    }
    o.path_len = w;
    if (qmark != 0xFFFFFFFFu) {
        for (uint32_t k = qmark + 1; k < t1; k++) dst[w++] = m[k];
        o.query_len = t1 - qmark - 1;
        if (o.query_len == 0) o.flags |= GOFR_REQ_FORCE_QUERY;
    }
    while (w & 3u) dst[w++] = 0;
    if (chunked) {
        uint32_t q = pos, at = w;
        for (;;) {
            uint32_t size = 0;
            while (m[q] != '\r') { size = size << 4 | (uint32_t)http_hex(m[q]); q++; }
            q += 2;
            if (size == 0) break;
            for (uint32_t k = 0; k < size; k++) dst[at + k] = m[q + k];
            at += size;
            q += size + 2;
        }
    } else
    for (uint32_t k = 0; k < body_len; k++) dst[w + k] = m[pos + k];
    o.data_len = body_len;
    o.method = method;
    o.spans[GOFR_HTTP_SPAN_METHOD] = (uint64_t)0 | (uint64_t)ml << 32;
    o.spans[GOFR_HTTP_SPAN_TARGET] = (uint64_t)t0 | (uint64_t)(t1 - t0) << 32;
    o.spans[GOFR_HTTP_SPAN_USER_AGENT] = sp_ua;
    o.spans[GOFR_HTTP_SPAN_XFF] = sp_xff;
    o.spans[GOFR_HTTP_SPAN_HOST] = sp_host;
    // a chunked body has no contiguous image in the message: the span covers the chunk stream as received
    o.spans[GOFR_HTTP_SPAN_BODY] = (uint64_t)pos | (uint64_t)(chunked ? n - pos : body_len) << 32;
    o.status = GOFR_HTTP_OK;
    *out = o;
}

}  // namespace gofr

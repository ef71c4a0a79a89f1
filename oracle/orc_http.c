/* orc_http.c — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h; parity unpinned: no Go toolchain in this image).
 *
 * CPU restatement of what Go 1.21's net/http server derives from one HTTP/1.1 request message before it calls the
 * handler (the step in front of pkg/gofr/http/router.go:14; the server is started at pkg/gofr/httpServer.go:29-33):
 *   http.readRequest          request line: strings.Cut on the first two spaces, ParseHTTPVersion, validMethod
 *   url.ParseRequestURI       origin form: Path = unescape(before '?', encodePath), RawQuery = after the first '?',
 *                             ForceQuery for a lone trailing '?'
 *   textproto.ReadMIMEHeader  one header per CRLF line, "name: value", value trimmed of spaces and tabs
 *   http.readTransfer         Content-Length delimits the body, or Transfer-Encoding: chunked does
 *   internal.chunkedReader    chunk = hex size line, data, CRLF; a zero-size chunk and a blank line end the body
 * for the conservative subset documented with gofr_http_parse_device (include/gofr_b200.h); every message outside the
 * subset is reported as DEFER.  Written line-first (split the head into lines, then classify) on purpose: the device
 * code walks the bytes once, so the two restatements do not share their structure. */
#include "gofr_oracle.h"
#include "orc_internal.h"

enum { SPAN_METHOD, SPAN_TARGET, SPAN_UA, SPAN_XFF, SPAN_HOST, SPAN_BODY, N_SPANS };

static int is_tchar(uint8_t c) {
    if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) return 1;
    return c != 0 && strchr("!#$%&'*+-.^_`|~", c) != NULL;
}
static int hexv(uint8_t c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}
static int ieq(const uint8_t* p, size_t n, const char* lit) {
    if (n != strlen(lit)) return 0;
    for (size_t k = 0; k < n; k++) {
        uint8_t c = p[k];
        if (c >= 'A' && c <= 'Z') c = (uint8_t)(c + 32);
        if (c != (uint8_t)lit[k]) return 0;
    }
    return 1;
}

typedef struct { size_t off, len; } span;

/* returns 0 (OK) or 1 (DEFER) */
static int parse_one(const uint8_t* m, size_t n, uint8_t* dst, uint32_t* path_len, uint32_t* query_len, uint32_t* data_len,
                     uint32_t* method, uint32_t* flags, span sp[N_SPANS]) {
    memset(sp, 0, sizeof(span) * N_SPANS);
    /* ---- split the head into CRLF-terminated lines ---- */
    static __thread span lines[4200];
    int nl = 0;
    size_t pos = 0, head_end = 0;
    for (;;) {
        size_t e = pos;
        while (e < n && m[e] != '\r' && m[e] != '\n') e++;
        if (e + 1 >= n || m[e] != '\r' || m[e + 1] != '\n') return 1; /* bare LF, bare CR, or no line end at all */
        if (e == pos) { head_end = e + 2; break; }                   /* the blank line */
        if (nl == 4200 || (nl > 0 && pos > 16384)) return 1; /* a header line may not start beyond 16 KiB */
        lines[nl].off = pos; lines[nl].len = e - pos; nl++;
        pos = e + 2;
    }
    if (nl == 0) return 1;
    /* ---- request line ---- */
    const uint8_t* rl = m + lines[0].off;
    size_t rn = lines[0].len;
    const uint8_t* s1 = memchr(rl, ' ', rn);
    if (!s1) return 1;
    size_t mlen = (size_t)(s1 - rl);
    const uint8_t* rest = s1 + 1;
    size_t restn = rn - mlen - 1;
    const uint8_t* s2 = memchr(rest, ' ', restn);
    if (!s2) return 1;
    size_t tlen = (size_t)(s2 - rest);
    const uint8_t* proto = s2 + 1;
    size_t pn = restn - tlen - 1;
    if (pn != 8 || memcmp(proto, "HTTP/1.1", 8) != 0) return 1;
    static const char* names[] = {"GET", "HEAD", "POST", "PUT", "PATCH", "DELETE", "", "OPTIONS"};
    int mcode = -1;
    for (int k = 0; k < 8; k++)
        if (names[k][0] && strlen(names[k]) == mlen && memcmp(rl, names[k], mlen) == 0) mcode = k;
    if (mcode < 0) return 1;
    if (tlen == 0 || tlen > 8192 || rest[0] != '/' || (tlen >= 2 && rest[1] == '/')) return 1;
    const uint8_t* q = NULL;
    for (size_t k = 0; k < tlen; k++) {
        uint8_t c = rest[k];
        if (c < 0x21 || c > 0x7E || c == '#') return 1;
        if (c == '?' && !q) q = rest + k;
    }
    size_t plen = q ? (size_t)(q - rest) : tlen;
    for (size_t k = 0; k < plen; k++)
        if (rest[k] == '%' && (k + 2 >= plen || hexv(rest[k + 1]) < 0 || hexv(rest[k + 2]) < 0)) return 1;
    /* ---- headers ---- */
    int n_host = 0, n_cl = 0, n_te = 0, have_ua = 0, have_xff = 0;
    uint64_t cl = 0;
    for (int li = 1; li < nl; li++) {
        const uint8_t* L = m + lines[li].off;
        size_t ln = lines[li].len;
        size_t c = 0;
        while (c < ln && is_tchar(L[c])) c++;
        if (c == 0 || c == ln || L[c] != ':') return 1; /* leading whitespace (obs-fold), empty name, junk in the name */
        size_t v0 = c + 1, v1 = ln;
        while (v0 < v1 && (L[v0] == ' ' || L[v0] == '\t')) v0++;
        while (v1 > v0 && (L[v1 - 1] == ' ' || L[v1 - 1] == '\t')) v1--;
        for (size_t k = c + 1; k < ln; k++)
            if ((L[k] < 0x20 && L[k] != '\t') || L[k] == 0x7F) return 1;
        span val = {lines[li].off + v0, v1 - v0};
        if (ieq(L, c, "host")) { n_host++; sp[SPAN_HOST] = val; }
        else if (ieq(L, c, "user-agent")) { if (!have_ua) { sp[SPAN_UA] = val; have_ua = 1; } }
        else if (ieq(L, c, "x-forwarded-for")) { if (!have_xff) { sp[SPAN_XFF] = val; have_xff = 1; } }
        else if (ieq(L, c, "content-length")) {
            n_cl++;
            if (val.len == 0 || val.len > 9) return 1;
            cl = 0;
            for (size_t k = 0; k < val.len; k++) {
                uint8_t d = m[val.off + k];
                if (d < '0' || d > '9') return 1;
                cl = cl * 10 + (uint64_t)(d - '0');
            }
        } else if (ieq(L, c, "connection")) {
            if (!ieq(m + val.off, val.len, "keep-alive")) return 1;
        } else if (ieq(L, c, "transfer-encoding")) {
            if (!ieq(m + val.off, val.len, "chunked")) return 1; /* other codings: readTransfer refuses them */
            n_te++;
        } else if (ieq(L, c, "expect") || ieq(L, c, "upgrade") || ieq(L, c, "trailer")) return 1;
    }
    if (n_host != 1 || n_cl > 1 || sp[SPAN_HOST].len == 0) return 1;
    if (n_te > 1 || (n_te && n_cl)) return 1;
    for (size_t k = 0; k < sp[SPAN_HOST].len; k++) {
        uint8_t c = m[sp[SPAN_HOST].off + k];
        if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || strchr(".:-_[]", c))) return 1;
    }
    size_t body = n - head_end;
    /* a chunked body is decoded into a scratch buffer first (chunk lines split off one by one) */
    uint8_t* dechunked = NULL;
    size_t raw_body = body;
    if (n_te) {
        dechunked = (uint8_t*)malloc(body + 1);
        size_t at = head_end, out = 0;
        int ok = 0;
        for (;;) {
            /* the size line: up to the next CRLF, hex digits only (no extension, no padding), 1..8 of them */
            size_t e = at;
            while (e < n && m[e] != '\r') e++;
            if (e + 1 >= n || m[e + 1] != '\n' || e == at || e - at > 8) break;
            uint64_t size = 0;
            int bad = 0;
            for (size_t k = at; k < e; k++) { int h = hexv(m[k]); if (h < 0) bad = 1; size = size * 16 + (uint64_t)(h < 0 ? 0 : h); }
            if (bad) break;
            at = e + 2;
            if (size == 0) { ok = at + 2 == n && m[at] == '\r' && m[at + 1] == '\n'; break; } /* no trailers, nothing after */
            if (size + 2 > n - at || m[at + size] != '\r' || m[at + size + 1] != '\n') break;
            memcpy(dechunked + out, m + at, (size_t)size);
            out += (size_t)size;
            at += (size_t)size + 2;
        }
        if (!ok) { free(dechunked); return 1; }
        body = out;
    } else if (n_cl ? (uint64_t)body != cl : body != 0) return 1;
    /* ---- URL.Path | URL.RawQuery | pad4 | body ---- */
    size_t w = 0;
    for (size_t k = 0; k < plen; k++) {
        uint8_t c = rest[k];
        if (c == '%') { c = (uint8_t)(hexv(rest[k + 1]) * 16 + hexv(rest[k + 2])); k += 2; }
        dst[w++] = c;
    }
    *path_len = (uint32_t)w;
    *query_len = 0;
    *flags = 0;
    if (q) {
        size_t qn = tlen - plen - 1;
        memcpy(dst + w, q + 1, qn);
        w += qn;
        *query_len = (uint32_t)qn;
        if (qn == 0) *flags = 1; /* ForceQuery */
    }
    while (w % 4) dst[w++] = 0;
    memcpy(dst + w, dechunked ? dechunked : m + head_end, body);
    free(dechunked);
    *data_len = (uint32_t)body;
    *method = (uint32_t)mcode;
    sp[SPAN_METHOD].off = 0; sp[SPAN_METHOD].len = mlen;
    sp[SPAN_TARGET].off = (size_t)(rest - m); sp[SPAN_TARGET].len = tlen;
    sp[SPAN_BODY].off = head_end; sp[SPAN_BODY].len = raw_body; /* chunked: the chunk stream as received */
    return 0;
}

int orc_http_parse(const uint8_t* raw, const uint32_t* raw_off, uint32_t n, void* desc_v, uint8_t* arena, uint32_t* status,
                   uint64_t* spans) {
    uint8_t* desc = (uint8_t*)desc_v;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t mo = raw_off[i], mn = raw_off[i + 1] - mo, a = (mo + 3u) & ~3u;
        uint32_t pl = 0, ql = 0, dl = 0, me = 0, fl = 0;
        span sp[N_SPANS];
        /* parse into a scratch copy first: a deferred message must leave the arena untouched beyond its own extent */
        uint8_t* tmp = (uint8_t*)malloc((size_t)mn + 16);
        int st = parse_one(raw + mo, mn, tmp, &pl, &ql, &dl, &me, &fl, sp);
        status[i] = (uint32_t)st;
        uint32_t d[4] = {0, 0, 0, 0};
        if (st == 0) {
            size_t total = (((size_t)pl + ql + 3) & ~(size_t)3) + dl;
            memcpy(arena + a, tmp, total);
            d[0] = a; d[1] = pl | ql << 16; d[2] = dl; d[3] = me | fl << 8;
        }
        free(tmp);
        memcpy(desc + (size_t)i * 16, d, 16);
        for (int k = 0; k < N_SPANS; k++)
            spans[(size_t)i * N_SPANS + k] = st == 0 ? ((uint64_t)(sp[k].off + mo) | (uint64_t)sp[k].len << 32) : 0;
    }
    return 0;
}

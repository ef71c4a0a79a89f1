// TEST INFRASTRUCTURE ONLY: the front-end's fan-in / fan-out (gofr_b200/csrc/frontend.cpp) on the CPU, under
// ThreadSanitizer, with a stub in place of the engine: the stub "serves" request i by writing a response derived from
// ITS path, body and trace id into slot i, so every producer can tell whether it got its own answer back.
//   frontend_stress <threads> <requests_per_thread> <max_batch> <max_wait_us>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../gofr_b200/csrc/frontend.cpp"

static std::atomic<uint64_t> g_batches{0}, g_largest{0};

void set_last_error(const char*, ...) {}
extern "C" {
void* gofr_alloc_pinned(size_t n) { return calloc(1, n ? n : 1); }
void gofr_free_pinned(void* p) { free(p); }
void gofr_format_http_date(int64_t t, char out[29]) { snprintf(out, 29, "%028lld", (long long)t); }
static uint32_t stub_response(const gofr_req_batch* in, uint32_t i, uint8_t* o) {
    const gofr_req_desc& d = in->desc[i];
    const uint8_t* p = in->arena + d.arena_off;
    const uint8_t* body = in->arena + ((d.arena_off + d.path_len + d.query_len + 3u) & ~3u);
    uint32_t w = 0;
    memcpy(o + w, in->date, 28); w += 28;
    memcpy(o + w, in->trace_ids + (size_t)i * 16, 16); w += 16;
    o[w++] = d.method; o[w++] = d.flags;
    memcpy(o + w, p, (size_t)d.path_len + d.query_len); w += d.path_len + d.query_len;
    memcpy(o + w, body, d.data_len); w += d.data_len;
    return w;
}
int gofr_batch_submit_slots(gofr_engine*, const gofr_req_batch* in, gofr_slot_batch* out, gofr_ticket* t) {
    g_batches++;
    uint64_t l = g_largest.load();
    while (in->n > l && !g_largest.compare_exchange_weak(l, in->n)) {}
    for (uint32_t i = 0; i < in->n; i++) {
        uint8_t tmp[512];
        const uint32_t w = stub_response(in, i, tmp);
        out->out_len[i] = w;
        if (w <= out->slot_bytes) memcpy(out->out + (size_t)i * out->slot_bytes, tmp, w);  // too long: length only, like the engine
        out->meta[i] = 200u | (uint32_t)in->desc[i].path_len << 16;
    }
    *t = 1;
    return 0;
}
static std::atomic<uint64_t> g_alone{0};
int gofr_batch_submit(gofr_engine*, const gofr_req_batch* in, gofr_resp_batch* out, gofr_ticket* t) {
    g_alone++;
    uint64_t pos = 0;
    for (uint32_t i = 0; i < in->n; i++) {
        out->out_off[i] = (uint32_t)pos;
        uint8_t tmp[512];
        const uint32_t w = stub_response(in, i, tmp);
        if (pos + w > out->out_cap) return GOFR_ERR_CAPACITY;
        memcpy(out->out + pos, tmp, w);
        out->meta[i] = 200u | (uint32_t)in->desc[i].path_len << 16;
        pos += w;
    }
    out->out_off[in->n] = (uint32_t)pos;
    out->out_bytes = pos;
    *t = 2;
    return 0;
}
int gofr_batch_wait(gofr_engine*, gofr_ticket) { return 0; }
}

int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 16, R = argc > 2 ? atoi(argv[2]) : 200;
    const uint32_t max_batch = argc > 3 ? (uint32_t)atoi(argv[3]) : 8, wait_us = argc > 4 ? (uint32_t)atoi(argv[4]) : 100;
    gofr_frontend* fe = nullptr;
    // a fifth argument of 64 makes every response longer than its slot: each one is then served alone (packed stub)
    const uint32_t slot = argc > 5 ? (uint32_t)atoi(argv[5]) : 256;
    if (gofr_frontend_create(&fe, (gofr_engine*)0x1, max_batch, wait_us, slot, 128)) return 2;
    gofr_frontend_set_clock(fe, 1700000000);
    std::atomic<uint64_t> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int r = 0; r < R; r++) {
                char path[64], body[64];
                uint8_t id[16], resp[256];
                const int pl = snprintf(path, sizeof path, "/t%d/r%d", t, r);
                const int ql = r % 3 ? snprintf(path + pl, sizeof path - pl, "k=%d", r) : 0;
                const int bl = r % 2 ? snprintf(body, sizeof body, "{\"t\":%d,\"r\":%d}", t, r) : 0;
                for (int k = 0; k < 16; k++) id[k] = (uint8_t)(t * 31 + r * 7 + k);
                uint32_t len = 0, meta = 0;
                int rc = gofr_frontend_serve(fe, (uint8_t)(r % 6), (const uint8_t*)path, (uint16_t)pl, (const uint8_t*)path + pl,
                                             (uint16_t)ql, (uint8_t)(r & 1), (const uint8_t*)body, (uint32_t)bl, id, resp, sizeof resp,
                                             &len, &meta);
                bool ok = rc == 0 && len == (uint32_t)(28 + 16 + 2 + pl + ql + bl) && meta == (200u | (uint32_t)pl << 16) &&
                          memcmp(resp, "0000000000000000001700000000", 28) == 0 && memcmp(resp + 28, id, 16) == 0 &&
                          resp[44] == (uint8_t)(r % 6) && resp[45] == (uint8_t)(r & 1) && memcmp(resp + 46, path, (size_t)pl + ql) == 0 &&
                          memcmp(resp + 46 + pl + ql, body, (size_t)bl) == 0;
                if (!ok) bad++;
            }
        });
    for (auto& x : th) x.join();
    uint64_t batches = 0, reqs = 0;
    gofr_frontend_stats(fe, &batches, &reqs);
    // capacity: a response that does not fit the caller's buffer is reported, not truncated
    {
        uint8_t id[16] = {0}, small[8];
        uint32_t len = 0;
        int rc = gofr_frontend_serve(fe, 0, (const uint8_t*)"/x", 2, nullptr, 0, 0, nullptr, 0, id, small, sizeof small, &len, nullptr);
        if (rc != GOFR_ERR_CAPACITY || len != 48) bad++;
        uint8_t big[512] = {0};
        rc = gofr_frontend_serve(fe, 0, big, 300, nullptr, 0, 0, nullptr, 0, id, big, sizeof big, &len, nullptr);  // > max_request_bytes
        if (rc != GOFR_ERR_CAPACITY) bad++;
    }
    gofr_frontend_destroy(fe);
    printf("{\"requests\": %llu, \"batches\": %llu, \"stub_batches\": %llu, \"largest\": %llu, \"alone\": %llu, \"bad\": %llu}\n",
           (unsigned long long)reqs, (unsigned long long)batches, (unsigned long long)g_batches.load(),
           (unsigned long long)g_largest.load(), (unsigned long long)g_alone.load(), (unsigned long long)bad.load());
    return bad.load() || reqs != (uint64_t)T * R ? 1 : 0;
}

// One process, several GPUs, no Python and no NCCL: the table is sealed ONCE and every device gets its own engine created
// from it (gofr_engine_create uploads the sealed image to that device) — what a Go host does with one goroutine group per
// GPU.  The reference has nothing to compare with (one process, one router, pkg/gofr/httpServer.go:29-33); the multi-process
// harness (bench.py) broadcasts the serialized image over NCCL instead (gofr_table_serialize / _deserialize).
// Each thread binds itself to its GPU's NUMA node, serves ITS contiguous shard of one request stream through host buffers,
// and the shards are concatenated and compared with the whole stream served by engine 0 alone.
//   g++ -std=c++17 -Iinclude examples/cpp/two_engines.cpp gofr_b200/libgofr_b200.so -lpthread -o two_engines
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gofr_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != GOFR_OK) { fprintf(stderr, "%s: %d %s\n", #x, rc_, gofr_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    const int n_engines = argc > 1 ? atoi(argv[1]) : 2;
    const int n_devices = argc > 2 ? atoi(argv[2]) : 1;  // engines k runs on device k % n_devices
    gofr_table* t = nullptr;
    CK(gofr_table_create(&t, GOFR_FRAME_WIRE));
    const char* pats[] = {"/hello", "/greet", "/items/{id}"};
    for (int k = 0; k < 3; k++) {
        gofr_handler_desc h;
        memset(&h, 0, sizeof h);
        h.kind = k == 0 ? GOFR_H_STATIC_STRING : k == 1 ? GOFR_H_PARAM_FORMAT : GOFR_H_PATHPARAM_FORMAT;
        h.s0 = k == 0 ? "Hello World!" : k == 1 ? "name" : "id";
        h.s1 = "World"; h.s2 = k == 2 ? "item " : "Hello "; h.s3 = "!";
        h.s0_len = (uint32_t)strlen(h.s0); h.s1_len = (uint32_t)strlen(h.s1); h.s2_len = (uint32_t)strlen(h.s2); h.s3_len = (uint32_t)strlen(h.s3);
        uint32_t id = 0;
        CK(gofr_table_add_route(t, GOFR_M_GET, pats[k], (uint32_t)strlen(pats[k]), &h, &id));
    }
    CK(gofr_table_add_default_routes(t, nullptr, 0));
    CK(gofr_table_seal(t));
    std::vector<gofr_engine*> eng(n_engines, nullptr);
    for (int k = 0; k < n_engines; k++) CK(gofr_engine_create(&eng[k], t, k % n_devices));

    // one request stream
    const uint32_t n = 40000, slot = 512;
    std::vector<gofr_req_desc> desc(n);
    std::vector<uint8_t> ids((size_t)n * 16);
    std::string arena;
    for (uint32_t i = 0; i < n; i++) {
        char path[64], query[64] = "";
        if (i % 3 == 0) snprintf(path, sizeof path, "/hello");
        else if (i % 3 == 1) { snprintf(path, sizeof path, "/greet"); snprintf(query, sizeof query, "name=user%u", i); }
        else snprintf(path, sizeof path, "/items/%u", i);
        memset(&desc[i], 0, sizeof desc[i]);
        desc[i].arena_off = (uint32_t)arena.size();
        desc[i].path_len = (uint16_t)strlen(path);
        desc[i].query_len = (uint16_t)strlen(query);
        desc[i].method = i % 50 == 7 ? GOFR_M_POST : GOFR_M_GET;
        arena += path; arena += query;
        arena.append((4 - arena.size() % 4) % 4, '\0');
        for (int k = 0; k < 16; k++) ids[(size_t)i * 16 + k] = (uint8_t)(i * 31 + k);
    }
    arena.append(64, '\0');
    char date[29];
    gofr_format_http_date(1700000000, date);

    auto serve = [&](gofr_engine* e, uint32_t lo, uint32_t hi, uint8_t* out, uint32_t* len, uint32_t* meta) -> int {
        gofr_req_batch in;
        memset(&in, 0, sizeof in);
        in.desc = desc.data() + lo; in.trace_ids = ids.data() + (size_t)lo * 16; in.arena = (const uint8_t*)arena.data();
        in.arena_bytes = arena.size(); in.n = hi - lo;
        memcpy(in.date, date, 29);
        gofr_slot_batch ob;
        memset(&ob, 0, sizeof ob);
        ob.out = out; ob.slot_bytes = slot; ob.out_len = len; ob.meta = meta;
        gofr_ticket tk = 0;
        int rc = gofr_batch_submit_slots(e, &in, &ob, &tk);
        return rc != GOFR_OK ? rc : gofr_batch_wait(e, tk);
    };
    std::vector<uint8_t> whole((size_t)n * slot), parts((size_t)n * slot);
    std::vector<uint32_t> wlen(n), wmeta(n), plen(n), pmeta(n);
    CK(serve(eng[0], 0, n, whole.data(), wlen.data(), wmeta.data()));
    std::vector<std::thread> th;
    std::vector<int> rcs(n_engines, 0);
    for (int k = 0; k < n_engines; k++)
        th.emplace_back([&, k] {
            gofr_bind_host_thread(k % n_devices, nullptr);  // best effort: the shard's buffers are caller memory here
            const uint32_t lo = (uint32_t)((uint64_t)n * k / n_engines), hi = (uint32_t)((uint64_t)n * (k + 1) / n_engines);
            for (int rep = 0; rep < 3 && !rcs[k]; rep++)
                rcs[k] = serve(eng[k], lo, hi, parts.data() + (size_t)lo * slot, plen.data() + lo, pmeta.data() + lo);
        });
    for (auto& x : th) x.join();
    for (int k = 0; k < n_engines; k++) if (rcs[k]) { fprintf(stderr, "engine %d: %d %s\n", k, rcs[k], gofr_last_error()); return 1; }
    size_t bad = 0;
    for (uint32_t i = 0; i < n; i++)
        if (wlen[i] != plen[i] || wmeta[i] != pmeta[i] || memcmp(&whole[(size_t)i * slot], &parts[(size_t)i * slot], wlen[i] <= slot ? wlen[i] : 0)) bad++;
    printf("%d engines on %d device(s): %u requests, %zu responses differ from the single-engine result; sample: %.*s\n", n_engines, n_devices, n, bad,
           (int)(wlen[1] < 60 ? wlen[1] : 60), (const char*)&whole[slot]);
    for (auto* e : eng) gofr_engine_destroy(e);
    gofr_table_destroy(t);
    return bad ? 1 : 0;
}

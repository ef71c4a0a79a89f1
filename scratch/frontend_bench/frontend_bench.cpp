// frontend_bench.cpp — closed-loop load on gofr_frontend_serve: T threads ("connections"), each sends its next request as
// soon as the previous response is back.  Reports requests/s and per-request latency percentiles.
//   frontend_bench <table.img> <desc.bin> <ids.bin> <arena.bin> <threads> <max_batch> <max_wait_us> <seconds>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/gofr_b200.h"

static std::vector<uint8_t> slurp(const char* p) {
    FILE* f = fopen(p, "rb");
    if (!f) { perror(p); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v((size_t)n);
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) { perror(p); exit(2); }
    fclose(f);
    return v;
}
#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, gofr_last_error()); exit(1); } } while (0)

int main(int argc, char** argv) {
    if (argc < 9) { fprintf(stderr, "usage\n"); return 2; }
    auto img = slurp(argv[1]), descb = slurp(argv[2]), ids = slurp(argv[3]), arena = slurp(argv[4]);
    const int T = atoi(argv[5]);
    const uint32_t max_batch = (uint32_t)atoi(argv[6]), wait_us = (uint32_t)atoi(argv[7]);
    const double secs = atof(argv[8]);
    const gofr_req_desc* desc = (const gofr_req_desc*)descb.data();
    const size_t n = descb.size() / sizeof(gofr_req_desc);
    gofr_table* tab = nullptr;
    CK(gofr_table_deserialize(&tab, img.data(), img.size()));
    gofr_engine* eng = nullptr;
    CK(gofr_engine_create(&eng, tab, 0));
    gofr_frontend* fe = nullptr;
    CK(gofr_frontend_create(&fe, eng, max_batch, wait_us, 1024, 2048));
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> bad{0};
    std::vector<std::vector<float>> lat((size_t)T);
    std::vector<std::vector<float>> slow_at((size_t)T);  // start offsets (ms) of requests slower than 20 ms
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            uint8_t resp[1024];
            size_t i = (size_t)t % n;
            while (!stop.load(std::memory_order_relaxed)) {
                const gofr_req_desc& d = desc[i];
                const uint8_t* a = arena.data() + d.arena_off;
                const uint8_t* body = arena.data() + ((d.arena_off + d.path_len + d.query_len + 3u) & ~3u);
                uint32_t len = 0, meta = 0;
                auto s = std::chrono::steady_clock::now();
                int rc = gofr_frontend_serve(fe, d.method, a, d.path_len, a + d.path_len, d.query_len, d.flags, body, d.data_len,
                                             ids.data() + i * 16, resp, sizeof resp, &len, &meta);
                auto e = std::chrono::steady_clock::now();
                if (rc || len < 17 || resp[0] != 'H') bad++;
                lat[(size_t)t].push_back(std::chrono::duration<float, std::micro>(e - s).count());
                if (e - s > std::chrono::milliseconds(20)) slow_at[(size_t)t].push_back(std::chrono::duration<float, std::milli>(s - t0).count());
                i = (i + (size_t)T) % n;
            }
        });
    std::this_thread::sleep_for(std::chrono::duration<double>(secs));
    stop = true;
    for (auto& x : th) x.join();
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<float> all;
    for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    if (getenv("FRONTEND_BENCH_SLOW")) {
        std::vector<float> sl;
        for (auto& v : slow_at) sl.insert(sl.end(), v.begin(), v.end());
        std::sort(sl.begin(), sl.end());
        fprintf(stderr, "slow (>20 ms) requests: %zu; start offsets (ms), every %zu-th:", sl.size(), sl.size() / 40 + 1);
        for (size_t k = 0; k < sl.size(); k += sl.size() / 40 + 1) fprintf(stderr, " %.0f", sl[k]);
        fprintf(stderr, "\n");
    }
    uint64_t batches = 0, reqs = 0;
    gofr_frontend_stats(fe, &batches, &reqs);
    auto pct = [&](double p) { return all.empty() ? 0.f : all[(size_t)(p * (all.size() - 1))]; };
    printf("{\"threads\": %d, \"max_batch\": %u, \"max_wait_us\": %u, \"requests\": %zu, \"req_per_s\": %.0f, \"batches\": %llu, "
           "\"mean_batch\": %.1f, \"lat_us_p50\": %.1f, \"lat_us_p99\": %.1f, \"bad\": %llu}\n",
           T, max_batch, wait_us, all.size(), all.size() / el, (unsigned long long)batches, batches ? (double)reqs / batches : 0.0,
           pct(0.5), pct(0.99), (unsigned long long)bad.load());
    gofr_frontend_destroy(fe);
    gofr_engine_destroy(eng);
    gofr_table_destroy(tab);
    return bad.load() ? 1 : 0;
}

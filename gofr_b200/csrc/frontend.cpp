// frontend.cpp — batching front-end over gofr_batch_submit_slots ("collect in-flight requests from
// conn goroutines into the ring, fan results back").
//
// The reference serves one request per goroutine (net/http conn.serve → router.ServeHTTP, pkg/gofr/httpServer.go:29-33);
// the GPU path wants batches.  This is the piece in between, host C++ only: any number of producer threads (connection
// goroutines through cgo, or the threads of the C++ stand-in) hand in single requests and block until their response is
// ready; one dispatcher thread closes a batch when it is full or when its oldest request has waited max_wait_us, runs it
// through the engine's slot-layout host path and wakes the producers, each of which copies its own slot out.
// Two pinned batches alternate: while one is in flight the other fills.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <time.h>

#include "../../include/gofr_b200.h"
#include "engine_internal.h"

using namespace gofr;

struct gofr_frontend_batch {
    // pinned staging in the ABI layout (gofr_req_batch / gofr_slot_batch)
    gofr_req_desc* desc = nullptr;
    uint8_t* ids = nullptr;
    uint8_t* arena = nullptr;
    uint8_t* out = nullptr;
    uint32_t* out_len = nullptr;
    uint32_t* meta = nullptr;
    enum State { FILLING, IN_FLIGHT, DONE } state = FILLING;
    uint32_t count = 0;      // requests appended
    std::atomic<uint32_t> collected{0};  // producers that have copied their response out (DONE state)
    std::condition_variable cv_ready;    // this batch's producers: "your round is DONE"
    uint32_t arena_used = 0;
    uint64_t generation = 0;  // bumped when the batch is recycled: producers wait for "their" generation to complete
    int rc = GOFR_OK;
    std::chrono::steady_clock::time_point first_arrival;
};
using Batch = gofr_frontend_batch;

struct gofr_frontend {
    gofr_engine* eng = nullptr;
    uint32_t max_batch = 0, max_wait_us = 0, slot_bytes = 0, arena_cap = 0;
    int64_t fixed_clock = 0;  // tests: the Date of every batch; 0 = wall clock
    Batch b[2];
    int filling = 0;          // index of the batch producers append to
    bool stop = false;
    std::mutex mu;
    std::condition_variable cv_dispatch;  // dispatcher: "a batch may be ready"
    std::condition_variable cv_room;      // producers without a place yet: "the filling batch changed / was recycled"
    std::thread worker;
    // counters
    uint64_t batches = 0, requests = 0;
};

static void free_batch(Batch& x) {
    gofr_free_pinned(x.desc); gofr_free_pinned(x.ids); gofr_free_pinned(x.arena);
    gofr_free_pinned(x.out); gofr_free_pinned(x.out_len); gofr_free_pinned(x.meta);
}

static void dispatcher(gofr_frontend* f) {
    std::unique_lock<std::mutex> lk(f->mu);
    for (;;) {
        Batch& cur = f->b[f->filling];
        if (cur.state != Batch::FILLING) { f->cv_dispatch.wait(lk); continue; }  // still being collected by its producers
        if (cur.count == 0) {
            if (f->stop) return;
            f->cv_dispatch.wait(lk);
            continue;
        }
        if (cur.count < f->max_batch && !f->stop) {
            // not full: wait until the oldest request has been here max_wait_us (or the batch fills up)
            auto deadline = cur.first_arrival + std::chrono::microseconds(f->max_wait_us);
            if (std::chrono::steady_clock::now() < deadline) { f->cv_dispatch.wait_until(lk, deadline); continue; }
        }
        // close the batch: producers now fill the other one (as soon as its previous occupants have left)
        cur.state = Batch::IN_FLIGHT;
        f->filling ^= 1;
        f->cv_room.notify_all();
        Batch& x = cur;
        gofr_req_batch in;
        memset(&in, 0, sizeof in);
        in.desc = x.desc; in.trace_ids = x.ids; in.arena = x.arena; in.arena_bytes = (x.arena_used + 15u) & ~15u; in.n = x.count;
        int64_t now = f->fixed_clock;
        if (!now) { struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts); now = (int64_t)ts.tv_sec; }
        gofr_format_http_date(now, in.date);
        gofr_slot_batch out;
        memset(&out, 0, sizeof out);
        out.out = x.out; out.slot_bytes = f->slot_bytes; out.out_len = x.out_len; out.meta = x.meta;
        lk.unlock();
        gofr_ticket t = 0;
        int rc = gofr_batch_submit_slots(f->eng, &in, &out, &t);
        if (rc == GOFR_OK) rc = gofr_batch_wait(f->eng, t);
        lk.lock();
        x.rc = rc;
        x.state = Batch::DONE;
        f->batches++;
        f->requests += x.count;
        x.cv_ready.notify_all();
    }
}

extern "C" {

int gofr_frontend_create(gofr_frontend** out, gofr_engine* e, uint32_t max_batch, uint32_t max_wait_us, uint32_t slot_bytes,
                         uint32_t max_request_bytes) {
    if (!out || !e || max_batch == 0 || slot_bytes == 0 || (slot_bytes & 15u)) return GOFR_ERR_INVALID;
    gofr_frontend* f = new gofr_frontend();
    f->eng = e;
    f->max_batch = max_batch;
    f->max_wait_us = max_wait_us;
    f->slot_bytes = slot_bytes;
    const uint64_t cap = (uint64_t)max_batch * (((uint64_t)max_request_bytes + 7u) & ~(uint64_t)3u) + 64;
    if (cap > 0xFFFFFFF0ull) { delete f; return GOFR_ERR_CAPACITY; }
    f->arena_cap = (uint32_t)cap;
    for (auto& x : f->b) {
        x.desc = (gofr_req_desc*)gofr_alloc_pinned((size_t)max_batch * sizeof(gofr_req_desc));
        x.ids = (uint8_t*)gofr_alloc_pinned((size_t)max_batch * 16);
        x.arena = (uint8_t*)gofr_alloc_pinned(f->arena_cap);
        x.out = (uint8_t*)gofr_alloc_pinned((size_t)max_batch * slot_bytes);
        x.out_len = (uint32_t*)gofr_alloc_pinned((size_t)max_batch * 4);
        x.meta = (uint32_t*)gofr_alloc_pinned((size_t)max_batch * 4);
        if (!x.desc || !x.ids || !x.arena || !x.out || !x.out_len || !x.meta) {
            for (auto& y : f->b) free_batch(y);
            delete f;
            return GOFR_ERR_NOMEM;
        }
    }
    f->worker = std::thread(dispatcher, f);
    *out = f;
    return GOFR_OK;
}

void gofr_frontend_destroy(gofr_frontend* f) {
    if (!f) return;
    {
        std::lock_guard<std::mutex> g(f->mu);
        f->stop = true;
    }
    f->cv_dispatch.notify_all();
    if (f->worker.joinable()) f->worker.join();
    for (auto& x : f->b) free_batch(x);
    delete f;
}

int gofr_frontend_set_clock(gofr_frontend* f, int64_t unix_seconds) {
    if (!f) return GOFR_ERR_INVALID;
    std::lock_guard<std::mutex> g(f->mu);
    f->fixed_clock = unix_seconds;
    return GOFR_OK;
}

int gofr_frontend_stats(gofr_frontend* f, uint64_t* batches, uint64_t* requests) {
    if (!f) return GOFR_ERR_INVALID;
    std::lock_guard<std::mutex> g(f->mu);
    if (batches) *batches = f->batches;
    if (requests) *requests = f->requests;
    return GOFR_OK;
}

int gofr_frontend_serve(gofr_frontend* f, uint8_t method, const uint8_t* path, uint16_t path_len, const uint8_t* query,
                        uint16_t query_len, uint8_t flags, const uint8_t* data, uint32_t data_len, const uint8_t trace_id[16],
                        uint8_t* resp, uint32_t resp_cap, uint32_t* resp_len, uint32_t* meta) {
    if (!f || !resp_len || !trace_id || (path_len && !path) || (query_len && !query) || (data_len && !data)) return GOFR_ERR_INVALID;
    const uint32_t need = (((uint32_t)path_len + query_len + 3u) & ~3u) + ((data_len + 3u) & ~3u);
    if (need > f->arena_cap / f->max_batch) { set_last_error("request of %u bytes exceeds the front-end's max_request_bytes", need); return GOFR_ERR_CAPACITY; }
    std::unique_lock<std::mutex> lk(f->mu);
    // room in the filling batch?  (it can be full, or still be handing out the responses of its previous round)
    for (;;) {
        if (f->stop) return GOFR_ERR_INVALID;
        Batch& cur = f->b[f->filling];
        if (cur.state == Batch::FILLING && cur.count < f->max_batch && cur.arena_used + need <= f->arena_cap) break;
        if (cur.state == Batch::FILLING) f->cv_dispatch.notify_one();  // full: the dispatcher should close it now
        f->cv_room.wait(lk);
    }
    Batch& x = f->b[f->filling];
    const uint32_t i = x.count++;
    if (i == 0) x.first_arrival = std::chrono::steady_clock::now();
    const uint64_t gen = x.generation;
    uint32_t a = x.arena_used;
    gofr_req_desc d;
    memset(&d, 0, sizeof d);
    d.arena_off = a; d.path_len = path_len; d.query_len = query_len; d.data_len = data_len; d.method = method; d.flags = flags;
    if (path_len) memcpy(x.arena + a, path, path_len);
    if (query_len) memcpy(x.arena + a + path_len, query, query_len);
    a = (a + path_len + query_len + 3u) & ~3u;
    if (data_len) memcpy(x.arena + a, data, data_len);
    x.arena_used = (a + data_len + 3u) & ~3u;
    x.desc[i] = d;
    memcpy(x.ids + (size_t)i * 16, trace_id, 16);
    if (i == 0 || x.count == f->max_batch) f->cv_dispatch.notify_one();  // start the batch's timer / close a full batch
    // wait for this round of the batch to come back; the copy-out happens outside the lock
    while (!(x.generation == gen && x.state == Batch::DONE)) x.cv_ready.wait(lk);
    int rc = x.rc;
    const uint32_t count = x.count;
    lk.unlock();
    const uint32_t len = x.out_len[i];
    if (meta) *meta = x.meta[i];
    *resp_len = len;
    if (rc == GOFR_OK) {
        if (len > f->slot_bytes || len > resp_cap) rc = GOFR_ERR_CAPACITY;  // the caller serves it through the packed path
        else if (len) memcpy(resp, x.out + (size_t)i * f->slot_bytes, len);
    }
    // the last producer to leave recycles the batch
    if (x.collected.fetch_add(1, std::memory_order_acq_rel) + 1 == count) {
        lk.lock();
        x.state = Batch::FILLING;
        x.count = 0;
        x.collected.store(0, std::memory_order_relaxed);
        x.arena_used = 0;
        x.generation++;
        f->cv_room.notify_all();
        f->cv_dispatch.notify_one();
    }
    return rc;
}

}  // extern "C"

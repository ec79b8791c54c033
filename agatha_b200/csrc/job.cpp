// Whole-job API: any number of pairs over any number of GPUs of one box (agatha_align_job, include/agatha_b200.h).
// The reference has no multi-GPU path (gasal_set_device exists, interfaces.cpp:86-116, but its only call is commented
// out, test_prog.cpp:31). Pairs are independent, so there is no collective: a host scheduler balances estimated work
// (cells inside the band) over the devices, one worker thread per device drives double-buffered streams, and results
// are scattered back by original index.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <numeric>
#include <parallel/algorithm>
#include <string>
#include <thread>
#include <vector>

#include "agatha_b200.h"
#include "engine_internal.h"

using namespace agatha;

namespace {

// Streams (pinned staging + device buffers) are expensive to create; keep them between jobs, per device.
std::mutex g_pool_mu;
std::map<int, std::vector<agatha_stream_t*>> g_pool;

agatha_stream_t* pool_get(int device, uint32_t batch_alns)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto& v = g_pool[device];
        if (!v.empty()) { agatha_stream_t* s = v.back(); v.pop_back(); return s; }
    }
    return agatha_stream_create(device, batch_alns, 1 << 20, 1 << 20);
}

void pool_put(int device, agatha_stream_t* s)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool[device].push_back(s);
}

struct Batch { agatha_stream_t* s = nullptr; std::vector<uint64_t> ids; bool busy = false; };

// The packing threads of one device worker. Threads SLEEP between batches (condition variable): with one process per GPU on
// one box there are as many of these pools as GPUs, next to the CUDA driver's and the caller's own threads, and an OpenMP
// team's spinning waiters (libgomp's default) take the cores the other ranks' packers need -- measured here: the same
// packing loop got slower with every thread added once the waiters of several teams had to share cores.
class PackPool {
public:
    explicit PackPool(int n_threads)
    {
        for (int i = 1; i < n_threads; i++) threads_.emplace_back([this] { loop(); });
    }
    ~PackPool()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_work_.notify_all();
        for (auto& t : threads_) t.join();
    }
    // fn(first, last) over [0, n) in chunks; the calling thread works too; returns when everything is done
    void run(uint64_t n, uint64_t chunk, const std::function<void(uint64_t, uint64_t)>& fn)
    {
        if (n == 0) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; chunk_ = std::max<uint64_t>(chunk, 1); next_.store(0, std::memory_order_relaxed);
            pending_ = (int)threads_.size();
            gen_++;
        }
        cv_work_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work()
    {
        for (;;) {
            const uint64_t at = next_.fetch_add(chunk_, std::memory_order_relaxed);
            if (at >= n_) break;
            (*fn_)(at, std::min(n_, at + chunk_));
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(uint64_t, uint64_t)>* fn_ = nullptr;
    uint64_t n_ = 0, chunk_ = 1, gen_ = 0;
    std::atomic<uint64_t> next_{0};
    int pending_ = 0;
    bool stop_ = false;
};

struct Worker {
    int device = 0;
    std::vector<uint64_t> pairs;       // pair indices of this device, most expensive first
    int rc = AGATHA_OK;
    std::string err;
    double kernel_ms = 0;
    uint64_t h2d = 0, d2h = 0;
    uint32_t batches = 0;
};

struct JobView {
    const uint8_t *qb, *tb;
    const uint64_t *qo, *to;
    const uint32_t *ql, *tl;
    const agatha_params_t* params;
    const uint8_t *qops, *tops;        // optional per-pair reverse/complement ops
    int32_t *score, *qend, *tend, *stop, *dstop;
};

// Stage one batch with gasal_host_batch_fill's layout (host_batch.cpp:79-154: each sequence at a multiple of 8, padded with
// 'N') but already in the packed device format, per-pair ops applied on the way: half the H2D bytes, no pack kernel.
int fill_batch(Batch& b, const JobView& jv, PackPool& pool, uint64_t& qbases, uint64_t& tbases)
{
    const uint64_t n = b.ids.size();
    const uint64_t qtot = agatha_staged_bytes(jv.ql, b.ids.data(), n), ttot = agatha_staged_bytes(jv.tl, b.ids.data(), n);
    if (qtot > 0xfffffff8ull || ttot > 0xfffffff8ull) return set_error(AGATHA_EINVAL, "batch exceeds 4 GiB of bases; lower batch_alns");
    int rc = agatha_stream_reserve(b.s, (uint32_t)n, qtot, ttot);
    if (rc) return rc;
    rc = pack_layout(jv.ql, b.ids.data(), n, qtot / 8, agatha_stream_query_offsets(b.s), agatha_stream_query_lens(b.s), &qbases);
    if (rc) return rc;
    rc = pack_layout(jv.tl, b.ids.data(), n, ttot / 8, agatha_stream_target_offsets(b.s), agatha_stream_target_lens(b.s), &tbases);
    if (rc) return rc;
    const PackView qv{jv.qb, jv.qo, jv.ql, b.ids.data(), jv.qops, 0, agatha_stream_query_packed(b.s), agatha_stream_query_offsets(b.s)};
    const PackView tv{jv.tb, jv.to, jv.tl, b.ids.data(), jv.tops, 1, agatha_stream_target_packed(b.s), agatha_stream_target_offsets(b.s)};
    // a batch without a single base still uploads one word: make it padding (a batch of one short sequence overwrites it)
    if (qbases == 8) pack_empty(0, agatha_stream_query_packed(b.s));
    if (tbases == 8) pack_empty(1, agatha_stream_target_packed(b.s));
    // one pass over both sides of the batch: items [0, n) are the reads, [n, 2n) the reference windows
    pool.run(2 * n, 16, [&](uint64_t a, uint64_t e) {
        if (a < n) pack_range(qv, a, std::min(e, n));
        if (e > n) pack_range(tv, std::max(a, n) - n, e - n);
    });
    return AGATHA_OK;
}

void collect(Batch& b, const JobView& jv, Worker& w)
{
    const int32_t *sc = agatha_stream_scores(b.s), *qe = agatha_stream_query_ends(b.s), *te = agatha_stream_target_ends(b.s);
    const int32_t *sp = agatha_stream_stops(b.s), *ds = agatha_stream_dstops(b.s);
    for (size_t j = 0; j < b.ids.size(); j++) {
        const uint64_t id = b.ids[j];
        jv.score[id] = sc[j]; jv.qend[id] = qe[j]; jv.tend[id] = te[j];
        if (jv.stop) jv.stop[id] = sp[j];
        if (jv.dstop) jv.dstop[id] = ds[j];
    }
    float ms[3];
    agatha_stream_timings(b.s, ms);
    w.kernel_ms += ms[1];
    w.d2h += 20ull * b.ids.size();
    b.busy = false;
}

void run_worker(Worker& w, const JobView& jv, uint32_t batch_alns, int n_streams, int fill_threads)
{
    std::vector<Batch> bs((size_t)n_streams);
    auto fail = [&](int rc) { w.rc = rc; w.err = agatha_last_error(); };
    for (auto& b : bs) {
        b.s = pool_get(w.device, batch_alns);
        if (!b.s) { fail(AGATHA_ECUDA); break; }
    }
    // size every stream for the largest batch of this job once, before the pipeline starts: no (re)allocation of pinned
    // or device memory while batches are in flight, and none at all when the cached streams already fit
    if (w.rc == AGATHA_OK) {
        uint64_t qmax = 8, tmax = 8;
        for (size_t at = 0; at < w.pairs.size(); at += batch_alns) {
            const size_t cnt = std::min<size_t>(batch_alns, w.pairs.size() - at);
            qmax = std::max(qmax, agatha_staged_bytes(jv.ql, w.pairs.data() + at, cnt));
            tmax = std::max(tmax, agatha_staged_bytes(jv.tl, w.pairs.data() + at, cnt));
        }
        if (qmax > 0xfffffff8ull || tmax > 0xfffffff8ull) { set_error(AGATHA_EINVAL, "batch exceeds 4 GiB of bases; lower batch_alns"); fail(AGATHA_EINVAL); }
        for (auto& b : bs) {
            if (w.rc != AGATHA_OK) break;
            int rc = agatha_stream_reserve(b.s, (uint32_t)std::min<size_t>(batch_alns, w.pairs.size()), qmax, tmax);
            if (rc) fail(rc);
        }
    }
    PackPool pool(fill_threads);
    size_t next = 0;
    int cur = 0;
    // The first batches are small (1/8, 1/4, 1/2 of batch_alns): the device starts after an eighth of a batch has been packed
    // instead of a whole one -- with the longest pairs first that is most of what the job waits for before its first kernel.
    uint32_t ramp = std::max<uint32_t>(batch_alns / 8, std::min<uint32_t>(batch_alns, 256));
    while (w.rc == AGATHA_OK && next < w.pairs.size()) {
        Batch& b = bs[(size_t)cur];
        if (b.busy) {                                  // oldest batch of this slot: wait for it, then reuse its buffers
            int rc = agatha_stream_wait(b.s);
            if (rc) { fail(rc); break; }
            collect(b, jv, w);
        }
        const size_t cnt = std::min<size_t>(std::min(ramp, batch_alns), w.pairs.size() - next);
        ramp = ramp >= batch_alns / 2 ? batch_alns : ramp * 2;
        b.ids.assign(w.pairs.begin() + (long)next, w.pairs.begin() + (long)(next + cnt));
        next += cnt;
        uint64_t qbytes = 0, tbytes = 0;
        int rc = fill_batch(b, jv, pool, qbytes, tbytes);
        if (!rc) rc = agatha_stream_submit_packed(b.s, qbytes, tbytes, (uint32_t)cnt, jv.params);
        if (rc) { fail(rc); break; }
        b.busy = true;
        w.h2d += (qbytes + tbytes) / 2 + 20ull * cnt;
        w.batches++;
        cur = (cur + 1) % n_streams;
    }
    for (auto& b : bs) {
        if (b.s && b.busy) {
            int rc = agatha_stream_wait(b.s);
            if (rc && w.rc == AGATHA_OK) fail(rc);
            if (!rc) collect(b, jv, w);
        }
    }
    // back to the pool (same stream, same slot next time) -- unless something failed: a stream whose submit or wait failed may
    // still have copies or kernels in flight on its staging, or be stuck in the "submitted" state; destroy those instead
    for (auto it = bs.rbegin(); it != bs.rend(); ++it) {
        if (!it->s) continue;
        if (w.rc == AGATHA_OK) pool_put(w.device, it->s); else agatha_stream_destroy(it->s);
    }
}

}  // namespace

extern "C" void agatha_release_cached(void)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& kv : g_pool) for (agatha_stream_t* s : kv.second) agatha_stream_destroy(s);
    g_pool.clear();
}

extern "C" int agatha_align_job(const uint8_t* query_bases, const uint64_t* query_offsets, const uint32_t* query_lens,
                                const uint8_t* target_bases, const uint64_t* target_offsets, const uint32_t* target_lens,
                                uint64_t n_alns, const agatha_params_t* params, const agatha_job_config_t* cfg,
                                int32_t* score, int32_t* query_end, int32_t* target_end, int32_t* stop, int32_t* dstop,
                                agatha_job_stats_t* stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (!params || !score || !query_end || !target_end) return set_error(AGATHA_EINVAL, "NULL argument");
    if (n_alns == 0) return AGATHA_OK;
    if (!query_bases || !target_bases || !query_offsets || !target_offsets || !query_lens || !target_lens) return set_error(AGATHA_EINVAL, "NULL argument");
    const int visible = agatha_device_count();
    if (visible == 0) return set_error(AGATHA_ENODEV, "no CUDA device (agatha_b200 has no CPU fallback)");
    int ndev = (cfg && cfg->n_devices > 0) ? cfg->n_devices : visible;
    std::vector<int> devs((size_t)ndev);
    for (int i = 0; i < ndev; i++) {
        devs[(size_t)i] = (cfg && cfg->devices) ? cfg->devices[i] : i;
        if (devs[(size_t)i] < 0 || devs[(size_t)i] >= visible) return set_error(AGATHA_EINVAL, "device %d not visible (%d devices)", devs[(size_t)i], visible);
    }
    const uint32_t batch_alns = (cfg && cfg->batch_alns) ? cfg->batch_alns : 8192u;   // the reference's kernel_align_num default (args_parser.cpp:23)
    // Batches in flight per device. Measured on 200 k ONT-like pairs (one B200, kernels of consecutive batches overlap on their
    // streams): 1 stream 933 ms, 2: 763, 3: 744, 4: 713-720, 5: 715, 6-12: 713-715; the same pairs resident in HBM, one launch:
    // 702 ms. With three, two kernels that share the device end together and the single worker thread refills one stream at a
    // time.
    const int n_streams = (cfg && cfg->streams_per_device > 0) ? cfg->streams_per_device : 5;

    // host scheduler: balance estimated cells over the devices (greedy longest-processing-time, the same rule as
    // agatha_shard_pairs), most expensive pairs first on every device. One parallel sort; at 1M pairs the two
    // single-threaded sorts this replaced cost a quarter of the whole 8-GPU job.
    std::vector<Worker> workers((size_t)ndev);
    for (int i = 0; i < ndev; i++) workers[(size_t)i].device = devs[(size_t)i];
    {
        struct Key { uint64_t cost, idx; };
        std::vector<Key> keys(n_alns);
        const int64_t W = params->band_width;
        const uint64_t width = 2 * (uint64_t)std::max<int64_t>(W, 0) + 1;
        // (explicit thread counts: launchers such as torchrun export OMP_NUM_THREADS=1, which would make this serial)
        const int sort_threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), (cfg && cfg->staging_threads > 0) ? (unsigned)(cfg->staging_threads * ndev) : 8u));
#pragma omp parallel for schedule(static) num_threads(sort_threads)
        for (int64_t i = 0; i < (int64_t)n_alns; i++) {
            const uint64_t lo = std::min(query_lens[i], target_lens[i]), hi = std::max(query_lens[i], target_lens[i]);
            keys[(size_t)i] = {lo * std::min<uint64_t>(width, hi) + 64, (uint64_t)i};
        }
        __gnu_parallel::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.cost != b.cost ? a.cost > b.cost : a.idx < b.idx; },
                             __gnu_parallel::default_parallel_tag((unsigned)sort_threads));
        std::vector<uint64_t> load((size_t)ndev, 0);
        for (auto& w : workers) w.pairs.reserve(n_alns / (size_t)ndev + 16);
        for (const Key& k : keys) {
            int best = 0;
            for (int s2 = 1; s2 < ndev; s2++) if (load[(size_t)s2] < load[(size_t)best]) best = s2;
            workers[(size_t)best].pairs.push_back(k.idx);
            load[(size_t)best] += k.cost;
        }
    }
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int fill_threads = (cfg && cfg->staging_threads > 0) ? cfg->staging_threads : std::max(1, std::min(8, hw / ndev));
    JobView jv{query_bases, target_bases, query_offsets, target_offsets, query_lens, target_lens, params,
               cfg ? cfg->query_ops : nullptr, cfg ? cfg->target_ops : nullptr, score, query_end, target_end, stop, dstop};
    std::vector<std::thread> threads;
    for (int i = 1; i < ndev; i++) threads.emplace_back(run_worker, std::ref(workers[(size_t)i]), std::cref(jv), batch_alns, n_streams, fill_threads);
    run_worker(workers[0], jv, batch_alns, n_streams, fill_threads);
    for (auto& t : threads) t.join();

    double kmax = 0; uint64_t h2d = 0, d2h = 0; uint32_t nb = 0;
    for (auto& w : workers) {
        if (w.rc != AGATHA_OK) return set_error(w.rc, "device %d: %s", w.device, w.err.c_str());
        kmax = std::max(kmax, w.kernel_ms); h2d += w.h2d; d2h += w.d2h; nb += w.batches;
    }
    if (stats) {
        stats->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats->seconds_kernel_max = kmax * 1e-3;
        stats->h2d_bytes = h2d; stats->d2h_bytes = d2h; stats->n_batches = nb; stats->n_devices = (uint32_t)ndev;
    }
    return AGATHA_OK;
}

// Start positions (gasal_res_t.query_batch_start / target_batch_start, gasal.h:89-90; the reference allocates neither,
// res.cpp:27-28, so there is nothing to be bit-compatible with -- the convention is GASAL2's WITH_START, gasal.h:36-39):
// the start of the best-scoring alignment that ENDS in the reported end cell, found by a second extension that runs
// backwards from that cell over the reversed prefixes query[0..qend], target[0..tend] -- same scoring and band, Z-drop off.
// The reverse pass is an ordinary job: the host packer reverses the truncated sequences on the way into pinned staging.
extern "C" int agatha_align_job_starts(const uint8_t* query_bases, const uint64_t* query_offsets, const uint32_t* query_lens,
                                       const uint8_t* target_bases, const uint64_t* target_offsets, const uint32_t* target_lens,
                                       uint64_t n_alns, const agatha_params_t* params, const agatha_job_config_t* cfg,
                                       int32_t* score, int32_t* query_end, int32_t* target_end, int32_t* stop, int32_t* dstop,
                                       int32_t* query_start, int32_t* target_start, agatha_job_stats_t* stats)
{
    if (!query_start || !target_start) return set_error(AGATHA_EINVAL, "NULL start arrays");
    if (cfg && (cfg->query_ops || cfg->target_ops)) return set_error(AGATHA_EUNSUPPORTED, "start positions with per-pair ops are not supported");
    int rc = agatha_align_job(query_bases, query_offsets, query_lens, target_bases, target_offsets, target_lens, n_alns, params, cfg,
                              score, query_end, target_end, stop, dstop, stats);
    if (rc || n_alns == 0) return rc;
    std::vector<uint32_t> ql(n_alns), tl(n_alns);
    std::vector<uint8_t> rev(n_alns, 1);                                // op bit 0: reverse
    std::vector<int32_t> s2(n_alns), q2(n_alns), t2(n_alns);
    for (uint64_t i = 0; i < n_alns; i++) {
        // empty pairs and pairs without a positive cell report (0, 0, 0): their alignment is empty, the start is the origin
        const bool none = query_lens[i] == 0 || target_lens[i] == 0;
        ql[i] = none ? 0u : (uint32_t)query_end[i] + 1u;
        tl[i] = none ? 0u : (uint32_t)target_end[i] + 1u;
    }
    agatha_params_t p2 = *params;
    p2.z_threshold = -1;
    agatha_job_config_t c2;
    if (cfg) c2 = *cfg; else std::memset(&c2, 0, sizeof(c2));
    c2.query_ops = rev.data(); c2.target_ops = rev.data();
    agatha_job_stats_t st2;
    rc = agatha_align_job(query_bases, query_offsets, ql.data(), target_bases, target_offsets, tl.data(), n_alns, &p2, &c2,
                          s2.data(), q2.data(), t2.data(), nullptr, nullptr, &st2);
    if (rc) return rc;
    for (uint64_t i = 0; i < n_alns; i++) {
        const bool none = ql[i] == 0 || tl[i] == 0 || score[i] <= 0;
        query_start[i] = none ? 0 : query_end[i] - q2[i];
        target_start[i] = none ? 0 : target_end[i] - t2[i];
    }
    if (stats) {
        stats->seconds_total += st2.seconds_total; stats->seconds_kernel_max += st2.seconds_kernel_max;
        stats->h2d_bytes += st2.h2d_bytes; stats->d2h_bytes += st2.d2h_bytes; stats->n_batches += st2.n_batches;
    }
    return AGATHA_OK;
}

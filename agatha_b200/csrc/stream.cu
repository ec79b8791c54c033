// Host-level batch pipeline of the C ABI: one agatha_stream_t = one CUDA stream with pinned staging and device buffers.
// Replaces gasal_init_streams / gasal_aln_async / gasal_is_aln_async_done / gasal_destroy_streams
// (AGAThA/src/ctors.cpp:26-167, gasal_align.cu:27-292). Differences by design:
//   * really asynchronous: no cudaStreamSynchronize + host sort inside submit (the reference's launcher blocks,
//     gasal_align.cu:14-18); bucketing is done on the host BEFORE the upload, lengths are host data;
//   * no per-subwarp global scratch (the reference allocates 3 strips x maxlen per subwarp, ctors.cpp:89): the kernel
//     keeps all DP state in registers;
//   * 64-bit byte counts (the reference's int sizes overflow at n*len >= 2^31, ctors.cpp:34-37).
#include <algorithm>
#include <cstdint>
#include <cstring>

#include "agatha_b200.h"
#include "engine_internal.h"

using namespace agatha;

struct agatha_stream {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start, packed, kernel done, results on host
    // pinned host staging
    uint8_t *h_q = nullptr, *h_t = nullptr;
    uint64_t hcap_q = 0, hcap_t = 0;
    uint32_t* h_meta = nullptr;     // [qoff | toff | qlen | tlen | order] x cap_n
    int32_t* h_res = nullptr;       // [score | qend | tend | stop | dstop] x cap_n
    uint8_t* h_ops = nullptr;       // [query op | target op] x cap_n, zero unless the caller fills them
    uint32_t cap_n = 0;
    // device
    uint8_t *d_q = nullptr, *d_t = nullptr;
    uint32_t *d_qp = nullptr, *d_tp = nullptr;
    uint64_t dcap_q = 0, dcap_t = 0;   // capacity in bases of the packed buffers
    uint64_t acap_q = 0, acap_t = 0;   // capacity of the ASCII mirrors (0 until an ASCII batch is submitted)
    bool ascii = false;                // this stream has been used with ASCII batches: keep the mirrors sized
    uint32_t* d_meta = nullptr;
    int32_t* d_res = nullptr;
    uint8_t* d_ops = nullptr;
    uint32_t dcap_n = 0;
    void* d_ws = nullptr;
    uint32_t cur_n = 0;
    int state = 0;                  // 0 idle, 1 submitted, 2 finished
    float ms[3] = {0, 0, 0};
};

#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_error(e_, what); } while (0)

static uint64_t round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }

static int grow_host_bases(uint8_t** p, uint64_t* cap, uint64_t need, uint64_t keep)
{
    if (need <= *cap) return AGATHA_OK;
    uint64_t ncap = std::max<uint64_t>(round_up(need, 4096), *cap * 2);
    uint8_t* np = nullptr;
    CK(cudaHostAlloc((void**)&np, ncap, cudaHostAllocDefault), "cudaHostAlloc(staging)");
    if (*p) { if (keep) std::memcpy(np, *p, keep); cudaFreeHost(*p); }
    *p = np; *cap = ncap;
    return AGATHA_OK;
}

extern "C" {

agatha_stream_t* agatha_stream_create(int device, uint32_t max_alns, uint64_t max_query_bytes, uint64_t max_target_bytes)
{
    int ndev = agatha_device_count();
    if (ndev == 0) { set_error(AGATHA_ENODEV, "no CUDA device (agatha_b200 has no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= ndev) { set_error(AGATHA_EINVAL, "device %d out of range (0..%d)", device, ndev - 1); return nullptr; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { cuda_error(e, "cudaSetDevice"); return nullptr; }
    auto* s = new agatha_stream();
    s->device = device;
    if ((e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking)) != cudaSuccess) { cuda_error(e, "cudaStreamCreate"); delete s; return nullptr; }
    // blocking waits: a host thread that waits for a batch sleeps instead of spinning -- with one process per GPU and the
    // packing threads of 8 ranks on 32 cores, spinning waiters take cores away from the packers
    for (auto& ev : s->ev) cudaEventCreateWithFlags(&ev, cudaEventBlockingSync);
    if ((e = cudaMalloc(&s->d_ws, AGATHA_WORKSPACE_BYTES)) != cudaSuccess) { cuda_error(e, "cudaMalloc(workspace)"); agatha_stream_destroy(s); return nullptr; }
    if (agatha_stream_reserve(s, std::max<uint32_t>(max_alns, 1), std::max<uint64_t>(max_query_bytes, 8), std::max<uint64_t>(max_target_bytes, 8)) != AGATHA_OK) {
        agatha_stream_destroy(s);
        return nullptr;
    }
    return s;
}

void agatha_stream_destroy(agatha_stream_t* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->st) cudaStreamSynchronize(s->st);
    cudaFreeHost(s->h_q); cudaFreeHost(s->h_t); cudaFreeHost(s->h_meta); cudaFreeHost(s->h_res); cudaFreeHost(s->h_ops);
    cudaFree(s->d_q); cudaFree(s->d_t); cudaFree(s->d_qp); cudaFree(s->d_tp); cudaFree(s->d_meta); cudaFree(s->d_res); cudaFree(s->d_ops); cudaFree(s->d_ws);
    for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

// Device buffers: the packed words always; the unpacked ASCII mirror (d_q / d_t) only for streams that were actually given
// an ASCII batch (agatha_stream_submit[_ops]) -- batches packed on the host never need it.
static int grow_device(agatha_stream_t* s, uint32_t n, uint64_t qbytes, uint64_t tbytes, bool need_ascii)
{
    if (qbytes > s->dcap_q) {
        const uint64_t cap = std::max<uint64_t>(round_up(qbytes, 4096), s->dcap_q * 2);
        cudaFree(s->d_q); cudaFree(s->d_qp); s->d_q = nullptr; s->d_qp = nullptr; s->dcap_q = 0; s->acap_q = 0;
        CK(cudaMalloc((void**)&s->d_qp, cap / 2 + 4 * AGATHA_PACK_SLACK_WORDS), "cudaMalloc(packed query)");
        s->dcap_q = cap;
    }
    if (tbytes > s->dcap_t) {
        const uint64_t cap = std::max<uint64_t>(round_up(tbytes, 4096), s->dcap_t * 2);
        cudaFree(s->d_t); cudaFree(s->d_tp); s->d_t = nullptr; s->d_tp = nullptr; s->dcap_t = 0; s->acap_t = 0;
        CK(cudaMalloc((void**)&s->d_tp, cap / 2 + 4 * AGATHA_PACK_SLACK_WORDS), "cudaMalloc(packed target)");
        s->dcap_t = cap;
    }
    if (need_ascii && s->acap_q < s->dcap_q) {
        cudaFree(s->d_q); s->d_q = nullptr; s->acap_q = 0;
        CK(cudaMalloc((void**)&s->d_q, s->dcap_q), "cudaMalloc(query bases)");
        s->acap_q = s->dcap_q;
    }
    if (need_ascii && s->acap_t < s->dcap_t) {
        cudaFree(s->d_t); s->d_t = nullptr; s->acap_t = 0;
        CK(cudaMalloc((void**)&s->d_t, s->dcap_t), "cudaMalloc(target bases)");
        s->acap_t = s->dcap_t;
    }
    if (n > s->dcap_n) {
        const uint32_t cap = std::max<uint32_t>(n, s->dcap_n * 2);
        cudaFree(s->d_meta); cudaFree(s->d_res); cudaFree(s->d_ops); s->d_meta = nullptr; s->d_res = nullptr; s->d_ops = nullptr; s->dcap_n = 0;
        CK(cudaMalloc((void**)&s->d_meta, sizeof(uint32_t) * 5ull * cap), "cudaMalloc(meta)");
        CK(cudaMalloc((void**)&s->d_res, sizeof(int32_t) * 5ull * cap), "cudaMalloc(results)");
        CK(cudaMalloc((void**)&s->d_ops, 2ull * cap), "cudaMalloc(ops)");
        s->dcap_n = cap;
    }
    return AGATHA_OK;
}

// Grow-on-demand like the reference (host pages x2, host_batch.cpp:107-126; device buffers, gasal_align.cu:71-133).
// Staged bytes/metadata already written are preserved.
int agatha_stream_reserve(agatha_stream_t* s, uint32_t n_alns, uint64_t query_bytes, uint64_t target_bytes)
{
    if (!s) return set_error(AGATHA_EINVAL, "stream is NULL");
    if (s->state == 1) return set_error(AGATHA_EINVAL, "reserve while a batch is in flight");
    CK(cudaSetDevice(s->device), "cudaSetDevice");
    int rc;
    if ((rc = grow_host_bases(&s->h_q, &s->hcap_q, round_up(query_bytes, 8), s->hcap_q))) return rc;
    if ((rc = grow_host_bases(&s->h_t, &s->hcap_t, round_up(target_bytes, 8), s->hcap_t))) return rc;
    if (n_alns > s->cap_n) {
        const uint32_t ncap = std::max<uint32_t>(n_alns, s->cap_n * 2);
        uint32_t* nm = nullptr; int32_t* nr = nullptr; uint8_t* no = nullptr;
        CK(cudaHostAlloc((void**)&nm, sizeof(uint32_t) * 5ull * ncap, cudaHostAllocDefault), "cudaHostAlloc(meta)");
        CK(cudaHostAlloc((void**)&nr, sizeof(int32_t) * 5ull * ncap, cudaHostAllocDefault), "cudaHostAlloc(results)");
        CK(cudaHostAlloc((void**)&no, 2ull * ncap, cudaHostAllocDefault), "cudaHostAlloc(ops)");
        std::memset(no, 0, 2ull * ncap);
        if (s->h_ops) {
            for (int k = 0; k < 2; k++) std::memcpy(no + (size_t)k * ncap, s->h_ops + (size_t)k * s->cap_n, s->cap_n);
            cudaFreeHost(s->h_ops);
        }
        if (s->h_meta) {
            for (int k = 0; k < 5; k++) std::memcpy(nm + (size_t)k * ncap, s->h_meta + (size_t)k * s->cap_n, sizeof(uint32_t) * s->cap_n);
            cudaFreeHost(s->h_meta);
        }
        if (s->h_res) {
            for (int k = 0; k < 5; k++) std::memcpy(nr + (size_t)k * ncap, s->h_res + (size_t)k * s->cap_n, sizeof(int32_t) * s->cap_n);
            cudaFreeHost(s->h_res);
        }
        s->h_meta = nm; s->h_res = nr; s->h_ops = no; s->cap_n = ncap;
    }
    // device side too, so that a submit of anything that fits the reservation never allocates (cudaMalloc/cudaFree
    // synchronise the whole device)
    return grow_device(s, n_alns, round_up(query_bytes, 8), round_up(target_bytes, 8), s->ascii);
}

void agatha_stream_capacity(agatha_stream_t* s, uint32_t* max_alns, uint64_t* query_bytes, uint64_t* target_bytes)
{
    if (max_alns) *max_alns = s->cap_n;
    if (query_bytes) *query_bytes = s->hcap_q;
    if (target_bytes) *target_bytes = s->hcap_t;
}
uint8_t* agatha_stream_query_bases(agatha_stream_t* s) { return s->h_q; }
uint8_t* agatha_stream_target_bases(agatha_stream_t* s) { return s->h_t; }
uint32_t* agatha_stream_query_packed(agatha_stream_t* s) { return (uint32_t*)s->h_q; }
uint32_t* agatha_stream_target_packed(agatha_stream_t* s) { return (uint32_t*)s->h_t; }
uint32_t* agatha_stream_query_offsets(agatha_stream_t* s) { return s->h_meta; }
uint32_t* agatha_stream_target_offsets(agatha_stream_t* s) { return s->h_meta + (size_t)s->cap_n; }
uint32_t* agatha_stream_query_lens(agatha_stream_t* s) { return s->h_meta + 2 * (size_t)s->cap_n; }
uint32_t* agatha_stream_target_lens(agatha_stream_t* s) { return s->h_meta + 3 * (size_t)s->cap_n; }
uint8_t* agatha_stream_query_ops(agatha_stream_t* s) { return s->h_ops; }
uint8_t* agatha_stream_target_ops(agatha_stream_t* s) { return s->h_ops + (size_t)s->cap_n; }

}  // extern "C"

static int submit(agatha_stream_t* s, uint64_t query_bytes, uint64_t target_bytes, uint32_t n_alns, const agatha_params_t* params, bool with_ops, bool packed)
{
    if (!s) return set_error(AGATHA_EINVAL, "stream is NULL");
    // the reference's argument checks, gasal_align.cu:33-68
    if (n_alns == 0) return set_error(AGATHA_EINVAL, "actual_n_alns <= 0");
    if (query_bytes == 0) return set_error(AGATHA_EINVAL, "actual_query_batch_bytes <= 0");
    if (target_bytes == 0) return set_error(AGATHA_EINVAL, "actual_target_batch_bytes <= 0");
    if (query_bytes % 8) return set_error(AGATHA_EINVAL, "actual_query_batch_bytes=%llu is not a multiple of 8", (unsigned long long)query_bytes);
    if (target_bytes % 8) return set_error(AGATHA_EINVAL, "actual_target_batch_bytes=%llu is not a multiple of 8", (unsigned long long)target_bytes);
    if (query_bytes > s->hcap_q) return set_error(AGATHA_EINVAL, "actual_query_batch_bytes(%llu) > host_max_query_batch_bytes(%llu)", (unsigned long long)query_bytes, (unsigned long long)s->hcap_q);
    if (target_bytes > s->hcap_t) return set_error(AGATHA_EINVAL, "actual_target_batch_bytes(%llu) > host_max_target_batch_bytes(%llu)", (unsigned long long)target_bytes, (unsigned long long)s->hcap_t);
    if (n_alns > s->cap_n) return set_error(AGATHA_EINVAL, "actual_n_alns(%u) > host_max_n_alns(%u)", n_alns, s->cap_n);
    if (s->state == 1) return set_error(AGATHA_EINVAL, "stream busy: poll or wait first");
    if (!params) return set_error(AGATHA_EINVAL, "params is NULL");
    CK(cudaSetDevice(s->device), "cudaSetDevice");
    if (!packed) s->ascii = true;
    int rc = grow_device(s, n_alns, query_bytes, target_bytes, !packed);
    if (rc) return rc;

    uint32_t* h_qoff = s->h_meta;
    uint32_t* h_toff = s->h_meta + (size_t)s->cap_n;
    uint32_t* h_qlen = s->h_meta + 2 * (size_t)s->cap_n;
    uint32_t* h_tlen = s->h_meta + 3 * (size_t)s->cap_n;
    uint32_t* h_order = s->h_meta + 4 * (size_t)s->cap_n;
    for (uint32_t i = 0; i < n_alns; i++) {
        if ((h_qoff[i] & 7u) || (h_toff[i] & 7u)) return set_error(AGATHA_EINVAL, "sequence offsets must be multiples of 8 (pair %u)", i);
        if ((uint64_t)h_qoff[i] + h_qlen[i] > query_bytes || (uint64_t)h_toff[i] + h_tlen[i] > target_bytes)
            return set_error(AGATHA_EINVAL, "pair %u runs past the end of the staged batch", i);
    }
    // length-aware bucketing on the host, before the upload (replaces agatha_sort + D2H + std::sort + H2D)
    if ((rc = agatha_bucket_order(h_qlen, h_tlen, n_alns, params->band_width, h_order))) return rc;

    cudaStream_t st = s->st;
    CK(cudaEventRecord(s->ev[0], st), "cudaEventRecord");
    for (int k = 0; k < 5; k++)
        CK(cudaMemcpyAsync(s->d_meta + (size_t)k * s->dcap_n, s->h_meta + (size_t)k * s->cap_n, sizeof(uint32_t) * n_alns, cudaMemcpyHostToDevice, st), "H2D batch metadata");
    const size_t dn = s->dcap_n;
    if (packed) {                                     // packed on the host (agatha_pack_batch): half the bytes, no pack kernel
        CK(cudaMemcpyAsync(s->d_qp, s->h_q, query_bytes / 2, cudaMemcpyHostToDevice, st), "H2D packed query");
        CK(cudaMemcpyAsync(s->d_tp, s->h_t, target_bytes / 2, cudaMemcpyHostToDevice, st), "H2D packed target");
    } else {
        CK(cudaMemcpyAsync(s->d_q, s->h_q, query_bytes, cudaMemcpyHostToDevice, st), "H2D query bases");
        CK(cudaMemcpyAsync(s->d_t, s->h_t, target_bytes, cudaMemcpyHostToDevice, st), "H2D target bases");
        if ((rc = agatha_pack_device(s->d_q, query_bytes, s->d_t, target_bytes, s->d_qp, s->d_tp, st))) return rc;
    }
    if (with_ops) {                                   // gasal_align.cu:199-212
        for (int k = 0; k < 2; k++)
            CK(cudaMemcpyAsync(s->d_ops + (size_t)k * dn, s->h_ops + (size_t)k * s->cap_n, n_alns, cudaMemcpyHostToDevice, st), "H2D ops");
        if ((rc = agatha_apply_ops_device(s->d_q, s->d_t, s->d_meta, s->d_meta + dn, s->d_meta + 2 * dn, s->d_meta + 3 * dn,
                                          s->d_ops, s->d_ops + dn, n_alns, s->d_qp, s->d_tp, st))) return rc;
    }
    CK(cudaEventRecord(s->ev[1], st), "cudaEventRecord");
    rc = agatha_extend_device(s->d_qp, s->d_tp, s->d_meta, s->d_meta + dn, s->d_meta + 2 * dn, s->d_meta + 3 * dn, s->d_meta + 4 * dn,
                              n_alns, params, s->d_res, s->d_res + dn, s->d_res + 2 * dn, s->d_res + 3 * dn, s->d_res + 4 * dn, s->d_ws, st);
    if (rc) return rc;
    CK(cudaEventRecord(s->ev[2], st), "cudaEventRecord");
    for (int k = 0; k < 5; k++)
        CK(cudaMemcpyAsync(s->h_res + (size_t)k * s->cap_n, s->d_res + (size_t)k * dn, sizeof(int32_t) * n_alns, cudaMemcpyDeviceToHost, st), "D2H results");
    CK(cudaEventRecord(s->ev[3], st), "cudaEventRecord");
    s->cur_n = n_alns;
    s->state = 1;
    return AGATHA_OK;
}

extern "C" {

int agatha_stream_submit(agatha_stream_t* s, uint64_t query_bytes, uint64_t target_bytes, uint32_t n_alns, const agatha_params_t* params)
{
    return submit(s, query_bytes, target_bytes, n_alns, params, false, false);
}

int agatha_stream_submit_packed(agatha_stream_t* s, uint64_t query_bases, uint64_t target_bases, uint32_t n_alns, const agatha_params_t* params)
{
    return submit(s, query_bases, target_bases, n_alns, params, false, true);
}

int agatha_stream_submit_ops(agatha_stream_t* s, uint64_t query_bytes, uint64_t target_bytes, uint32_t n_alns, const agatha_params_t* params)
{
    return submit(s, query_bytes, target_bytes, n_alns, params, true, false);
}

static void finish(agatha_stream_t* s)
{
    cudaEventElapsedTime(&s->ms[0], s->ev[0], s->ev[1]);
    cudaEventElapsedTime(&s->ms[1], s->ev[1], s->ev[2]);
    cudaEventElapsedTime(&s->ms[2], s->ev[0], s->ev[3]);
    s->state = 2;
}

int agatha_stream_poll(agatha_stream_t* s)
{
    if (!s) { set_error(AGATHA_EINVAL, "stream is NULL"); return -6; }   // errors are < -2: -1/-2 are taken by the protocol
    if (s->state != 1) return -2;                     // nothing launched, gasal_align.cu:279
    cudaError_t e = cudaEventQuery(s->ev[3]);
    if (e == cudaErrorNotReady) return -1;            // gasal_align.cu:282
    if (e != cudaSuccess) { cuda_error(e, "cudaEventQuery"); return AGATHA_ECUDA; }
    finish(s);
    return 0;
}

int agatha_stream_wait(agatha_stream_t* s)
{
    if (!s) return set_error(AGATHA_EINVAL, "stream is NULL");
    if (s->state != 1) return AGATHA_OK;
    CK(cudaEventSynchronize(s->ev[3]), "cudaEventSynchronize");
    finish(s);
    return AGATHA_OK;
}

int agatha_stream_timings(agatha_stream_t* s, float ms[3])
{
    if (!s || !ms) return set_error(AGATHA_EINVAL, "NULL argument");
    ms[0] = s->ms[0]; ms[1] = s->ms[1]; ms[2] = s->ms[2];
    return AGATHA_OK;
}

const int32_t* agatha_stream_scores(agatha_stream_t* s) { return s->h_res; }
const int32_t* agatha_stream_query_ends(agatha_stream_t* s) { return s->h_res + (size_t)s->cap_n; }
const int32_t* agatha_stream_target_ends(agatha_stream_t* s) { return s->h_res + 2 * (size_t)s->cap_n; }
const int32_t* agatha_stream_stops(agatha_stream_t* s) { return s->h_res + 3 * (size_t)s->cap_n; }
const int32_t* agatha_stream_dstops(agatha_stream_t* s) { return s->h_res + 4 * (size_t)s->cap_n; }

}  // extern "C"

// Implementation of include/gasal_compat/gasal_compat.h: the reference's GASAL2-style host API on top of the C ABI.
// Written from the interface description in SURVEY.md section 8(b) and the declarations in the reference headers;
// behaviour notes cite the reference implementation they reproduce.
#include <algorithm>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "gasal_compat/gasal_compat.h"

namespace {

// scoring is process-global in the reference (__constant__ symbols, gasal_kernels.h:29-36); same here
agatha_params_t g_scores = {1, 4, 6, 2, 3, 400, 751};
bool g_scores_set = false;

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "[GASAL ERROR:] %s%s%s\n", what, *agatha_last_error() ? ": " : "", agatha_last_error());
    exit(EXIT_FAILURE);
}

agatha_stream_t* handle(gasal_gpu_storage_t* g) { return reinterpret_cast<agatha_stream_t*>(g->global_buffer); }

// (re)derive every pointer the caller may touch from the engine's pinned buffers
void refresh_views(gasal_gpu_storage_t* g)
{
    agatha_stream_t* s = handle(g);
    g->host_query_batch_offsets = agatha_stream_query_offsets(s);
    g->host_target_batch_offsets = agatha_stream_target_offsets(s);
    g->host_query_batch_lens = agatha_stream_query_lens(s);
    g->host_target_batch_lens = agatha_stream_target_lens(s);
    g->extensible_host_unpacked_query_batch->data = agatha_stream_query_bases(s);
    g->extensible_host_unpacked_target_batch->data = agatha_stream_target_bases(s);
    g->host_res->aln_score = const_cast<int32_t*>(agatha_stream_scores(s));
    g->host_res->query_batch_end = const_cast<int32_t*>(agatha_stream_query_ends(s));
    g->host_res->target_batch_end = const_cast<int32_t*>(agatha_stream_target_ends(s));
}

host_batch_t* new_page_view(uint32_t bytes)
{
    host_batch_t* p = (host_batch_t*)calloc(1, sizeof(host_batch_t));
    p->page_size = bytes;
    return p;
}

}  // namespace

// ---------------------------------------------------------------- ctors.h
gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams)
{
    gasal_gpu_storage_v v;
    v.a = (gasal_gpu_storage_t*)calloc((size_t)n_streams, sizeof(gasal_gpu_storage_t));
    v.n = n_streams;
    return v;
}

// Same sizing rule as the reference (ctors.cpp:29-39): room for kernel_align_num sequences of the longest length.
void gasal_init_streams(gasal_gpu_storage_v* vec, int max_query_len, int max_target_len, int32_t maximum_sequence_length, Parameters* params)
{
    const uint64_t n = (uint64_t)params->kernel_align_num;
    const uint64_t q8 = ((uint64_t)max_query_len + 7) & ~7ull, t8 = ((uint64_t)max_target_len + 7) & ~7ull;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < vec->n; i++) {
        gasal_gpu_storage_t* g = &vec->a[i];
        // start smaller than the reference's worst case (it never grows back); staging grows on demand
        const uint64_t qcap = std::min<uint64_t>(n * q8, 0xfffffff8ull), tcap = std::min<uint64_t>(n * t8, 0xfffffff8ull);
        agatha_stream_t* s = agatha_stream_create(dev, (uint32_t)n, std::min<uint64_t>(qcap, 64ull << 20), std::min<uint64_t>(tcap, 64ull << 20));
        if (!s) die("gasal_init_streams");
        g->global_buffer = reinterpret_cast<short2*>(s);
        g->extensible_host_unpacked_query_batch = new_page_view((uint32_t)qcap);
        g->extensible_host_unpacked_target_batch = new_page_view((uint32_t)tcap);
        g->host_query_op = (uint8_t*)calloc(n, 1);
        g->host_target_op = (uint8_t*)calloc(n, 1);
        g->host_res = (gasal_res_t*)calloc(1, sizeof(gasal_res_t));
        g->host_max_query_batch_bytes = g->gpu_max_query_batch_bytes = (uint32_t)qcap;
        g->host_max_target_batch_bytes = g->gpu_max_target_batch_bytes = (uint32_t)tcap;
        g->host_max_n_alns = g->gpu_max_n_alns = (uint32_t)n;
        g->current_n_alns = 0;
        g->is_free = 1;
        g->slice_width = params->slice_width;
        g->maximum_sequence_length = (uint32_t)maximum_sequence_length;
        refresh_views(g);
    }
}

void gasal_destroy_streams(gasal_gpu_storage_v* vec, Parameters*)
{
    for (int i = 0; i < vec->n; i++) {
        gasal_gpu_storage_t* g = &vec->a[i];
        agatha_stream_destroy(handle(g));
        g->global_buffer = nullptr;
        free(g->extensible_host_unpacked_query_batch); free(g->extensible_host_unpacked_target_batch);
        free(g->host_query_op); free(g->host_target_op); free(g->host_res);
    }
}

void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v* vec)
{
    if (vec->a) free(vec->a);
    vec->a = nullptr;
}

// ---------------------------------------------------------------- host_batch.h
// The reference keeps a linked list of pinned pages that doubles when full (host_batch.cpp:107-126); the engine keeps
// one pinned buffer per side that grows in place, so idx is simply the write position. Padding to a multiple of 8
// with 'N' and the returned next index are the reference's (host_batch.cpp:95-153).
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t* g, uint32_t idx, const char* data, uint32_t size, data_source SRC)
{
    if (SRC != QUERY && SRC != TARGET) die("gasal_host_batch_fill: SRC must be QUERY or TARGET");
    const uint32_t padded = (size + 7u) & ~7u;
    host_batch_t* page = SRC == QUERY ? g->extensible_host_unpacked_query_batch : g->extensible_host_unpacked_target_batch;
    const uint64_t need = (uint64_t)idx + padded;
    if (need > 0xfffffff8ull) die("gasal_host_batch_fill: batch exceeds 32-bit offsets");
    agatha_stream_t* s = handle(g);
    uint64_t qcap = 0, tcap = 0;
    agatha_stream_capacity(s, nullptr, &qcap, &tcap);
    if (need > (SRC == QUERY ? qcap : tcap)) {
        if (agatha_stream_reserve(s, g->host_max_n_alns, SRC == QUERY ? need : qcap, SRC == QUERY ? tcap : need)) die("gasal_host_batch_fill");
        refresh_views(g);
    }
    uint8_t* dst = page->data + idx;
    memcpy(dst, data, size);
    memset(dst + size, 'N', padded - size);
    if (need > page->data_size) page->data_size = (uint32_t)need;
    if (need > page->page_size) {
        page->page_size = (uint32_t)need;
        if (SRC == QUERY) g->host_max_query_batch_bytes = (uint32_t)need; else g->host_max_target_batch_bytes = (uint32_t)need;
    }
    return idx + padded;
}

uint32_t gasal_host_batch_add(gasal_gpu_storage_t* g, uint32_t idx, const char* data, uint32_t size, data_source SRC)
{
    // like the reference's _add: raw append, no padding (host_batch.cpp:163-222)
    if (SRC != QUERY && SRC != TARGET) die("gasal_host_batch_add: SRC must be QUERY or TARGET");
    host_batch_t* page = SRC == QUERY ? g->extensible_host_unpacked_query_batch : g->extensible_host_unpacked_target_batch;
    const uint64_t need = (uint64_t)idx + size;
    uint64_t qcap = 0, tcap = 0;
    agatha_stream_capacity(handle(g), nullptr, &qcap, &tcap);
    if (need > (SRC == QUERY ? qcap : tcap)) {
        if (agatha_stream_reserve(handle(g), g->host_max_n_alns, SRC == QUERY ? need : qcap, SRC == QUERY ? tcap : need)) die("gasal_host_batch_add");
        refresh_views(g);
    }
    memcpy(page->data + idx, data, size);
    if (need > page->data_size) page->data_size = (uint32_t)need;
    return idx + size;
}

uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t* g, uint32_t idx, const char base, data_source SRC)
{
    return gasal_host_batch_add(g, idx, &base, 1, SRC);
}

void gasal_host_batch_reset(gasal_gpu_storage_t* g)
{
    g->extensible_host_unpacked_query_batch->data_size = 0;
    g->extensible_host_unpacked_target_batch->data_size = 0;
}

host_batch_t* gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset)
{
    host_batch_t* p = new_page_view(batch_bytes);
    cudaError_t err;
    CHECKCUDAERROR(cudaHostAlloc((void**)&p->data, batch_bytes, cudaHostAllocDefault));
    p->offset = offset;
    return p;
}

void gasal_host_batch_destroy(host_batch_t* res)
{
    if (!res) { fprintf(stderr, "[GASAL ERROR] Trying to free a NULL pointer\n"); exit(1); }
    if (res->next) gasal_host_batch_destroy(res->next);
    if (res->data) cudaFreeHost(res->data);
    free(res);
}

host_batch_t* gasal_host_batch_getlast(host_batch_t* arg) { return arg->next == NULL ? arg : gasal_host_batch_getlast(arg->next); }

void gasal_host_batch_print(host_batch_t* res)
{
    fprintf(stderr, "[GASAL PRINT] Page data: offset=%d, next_offset=%d, data size=%d, page size=%d\n", res->offset,
            (res->next != NULL ? (int)res->next->offset : -1), res->data_size, res->page_size);
}

void gasal_host_batch_printall(host_batch_t* res)
{
    gasal_host_batch_print(res);
    if (res->next) { fprintf(stderr, "+--->"); gasal_host_batch_printall(res->next); }
}

// ---------------------------------------------------------------- interfaces.h
void gasal_host_alns_resize(gasal_gpu_storage_t* g, int new_max_alns, Parameters*)
{
    fprintf(stderr, "[GASAL WARNING] Resizing gpu_storage from %d sequences to %d sequences... ", g->host_max_n_alns, new_max_alns);
    if (new_max_alns < (int)g->host_max_n_alns) { fprintf(stderr, "[GASAL ERROR] cudoHostRealloc: invalid sizes. New size < old size (%d < %d)", new_max_alns, g->host_max_n_alns); exit(EXIT_FAILURE); }
    if (agatha_stream_reserve(handle(g), (uint32_t)new_max_alns, 8, 8)) die("gasal_host_alns_resize");
    g->host_query_op = (uint8_t*)realloc(g->host_query_op, (size_t)new_max_alns);
    g->host_target_op = (uint8_t*)realloc(g->host_target_op, (size_t)new_max_alns);
    memset(g->host_query_op + g->host_max_n_alns, 0, (size_t)new_max_alns - g->host_max_n_alns);
    memset(g->host_target_op + g->host_max_n_alns, 0, (size_t)new_max_alns - g->host_max_n_alns);
    g->host_max_n_alns = (uint32_t)new_max_alns;
    refresh_views(g);
    fprintf(stderr, " done. This can harm performance.\n");
}

void gasal_op_fill(gasal_gpu_storage_t* g, uint8_t* data, uint32_t nbr_seqs_in_stream, data_source SRC)
{
    // only applied when the caller sets params->isReverseComplement (the reference's driver never does, args_parser.cpp:28)
    uint8_t* dst = SRC == QUERY ? g->host_query_op : (SRC == TARGET ? g->host_target_op : nullptr);
    if (dst) memcpy(dst, data, nbr_seqs_in_stream);
}

void gasal_set_device(int gpu_select, bool isPrintingProp)
{
    int n = agatha_device_count();
    if (isPrintingProp) {
        fprintf(stderr, "Found %d GPUs\n", n);
        if (gpu_select > n - 1) {
            fprintf(stderr, "Error: can't select device %d when only %d devices are selected (range from 0 to %d)\n", gpu_select, n, n - 1);
            exit(EXIT_FAILURE);
        }
        for (int d = 0; d < n; d++) { cudaDeviceProp p; cudaGetDeviceProperties(&p, d); fprintf(stderr, "\tGPU %d: %s\n", d, p.name); }
        if (n > 0) { cudaDeviceProp p; cudaGetDeviceProperties(&p, gpu_select); fprintf(stderr, "Selected device %d : %s\n", gpu_select, p.name); }
    }
    if (n > 0) cudaSetDevice(gpu_select);
}

// ---------------------------------------------------------------- gasal_align.h
void gasal_copy_subst_scores(gasal_subst_scores* subst)
{
    g_scores.match = subst->match; g_scores.mismatch = subst->mismatch;
    g_scores.gap_open = subst->gap_open; g_scores.gap_extend = subst->gap_extend;
    g_scores.slice_width = subst->slice_width; g_scores.z_threshold = subst->z_threshold; g_scores.band_width = subst->band_width;
    g_scores_set = true;
}

void gasal_aln_async(gasal_gpu_storage_t* g, const uint32_t actual_query_batch_bytes, const uint32_t actual_target_batch_bytes, const uint32_t actual_n_alns, Parameters* params)
{
    if (!g_scores_set && params) {   // a caller that never called gasal_copy_subst_scores gets the driver's options
        g_scores.match = params->sa; g_scores.mismatch = params->sb; g_scores.gap_open = params->gapo; g_scores.gap_extend = params->gape;
        g_scores.slice_width = params->slice_width; g_scores.z_threshold = params->z_threshold; g_scores.band_width = params->band_width;
    }
    agatha_stream_t* s = handle(g);
    const bool with_ops = params && params->isReverseComplement;         // gasal_align.cu:199
    if (with_ops && actual_n_alns <= g->host_max_n_alns) {
        memcpy(agatha_stream_query_ops(s), g->host_query_op, actual_n_alns);
        memcpy(agatha_stream_target_ops(s), g->host_target_op, actual_n_alns);
    }
    if (with_ops ? agatha_stream_submit_ops(s, actual_query_batch_bytes, actual_target_batch_bytes, actual_n_alns, &g_scores)
                 : agatha_stream_submit(s, actual_query_batch_bytes, actual_target_batch_bytes, actual_n_alns, &g_scores)) {
        fprintf(stderr, "[GASAL ERROR:] %s\n", agatha_last_error());      // same checks/messages as gasal_align.cu:33-68
        exit(EXIT_FAILURE);
    }
    if (params && params->print_out) {
        // -p makes every batch synchronous and logs its kernel time in ms to raw_file (gasal_align.cu:219-236)
        if (agatha_stream_wait(s)) die("gasal_aln_async");
        float ms[3];
        agatha_stream_timings(s, ms);
        // one whole line per batch even when several driver threads (-n) share the stream object (the reference writes from
        // its threads without a lock, gasal_align.cu:233, and tears lines now and then)
        static std::mutex raw_mu;
        std::lock_guard<std::mutex> lk(raw_mu);
        char line[64];
        const int len = snprintf(line, sizeof(line), "%g\n", (double)ms[1]);
        params->raw_file.write(line, len);
        params->raw_file.flush();
    }
    g->is_free = 0;
}

int gasal_is_aln_async_done(gasal_gpu_storage_t* g)
{
    if (g->is_free == 1) return -2;
    int rc = agatha_stream_poll(handle(g));
    if (rc == -1) return -1;
    if (rc == -2) rc = 0;                            // already finished through the synchronous -p path
    if (rc < 0) die("gasal_is_aln_async_done");
    refresh_views(g);
    gasal_host_batch_reset(g);
    g->is_free = 1;
    g->current_n_alns = 0;
    return 0;
}

// ---------------------------------------------------------------- res.h (host side only; device triples are internal to the engine)
gasal_res_t* gasal_res_new_host(uint32_t max_n_alns, Parameters*)
{
    gasal_res_t* r = (gasal_res_t*)calloc(1, sizeof(gasal_res_t));
    cudaError_t err;
    CHECKCUDAERROR(cudaHostAlloc((void**)&r->aln_score, max_n_alns * sizeof(int32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc((void**)&r->query_batch_end, max_n_alns * sizeof(int32_t), cudaHostAllocDefault));
    CHECKCUDAERROR(cudaHostAlloc((void**)&r->target_batch_end, max_n_alns * sizeof(int32_t), cudaHostAllocDefault));
    return r;
}

// Device-side result triples (res.cpp:30-82, :97-115). The engine keeps its own result arrays inside the stream object, so
// nothing in this library needs these; they are exported, with the reference's behaviour, for callers that manage device
// results themselves (start fields NULL as in res.cpp:76-77).
gasal_res_t* gasal_res_new_device_cpy(uint32_t max_n_alns, Parameters*)
{
    gasal_res_t* r = (gasal_res_t*)calloc(1, sizeof(gasal_res_t));
    cudaError_t err;
    CHECKCUDAERROR(cudaMalloc((void**)&r->aln_score, max_n_alns * sizeof(int32_t)));
    CHECKCUDAERROR(cudaMalloc((void**)&r->query_batch_end, max_n_alns * sizeof(int32_t)));
    CHECKCUDAERROR(cudaMalloc((void**)&r->target_batch_end, max_n_alns * sizeof(int32_t)));
    return r;
}

gasal_res_t* gasal_res_new_device(gasal_res_t* device_cpy)
{
    gasal_res_t* d_c = nullptr;
    cudaError_t err;
    CHECKCUDAERROR(cudaMalloc((void**)&d_c, sizeof(gasal_res_t)));
    CHECKCUDAERROR(cudaMemcpy(d_c, device_cpy, sizeof(gasal_res_t), cudaMemcpyHostToDevice));   // the pointers ARE device pointers
    return d_c;
}

void gasal_res_destroy_device(gasal_res_t* device_res, gasal_res_t* device_cpy)
{
    if (device_cpy) {
        if (device_cpy->aln_score) cudaFree(device_cpy->aln_score);
        if (device_cpy->query_batch_start) cudaFree(device_cpy->query_batch_start);
        if (device_cpy->target_batch_start) cudaFree(device_cpy->target_batch_start);
        if (device_cpy->query_batch_end) cudaFree(device_cpy->query_batch_end);
        if (device_cpy->target_batch_end) cudaFree(device_cpy->target_batch_end);
        free(device_cpy);
    }
    if (device_res) cudaFree(device_res);
}

// ctors.h:9-11. The reference (re)allocates the unpacked and packed device batches of a storage here (ctors.cpp:178-226);
// in this library they belong to the stream object, which grows on demand: alloc = reserve, free = nothing to do until
// gasal_destroy_streams.
void gasal_gpu_mem_alloc(gasal_gpu_storage_t* g, int gpu_max_query_batch_bytes, int gpu_max_target_batch_bytes, Parameters*)
{
    if (gpu_max_query_batch_bytes % 8 || gpu_max_target_batch_bytes % 8) { fprintf(stderr, "[GASAL ERROR:] batch bytes must be multiples of 8\n"); exit(EXIT_FAILURE); }
    if (agatha_stream_reserve(handle(g), g->host_max_n_alns, (uint64_t)gpu_max_query_batch_bytes, (uint64_t)gpu_max_target_batch_bytes)) die("gasal_gpu_mem_alloc");
    g->gpu_max_query_batch_bytes = (uint32_t)gpu_max_query_batch_bytes;
    g->gpu_max_target_batch_bytes = (uint32_t)gpu_max_target_batch_bytes;
    refresh_views(g);
}

void gasal_gpu_mem_free(gasal_gpu_storage_t*, Parameters*) {}

void gasal_res_destroy_host(gasal_res_t* r)
{
    if (!r) return;
    if (r->aln_score) cudaFreeHost(r->aln_score);
    if (r->query_batch_end) cudaFreeHost(r->query_batch_end);
    if (r->target_batch_end) cudaFreeHost(r->target_batch_end);
    free(r);
}

// ---------------------------------------------------------------- args_parser.h
Parameters::Parameters(int argc_, char** argv_)
    : sa(2), sb(4), gapo(4), gape(2), print_out(0), n_threads(1), slice_width(3), z_threshold(400), band_width(751),
      kernel_block_num(256), kernel_thread_num(256), kernel_align_num(8192), isPacked(false), isReverseComplement(false),
      argc(argc_), argv(argv_)   // defaults: args_parser.cpp:12-28
{
}

Parameters::~Parameters()
{
    query_batch_fasta.close();
    target_batch_fasta.close();
    raw_file.close();
}

void Parameters::print()
{
    std::cerr << "sa=" << sa << " , sb=" << sb << " , gapo=" << gapo << " , gape=" << gape << std::endl;
    std::cerr << "slice_width=" << slice_width << ", z_threshold=" << z_threshold << ", band_width=" << band_width << std::endl;
    std::cerr << "kernel launch: block_num=" << kernel_block_num << ", thread_num=" << kernel_thread_num << ", align_num=" << kernel_align_num << std::endl;
    std::cerr << "print_out=" << print_out << " , n_threads=" << n_threads << std::endl;
    std::cerr << "query_batch_fasta_filename=" << query_batch_fasta_filename << " , target_batch_fasta_filename=" << target_batch_fasta_filename << std::endl;
}

void Parameters::failure(fail_type f)
{
    if (f == NOT_ENOUGH_ARGS) std::cerr << "Not enough Parameters. Required: file1.fasta file2.fasta. See help (--help, -h) for usage. " << std::endl;
    else if (f == WRONG_ARG) std::cerr << "Wrong argument. See help (--help, -h) for usage. " << std::endl;
    else if (f == WRONG_FILES) std::cerr << "File error: either a file doesn't exist, or cannot be opened." << std::endl;
    exit(1);
}

void Parameters::help()
{
    std::cerr << "Usage: manual [-m] [-x] [-q] [-r] [-s] [-z] [-w] [-b] [-t] [-a] [-p] [-n] <query_batch.fasta> <target_batch.fasta> [raw_file]\n"
              << "Options: -m INT    match score [" << sa << "]\n"
              << "         -x INT    mismatch penalty [" << sb << "]\n"
              << "         -q INT    gap open penalty [" << gapo << "]\n"
              << "         -r INT    gap extension penalty [" << gape << "]\n"
              << "         -s INT    slice width [" << slice_width << "]\n"
              << "         -z INT    z-drop threshold [" << z_threshold << "]\n"
              << "         -w INT    band width [" << band_width << "]\n"
              << "         -b/-t INT accepted, ignored (the engine sizes its own persistent grid)\n"
              << "         -a INT    alignments per batch [" << kernel_align_num << "]\n"
              << "         -p        print the alignment results; kernel ms per batch go to raw_file\n"
              << "         -n INT    number of CPU threads [" << n_threads << "]\n"
              << "         --help, -h : displays this message." << std::endl;
}

// Contract of args_parser.cpp:93-229: needs argc >= 4; the last two arguments (three with -p) are positional.
void Parameters::parse()
{
    for (int c = 1; c < argc; c++) {
        const std::string a(argv[c]);
        if (a == "--help" || a == "-h") { help(); exit(0); }
    }
    if (argc < 4) failure(NOT_ENOUGH_ARGS);
    int c = 1;
    for (; c < argc - 3; c++) {
        const std::string a(argv[c]);
        if (a.size() < 2 || a[0] != '-') failure(WRONG_ARG);
        if (a[1] == '-') continue;
        if (a.size() > 2) failure(WRONG_ARG);
        int* dst = nullptr;
        switch (a[1]) {
            case 'm': dst = &sa; break;
            case 'x': dst = &sb; break;
            case 'q': dst = &gapo; break;
            case 'r': dst = &gape; break;
            case 'n': dst = &n_threads; break;
            case 's': dst = &slice_width; break;
            case 'z': dst = &z_threshold; break;
            case 'w': dst = &band_width; break;
            case 'b': dst = &kernel_block_num; break;
            case 't': dst = &kernel_thread_num; break;
            case 'a': dst = &kernel_align_num; break;
            case 'p': print_out = 1; break;
            default: break;
        }
        if (dst) { c++; *dst = std::stoi(std::string(argv[c])); }
    }
    query_batch_fasta_filename = argv[c++];
    target_batch_fasta_filename = argv[c];
    if (print_out) { c++; raw_filename = argv[c]; }
    fileopen();
}

void Parameters::fileopen()
{
    query_batch_fasta.open(query_batch_fasta_filename, std::ifstream::in);
    if (!query_batch_fasta) failure(WRONG_FILES);
    target_batch_fasta.open(target_batch_fasta_filename);
    if (!target_batch_fasta) failure(WRONG_FILES);
    if (print_out) raw_file.open(raw_filename, std::ios::app);
}

// FASTA reader for the reference driver's input format. Replaces the lock-step getline loop of
// AGAThA/test_prog/test_prog.cpp:94-149 with the same semantics, but mmap'd and multi-threaded (SURVEY.md 8f item 1:
// at this kernel speed the driver's wall clock is dominated by parsing):
//   * both files are read line by line TOGETHER; line i of one file is only ever looked at with line i of the other;
//   * a line whose first character is one of "></+" in BOTH files starts a record (the character encodes the -- unused --
//     reverse/complement op, test_prog.cpp:83-92); every other line is appended to the current record of its file;
//   * reading stops at the end of the shorter file; a sequence line before the first header is an error (test_prog.cpp:137).
// Like std::getline, a trailing '\r' is NOT stripped. One deliberate difference: an EMPTY line in both files is skipped here,
// while the reference takes it for a record header (its strchr("></+", line[0]) also matches the terminating NUL of an
// empty string, test_prog.cpp:103-106) and then starts an empty record. Records longer than 4 GiB are refused (lengths
// are 32-bit throughout the API). No GPU needed.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "agatha_b200.h"
#include "engine_internal.h"

struct agatha_fasta_pairs {
    uint8_t *qb = nullptr, *tb = nullptr;    // malloc'd, not value-initialised: the bases are written exactly once, in parallel
    ~agatha_fasta_pairs() { free(qb); free(tb); }
    std::vector<uint64_t> qo, to;
    std::vector<uint32_t> ql, tl;
    std::vector<uint8_t> qop, top;
    uint32_t max_len = 0;
};

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char* path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) { p = ""; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char*)m;
        return true;
    }
    ~Mapped()
    {
        if (p && n) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
};

// start offset of every line (std::getline semantics: a final line without '\n' counts, an empty tail does not)
void index_lines(const Mapped& f, std::vector<uint64_t>& starts, int threads)
{
    starts.clear();
    if (f.n == 0) return;
    std::vector<std::vector<uint64_t>> part((size_t)threads);
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int t = 0, nt = 1;
#endif
        const size_t lo = f.n * (size_t)t / (size_t)nt, hi = f.n * (size_t)(t + 1) / (size_t)nt;
        auto& v = part[(size_t)t];
        const char* s = f.p + lo;
        const char* e = f.p + hi;
        while (s < e) {
            const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
            if (!nl) break;
            v.push_back((uint64_t)(nl - f.p) + 1);          // the next line starts after this newline
            s = nl + 1;
        }
    }
    starts.push_back(0);
    for (auto& v : part) starts.insert(starts.end(), v.begin(), v.end());
    if (starts.back() >= f.n) starts.pop_back();            // file ends with '\n': no extra empty line
}

inline int op_of(char c)
{
    switch (c) { case '>': return 0; case '<': return 1; case '/': return 2; case '+': return 3; default: return -1; }
}

}  // namespace

extern "C" {

agatha_fasta_pairs_t* agatha_fasta_load(const char* query_path, const char* target_path)
{
    using namespace agatha;
    Mapped fq, ft;
    if (!fq.open(query_path) || !ft.open(target_path)) {
        set_error(AGATHA_EINVAL, "File error: either a file doesn't exist, or cannot be opened.");   // args_parser.cpp:66
        return nullptr;
    }
    int threads = 1;
#ifdef _OPENMP
    threads = std::max(1, std::min(16, omp_get_max_threads()));
#endif
    std::vector<uint64_t> lq, lt;
    index_lines(fq, lq, threads);
    index_lines(ft, lt, threads);
    const size_t nlines = std::min(lq.size(), lt.size());   // lock-step: stops with the shorter file
    auto line_len = [](const Mapped& f, const std::vector<uint64_t>& ls, size_t i) -> uint64_t {
        const uint64_t b = ls[i], e = (i + 1 < ls.size()) ? ls[i + 1] - 1 : ((f.n && f.p[f.n - 1] == '\n') ? f.n - 1 : f.n);
        return e - b;
    };

    // record headers: lines that start with an op character in BOTH files
    std::vector<size_t> hdr;
    for (size_t i = 0; i < nlines; i++) {
        const uint64_t a = line_len(fq, lq, i), b = line_len(ft, lt, i);
        if (a && b && op_of(fq.p[lq[i]]) >= 0 && op_of(ft.p[lt[i]]) >= 0) hdr.push_back(i);
        else if (hdr.empty()) {
            set_error(AGATHA_EINVAL, "Batch1 and target_batch files should be fasta having same number of sequences");   // test_prog.cpp:137
            return nullptr;
        }
    }
    auto* f = new agatha_fasta_pairs();
    const size_t n = hdr.size();
    f->ql.resize(n); f->tl.resize(n); f->qo.resize(n); f->to.resize(n); f->qop.resize(n); f->top.resize(n);
    bool too_long = false;                      // lengths are 32-bit throughout the ABI (gasal.h:120-123): refuse, never truncate
#pragma omp parallel for schedule(static) num_threads(threads) reduction(|| : too_long)
    for (int64_t r = 0; r < (int64_t)n; r++) {
        const size_t b = hdr[(size_t)r] + 1, e = ((size_t)r + 1 < n) ? hdr[(size_t)r + 1] : nlines;
        uint64_t a = 0, c = 0;
        for (size_t i = b; i < e; i++) { a += line_len(fq, lq, i); c += line_len(ft, lt, i); }
        too_long = too_long || a > 0xffffffffull || c > 0xffffffffull;
        f->ql[(size_t)r] = (uint32_t)a; f->tl[(size_t)r] = (uint32_t)c;
        f->qop[(size_t)r] = (uint8_t)op_of(fq.p[lq[hdr[(size_t)r]]]);
        f->top[(size_t)r] = (uint8_t)op_of(ft.p[lt[hdr[(size_t)r]]]);
    }
    if (too_long) { delete f; set_error(AGATHA_EINVAL, "a record is longer than 4 GiB"); return nullptr; }
    uint64_t qo = 0, to = 0;
    uint32_t mx = 0;
    for (size_t r = 0; r < n; r++) {
        f->qo[r] = qo; f->to[r] = to; qo += f->ql[r]; to += f->tl[r];
        mx = std::max(mx, std::max(f->ql[r], f->tl[r]));
    }
    f->max_len = mx;
    f->qb = (uint8_t*)malloc(std::max<uint64_t>(qo, 1)); f->tb = (uint8_t*)malloc(std::max<uint64_t>(to, 1));
    if (!f->qb || !f->tb) { delete f; set_error(AGATHA_ENOMEM, "out of memory"); return nullptr; }
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t r = 0; r < (int64_t)n; r++) {
        const size_t b = hdr[(size_t)r] + 1, e = ((size_t)r + 1 < n) ? hdr[(size_t)r + 1] : nlines;
        uint8_t* dq = f->qb + f->qo[(size_t)r];
        uint8_t* dt = f->tb + f->to[(size_t)r];
        for (size_t i = b; i < e; i++) {
            const uint64_t a = line_len(fq, lq, i), c = line_len(ft, lt, i);
            memcpy(dq, fq.p + lq[i], a); dq += a;
            memcpy(dt, ft.p + lt[i], c); dt += c;
        }
    }
    return f;
}

void agatha_fasta_free(agatha_fasta_pairs_t* f) { delete f; }
uint64_t agatha_fasta_count(const agatha_fasta_pairs_t* f) { return f ? f->ql.size() : 0; }
uint32_t agatha_fasta_max_len(const agatha_fasta_pairs_t* f) { return f ? f->max_len : 0; }
const uint8_t* agatha_fasta_query_bases(const agatha_fasta_pairs_t* f) { return f->qb; }
const uint8_t* agatha_fasta_target_bases(const agatha_fasta_pairs_t* f) { return f->tb; }
const uint64_t* agatha_fasta_query_offsets(const agatha_fasta_pairs_t* f) { return f->qo.data(); }
const uint64_t* agatha_fasta_target_offsets(const agatha_fasta_pairs_t* f) { return f->to.data(); }
const uint32_t* agatha_fasta_query_lens(const agatha_fasta_pairs_t* f) { return f->ql.data(); }
const uint32_t* agatha_fasta_target_lens(const agatha_fasta_pairs_t* f) { return f->tl.data(); }
const uint8_t* agatha_fasta_query_ops(const agatha_fasta_pairs_t* f) { return f->qop.data(); }
const uint8_t* agatha_fasta_target_ops(const agatha_fasta_pairs_t* f) { return f->top.data(); }

}  // extern "C"

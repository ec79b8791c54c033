// FASTA reader for the reference driver's input format. Replaces the lock-step getline loop of
// AGAThA/test_prog/test_prog.cpp:94-149: both files are read line by line together; a line whose first character is
// one of "></+" in BOTH files starts a record (the character encodes the -- unused -- reverse/complement op,
// test_prog.cpp:83-92), other lines are appended to the current record. No GPU needed.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "agatha_b200.h"
#include "engine_internal.h"

struct agatha_fasta_pairs {
    std::vector<uint8_t> qb, tb;
    std::vector<uint64_t> qo, to;
    std::vector<uint32_t> ql, tl;
    std::vector<uint8_t> qop, top;
    uint32_t max_len = 0;
};

extern "C" {

agatha_fasta_pairs_t* agatha_fasta_load(const char* query_path, const char* target_path)
{
    using namespace agatha;
    std::ifstream fq(query_path), ft(target_path);
    if (!fq || !ft) { set_error(AGATHA_EINVAL, "File error: either a file doesn't exist, or cannot be opened."); return nullptr; }   // args_parser.cpp:66
    auto* f = new agatha_fasta_pairs();
    static const char starts[5] = "></+";
    std::string lq, lt;
    int state = 0;   // 0 = before first header, 1 = header seen, 2 = inside sequence
    auto close_record = [&]() {
        const uint32_t a = (uint32_t)(f->qb.size() - f->qo.back()), b = (uint32_t)(f->tb.size() - f->to.back());
        f->ql.push_back(a); f->tl.push_back(b);
        if (a > f->max_len) f->max_len = a;
        if (b > f->max_len) f->max_len = b;
    };
    while (std::getline(fq, lq) && std::getline(ft, lt)) {
        if (!lq.empty() && lq.back() == '\r') lq.pop_back();
        if (!lt.empty() && lt.back() == '\r') lt.pop_back();
        const char* q = lq.empty() ? nullptr : std::strchr(starts, lq[0]);
        const char* t = lt.empty() ? nullptr : std::strchr(starts, lt[0]);
        if (q && *q && t && *t) {
            if (state != 0) close_record();
            f->qop.push_back((uint8_t)(q - starts)); f->top.push_back((uint8_t)(t - starts));
            f->qo.push_back(f->qb.size()); f->to.push_back(f->tb.size());
            state = 1;
        } else if (state >= 1) {
            f->qb.insert(f->qb.end(), lq.begin(), lq.end());
            f->tb.insert(f->tb.end(), lt.begin(), lt.end());
            state = 2;
        } else {
            set_error(AGATHA_EINVAL, "Batch1 and target_batch files should be fasta having same number of sequences");   // test_prog.cpp:137
            delete f;
            return nullptr;
        }
    }
    if (state != 0) close_record();
    return f;
}

void agatha_fasta_free(agatha_fasta_pairs_t* f) { delete f; }
uint64_t agatha_fasta_count(const agatha_fasta_pairs_t* f) { return f ? f->ql.size() : 0; }
uint32_t agatha_fasta_max_len(const agatha_fasta_pairs_t* f) { return f ? f->max_len : 0; }
const uint8_t* agatha_fasta_query_bases(const agatha_fasta_pairs_t* f) { return f->qb.data(); }
const uint8_t* agatha_fasta_target_bases(const agatha_fasta_pairs_t* f) { return f->tb.data(); }
const uint64_t* agatha_fasta_query_offsets(const agatha_fasta_pairs_t* f) { return f->qo.data(); }
const uint64_t* agatha_fasta_target_offsets(const agatha_fasta_pairs_t* f) { return f->to.data(); }
const uint32_t* agatha_fasta_query_lens(const agatha_fasta_pairs_t* f) { return f->ql.data(); }
const uint32_t* agatha_fasta_target_lens(const agatha_fasta_pairs_t* f) { return f->tl.data(); }
const uint8_t* agatha_fasta_query_ops(const agatha_fasta_pairs_t* f) { return f->qop.data(); }
const uint8_t* agatha_fasta_target_ops(const agatha_fasta_pairs_t* f) { return f->top.data(); }

}  // extern "C"

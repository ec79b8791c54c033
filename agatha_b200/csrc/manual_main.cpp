// agatha_manual: command-line driver with the reference's contract (AGAThA/test_prog/test_prog.cpp, args_parser.cpp):
//   agatha_manual [-m -x -q -r -s -z -w INT] [-a INT] [-n INT] [-g INT] [-p] <query_batch.fasta> <target_batch.fasta> [raw_file]
// Same positional rules (the last two arguments, three with -p, are files; argc >= 4), same stdout line format
// "score\tquery_batch_end=..\ttarget_batch_end=.." under -p (test_prog.cpp:363-368), and one line of kernel milliseconds
// appended to raw_file under -p, which misc/avg_time.py sums into time.json. Differences: the whole input goes through
// agatha_align_job (any number of GPUs, -g), -b/-t are accepted and ignored, -n is the number of staging threads' worth of
// devices (kept for compatibility).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "agatha_b200.h"

static void usage()
{
    fprintf(stderr, "Usage: agatha_manual [-m] [-x] [-q] [-r] [-s] [-z] [-w] [-b] [-t] [-a] [-g] [-p] [-n] [-R] <query_batch.fasta> <target_batch.fasta> [raw_file]\n");
}

int main(int argc, char** argv)
{
    agatha_params_t p = {2, 4, 4, 2, 3, 400, 751};   // args_parser.cpp:12-22
    int print_out = 0, gpus = 0, batch = 0, apply_ops = 0;
    for (int c = 1; c < argc; c++) if (!strcmp(argv[c], "--help") || !strcmp(argv[c], "-h")) { usage(); return 0; }
    if (argc < 4) { fprintf(stderr, "Not enough Parameters. Required: file1.fasta file2.fasta. See help (--help, -h) for usage. \n"); return 1; }
    int c = 1;
    for (; c < argc - 3; c++) {
        const char* a = argv[c];
        if (a[0] != '-' || strlen(a) != 2) { fprintf(stderr, "Wrong argument. See help (--help, -h) for usage. \n"); return 1; }
        int* dst = nullptr; int dummy = 0;
        switch (a[1]) {
            case 'm': dst = &p.match; break;
            case 'x': dst = &p.mismatch; break;
            case 'q': dst = &p.gap_open; break;
            case 'r': dst = &p.gap_extend; break;
            case 's': dst = &p.slice_width; break;
            case 'z': dst = &p.z_threshold; break;
            case 'w': dst = &p.band_width; break;
            case 'g': dst = &gpus; break;
            case 'a': dst = &batch; break;
            case 'b': case 't': case 'n': dst = &dummy; break;
            case 'p': print_out = 1; break;
            case 'R': apply_ops = 1; break;   // extension: honour the header characters '<' '/' '+' (test_prog.cpp:83-92)
            default: break;
        }
        if (dst) { c++; *dst = atoi(argv[c]); }
    }
    const char* qpath = argv[c++];
    const char* tpath = argv[c];
    const char* rawpath = nullptr;
    if (print_out) { c++; if (c < argc) rawpath = argv[c]; }

    agatha_fasta_pairs_t* f = agatha_fasta_load(qpath, tpath);
    if (!f) { fprintf(stderr, "%s\n", agatha_last_error()); return 1; }
    const uint64_t n = agatha_fasta_count(f);
    std::vector<int32_t> score(n), qend(n), tend(n);
    agatha_job_config_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.n_devices = gpus; cfg.batch_alns = batch > 0 ? (uint32_t)batch : 0u;
    if (apply_ops) { cfg.query_ops = agatha_fasta_query_ops(f); cfg.target_ops = agatha_fasta_target_ops(f); }
    agatha_job_stats_t st;
    int rc = agatha_align_job(agatha_fasta_query_bases(f), agatha_fasta_query_offsets(f), agatha_fasta_query_lens(f),
                              agatha_fasta_target_bases(f), agatha_fasta_target_offsets(f), agatha_fasta_target_lens(f),
                              n, &p, &cfg, score.data(), qend.data(), tend.data(), nullptr, nullptr, &st);
    if (rc) { fprintf(stderr, "[GASAL ERROR:] %s\n", agatha_last_error()); return 1; }
    if (print_out) {
        std::string out;
        out.reserve(n * 48);
        char line[96];
        for (uint64_t i = 0; i < n; i++) {
            int k = snprintf(line, sizeof(line), "%d\tquery_batch_end=%d\ttarget_batch_end=%d\n", score[i], qend[i], tend[i]);
            out.append(line, (size_t)k);
        }
        fwrite(out.data(), 1, out.size(), stdout);
        if (rawpath) { std::ofstream raw(rawpath, std::ios::app); raw << st.seconds_kernel_max * 1e3 << std::endl; }
    }
    agatha_fasta_free(f);
    agatha_release_cached();
    return 0;
}

/*
 * Source-compatible shim: the host API of readwrite112/AGAThA (GASAL2 style) implemented on libagatha_b200.
 *
 * A caller written against the reference -- AGAThA/test_prog/test_prog.cpp is the only one in the reference tree --
 * compiles unchanged against this directory (put it on the include path as `include/`, it provides gasal_header.h)
 * and links libagatha_b200.so instead of libgasal.a. Same names, argument meaning, call protocol and error behaviour
 * (message on stderr + exit(EXIT_FAILURE), gasal.h:14-21). Reference declarations mirrored here:
 *   gasal.h:36-173 (enums, host_batch_t, gasal_res_t, gasal_gpu_storage_t/_v, gasal_subst_scores)
 *   args_parser.h:16-68 (Parameters), ctors.h, host_batch.h, interfaces.h, gasal_align.h, res.h
 * Field ORDER of the public structs follows the reference so that code poking at them keeps working; fields that
 * only made sense for the reference's kernel (device strips, packed buffers, op arrays on the device) stay NULL.
 */
#ifndef AGATHA_GASAL_COMPAT_H
#define AGATHA_GASAL_COMPAT_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include <fstream>
#include <iostream>
#include <string>

#include <cuda_runtime.h>

#include "../agatha_b200.h"

#ifndef HOST_MALLOC_SAFETY_FACTOR
#define HOST_MALLOC_SAFETY_FACTOR 5
#endif

#define CHECKCUDAERROR(error) \
    do { \
        err = error; \
        if (cudaSuccess != err) { \
            fprintf(stderr, "[GASAL CUDA ERROR:] %s(CUDA error no.=%d). Line no. %d in file %s\n", cudaGetErrorString(err), err, __LINE__, __FILE__); \
            exit(EXIT_FAILURE); \
        } \
    } while (0)

inline int CudaCheckKernelLaunch() { return cudaGetLastError() == cudaSuccess ? 0 : -1; }

enum comp_start { WITHOUT_START, WITH_START, WITH_TB };
enum Bool { FALSE, TRUE };
enum data_source { NONE, QUERY, TARGET, BOTH };
enum algo_type { UNKNOWN, GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW };
enum operation_on_seq { FORWARD_NATURAL, REVERSE_NATURAL, FORWARD_COMPLEMENT, REVERSE_COMPLEMENT };

/* One pinned staging page. The shim keeps exactly one page per side (it grows in place), so `next` is always NULL. */
struct host_batch {
    uint8_t *data;
    uint32_t page_size;
    uint32_t data_size;
    uint32_t offset;
    int is_locked;
    struct host_batch *next;
};
typedef struct host_batch host_batch_t;

struct gasal_res {
    int32_t *aln_score;
    int32_t *query_batch_end;
    int32_t *target_batch_end;
    int32_t *query_batch_start;   /* never filled, as in the reference (res.cpp:27-28) */
    int32_t *target_batch_start;
    uint8_t *cigar;
    uint32_t *n_cigar_ops;
};
typedef struct gasal_res gasal_res_t;

typedef struct {
    uint8_t *unpacked_query_batch;      /* device side: owned by the engine, not exposed */
    uint8_t *unpacked_target_batch;
    uint32_t *packed_query_batch;
    uint32_t *packed_target_batch;
    uint32_t *query_batch_offsets;
    uint32_t *target_batch_offsets;
    uint32_t *query_batch_lens;
    uint32_t *target_batch_lens;

    uint32_t *host_seed_scores;
    uint32_t *seed_scores;

    host_batch_t *extensible_host_unpacked_query_batch;
    host_batch_t *extensible_host_unpacked_target_batch;

    uint8_t *host_query_op;
    uint8_t *host_target_op;
    uint8_t *query_op;
    uint8_t *target_op;

    uint32_t *host_query_batch_offsets;   /* caller writes these directly (test_prog.cpp:287-323) */
    uint32_t *host_target_batch_offsets;
    uint32_t *host_query_batch_lens;
    uint32_t *host_target_batch_lens;

    gasal_res_t *host_res;                /* caller reads aln_score / query_batch_end / target_batch_end (test_prog.cpp:363-368) */
    gasal_res_t *device_cpy;
    gasal_res_t *device_res;

    gasal_res_t *host_res_second;
    gasal_res_t *device_res_second;
    gasal_res_t *device_cpy_second;

    uint32_t gpu_max_query_batch_bytes;
    uint32_t gpu_max_target_batch_bytes;

    uint32_t host_max_query_batch_bytes;
    uint32_t host_max_target_batch_bytes;

    uint32_t gpu_max_n_alns;
    uint32_t host_max_n_alns;
    uint32_t current_n_alns;

    uint64_t packed_tb_matrix_size;
    uint4 *packed_tb_matrices;

    int32_t slice_width;
    uint32_t maximum_sequence_length;
    short2 *global_buffer;                /* the reference's per-subwarp strips: not needed, holds the engine handle */
    short2 *host_buffer;

    cudaStream_t str;
    int is_free;
    int id;
} gasal_gpu_storage_t;

typedef struct {
    int n;
    gasal_gpu_storage_t *a;
} gasal_gpu_storage_v;

typedef struct {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open;
    int32_t gap_extend;
    int32_t slice_width;
    int32_t z_threshold;
    int32_t band_width;
} gasal_subst_scores;

enum fail_type { NOT_ENOUGH_ARGS, TOO_MANY_ARGS, WRONG_ARG, WRONG_FILES, WRONG_ALGO };

/* The driver's option parser (args_parser.h:24-68): -m -x -q -r -s -z -w -b -t -a -n -p <query_batch.fasta> <target_batch.fasta> [raw_file] */
class Parameters {
public:
    Parameters(int argc, char **argv);
    ~Parameters();
    void print();
    void failure(fail_type f);
    void help();
    void parse();
    void fileopen();

    int32_t sa;
    int32_t sb;
    int32_t gapo;
    int32_t gape;

    int print_out;
    int n_threads;

    int slice_width;
    int z_threshold;
    int band_width;

    int32_t kernel_block_num;    /* accepted for compatibility; the engine sizes its own persistent grid */
    int32_t kernel_thread_num;
    int32_t kernel_align_num;

    bool isPacked;
    bool isReverseComplement;

    std::string query_batch_fasta_filename;
    std::string target_batch_fasta_filename;
    std::string raw_filename;

    std::ifstream query_batch_fasta;
    std::ifstream target_batch_fasta;
    std::ofstream raw_file;

private:
    int argc;
    char **argv;
};

/* ctors.h */
gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams);
void gasal_init_streams(gasal_gpu_storage_v *gpu_storage_vec, int max_query_len, int max_target_len, int32_t maximum_sequence_length, Parameters *params);
void gasal_destroy_streams(gasal_gpu_storage_v *gpu_storage_vec, Parameters *params);
void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v *gpu_storage_vec);

/* host_batch.h */
host_batch_t *gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset);
void gasal_host_batch_destroy(host_batch_t *res);
host_batch_t *gasal_host_batch_getlast(host_batch_t *arg);
void gasal_host_batch_reset(gasal_gpu_storage_t *gpu_storage);
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char *data, uint32_t size, data_source SRC);
uint32_t gasal_host_batch_add(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char *data, uint32_t size, data_source SRC);
uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char base, data_source SRC);
void gasal_host_batch_print(host_batch_t *res);
void gasal_host_batch_printall(host_batch_t *res);

/* interfaces.h */
void gasal_host_alns_resize(gasal_gpu_storage_t *gpu_storage, int new_max_alns, Parameters *params);
void gasal_op_fill(gasal_gpu_storage_t *gpu_storage_t, uint8_t *data, uint32_t nbr_seqs_in_stream, data_source SRC);
void gasal_set_device(int gpu_select = 0, bool isPrintingProp = true);

/* gasal_align.h */
void gasal_copy_subst_scores(gasal_subst_scores *subst);
void gasal_aln_async(gasal_gpu_storage_t *gpu_storage, const uint32_t actual_query_batch_bytes, const uint32_t actual_target_batch_bytes, const uint32_t actual_n_alns, Parameters *params);
int gasal_is_aln_async_done(gasal_gpu_storage_t *gpu_storage);

/* res.h */
gasal_res_t *gasal_res_new_host(uint32_t max_n_alns, Parameters *params);
void gasal_res_destroy_host(gasal_res_t *res);
gasal_res_t *gasal_res_new_device(gasal_res_t *device_cpy);                         /* res.h:5 */
gasal_res_t *gasal_res_new_device_cpy(uint32_t max_n_alns, Parameters *params);     /* res.h:6 */
void gasal_res_destroy_device(gasal_res_t *device_res, gasal_res_t *device_cpy);    /* res.h:9 */
void gasal_gpu_mem_alloc(gasal_gpu_storage_t *gpu_storage, int gpu_max_query_batch_bytes, int gpu_max_target_batch_bytes, Parameters *params);   /* ctors.h:9 */
void gasal_gpu_mem_free(gasal_gpu_storage_t *gpu_storage, Parameters *params);      /* ctors.h:11 */

#endif

/*
 * agatha_b200 -- C ABI of the B200-native guided-alignment engine.
 *
 * Drop-in boundary for the seed-extension path of readwrite112/AGAThA (GASAL2-style host API). Every entry point
 * names the reference interface it replaces (paths relative to /root/reference/AGAThA/src). Plain pointers and
 * sizes only; no C++ or torch types cross this boundary. A source-compatible C++ shim that keeps the reference's
 * own names (gasal_init_streams, gasal_aln_async, ...) on top of this ABI is in include/gasal_compat.h.
 *
 * Error behaviour: functions return 0 on success and a negative AGATHA_E* code on failure (the reference prints and
 * exit(1)s, gasal.h:14-21; the compat shim restores that behaviour). There is NO CPU fallback: without a CUDA device
 * every compute entry point fails with AGATHA_ENODEV.
 */
#ifndef AGATHA_B200_H
#define AGATHA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGATHA_B200_ABI_VERSION 2

enum {
    AGATHA_OK = 0,
    AGATHA_EINVAL = -1,     /* bad argument (the reference's arg checks, gasal_align.cu:33-68) */
    AGATHA_ENODEV = -2,     /* no usable CUDA device */
    AGATHA_ECUDA = -3,      /* a CUDA call failed; see agatha_last_error() */
    AGATHA_ENOMEM = -4,
    AGATHA_EUNSUPPORTED = -5 /* band width outside what the kernels cover */
};

/* How an alignment ended. The reference computes this (`terminated`, agatha_kernel.h:62,190,306) but never
 * exports it; it is checked against the CPU oracle. */
enum {
    AGATHA_STOP_END = 0,      /* all anti-diagonals scanned */
    AGATHA_STOP_ZDROP = 1,    /* Z-drop fired */
    AGATHA_STOP_BANDEXIT = 2  /* band left the DP matrix at a slice boundary (agatha_kernel.h:189-191) */
};

/* Scoring and band parameters: same fields, order and meaning as gasal_subst_scores (gasal.h:165-173),
 * i.e. the -m -x -q -r -s -z -w options of the reference driver (args_parser.cpp:135-171). */
typedef struct {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open;
    int32_t gap_extend;
    int32_t slice_width;
    int32_t z_threshold;
    int32_t band_width;
} agatha_params_t;

/* Last error message of the calling thread ("" if none). */
const char *agatha_last_error(void);

/* Number of CUDA devices visible (0 when there is none: nothing else will work). */
int agatha_device_count(void);

/* Largest band width the compiled kernels support. */
int agatha_max_band_width(void);

/* ------------------------------------------------------------------------------------------------------------
 * Device-level entry points: device pointers, caller's stream (a cudaStream_t passed as void*), asynchronous.
 * ------------------------------------------------------------------------------------------------------------ */

/* ASCII bases -> 4-bit packed words, 8 bases per word. Replaces gasal_pack_kernel (kernels/pack_rc_seqs.h:13-53).
 * Input layout is the reference's unpacked batch: every sequence starts at a multiple of 8 and is padded to a
 * multiple of 8 with 'N' (host_batch.cpp:143-146). n_bytes must be a multiple of 8 (gasal_align.cu:46-53).
 * The packed encoding is private to this library; d_query_packed / d_target_packed need n_bytes/2
 * bytes each plus AGATHA_PACK_SLACK_WORDS words of slack. */
#define AGATHA_PACK_SLACK_WORDS 64
int agatha_pack_device(const uint8_t *d_query_bases, uint64_t query_bytes,
                       const uint8_t *d_target_bases, uint64_t target_bytes,
                       uint32_t *d_query_packed, uint32_t *d_target_packed, void *stream);

/* Per-sequence reverse / complement, GASAL2's "op" byte: bit 0 = reverse, bit 1 = complement (A<->T, C<->G on the 4-bit
 * code). Replaces gasal_reversecomplement_kernel (kernels/pack_rc_seqs.h:56-212), which gasal_aln_async launches after
 * packing when params->isReverseComplement is set (gasal_align.cu:199-212). Call it AFTER agatha_pack_device on the same
 * stream, with the same ASCII batch: sequences whose op is non-zero are packed again with the operation applied to their
 * real bases (the 'N' padding stays behind the sequence); sequences with op 0 are left alone.
 * NOTE: bit 0 implements the INTENDED reverse. The reference's kernel, as compiled, does not reverse anything useful
 * (its padding counter compares 4-bit codes with N_CODE = 0x4E and is always 0), so results on
 * reversed sequences are pinned by this repository's oracle, not by the reference GPU binary. The complement bit is identical. */
int agatha_apply_ops_device(const uint8_t *d_query_bases, const uint8_t *d_target_bases,
                            const uint32_t *d_query_offsets, const uint32_t *d_target_offsets,
                            const uint32_t *d_query_lens, const uint32_t *d_target_lens,
                            const uint8_t *d_query_ops, const uint8_t *d_target_ops, uint32_t n_alns,
                            uint32_t *d_query_packed, uint32_t *d_target_packed, void *stream);

/* The alignment itself. Replaces agatha_sort + host std::sort + agatha_kernel (gasal_align.cu:10-23,
 * kernels/agatha_kernel.h:49-458).
 *   d_*_offsets  start of each sequence in BASES (multiples of 8), as the reference's *_batch_offsets
 *   d_*_lens     sequence lengths in bases
 *   d_order      optional processing order (job -> pair index), longest first; NULL = input order.
 *                Scheduling only, never changes results.
 *   d_score/d_query_end/d_target_end   results, indexed like the inputs (gasal_res_t fields, gasal.h:85-94)
 *   d_stop, d_dstop  optional (may be NULL): AGATHA_STOP_* and the number of leading anti-diagonals needed
 *   d_workspace  at least AGATHA_WORKSPACE_BYTES bytes, private to this call until it completes */
#define AGATHA_WORKSPACE_BYTES 256
int agatha_extend_device(const uint32_t *d_query_packed, const uint32_t *d_target_packed,
                         const uint32_t *d_query_offsets, const uint32_t *d_target_offsets,
                         const uint32_t *d_query_lens, const uint32_t *d_target_lens,
                         const uint32_t *d_order, uint32_t n_alns, const agatha_params_t *params,
                         int32_t *d_score, int32_t *d_query_end, int32_t *d_target_end,
                         int32_t *d_stop, int32_t *d_dstop, void *d_workspace, void *stream);

/* Measured integer issue rates of `device`, in 1e12 lane-operations per second: a stream of VIADDMNMX.U16x2 (the ALU pipe,
 * which bounds the extension kernels), of IMAD (the FMA pipe), and of both interleaved. A few milliseconds; bench.py calls it
 * in the same run as the measurement so that the roofline denominator is not a number from another day. Any pointer may be NULL. */
int agatha_measure_int_peak(int device, double *alu_tera_lane_ops, double *fma_tera_lane_ops, double *mixed_tera_lane_ops);

/* Number of kernel launches issued by this library in the calling process (pack + extend), for bench accounting. */
uint64_t agatha_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Host-level batch API: host buffers in, host results out. One agatha_stream_t is the equivalent of one
 * gasal_gpu_storage_t (gasal.h:97-155): a CUDA stream with its pinned staging and device buffers.
 * Replaces gasal_init_streams / gasal_aln_async / gasal_is_aln_async_done / gasal_destroy_streams
 * (ctors.cpp:26-167, gasal_align.cu:27-292).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct agatha_stream agatha_stream_t;

/* device: CUDA ordinal (the reference's gasal_set_device, interfaces.cpp:86-116). Capacities grow on demand. */
agatha_stream_t *agatha_stream_create(int device, uint32_t max_alns, uint64_t max_query_bytes, uint64_t max_target_bytes);
void agatha_stream_destroy(agatha_stream_t *s);

/* Pinned staging the caller fills directly (like host_batch_t pages + host_*_offsets/lens, gasal.h:74-82,120-123).
 * agatha_stream_reserve grows them; pointers returned earlier are invalidated by a growing reserve. */
int agatha_stream_reserve(agatha_stream_t *s, uint32_t n_alns, uint64_t query_bytes, uint64_t target_bytes);
/* Current capacities of the pinned staging (any pointer may be NULL). */
void agatha_stream_capacity(agatha_stream_t *s, uint32_t *max_alns, uint64_t *query_bytes, uint64_t *target_bytes);
uint8_t *agatha_stream_query_bases(agatha_stream_t *s);
uint8_t *agatha_stream_target_bases(agatha_stream_t *s);
uint32_t *agatha_stream_query_offsets(agatha_stream_t *s);
uint32_t *agatha_stream_target_offsets(agatha_stream_t *s);
uint32_t *agatha_stream_query_lens(agatha_stream_t *s);
uint32_t *agatha_stream_target_lens(agatha_stream_t *s);
/* Per-sequence op bytes (host_query_op / host_target_op, gasal.h:113-114; filled by gasal_op_fill, interfaces.cpp:69-84).
 * Zero-initialised; only read by agatha_stream_submit_ops. */
uint8_t *agatha_stream_query_ops(agatha_stream_t *s);
uint8_t *agatha_stream_target_ops(agatha_stream_t *s);

/* Pinned staging for batches that are packed on the host (agatha_pack_batch): the same buffers seen as 32-bit words,
 * capacity = reserved bytes / 4 words each. */
uint32_t *agatha_stream_query_packed(agatha_stream_t *s);
uint32_t *agatha_stream_target_packed(agatha_stream_t *s);

/* Asynchronous: H2D of the staged batch, pack, length-aware bucketing, extension kernel, D2H of the results.
 * query_bytes/target_bytes > 0 and multiples of 8, n_alns > 0 (gasal_align.cu:33-68). */
int agatha_stream_submit(agatha_stream_t *s, uint64_t query_bytes, uint64_t target_bytes, uint32_t n_alns,
                         const agatha_params_t *params);
/* Same, with the staged op bytes applied between packing and alignment (the reference's isReverseComplement path). */
int agatha_stream_submit_ops(agatha_stream_t *s, uint64_t query_bytes, uint64_t target_bytes, uint32_t n_alns,
                             const agatha_params_t *params);
/* Same for a batch staged in packed form (agatha_stream_*_packed() filled by agatha_pack_batch): query_bases / target_bases
 * are the staged BASES (multiples of 8), half as many bytes are uploaded and no pack kernel runs. Ops, if any, were
 * applied by agatha_pack_batch. */
int agatha_stream_submit_packed(agatha_stream_t *s, uint64_t query_bases, uint64_t target_bases, uint32_t n_alns,
                                const agatha_params_t *params);
/* 0 = finished (results valid until the next submit), -1 = still running, -2 = nothing submitted
 * (the three return values of gasal_is_aln_async_done, gasal_align.cu:276-292). */
int agatha_stream_poll(agatha_stream_t *s);
int agatha_stream_wait(agatha_stream_t *s);
/* Device milliseconds of the last finished batch: [0] H2D+pack, [1] extension kernel, [2] whole batch incl. D2H. */
int agatha_stream_timings(agatha_stream_t *s, float ms[3]);

const int32_t *agatha_stream_scores(agatha_stream_t *s);
const int32_t *agatha_stream_query_ends(agatha_stream_t *s);
const int32_t *agatha_stream_target_ends(agatha_stream_t *s);
const int32_t *agatha_stream_stops(agatha_stream_t *s);
const int32_t *agatha_stream_dstops(agatha_stream_t *s);

/* ------------------------------------------------------------------------------------------------------------
 * Whole-job API: any number of pairs, any number of GPUs of one box. Pairs are independent, so they are
 * sharded over the devices by a cost-balancing host scheduler (no collective); each device runs double-buffered
 * streams; results come back in input order. Sequences are plain ASCII, NOT padded; offsets in bytes.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t n_devices;        /* <= 0: use all visible devices */
    const int32_t *devices;   /* optional list of ordinals, NULL = 0..n_devices-1 */
    uint32_t batch_alns;      /* alignments per batch, 0 = default (8192) */
    int32_t streams_per_device; /* 0 = default (5): batches in flight per device; below 4 the device runs short of launched kernels */
    int32_t staging_threads;  /* host threads per device that copy sequences into pinned staging, 0 = min(8, cores / devices) */
    const uint8_t *query_ops; /* optional per-pair op bytes (bit 0 reverse, bit 1 complement); both NULL = no ops */
    const uint8_t *target_ops;
} agatha_job_config_t;

typedef struct {
    double seconds_total;     /* wall clock of the call */
    double seconds_kernel_max;/* largest per-device sum of extension-kernel time */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t n_batches, n_devices;
} agatha_job_stats_t;

int agatha_align_job(const uint8_t *query_bases, const uint64_t *query_offsets, const uint32_t *query_lens,
                     const uint8_t *target_bases, const uint64_t *target_offsets, const uint32_t *target_lens,
                     uint64_t n_alns, const agatha_params_t *params, const agatha_job_config_t *cfg,
                     int32_t *score, int32_t *query_end, int32_t *target_end, int32_t *stop, int32_t *dstop,
                     agatha_job_stats_t *stats);

/* The same, plus start positions: gasal_res_t.query_batch_start / target_batch_start (gasal.h:89-90), which the reference
 * declares but never allocates or fills (res.cpp:27-28; nothing to be bit-compatible with). Convention: GASAL2's WITH_START
 * (gasal.h:36-39) -- the start of the best-scoring alignment that ends in the reported end cell, found by a second extension
 * running backwards from that cell over the reversed prefixes (same scoring and band, Z-drop off). A pair whose score is 0
 * reports the origin. Costs a second pass over the aligned prefixes. Not available together with per-pair ops. */
int agatha_align_job_starts(const uint8_t *query_bases, const uint64_t *query_offsets, const uint32_t *query_lens,
                            const uint8_t *target_bases, const uint64_t *target_offsets, const uint32_t *target_lens,
                            uint64_t n_alns, const agatha_params_t *params, const agatha_job_config_t *cfg,
                            int32_t *score, int32_t *query_end, int32_t *target_end, int32_t *stop, int32_t *dstop,
                            int32_t *query_start, int32_t *target_start, agatha_job_stats_t *stats);

/* agatha_align_job keeps its streams (pinned staging, device buffers) for the next call; this frees them. */
void agatha_release_cached(void);

/* ------------------------------------------------------------------------------------------------------------
 * Host utilities (no GPU needed).
 * ------------------------------------------------------------------------------------------------------------ */

/* Length-aware bucketing: job order with the most expensive pairs first (cost = cells inside the band).
 * Replaces agatha_sort + std::sort (agatha_kernel.h:434-458, gasal_align.cu:14-18). */
int agatha_bucket_order(const uint32_t *query_lens, const uint32_t *target_lens, uint32_t n, int32_t band_width,
                        uint32_t *order_out);

/* Cost-balanced sharding of n pairs over n_shards devices (greedy longest-processing-time). shard_out[i] in [0,n_shards). */
int agatha_shard_pairs(const uint32_t *query_lens, const uint32_t *target_lens, uint64_t n, int32_t band_width,
                       int32_t n_shards, int32_t *shard_out);

/* In-band real cells on anti-diagonals < dstop for each pair (the GCUPS numerator).
 * dstop may be NULL (= whole matrix). */
int agatha_count_cells(const uint32_t *query_lens, const uint32_t *target_lens, const int32_t *dstop, uint64_t n,
                       int32_t band_width, uint64_t *cells_out, uint64_t *total_out);

/* Stage many sequences the way gasal_host_batch_fill does one (host_batch.cpp:79-154): sequence i is copied to the next
 * multiple of 8 in dst and padded to a multiple of 8 with 'N'. dst_offsets (in bases == bytes) and *bytes_out follow the
 * reference's conventions; dst may be pinned staging obtained from agatha_stream_*_bases(). Fails if the batch does not
 * fit dst_capacity or 32-bit offsets. ids (optional) selects and orders the sequences: sequence j is ids[j]. */
int agatha_stage_batch(const uint8_t *bases, const uint64_t *offsets, const uint32_t *lens, const uint64_t *ids, uint64_t n,
                       uint8_t *dst, uint64_t dst_capacity, uint32_t *dst_offsets, uint32_t *dst_lens, uint64_t *bytes_out,
                       int32_t n_threads);
/* Bytes agatha_stage_batch will need. */
uint64_t agatha_staged_bytes(const uint32_t *lens, const uint64_t *ids, uint64_t n);

/* The same staging, but straight into the packed device format (8 bases per 32-bit word, the library's private 4-bit codes):
 * host_batch.cpp:79-154 and the pack kernel (kernels/pack_rc_seqs.h:13-53) in one pass over the bases, so that only half the
 * bytes cross PCIe and nothing is packed on the device. is_target selects the target word layout (first base in the bottom
 * nibble; queries carry it in the top nibble). ops (optional, indexed like lens) applies reverse (bit 0) / complement
 * (bit 1) to a sequence on the way, as gasal_reversecomplement_kernel would (pack_rc_seqs.h:56-212). dst_offsets are in BASES
 * (multiples of 8; word index = offset / 8), *bases_out = staged bases (a multiple of 8): the values agatha_stream_submit_packed
 * and agatha_extend_device expect. dst_words may be pinned staging from agatha_stream_*_packed(). */
int agatha_pack_batch(const uint8_t *bases, const uint64_t *offsets, const uint32_t *lens, const uint64_t *ids, const uint8_t *ops,
                      uint64_t n, int32_t is_target, uint32_t *dst_words, uint64_t dst_capacity_words,
                      uint32_t *dst_offsets, uint32_t *dst_lens, uint64_t *bases_out, int32_t n_threads);

/* FASTA reader for the reference's input format: records ">>> idx" + sequence lines, both files read in
 * lock-step (test_prog.cpp:94-149). Returns an opaque handle or NULL. */
typedef struct agatha_fasta_pairs agatha_fasta_pairs_t;
agatha_fasta_pairs_t *agatha_fasta_load(const char *query_path, const char *target_path);
void agatha_fasta_free(agatha_fasta_pairs_t *f);
uint64_t agatha_fasta_count(const agatha_fasta_pairs_t *f);
uint32_t agatha_fasta_max_len(const agatha_fasta_pairs_t *f);
const uint8_t *agatha_fasta_query_bases(const agatha_fasta_pairs_t *f);
const uint8_t *agatha_fasta_target_bases(const agatha_fasta_pairs_t *f);
const uint64_t *agatha_fasta_query_offsets(const agatha_fasta_pairs_t *f);
const uint64_t *agatha_fasta_target_offsets(const agatha_fasta_pairs_t *f);
const uint32_t *agatha_fasta_query_lens(const agatha_fasta_pairs_t *f);
const uint32_t *agatha_fasta_target_lens(const agatha_fasta_pairs_t *f);
const uint8_t *agatha_fasta_query_ops(const agatha_fasta_pairs_t *f);   /* header char -> 0..3 (test_prog.cpp:83-92) */
const uint8_t *agatha_fasta_target_ops(const agatha_fasta_pairs_t *f);

#ifdef __cplusplus
}
#endif
#endif /* AGATHA_B200_H */

/* sb200.h — C ABI of libspades_b200.so: the B200-native (sm_100a) graph-construction front end of SPAdes 3.15.4.
 *
 * The reference has no plugin/FFI layer on this path: the boundary is the set of C++ template interfaces its three
 * callers use (spades-core's Construction stage, spades-gbuilder, spades-kmercount).  Each entry point below names
 * the reference interface it stands in for (paths relative to /root/reference/assembler/src/common); the C++ adapter
 * that plugs them back into SPAdes is include/sb200_adapters.hpp, and INTEGRATION.md shows the
 * reference-side edit.
 *
 * Conventions
 *   - plain pointers and sizes only; every handle is opaque and owned by the caller until its *_free.
 *   - return value 0 = ok; nonzero = error, message via sb200_last_error(ctx).  The reference aborts with
 *     FATAL_ERROR / VERIFY on the same conditions (utils/logger/logger.hpp:177-190); the adapter maps nonzero to that.
 *   - every call is blocking and must come from one host thread per context (the reference calls these interfaces
 *     from a single stage thread, pipeline/stage.cpp:143-204).
 *   - there is NO CPU fallback: sb200_create fails when no sm_100 device is present.
 *   - k-mer records: W = ceil(K/32) little-endian uint64 words, base i (A,C,G,T = 0..3) at bits 2(i%32) of word i/32
 *     (sequence/rtseq.hpp:34-131).  "File order" = hash bucket (utils/kmer_mph/kmer_buckets.hpp:28-41), then word-wise
 *     ascending from word 0 (adt/array_vector.hpp:247-256) — the order of KMerDiskStorage's kmers<i> files and of
 *     final_kmers.
 */
#ifndef SB200_H
#define SB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb200_ctx sb200_ctx;
typedef struct sb200_reads sb200_reads;       /* packed reads resident in HBM                                        */
typedef struct sb200_kmers sb200_kmers;       /* kmers::KMerDiskStorage<RtSeq>  (kmer_mph/kmer_index_builder.hpp:48-191) */
typedef struct sb200_mphf sb200_mphf;         /* kmers::KMerIndex<traits>       (kmer_mph/kmer_index.hpp:25-147)      */
typedef struct sb200_ext sb200_ext;           /* utils::DeBruijnExtensionIndex<> payload (extension_index/kmer_extension_index.hpp:242-339) */
typedef struct sb200_unitigs sb200_unitigs;   /* std::vector<Sequence> of UnbranchingPathExtractor                   */

/* ---- context ------------------------------------------------------------------------------------------------------ */
int  sb200_create(int device, sb200_ctx **out);          /* fails (nonzero, *out = NULL) without an sm_100 GPU      */
void sb200_destroy(sb200_ctx *ctx);
const char *sb200_last_error(const sb200_ctx *ctx);      /* ctx may be NULL: error of the last failed sb200_create  */
int  sb200_synchronize(sb200_ctx *ctx);
void *sb200_stream(sb200_ctx *ctx);                       /* cudaStream_t all work of this context is launched on    */
uint64_t sb200_kernel_launches(sb200_ctx *ctx, int reset);/* kernels launched by this library since the last reset   */
/* Per-kernel device timing: CUDA events around every launch on the context's stream (measurement aid for bench.py).  */
int  sb200_profile(sb200_ctx *ctx, int enable);           /* clears collected records                                 */
int  sb200_profile_report(sb200_ctx *ctx, char *out, uint64_t cap, uint64_t *needed);  /* "name\tlaunches\tms\n" lines */

/* ---- reads: io::ReadStreamList<io::SingleReadSeq> (io/reads/read_stream.hpp:63-87; binary layout of
 *      io/reads/single_read.hpp:279-299 / sequence/sequence.hpp:399-428: every read word-aligned, 2 bits per base) -- */
int  sb200_reads_upload(sb200_ctx *ctx, const uint64_t *words, const uint64_t *word_off /* n_reads+1 */,
                        const uint32_t *len, uint64_t n_reads, sb200_reads **out);           /* host -> HBM             */
int  sb200_reads_wrap_device(sb200_ctx *ctx, const uint64_t *d_words, const uint64_t *d_word_off, const uint32_t *d_len,
                             uint64_t n_reads, uint64_t n_words, sb200_reads **out);         /* HBM -> HBM copy         */
void sb200_reads_free(sb200_reads *r);

/* ---- KMerDiskCounter<RtSeq>::Count over DeBruijnReadKMerSplitter (kmer_mph/kmer_index_builder.hpp:241-267,
 *      kmer_mph/kmer_splitters.hpp:25-41,109-133, kmer_mph/kmer_splitter.hpp:120-167).
 *      canonical_only = StoringTypeFilter<InvertableStoring> (ph_map/storing_traits.hpp:88-101);
 *      add_rc = the streams are RC-wrapped (io/dataset_support/read_converter.cpp:212-242).
 *      gbuilder / spades-core: K = k+1, canonical_only = 1, add_rc = 1, num_buckets = 10*nthreads;
 *      spades-kmercount (projects/kmercount/main.cpp:186-228): canonical_only = 0, add_rc = 1, num_buckets = 16.
 *      Fails with the reference's message when no k-mer is extracted (kmer_index_builder.hpp:261-264). ---------------- */
int  sb200_count(sb200_ctx *ctx, const sb200_reads *reads, unsigned K, int canonical_only, int add_rc,
                 unsigned num_buckets, sb200_kmers **out);
/* KMerDiskCounter over DeBruijnKMerKMerSplitter(K_target = K_source-1, add_rc = true)
 * (kmer_mph/kmer_splitters.hpp:135-204; extension_index/kmer_extension_index_builder.hpp:88-97) */
int  sb200_derive_kmers(sb200_ctx *ctx, const sb200_kmers *kpomers, unsigned num_buckets, sb200_kmers **out);

/* KMerDiskStorage accessors: k(), num_buckets(), total_kmers(), bucket_size(i), bucket_begin/end(i), final_kmers() */
unsigned sb200_kmers_k(const sb200_kmers *s);
unsigned sb200_kmers_words(const sb200_kmers *s);
unsigned sb200_kmers_num_buckets(const sb200_kmers *s);
uint64_t sb200_kmers_size(const sb200_kmers *s);
uint64_t sb200_kmers_instances(const sb200_kmers *s);     /* k-mer instances that went into the count (0 if derived)  */
int  sb200_kmers_bucket_starts(const sb200_kmers *s, uint64_t *out /* num_buckets+1 */);
int  sb200_kmers_download(const sb200_kmers *s, uint64_t first, uint64_t count, uint64_t *records_out /* count*words */);
int  sb200_kmers_counts_download(const sb200_kmers *s, uint64_t first, uint64_t count, uint32_t *counts_out);
const uint64_t *sb200_kmers_device_records(const sb200_kmers *s);   /* device pointers, for zero-copy consumers        */
const uint32_t *sb200_kmers_device_counts(const sb200_kmers *s);
void sb200_kmers_free(sb200_kmers *s);

/* ---- KMerIndexBuilder<Index>::BuildIndex(index, storage) (kmer_mph/kmer_index_builder.hpp:383-433): one BooPHF
 *      (gamma 4, 25 levels, XXH3_128; ext/include/boomphf/BooPHF.h) per bucket + segment_starts_ ------------------ */
int  sb200_mphf_build(sb200_ctx *ctx, const sb200_kmers *kmers, sb200_mphf **out);
uint64_t sb200_mphf_size(const sb200_mphf *m);                                  /* KMerIndex::size()                   */
uint64_t sb200_mphf_mem_size(const sb200_mphf *m);                              /* KMerIndex::mem_size(), bytes        */
/* KMerIndex::seq_idx for n host records (kmer_index.hpp:85-90); idx_out[i] = ~0 if the key falls through all levels */
int  sb200_mphf_lookup(sb200_ctx *ctx, const sb200_mphf *m, const uint64_t *records, uint64_t n, uint64_t *idx_out);
/* KMerIndex::seq_idx(const Seq &) for ONE key, evaluated on the HOST (the bit-vectors + rank samples are copied out once): for the
 * reference's per-key consumers after the path (link records, edge index); thread-safe; *idx_out = ~0 if the key falls through */
int  sb200_mphf_seq_idx(const sb200_mphf *m, const uint64_t *record, uint64_t *idx_out);
/* KMerIndex::serialize bytes (kmer_index.hpp:99-105); out == NULL: size query */
int  sb200_mphf_serialize(const sb200_mphf *m, uint8_t *out, uint64_t *size);
void sb200_mphf_free(sb200_mphf *m);

/* ---- DeBruijnExtensionIndexBuilder::BuildExtensionIndexFromKPOMers, mask part
 *      (extension_index/kmer_extension_index_builder.hpp:82-106,44-59) ------------------------------------------------ */
int  sb200_ext_build(sb200_ctx *ctx, const sb200_kmers *kpomers, const sb200_kmers *kmers, const sb200_mphf *mphf,
                     sb200_ext **out);
int  sb200_ext_masks_download(const sb200_ext *e, uint8_t *masks_out /* size bytes, MPHF-index order = data_ */);
int  sb200_ext_idx_download(const sb200_ext *e, uint32_t *idx_out /* size, file order */);
void sb200_ext_free(sb200_ext *e);

/* ---- EarlyTipClipperProcessor(index, length_bound).ClipTips()
 *      (assembly_graph/construction/early_simplification.hpp:37-160); *removed = its return value -------------------- */
int  sb200_tipclip(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext,
                   uint64_t length_bound, uint64_t *removed);

/* ---- UnbranchingPathExtractor(index, k).ExtractUnbranchingPaths / ...AndLoops
 *      (assembly_graph/construction/debruijn_graph_constructor.hpp:182-388).  Sequences come in the reference's
 *      output order.  Unlike the reference the masks are left untouched (it isolates every consumed vertex). -------- */
int  sb200_unitigs_extract(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext,
                           int with_loops, sb200_unitigs **out);
uint64_t sb200_unitigs_count(const sb200_unitigs *u);
uint64_t sb200_unitigs_loops(const sb200_unitigs *u);
uint64_t sb200_unitigs_total_bases(const sb200_unitigs *u);
uint64_t sb200_unitigs_total_words(const sb200_unitigs *u);
/* packed like reads: sequence i = words[word_off[i] .. word_off[i+1]), len[i] bases */
int  sb200_unitigs_download(const sb200_unitigs *u, uint64_t *words_out, uint64_t *word_off_out /* count+1 */,
                            uint32_t *len_out);
void sb200_unitigs_free(sb200_unitigs *u);

/* ---- coverage: CoverageHashMapBuilder::BuildIndex (ph_map/coverage_hash_map_builder.hpp:15-54) — PerfectHashMap<RtSeq, uint32_t> over
 *      the (k+1)-mers = a KMerIndex of the (k+1)-mer storage + one multiplicity per (k+1)-mer in ITS index order (no second pass over
 *      the reads: the multiplicities are the run lengths of sb200_count) — and GraphCoverageFiller / FillCoverageAndFlankingFromPHM
 *      (assembly_graph/graph_support/coverage_filling.hpp:16-95) over the unitigs: kc[i] = sum of the multiplicities of the (k+1)-mers
 *      of sequence i (CoverageIndex raw coverage, GFA KC:i; DP:f = kc / (length - k)); flank[2i], flank[2i+1] = the same over its first /
 *      last `averaging_range` (k+1)-mers (FlankingCoverage raw coverage of the edge / of its conjugate; spades-gbuilder -c uses 50,
 *      projects/gbuilder/main.cpp:200-211). --------------------------------------------------------------------------------------- */
typedef struct sb200_covmap sb200_covmap;
int  sb200_coverage_map_build(sb200_ctx *ctx, const sb200_kmers *kpomers, sb200_covmap **out);
const sb200_mphf *sb200_coverage_map_index(const sb200_covmap *c);          /* borrowed: KMerIndex over the (k+1)-mers            */
uint64_t sb200_coverage_map_size(const sb200_covmap *c);
int  sb200_coverage_map_values_download(const sb200_covmap *c, uint32_t *values_out /* size, index order = data_ */);
void sb200_coverage_map_free(sb200_covmap *c);
int  sb200_unitigs_coverage(sb200_ctx *ctx, const sb200_covmap *c, const sb200_unitigs *u, uint32_t averaging_range,
                            uint64_t *kc_out /* count */, uint64_t *flank_out /* 2 * count, or NULL */);

/* ---- hash-sharded path (several GPUs, one process each): the same stages with the shuffle points exposed.
 *      GPU g of G owns the buckets [g*B/G, (g+1)*B/G) of KMerSegmentPolicy — a contiguous range of the reference's file
 *      order — so the shards concatenated in rank order are the single-GPU (= reference) result.  The reference has no
 *      counterpart (it shuffles through kmers_raw<i> files, kmer_splitter.hpp:140-161); the caller moves the byte ranges
 *      between GPUs when using these building blocks directly (any transport; sb200_construct_sharded below is the complete path,
 *      with the exchanges inside the library).
 *        records_extract_partitioned | records_derive -> records_partition   (grouped by owner, counts per owner)
 *        [exchange into records_alloc] -> count_records_owned            : this GPU's shard of KMerDiskStorage
 *        mphf_build_sharded + [sum bits/ranks over GPUs] + mphf_complete  : the whole KMerIndex on every GPU
 *        ext_build(local (k+1)-mers, local k-mers, whole index) + [sum masks over GPUs]
 *        unitigs_extract_local                                            : sequences whose start junction is in the shard */
typedef struct sb200_records sb200_records;   /* unsorted k-mer instances on their way through the shuffle */
int  sb200_records_extract(sb200_ctx *ctx, const sb200_reads *reads, unsigned K, int canonical_only, int add_rc, sb200_records **out);
int  sb200_records_derive(sb200_ctx *ctx, const sb200_kmers *kpomers, sb200_records **out);
int  sb200_records_partition(sb200_ctx *ctx, sb200_records *r, unsigned num_buckets, unsigned n_owners, uint64_t *counts_out /* n_owners */);
/* records_extract + records_partition in one step: for the canonical modes the extraction itself stores every record at its owner's
 * cursor (count pass, scan, write pass) and the instances cross HBM once; the order inside an owner's group is unspecified. */
int  sb200_records_extract_partitioned(sb200_ctx *ctx, const sb200_reads *reads, unsigned K, int canonical_only, int add_rc, unsigned num_buckets,
                                       unsigned n_owners, uint64_t *counts_out /* n_owners */, sb200_records **out);
int  sb200_records_alloc(sb200_ctx *ctx, uint64_t n, unsigned K, int flags /* = sb200_records_flags of the senders */, sb200_records **out);
uint64_t sb200_records_size(const sb200_records *r);
unsigned sb200_records_words(const sb200_records *r);
unsigned sb200_records_k(const sb200_records *r);
int  sb200_records_flags(const sb200_records *r);        /* bit 0: double_palindromes, bit 1: marker, bit 2: mask-bit payload in the padding */
uint64_t *sb200_records_device(sb200_records *r);        /* device pointer to n * words uint64 */
void sb200_records_free(sb200_records *r);
int  sb200_count_records(sb200_ctx *ctx, sb200_records *r /* consumed */, unsigned num_buckets, int want_counts, sb200_kmers **out);
/* The same when every record lies in the buckets [first_bucket, first_bucket + n_owned) — the owner's share after the exchange: the
 * grouping key then spends its bits on the value prefix instead of buckets this GPU does not own (same result, fewer oversize groups). */
int  sb200_count_records_owned(sb200_ctx *ctx, sb200_records *r /* consumed */, unsigned num_buckets, unsigned first_bucket, unsigned n_owned,
                               int want_counts, sb200_kmers **out);
int  sb200_mphf_build_sharded(sb200_ctx *ctx, const sb200_kmers *local_kmers, const uint64_t *global_bucket_sizes /* B */, sb200_mphf **out);
int  sb200_mphf_arrays(const sb200_mphf *m, uint64_t **bits, uint64_t *n_words, uint64_t **ranks, uint64_t *n_ranks);   /* device */
/* Call once the arrays above hold the sum over all GPUs: builds the lookup acceleration table (set bits before every word) that a
 * whole-table build on one GPU gets for free. */
int  sb200_mphf_complete(sb200_ctx *ctx, sb200_mphf *m);
int  sb200_ext_masks_device(const sb200_ext *e, uint8_t **masks, uint64_t *size_padded);                                /* device */
int  sb200_unitigs_extract_local(sb200_ctx *ctx, const sb200_kmers *local_kmers, const sb200_mphf *mphf, const sb200_ext *ext,
                                 uint64_t *stats /* 6: chain vertices, long chains, edges, kept, bases, non-junction k-mers */,
                                 sb200_unitigs **out /* NULL if a chain exceeded the walk limit */);
int  sb200_unitigs_device(const sb200_unitigs *u, uint64_t **words, uint64_t **word_off, uint32_t **len);               /* device */

/* ---- the hash-sharded path as ONE call per rank: host code in C++ (csrc/shard.cu, csrc/comm.cu).  The two record exchanges are stores
 *      of the grouping kernel into the owners' receive buffers over NVLink (peer memory: CUDA IPC mappings between processes, plain
 *      pointers inside one); the slice exchanges of index and masks and the final gather go through the NCCL C API.
 *      A communicator is either NCCL (one rank per GPU; rank 0 makes the id, the launcher hands its 128 bytes to the others) or
 *      "local" (G virtual ranks = G threads of one process, any placement of contexts on GPUs — the same orchestration on a one-GPU
 *      box).  Every rank calls sb200_construct_sharded with ITS slice of the reads; the result holds this rank's shards of both k-mer
 *      tables, the whole KMerIndex and mask array, and its slice of the unitigs (rank order = the reference's order); gather_to >= 0
 *      also moves the packed unitigs to that rank.  Perfect loops and chains beyond the direct-walk limit are extracted by rank 0 for
 *      everybody (info.whole_set_fallback).  No reference counterpart: SPAdes shuffles through kmers_raw<i> files
 *      (kmer_mph/kmer_splitter.hpp:140-161). -------------------------------------------------------------------------------------- */
typedef struct sb200_comm sb200_comm;
typedef struct sb200_shard sb200_shard;
int  sb200_comm_unique_id(uint8_t *id128);                       /* ncclGetUniqueId; nonzero when libnccl.so.2 cannot be loaded        */
const char *sb200_comm_last_error(void);
int  sb200_comm_create_nccl(sb200_ctx *ctx, int rank, int world, const uint8_t *id128, sb200_comm **out);
int  sb200_comm_create_local(int world, sb200_comm **out /* world handles, rank r = out[r] */);
int  sb200_comm_rank(const sb200_comm *c);
int  sb200_comm_size(const sb200_comm *c);
void sb200_comm_free(sb200_comm *c);
typedef struct {
    int rank, size;
    int whole_set_fallback;      /* rank 0 extracted all unitigs (long chains / perfect loops): the other ranks' slices are empty      */
    int gathered;                /* this rank's unitig handle holds the slices of all ranks                                           */
    uint64_t total_kpomers, total_kmers, total_instances, total_unitigs, total_unitig_bases, n_loops, clipped;
    uint64_t bytes_sent;         /* bytes this rank handed to other ranks (NVLink roofline)                                           */
    double exchange_ms;          /* device time of the two record exchanges: the pass-1 kernels that store into the owners' buffers
                                    over NVLink (or, without peer mappings, the two all-to-alls)                                      */
    double stage_ms[8];          /* count_kpomers, count_kmers, mphf, masks, tipclip, unitigs, gather, total (host wall, stages block) */
    uint64_t record_bytes;       /* the part of bytes_sent that the two record exchanges moved (what exchange_ms timed)               */
} sb200_shard_info_t;
struct sb200_construct_params_s;
int  sb200_construct_sharded(sb200_ctx *ctx, sb200_comm *comm, const sb200_reads *my_reads, const struct sb200_construct_params_s *params,
                             int gather_to /* rank, or -1: leave the unitigs sharded */, sb200_shard **out);
const sb200_kmers   *sb200_shard_kpomers(const sb200_shard *s);   /* borrowed handles: valid until sb200_shard_free               */
const sb200_kmers   *sb200_shard_kmers(const sb200_shard *s);
const sb200_mphf    *sb200_shard_mphf(const sb200_shard *s);
const sb200_ext     *sb200_shard_ext(const sb200_shard *s);
const sb200_unitigs *sb200_shard_unitigs(const sb200_shard *s);
int  sb200_shard_info(const sb200_shard *s, sb200_shard_info_t *info);
int  sb200_shard_walk_stats(const sb200_shard *s, uint64_t *out /* size x 6, see sb200_unitigs_extract_local */);
void sb200_shard_free(sb200_shard *s);

/* ---- whole path, host buffers in / host buffers out: what spades-gbuilder does between read conversion and output
 *      (projects/gbuilder/main.cpp:165-181) and spades-core's Construction stage (stages/construction.cpp:469-483).
 *      Result buffers are pinned host memory owned by the graph handle. ---------------------------------------------- */
typedef struct sb200_graph sb200_graph;
typedef struct sb200_construct_params_s {
    unsigned k;                 /* odd, 1 <= k < 128                                                                   */
    unsigned num_buckets;       /* 10 * nthreads in the reference                                                      */
    int      tip_clip;          /* run EarlyTipClipper (spades-core, !gap_closer)                                      */
    uint64_t tip_length_bound;  /* RL - k (stages/construction.cpp:300-304)                                            */
    int      with_loops;        /* keep_perfect_loops                                                                  */
    int      fetch_kmers;       /* copy final_kmers + (k+1)-mers + counts back as well                                 */
} sb200_construct_params;
typedef struct {
    uint64_t n_kpomers, n_kmers, n_unitigs, n_loops, unitig_bases, n_unitig_words, kpomer_instances, clipped;
    const uint64_t *kpomers; const uint32_t *kpomer_counts; const uint64_t *kpomer_bucket_starts;   /* fetch_kmers */
    const uint64_t *kmers;   const uint64_t *kmer_bucket_starts;                                     /* fetch_kmers */
    const uint8_t  *masks;                 /* n_kmers, MPHF-index order                                               */
    const uint8_t  *index_bytes; uint64_t index_size;   /* KMerIndex::serialize                                      */
    const uint64_t *unitig_words; const uint64_t *unitig_word_off; const uint32_t *unitig_len;
    uint64_t h2d_bytes, d2h_bytes;
} sb200_graph_view;
int  sb200_construct(sb200_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len,
                     uint64_t n_reads, const sb200_construct_params *params, sb200_graph **out);
int  sb200_graph_get(const sb200_graph *g, sb200_graph_view *view);
void sb200_graph_free(sb200_graph *g);

/* ---- several GPUs behind ONE host thread: what SURVEY.md 8(b) writes as sb200_create(n_gpus, device_ids).  The library drives one
 *      rank per entry of device_ids from its own host threads (NCCL between distinct devices; a device id that repeats gives virtual
 *      ranks with the local communicator, for one-GPU boxes), splits the host reads by index, runs sb200_construct_sharded on every rank
 *      and returns ONE graph in the layout of sb200_construct: the shards concatenated in rank order = the single-GPU result (every GPU
 *      copies its shard of the k-mer tables home over its own PCIe link; masks, index and unitigs come from rank 0).  The graph is
 *      released with sb200_graph_free before sb200_multi_destroy. ---------------------------------------------------------------------- */
typedef struct sb200_multi sb200_multi;
int  sb200_multi_create(int n_gpus, const int *device_ids, sb200_multi **out);
void sb200_multi_destroy(sb200_multi *m);
const char *sb200_multi_last_error(const sb200_multi *m);     /* m may be NULL: error of the last failed sb200_multi_create         */
int  sb200_multi_size(const sb200_multi *m);
sb200_ctx *sb200_multi_context(sb200_multi *m, int rank);     /* borrowed: e.g. sb200_mphf_lookup on rank 0 for the link records      */
int  sb200_multi_construct(sb200_multi *m, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                           const sb200_construct_params *params, sb200_graph **out);

#ifdef __cplusplus
}
#endif
#endif

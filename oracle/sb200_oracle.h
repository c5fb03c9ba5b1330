/* oracle/sb200_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the SPAdes 3.15.4 graph-construction front end (reads -> canonical (k+1)-mers ->
 * dedup -> k-mers -> BooPHF MPHF -> in/out extension masks -> [tip clipping] -> unbranching paths).  Only tests/,
 * bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load this library; the product
 * library (libspades_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked byte-for-byte against the unmodified reference compiled
 * from /root/reference by oracle/Makefile (oracle/_ref/ref_driver) — see tests/test_oracle_vs_reference.py — and
 * against the committed fixtures under tests/golden/ (generated from the reference by tests/golden/make_golden.py).
 *
 * Paths in the citations are relative to /root/reference/assembler: C/ = src/common/, E/ = ext/include/.
 */
#ifndef SB200_ORACLE_H
#define SB200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- primitives ---------------------------------------------------------------------------------------------- */
uint64_t ora_xxh3_64(const uint64_t *words, unsigned nwords);                       /* E/xxh/xxhash.h:2781-2909, seed 0 */
void     ora_xxh3_128(const uint64_t *words, unsigned nwords, uint64_t *hi, uint64_t *lo); /* xxhash.h:4271-4432 */
void     ora_kmer_rc(const uint64_t *x, unsigned k, uint64_t *out);                 /* C/sequence/rtseq.hpp:79-115 */
int      ora_kmer_is_minimal(const uint64_t *x, unsigned k);                        /* rtseq.hpp:407-415         */
unsigned ora_bucket(const uint64_t *x, unsigned nwords, unsigned num_buckets);      /* C/utils/kmer_mph/kmer_buckets.hpp:28-41 */

/* ---- read packing (C/sequence/sequence.hpp:71-122 layout; C/io/reads/longest_valid_wrapper.hpp:15-52) ----------- */
/* reads: concatenated ASCII, read i = ascii[off[i] .. off[i+1]).  Each read is replaced by its longest ACGT run
 * (first on ties) and packed 2 bits per base, base j at bits 2(j%32) of word j/32, every read starting on a word
 * boundary.  words_out may be NULL to query the number of words needed (returned). */
uint64_t ora_pack_reads(const char *ascii, const uint64_t *off, uint64_t n_reads,
                        uint64_t *words_out, uint64_t *word_off_out /* n_reads+1 */, uint32_t *len_out);

/* ---- k-mer sets ------------------------------------------------------------------------------------------------ */
typedef struct ora_kmers ora_kmers;
/* C/utils/kmer_mph/kmer_splitters.hpp:25-41,109-133 + kmer_splitter.hpp:120-167 + kmer_index_builder.hpp:241-365.
 * add_rc: also stream rc(read) (RCWrap).  canonical_only: keep a window iff IsMinimal (StoringTypeFilter).
 * counts[i] = number of kept window instances equal to record i (== coverage_hash_map_builder.hpp:15-38 when
 * canonical_only=add_rc=1).  Returns NULL if no k-mer was kept (the reference FATALs, kmer_index_builder.hpp:261). */
ora_kmers *ora_count_reads(const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                           unsigned K, int canonical_only, int add_rc, unsigned num_buckets);
/* kmer_splitters.hpp:159-204 with K_target = K_source-1, filter = IsMinimal, add_rc = true. */
ora_kmers *ora_derive_kmers(const ora_kmers *kpomers, unsigned num_buckets);
unsigned        ora_kmers_k(const ora_kmers *);
unsigned        ora_kmers_words(const ora_kmers *);
unsigned        ora_kmers_num_buckets(const ora_kmers *);
uint64_t        ora_kmers_size(const ora_kmers *);
const uint64_t *ora_kmers_data(const ora_kmers *);           /* size*words, file order                   */
const uint64_t *ora_kmers_bucket_starts(const ora_kmers *);  /* num_buckets+1                            */
const uint32_t *ora_kmers_counts(const ora_kmers *);         /* size (NULL for derived sets)             */
void            ora_kmers_free(ora_kmers *);

/* ---- MPHF (E/boomphf/BooPHF.h:409-441,575-675; kmer_index_builder.hpp:383-433) ---------------------------------- */
typedef struct ora_mphf ora_mphf;
ora_mphf *ora_mphf_build(const ora_kmers *);
uint64_t  ora_mphf_lookup(const ora_mphf *, const uint64_t *rec);   /* KMerIndex::seq_idx, kmer_index.hpp:85-90 */
uint64_t  ora_mphf_final_level_keys(const ora_mphf *);             /* keys that reached the exact-map level    */
/* KMerIndex::serialize bytes (kmer_index.hpp:99-105 + BooPHF.h:514-532,316-323); out==NULL -> size only.
 * Buckets with 0 keys serialise 25 empty bit-vectors whose _nchar field is uninitialised memory in the
 * reference (BooPHF.h:138-141); the oracle writes 0 there. */
uint64_t  ora_mphf_serialize(const ora_mphf *, uint8_t *out);
void      ora_mphf_free(ora_mphf *);

/* ---- extension masks (C/utils/extension_index/kmer_extension_index_builder.hpp:44-59; kmer_extension_index.hpp:92-106) */
void ora_fill_masks(const ora_kmers *kpomers, const ora_mphf *, uint8_t *data /* U0 bytes, zeroed by callee */);

/* ---- early tip clipper (C/assembly_graph/construction/early_simplification.hpp:20-152), single-thread order ----- */
uint64_t ora_tipclip(const ora_kmers *kmers, const ora_mphf *, uint8_t *data, uint64_t length_bound);

/* ---- unbranching paths + loops (C/assembly_graph/construction/debruijn_graph_constructor.hpp:193-384) ----------- */
typedef struct ora_seqs ora_seqs;
ora_seqs       *ora_unitigs(const ora_kmers *kmers, const ora_mphf *, uint8_t *data /* mutated: isolated */,
                            int with_loops);
uint64_t        ora_seqs_count(const ora_seqs *);
uint64_t        ora_seqs_n_loops(const ora_seqs *);
const uint64_t *ora_seqs_offsets(const ora_seqs *);   /* count+1 offsets into chars */
const char     *ora_seqs_chars(const ora_seqs *);     /* ACGT                      */
void            ora_seqs_free(ora_seqs *);

#ifdef __cplusplus
}
#endif
#endif

// shard_api.cu — C ABI of the hash-sharded (multi-GPU) path: the single-GPU stages with the shuffle points exposed.
//
// One process drives one GPU (torch.distributed / NCCL moves the bytes, spades_for_blackbird_b200/host/distributed.py).
// GPU g of G owns the buckets [g*B/G, (g+1)*B/G) of the reference's bucket function (kmer_buckets.hpp:28-41), i.e. a
// contiguous range of the reference's file order, so concatenating the shards in rank order IS the single-GPU result.
#include <string.h>

#include "../../include/sb200.h"
#include "common.cuh"
#include "kmer_set.cuh"

namespace sb200 {
sb200_records *extract_records(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc);
sb200_records *extract_records_partitioned(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc, unsigned B, unsigned G,
                                           uint64_t *counts_out);
sb200_records *derive_records(sb200_ctx *ctx, const sb200_kmers *kp);
void partition_records(sb200_ctx *ctx, sb200_records *r, unsigned B, unsigned n_parts, uint64_t *counts_out);
sb200_kmers *count_records(sb200_ctx *ctx, sb200_records *r, unsigned B, int want_counts, unsigned first_bucket, unsigned n_owned);
sb200_mphf *mphf_build(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes);
void mphf_complete(sb200_ctx *ctx, sb200_mphf *m);
sb200_unitigs *extract_unitigs_local(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, uint64_t *stats);
}  // namespace sb200

template<class F>
static int guarded(sb200_ctx *ctx, F &&f) {
    try {
        CUDA_CHECK(cudaSetDevice(ctx->device));
        f();
        return 0;
    } catch (const sb200_error &e) {
        ctx->last_error = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::exception &e) {
        ctx->last_error = e.what();
        return 3;
    }
}

extern "C" {

int sb200_records_extract(sb200_ctx *ctx, const sb200_reads *reads, unsigned K, int canonical_only, int add_rc, sb200_records **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::extract_records(ctx, reads, K, canonical_only, add_rc); });
}
int sb200_records_extract_partitioned(sb200_ctx *ctx, const sb200_reads *reads, unsigned K, int canonical_only, int add_rc, unsigned num_buckets,
                                      unsigned n_owners, uint64_t *counts_out, sb200_records **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::extract_records_partitioned(ctx, reads, K, canonical_only, add_rc, num_buckets, n_owners, counts_out); });
}
int sb200_records_derive(sb200_ctx *ctx, const sb200_kmers *kpomers, sb200_records **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::derive_records(ctx, kpomers); });
}
int sb200_records_partition(sb200_ctx *ctx, sb200_records *r, unsigned num_buckets, unsigned n_owners, uint64_t *counts_out) {
    return guarded(ctx, [&] { sb200::partition_records(ctx, r, num_buckets, n_owners, counts_out); });
}
int sb200_records_alloc(sb200_ctx *ctx, uint64_t n, unsigned K, int flags, sb200_records **out) {
    *out = nullptr;
    return guarded(ctx, [&] {
        SB200_REQUIRE(K >= 1 && K <= 128, "K out of range [1,128]");
        sb200_records *r = new sb200_records();
        r->ctx = ctx; r->k = K; r->words = (K + 31) / 32; r->n = n;
        r->double_palindromes = (flags & 1) != 0; r->marker = (flags & 2) != 0; r->mask_payload = (flags & 4) != 0;
        r->data.alloc(ctx, n * r->words);
        *out = r;
    });
}
uint64_t sb200_records_size(const sb200_records *r) { return r->n; }
unsigned sb200_records_words(const sb200_records *r) { return r->words; }
unsigned sb200_records_k(const sb200_records *r) { return r->k; }
int sb200_records_flags(const sb200_records *r) { return (r->double_palindromes ? 1 : 0) | (r->marker ? 2 : 0) | (r->mask_payload ? 4 : 0); }
uint64_t *sb200_records_device(sb200_records *r) { return r->data.p; }
void sb200_records_free(sb200_records *r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    delete r;
}
int sb200_count_records(sb200_ctx *ctx, sb200_records *r, unsigned num_buckets, int want_counts, sb200_kmers **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::count_records(ctx, r, num_buckets, want_counts, 0, 0); });
}
int sb200_count_records_owned(sb200_ctx *ctx, sb200_records *r, unsigned num_buckets, unsigned first_bucket, unsigned n_owned, int want_counts,
                              sb200_kmers **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::count_records(ctx, r, num_buckets, want_counts, first_bucket, n_owned); });
}

int sb200_mphf_build_sharded(sb200_ctx *ctx, const sb200_kmers *local_kmers, const uint64_t *global_bucket_sizes, sb200_mphf **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::mphf_build(ctx, local_kmers, global_bucket_sizes); });
}
int sb200_mphf_arrays(const sb200_mphf *m, uint64_t **bits, uint64_t *n_words, uint64_t **ranks, uint64_t *n_ranks) {
    *bits = m->bits.p; *n_words = m->total_words; *ranks = m->ranks.p; *n_ranks = m->total_ranks;
    return 0;
}
int sb200_mphf_complete(sb200_ctx *ctx, sb200_mphf *m) {
    return guarded(ctx, [&] { sb200::mphf_complete(ctx, m); });
}
int sb200_ext_masks_device(const sb200_ext *e, uint8_t **masks, uint64_t *size_padded) {
    *masks = e->masks.p; *size_padded = (e->size + 3) & ~3ULL;
    return 0;
}
int sb200_unitigs_extract_local(sb200_ctx *ctx, const sb200_kmers *local_kmers, const sb200_mphf *mphf, const sb200_ext *ext, uint64_t *stats,
                                sb200_unitigs **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::extract_unitigs_local(ctx, local_kmers, mphf, ext, stats); });
}
int sb200_unitigs_device(const sb200_unitigs *u, uint64_t **words, uint64_t **word_off, uint32_t **len) {
    *words = u->words.p; *word_off = u->word_off.p; *len = u->len.p;
    return 0;
}

}  // extern "C"

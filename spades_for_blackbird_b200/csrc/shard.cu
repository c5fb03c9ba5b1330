// shard.cu — the hash-sharded path: one rank per GPU, the k-mer space split by the reference's own bucket function.
//
// Rank g of G owns the buckets [g*B/G, (g+1)*B/G) of KMerSegmentPolicy (C/utils/kmer_mph/kmer_buckets.hpp:28-41) — a contiguous
// range of the reference's file order — so the shards concatenated in rank order ARE the single-GPU (= reference) result.  The
// reference does the same shuffle through kmers_raw<i> files (kmer_splitter.hpp:140-161); here the shuffle is part of the grouping kernels:
//   1. reads are split by index; a count pass over the (owner, coarse bin) bins and one tiny all-gather fix every size
//   2. pass 1 of the grouping stores every owner's runs of canonical (k+1)-mer instances straight into the OWNER's receive buffer over
//      NVLink (peer_exchange below; peer memory mapped once, sb200_comm::peer_buffers); a barrier; the owners group / deduplicate / count:
//      their shard of the (k+1)-mer KMerDiskStorage
//   3. owners derive k-mer candidates (mask bit in the padding, or a byte beside the record)
//   4. the same peer-store pass 1 by the owner of the K-MER -> owners deduplicate + OR the mask bits: their shard of final_kmers
//      (without peer mappings, or with SB200_NO_PEER_STORES=1: send buffers + two NCCL all-to-alls, exchange_and_receive)
//   5. all-gather of bucket sizes (tiny)     -> segment starts / level geometry of the whole KMerIndex on every rank
//   6. every rank builds the BooPHF levels of its own buckets inside the global layout (only its own slice of the arrays is touched) and
//      moves its k-mers' mask bytes to MPHF order at once; an owner's bit-vectors, rank samples and MPHF indices are CONTIGUOUS ranges,
//      so index and masks are completed by ONE exchange of slices (no zero-padded all-reduce)
//   7. only a k-mer set without the mask payload falls back to lookups + an all-reduce (decided jointly)
//   8. [EarlyTipClipper: find over the own junctions, kill lists OR-ed over the ranks, forward links of the own junctions, slices again]
//   9. every rank walks the start edges of the junctions in its own shard through the walk blocks of the global index + masks: its slice of the unitig
//      list, already in the reference's global order.  Chains beyond the direct-walk limit or perfect loops (no junction to start from)
//      send the extraction to rank 0, which gathers the k-mer shards and runs the whole-set path (pointer jumping + CollectLoops,
//      debruijn_graph_constructor.hpp:248-265,308-344) — complete for every input, at one GPU's speed for that stage;
//  10. optional gather of the packed unitig slices to rank 0 in rank order (= reference order).
// Host code is C++ (NCCL C API behind comm.cuh); Python only launches ranks and hands the NCCL id around.
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>

#include "../../include/sb200.h"
#include "comm.cuh"
#include "common.cuh"
#include "graph.cuh"
#include "scan.cuh"
#include "kmer_set.cuh"

namespace sb200 {
sb200_records *extract_records_partitioned(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc, unsigned B, unsigned G,
                                           uint64_t *counts_out);
sb200_records *derive_records(sb200_ctx *ctx, const sb200_kmers *kp);
void partition_records(sb200_ctx *ctx, sb200_records *r, unsigned B, unsigned n_parts, uint64_t *counts_out);
sb200_kmers *count_records(sb200_ctx *ctx, sb200_records *r, unsigned B, int want_counts, unsigned first_bucket, unsigned n_owned);
bool shard_plan(sb200_ctx *ctx, unsigned W, uint64_t n_owner_est, unsigned B, unsigned G, unsigned K, ShardPlan *pl);
sb200_records *shard_send_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts);
sb200_records *shard_send_derive(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts);
ShardPeerSend *shard_peer_count_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts);
ShardPeerSend *shard_peer_count_derive(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B, unsigned G, const ShardPlan *pl, uint64_t *owner_counts);
void shard_peer_scatter_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned B, unsigned G, ShardPeerSend *st, const SpPeers *peers);
void shard_peer_scatter_derive(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B, unsigned G, ShardPeerSend *st, const SpPeers *peers);
void shard_peer_send_counts(sb200_ctx *ctx, unsigned G, const ShardPeerSend *st, uint32_t *const *dst);
sb200_kmers *shard_receive(sb200_ctx *ctx, sb200_records *got, const uint64_t *run_start, unsigned G, const ShardPlan *pl, unsigned B, unsigned first_bucket,
                           unsigned n_owned, int want_counts);
sb200_mphf *mphf_build(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes);
sb200_mphf *mphf_build_slice(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes);
void mphf_complete(sb200_ctx *ctx, sb200_mphf *m);
sb200_ext *build_ext(sb200_ctx *ctx, const sb200_kmers *kpomers, const sb200_kmers *kmers, const sb200_mphf *mphf);
sb200_ext *build_ext_from_masks(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const uint8_t *masks_dev);
void tipclip_find(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, uint64_t bound, TipClipState &st);
void tipclip_apply(sb200_ctx *ctx, sb200_ext *ext, TipClipState &st);
uint64_t tipclip_links(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, TipClipState &st);
sb200_unitigs *extract_unitigs_local(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, uint64_t *stats);
sb200_unitigs *extract_unitigs(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops);
}  // namespace sb200

struct sb200_shard {
    sb200_ctx *ctx = nullptr;
    int rank = 0, size = 1;
    sb200_kmers *kpomers = nullptr, *kmers = nullptr;   // this rank's shards
    sb200_mphf *mphf = nullptr;                         // the whole index
    sb200_ext *ext = nullptr;                           // masks of the whole index, idx of the own k-mers
    sb200_unitigs *unitigs = nullptr;                   // own slice (or everything on rank 0, see whole_set_fallback / gathered)
    std::vector<uint64_t> walk_stats;                   // size x 6 (sb200_unitigs_extract_local's stats of every rank)
    std::vector<uint64_t> unitig_counts;                // sequences per rank before any gather
    uint64_t clipped = 0, total_kpomers = 0, total_kmers = 0, total_instances = 0, total_unitigs = 0, total_unitig_bases = 0, n_loops = 0;
    bool whole_set_fallback = false, gathered = false;
    double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // count_kpomers, count_kmers, mphf, masks, tipclip, unitigs, gather, total
    uint64_t bytes_sent = 0, record_bytes = 0;
    double exchange_ms = 0;
    ~sb200_shard() {
        delete unitigs; delete ext; delete mphf; delete kmers; delete kpomers;
    }
};

namespace sb200 {

static sb200_records *alloc_records(sb200_ctx *ctx, uint64_t n, const sb200_records *like) {
    sb200_records *r = new sb200_records();
    r->ctx = ctx; r->k = like->k; r->words = like->words; r->n = n;
    r->double_palindromes = like->double_palindromes; r->marker = like->marker; r->mask_payload = like->mask_payload;
    r->data.alloc(ctx, n * r->words);
    return r;
}

// records grouped by owner (counts[g] of them for owner g) -> exchange -> this rank's shard, grouped / deduplicated / counted
static sb200_kmers *exchange_and_count(sb200_ctx *ctx, sb200_comm *cm, std::unique_ptr<sb200_records> rec, const uint64_t *counts, unsigned B,
                                       bool want_counts) {
    const int G = cm->size;
    std::vector<uint64_t> all((size_t) G * G);
    cm->all_gather_host(ctx, counts, (size_t) G, all.data());
    const uint64_t rb = (uint64_t) rec->words * 8;
    std::vector<uint64_t> send_off((size_t) G + 1, 0), recv_off((size_t) G + 1, 0);
    for (int g = 0; g < G; ++g) {
        send_off[(size_t) g + 1] = send_off[(size_t) g] + counts[g] * rb;
        recv_off[(size_t) g + 1] = recv_off[(size_t) g] + all[(size_t) g * G + cm->rank] * rb;
    }
    const uint64_t n_recv = recv_off[(size_t) G] / rb;
    SB200_REQUIRE(n_recv < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: use more GPUs");
    std::unique_ptr<sb200_records> got(alloc_records(ctx, n_recv, rec.get()));
    cm->all_to_all_v(ctx, rec->data.p, send_off.data(), got->data.p, recv_off.data());
    rec.reset();   // the exchange has completed (all_to_all_v is blocking)
    const unsigned n_owned = B / (unsigned) G;
    return count_records(ctx, got.get(), B, want_counts ? 1 : 0, (unsigned) cm->rank * n_owned, n_owned);
}

// The same on the staged kernels (count.cu shard_send_* / shard_receive): `rec` is grouped by (owner, coarse bin) and carries the run sizes
// (and, for k-mers that fill their last word, the mask bits as bytes beside the records); all of it travels, and the owner's pass 2 takes its
// tiles straight from the receive buffer.
static sb200_kmers *exchange_and_receive(sb200_ctx *ctx, sb200_comm *cm, std::unique_ptr<sb200_records> rec, const uint64_t *counts, unsigned B,
                                         const ShardPlan &pl, bool want_counts) {
    const int G = cm->size;
    std::vector<uint64_t> all((size_t) G * G);
    cm->all_gather_host(ctx, counts, (size_t) G, all.data());
    const uint64_t rb = (uint64_t) rec->words * 8;
    std::vector<uint64_t> send_off((size_t) G + 1, 0), recv_off((size_t) G + 1, 0), run_start((size_t) G + 1, 0);
    std::vector<uint64_t> psend((size_t) G + 1, 0), precv((size_t) G + 1, 0), csend((size_t) G + 1, 0), crecv((size_t) G + 1, 0);
    for (int g = 0; g < G; ++g) {
        const uint64_t out = counts[g], in = all[(size_t) g * G + cm->rank];
        send_off[(size_t) g + 1] = send_off[(size_t) g] + out * rb;
        recv_off[(size_t) g + 1] = recv_off[(size_t) g] + in * rb;
        psend[(size_t) g + 1] = psend[(size_t) g] + out;
        precv[(size_t) g + 1] = precv[(size_t) g] + in;
        run_start[(size_t) g + 1] = run_start[(size_t) g] + in;
        csend[(size_t) g + 1] = csend[(size_t) g] + (uint64_t) pl.n_co * 4;
        crecv[(size_t) g + 1] = crecv[(size_t) g] + (uint64_t) pl.n_co * 4;
    }
    const uint64_t n_recv = run_start[(size_t) G];
    SB200_REQUIRE(n_recv < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: use more GPUs");
    std::unique_ptr<sb200_records> got(alloc_records(ctx, n_recv, rec.get()));
    got->coarse_counts.alloc(ctx, (uint64_t) G * pl.n_co);
    cm->all_to_all_v(ctx, rec->data.p, send_off.data(), got->data.p, recv_off.data());
    cm->all_to_all_v(ctx, rec->coarse_counts.p, csend.data(), got->coarse_counts.p, crecv.data());
    if (rec->pay.p) {
        got->pay.alloc(ctx, n_recv + 1);
        cm->all_to_all_v(ctx, rec->pay.p, psend.data(), got->pay.p, precv.data());
    }
    rec.reset();
    const unsigned n_owned = B / (unsigned) G;
    return shard_receive(ctx, got.get(), run_start.data(), (unsigned) G, &pl, B, (unsigned) cm->rank * n_owned, n_owned, want_counts ? 1 : 0);
}

// A record exchange with no exchange step: after the count pass every rank knows every size (one tiny all-gather), the owners' receive
// buffers are peer-visible (sb200_comm::peer_buffers), and pass 1 of the grouping writes each owner's runs straight into the owner's HBM
// over NVLink (SpPeers, staged_partition.cuh).  What remains is a barrier.  `reads` (first exchange: (k+1)-mer instances) or `kp` (second
// exchange: k-mer candidates of the own (k+1)-mers).  *done = false: some pair of GPUs cannot map each other — the caller takes the
// all-to-all path (and keeps taking it: the communicator remembers).
static sb200_kmers *peer_exchange(sb200_ctx *ctx, sb200_comm *cm, int slot, const sb200_reads *reads, const sb200_kmers *kp, unsigned K, unsigned B,
                                  const ShardPlan &pl, bool want_counts, bool *done) {
    const int G = cm->size, me = cm->rank;
    *done = false;
    if (cm->peer_failed || ctx->no_peer_stores) return nullptr;
    std::vector<uint64_t> counts((size_t) G, 0), all((size_t) G * G);
    std::unique_ptr<ShardPeerSend> st(reads ? shard_peer_count_reads(ctx, reads, K, B, (unsigned) G, &pl, counts.data())
                                            : shard_peer_count_derive(ctx, kp, B, (unsigned) G, &pl, counts.data()));
    cm->all_gather_host(ctx, counts.data(), (size_t) G, all.data());
    const uint64_t rb = (uint64_t) st->words * 8;
    auto align256 = [](uint64_t v) { return (v + 255) & ~(uint64_t) 255; };
    const uint64_t cc_bytes = align256((uint64_t) G * pl.n_co * 4);
    uint64_t n_recv[64], dst_start[64], need[64];
    for (int g = 0; g < G; ++g) {
        n_recv[g] = 0; dst_start[g] = 0;
        for (int src = 0; src < G; ++src) {
            if (src == me) dst_start[g] = n_recv[g];
            n_recv[g] += all[(size_t) src * G + g];
        }
        SB200_REQUIRE(n_recv[g] < (1ull << 32), "more than 2^32-1 k-mer instances on one GPU: use more GPUs");
        need[g] = cc_bytes + align256(n_recv[g] * rb) + (st->side_pay ? align256(n_recv[g] + 4) : 0);
    }
    void *ptrs[64];
    if (!cm->peer_buffers(ctx, slot, need, ptrs)) return nullptr;
    SpPeers peers;
    memset(&peers, 0, sizeof peers);
    peers.n_co = pl.n_co;
    uint32_t *cdst[64];
    uint64_t out_bytes = 0;
    for (int g = 0; g < G; ++g) {
        uint8_t *base = (uint8_t *) ptrs[g];
        cdst[g] = reinterpret_cast<uint32_t *>(base) + (uint64_t) me * pl.n_co;
        peers.rec[g] = reinterpret_cast<uint64_t *>(base + cc_bytes + dst_start[g] * rb);
        peers.pay[g] = st->side_pay ? base + cc_bytes + align256(n_recv[g] * rb) + dst_start[g] : nullptr;
        if (g != me) out_bytes += counts[(size_t) g] * (rb + (st->side_pay ? 1 : 0)) + (uint64_t) pl.n_co * 4;
    }
    cudaEvent_t e0 = ctx->get_event(), e1 = ctx->get_event();
    CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
    shard_peer_send_counts(ctx, (unsigned) G, st.get(), cdst);
    if (reads) shard_peer_scatter_reads(ctx, reads, B, (unsigned) G, st.get(), &peers);
    else shard_peer_scatter_derive(ctx, kp, B, (unsigned) G, st.get(), &peers);
    CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(e1));   // my stores have been performed at their owners
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ctx->event_pool.push_back(e0);
    ctx->event_pool.push_back(e1);
    cm->bytes_sent += out_bytes;
    cm->record_bytes += out_bytes;
    cm->exchange_ms += ms;
    cm->barrier(ctx);                        // ... and everybody else's at mine
    std::unique_ptr<sb200_records> got(new sb200_records());
    got->ctx = ctx; got->k = st->k; got->words = st->words; got->n = n_recv[me];
    got->double_palindromes = st->double_palindromes; got->mask_payload = st->mask_payload;
    uint8_t *mine = (uint8_t *) ptrs[me];
    got->coarse_counts.borrow(ctx, reinterpret_cast<uint32_t *>(mine), (size_t) G * pl.n_co);
    got->data.borrow(ctx, reinterpret_cast<uint64_t *>(mine + cc_bytes), (size_t) (n_recv[me] * st->words));
    if (st->side_pay) got->pay.borrow(ctx, mine + cc_bytes + align256(n_recv[me] * rb), (size_t) n_recv[me] + 1);
    std::vector<uint64_t> run_start((size_t) G + 1, 0);
    for (int src = 0; src < G; ++src) run_start[(size_t) src + 1] = run_start[(size_t) src] + all[(size_t) src * G + me];
    st.reset();
    *done = true;
    const unsigned n_owned = B / (unsigned) G;
    return shard_receive(ctx, got.get(), run_start.data(), (unsigned) G, &pl, B, (unsigned) me * n_owned, n_owned, want_counts ? 1 : 0);
}

static double now_ms() { return sb200_ctx::now_s() * 1e3; }

static sb200_shard *construct_sharded(sb200_ctx *ctx, sb200_comm *cm, const sb200_reads *reads, const sb200_construct_params *p, int gather_to) {
    const int G = cm->size, me = cm->rank;
    const unsigned B = p->num_buckets, k = p->k;
    SB200_REQUIRE((k & 1) && k >= 1 && k < 128, "k must be odd and in [1,128)");
    SB200_REQUIRE(B >= 1 && B <= 65536 && B % (unsigned) G == 0, "num_buckets must be a multiple of the number of GPUs");
    SB200_REQUIRE(G <= 64, "at most 64 ranks");
    std::unique_ptr<sb200_shard> res(new sb200_shard());
    res->ctx = ctx; res->rank = me; res->size = G;
    cm->bytes_sent = 0; cm->record_bytes = 0; cm->exchange_ms = 0;
    const unsigned n_owned = B / (unsigned) G, first_bucket = (unsigned) me * n_owned;
    double t0 = now_ms(), t_start = t0;
    auto lap = [&](int stage) {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const double t = now_ms();
        res->stage_ms[stage] += t - t0;
        t0 = t;
    };
    // SB200_SHARD_TRACE=1: rank 0 prints finer laps (each one synchronises the stream: a diagnostic, not the timed configuration)
    const bool trace = me == 0 && getenv("SB200_SHARD_TRACE") != nullptr;
    double tt = t0;
    auto mark = [&](const char *what) {
        if (!trace) return;
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const double t = now_ms();
        fprintf(stderr, "[sb200 shard trace] %-28s %8.3f ms\n", what, t - tt);
        tt = t;
    };

    // ---- 1-2: (k+1)-mers ------------------------------------------------------------------------------------------------------
    // every rank must take the same grouping plan: it is sized from the job's instance estimate (tiny all-gather)
    {
        uint64_t mine[2] = {reads->n_bases ? reads->n_bases : reads->n_words * 32, reads->n_reads}, all[2 * 64];
        cm->all_gather_host(ctx, mine, 2, all);
        uint64_t bases = 0, nr = 0;
        for (int g = 0; g < G; ++g) { bases += all[2 * g]; nr += all[2 * g + 1]; }
        const uint64_t shorter = nr * (uint64_t) k;
        const uint64_t n_owner_est = std::max<uint64_t>((bases > shorter ? bases - shorter : 1) / (uint64_t) G, 1);
        std::vector<uint64_t> counts((size_t) G, 0);
        ShardPlan pl;
        if (shard_plan(ctx, (k + 1 + 31) / 32, n_owner_est, B, (unsigned) G, k + 1, &pl)) {
            mark("plan (k+1)");
            bool done = false;
            if (G > 1) res->kpomers = peer_exchange(ctx, cm, 0, reads, nullptr, k + 1, B, pl, true, &done);
            if (!done) {
                std::unique_ptr<sb200_records> rec(shard_send_reads(ctx, reads, k + 1, B, (unsigned) G, &pl, counts.data()));
                mark("send side (k+1)");
                res->kpomers = exchange_and_receive(ctx, cm, std::move(rec), counts.data(), B, pl, true);
            }
            mark("exchange + receive (k+1)");
        } else {
            std::unique_ptr<sb200_records> rec(extract_records_partitioned(ctx, reads, k + 1, 1, 1, B, (unsigned) G, counts.data()));
            res->kpomers = exchange_and_count(ctx, cm, std::move(rec), counts.data(), B, true);
        }
    }
    lap(0);
    // ---- 3-4: k-mers --------------------------------------------------------------------------------------------------------------
    {
        uint64_t mine = res->kpomers->size, all[64], total = 0;
        cm->all_gather_host(ctx, &mine, 1, all);
        for (int g = 0; g < G; ++g) total += all[g];
        std::vector<uint64_t> counts((size_t) G, 0);
        ShardPlan pl;
        if (shard_plan(ctx, (k + 31) / 32, std::max<uint64_t>(2 * total / (uint64_t) G, 1), B, (unsigned) G, k, &pl)) {
            mark("plan k");
            bool done = false;
            if (G > 1) res->kmers = peer_exchange(ctx, cm, 1, nullptr, res->kpomers, k, B, pl, false, &done);
            if (!done) {
                std::unique_ptr<sb200_records> rec(shard_send_derive(ctx, res->kpomers, B, (unsigned) G, &pl, counts.data()));
                mark("send side k");
                res->kmers = exchange_and_receive(ctx, cm, std::move(rec), counts.data(), B, pl, false);
            }
            mark("exchange + receive k");
        } else {
            std::unique_ptr<sb200_records> rec(derive_records(ctx, res->kpomers));
            partition_records(ctx, rec.get(), B, (unsigned) G, counts.data());
            res->kmers = exchange_and_count(ctx, cm, std::move(rec), counts.data(), B, false);
        }
    }
    lap(1);
    // ---- 5: global sizes -------------------------------------------------------------------------------------------------------------
    std::vector<uint64_t> sizes((size_t) B, 0);   // k-mers per bucket, all buckets
    bool all_payload = true;
    {
        std::vector<uint64_t> mine((size_t) n_owned + 4, 0), all(((size_t) n_owned + 4) * G);
        mine[(size_t) n_owned + 3] = res->kmers->masks_file.p ? 1 : 0;   // (see 6-7)
        for (unsigned b = 0; b < n_owned; ++b)
            mine[b] = res->kmers->bucket_starts_host[(size_t) first_bucket + b + 1] - res->kmers->bucket_starts_host[(size_t) first_bucket + b];
        mine[n_owned] = res->kpomers->size; mine[(size_t) n_owned + 1] = res->kpomers->instances; mine[(size_t) n_owned + 2] = res->kmers->size;
        cm->all_gather_host(ctx, mine.data(), mine.size(), all.data());
        for (int g = 0; g < G; ++g) {
            const uint64_t *row = all.data() + (size_t) g * mine.size();
            for (unsigned b = 0; b < n_owned; ++b) sizes[(size_t) g * n_owned + b] = row[b];
            res->total_kpomers += row[n_owned]; res->total_instances += row[(size_t) n_owned + 1]; res->total_kmers += row[(size_t) n_owned + 2];
            all_payload = all_payload && row[(size_t) n_owned + 3];
        }
        SB200_REQUIRE(res->total_kmers > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
    }
    // ---- 6-7: index + masks ----------------------------------------------------------------------------------------------------------
    // Every rank builds the BooPHF levels of its own buckets inside the global layout (only its own slice of the arrays is touched) and
    // — the slice's ranks being final at once — moves its k-mers' mask bytes to MPHF order right away; ONE exchange then completes
    // bit-vectors, rank samples and masks on every GPU.
    // The masks a rank's k-mer sort OR-ed together are complete for its own k-mers; if ANY rank's set came without them (no padding bits
    // for this k, or its groups overflowed into the LSD path), every rank fills by lookups instead — a mix would leave holes.
    if (!all_payload) res->kmers->masks_file.release();
    mark("sizes + payload consensus");
    res->mphf = mphf_build_slice(ctx, res->kmers, sizes.data());
    sb200_mphf *m = res->mphf;
    mark("mphf build");
    std::vector<uint64_t> woff((size_t) G + 1), roff((size_t) G + 1), idx_off((size_t) G + 1, 0);
    for (int g = 0; g <= G; ++g) {
        const size_t t = (size_t) g * n_owned * 25;   // first (bucket, level) entry of rank g's buckets (mphf.cuh MPHF_LEVELS)
        woff[(size_t) g] = 8 * (g < G ? m->word_off_host[t] : m->total_words);
        roff[(size_t) g] = 8 * (g < G ? m->rank_off_host[t] : m->total_ranks);
    }
    for (int g = 0; g < G; ++g) {   // rank g's k-mers hold the MPHF indices [idx_off[g], idx_off[g + 1])
        uint64_t s = 0;
        for (unsigned b = 0; b < n_owned; ++b) s += sizes[(size_t) g * n_owned + b];
        idx_off[(size_t) g + 1] = idx_off[(size_t) g] + s;
    }
    if (all_payload) {
        res->ext = build_ext(ctx, res->kpomers, res->kmers, m);   // placement record + the slice's ranks: no lookups, no foreign data
        mark("build_ext (own slice)");
        void *bufs[3] = {m->bits.p, m->ranks.p, res->ext->masks.p};
        const uint64_t *offs[3] = {woff.data(), roff.data(), idx_off.data()};
        cm->all_gather_v_inplace_multi(ctx, 3, bufs, offs);
        mark("index + mask slices");
        mphf_complete(ctx, m);
        mark("mphf complete");
        lap(2);
    } else {
        void *bufs[2] = {m->bits.p, m->ranks.p};
        const uint64_t *offs[2] = {woff.data(), roff.data()};
        cm->all_gather_v_inplace_multi(ctx, 2, bufs, offs);
        mphf_complete(ctx, m);
        lap(2);
        res->ext = build_ext(ctx, res->kpomers, res->kmers, m);
        if (G > 1) cm->all_reduce_or_bytes(ctx, res->ext->masks.p, res->ext->size, false);
    }
    lap(3);
    // ---- 8: tip clipper --------------------------------------------------------------------------------------------------------------------
    if (p->tip_clip) {
        TipClipState st;
        tipclip_find(ctx, res->kmers, m, res->ext, p->tip_length_bound, st);
        if (G > 1) cm->all_reduce_or_bytes(ctx, st.kill.p, res->ext->size, true);
        tipclip_apply(ctx, res->ext, st);
        uint64_t removed = tipclip_links(ctx, res->kmers, m, res->ext, st), all[64];
        if (G > 1) cm->all_gather_v_inplace(ctx, res->ext->masks.p, idx_off.data());   // a junction's mask lies in its owner's slice
        cm->all_gather_host(ctx, &removed, 1, all);
        for (int g = 0; g < G; ++g) res->clipped += all[g];
        if (res->clipped) { res->ext->succ_valid = false; res->ext->masks_edited = true; }
        lap(4);
    }
    // ---- 9: unitigs -------------------------------------------------------------------------------------------------------------------------
    uint64_t stats[6] = {0, 0, 0, 0, 0, 0};
    res->unitigs = extract_unitigs_local(ctx, res->kmers, m, res->ext, stats);
    mark("local walks");
    res->walk_stats.assign((size_t) G * 6, 0);
    cm->all_gather_host(ctx, stats, 6, res->walk_stats.data());
    uint64_t chain_seen = 0, long_chains = 0;
    for (int g = 0; g < G; ++g) { chain_seen += res->walk_stats[(size_t) g * 6]; long_chains += res->walk_stats[(size_t) g * 6 + 1]; }
    const bool loops = p->with_loops && chain_seen != 2 * res->walk_stats[5];   // chain vertices no walk reached: perfect loops
    if (long_chains || loops) {
        // The direct walks cannot finish this input: rank 0 gathers the k-mer shards (rank order = file order) and runs the whole-set
        // extraction (pointer jumping for the long chains, CollectLoops for the loops); the other ranks contribute an empty slice.
        res->whole_set_fallback = true;
        delete res->unitigs;
        res->unitigs = nullptr;
        const uint64_t rb = (uint64_t) res->kmers->words * 8;
        std::vector<uint64_t> roff((size_t) G + 1, 0);
        for (int g = 0; g < G; ++g) roff[(size_t) g + 1] = idx_off[(size_t) g + 1] * rb;
        std::unique_ptr<sb200_kmers> full;
        if (me == 0) {
            full.reset(new sb200_kmers());
            full->ctx = ctx; full->k = k; full->words = res->kmers->words; full->num_buckets = B; full->size = res->total_kmers;
            full->data.alloc(ctx, full->size * full->words);
            full->bucket_starts_host.assign((size_t) B + 1, 0);
            for (unsigned b = 0; b < B; ++b) full->bucket_starts_host[(size_t) b + 1] = full->bucket_starts_host[b] + sizes[b];
            full->bucket_starts.alloc(ctx, (size_t) B + 1);
            CUDA_CHECK(cudaMemcpyAsync(full->bucket_starts.p, full->bucket_starts_host.data(), ((size_t) B + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
        cm->gather_v(ctx, res->kmers->data.p, res->kmers->size * rb, me == 0 ? full->data.p : nullptr, roff.data(), 0);
        if (me == 0) {
            std::unique_ptr<sb200_ext> fext(build_ext_from_masks(ctx, full.get(), m, res->ext->masks.p));
            res->unitigs = extract_unitigs(ctx, full.get(), m, fext.get(), p->with_loops);
        } else {
            res->unitigs = new sb200_unitigs();
            res->unitigs->ctx = ctx; res->unitigs->k = k;
            res->unitigs->word_off.alloc(ctx, 1); res->unitigs->word_off.zero();
            res->unitigs->len.alloc(ctx, 1); res->unitigs->words.alloc(ctx, 1);
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    {
        uint64_t mine[3] = {res->unitigs->count, res->unitigs->total_bases, res->unitigs->n_loops}, all[3 * 64];
        cm->all_gather_host(ctx, mine, 3, all);
        res->unitig_counts.assign((size_t) G, 0);
        for (int g = 0; g < G; ++g) {
            res->unitig_counts[(size_t) g] = all[3 * g];
            res->total_unitigs += all[3 * g]; res->total_unitig_bases += all[3 * g + 1]; res->n_loops += all[3 * g + 2];
        }
    }
    mark("walk consensus");
    lap(5);
    // ---- 10: gather ---------------------------------------------------------------------------------------------------------------------------
    if (gather_to >= 0 && G > 1 && !res->whole_set_fallback) {
        sb200_unitigs *u = res->unitigs;
        uint64_t mine[2] = {u->count, u->total_words}, all[2 * 64];
        cm->all_gather_host(ctx, mine, 2, all);
        std::vector<uint64_t> woff((size_t) G + 1, 0), loff((size_t) G + 1, 0), ooff((size_t) G + 1, 0);
        for (int g = 0; g < G; ++g) {
            woff[(size_t) g + 1] = woff[(size_t) g] + all[2 * g + 1] * 8;
            loff[(size_t) g + 1] = loff[(size_t) g] + all[2 * g] * 4;
            ooff[(size_t) g + 1] = ooff[(size_t) g] + all[2 * g] * 8;
        }
        std::unique_ptr<sb200_unitigs> tot;
        if (me == gather_to) {
            tot.reset(new sb200_unitigs());
            tot->ctx = ctx; tot->k = k; tot->count = loff[(size_t) G] / 4; tot->total_words = woff[(size_t) G] / 8;
            tot->total_bases = res->total_unitig_bases; tot->n_loops = res->n_loops;
            tot->words.alloc(ctx, tot->total_words + 1); tot->len.alloc(ctx, tot->count + 1); tot->word_off.alloc(ctx, tot->count + 1);
        }
        cm->gather_v(ctx, u->words.p, u->total_words * 8, tot ? tot->words.p : nullptr, woff.data(), gather_to);
        cm->gather_v(ctx, u->len.p, u->count * 4, tot ? tot->len.p : nullptr, loff.data(), gather_to);
        if (tot) {   // every sequence is word-aligned and the slices follow each other: the word offsets are a scan of the lengths (no transfer)
            void word_offsets_from_lengths(sb200_ctx *, const uint32_t *, uint64_t, uint64_t *);
            word_offsets_from_lengths(ctx, tot->len.p, tot->count, tot->word_off.p);
            delete res->unitigs;
            res->unitigs = tot.release();
            res->gathered = true;
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        lap(6);
    }
    res->stage_ms[7] = now_ms() - t_start;
    res->bytes_sent = cm->bytes_sent;
    res->record_bytes = cm->record_bytes;
    res->exchange_ms = cm->exchange_ms;
    return res.release();
}

__global__ void words_of_lengths_kernel(const uint32_t *__restrict__ len, uint64_t n, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((uint64_t) len[i] + 31) >> 5;
    else if (i == n) out[i] = 0;
}

// word_off[i] = first word of sequence i, word_off[n] = total words
void word_offsets_from_lengths(sb200_ctx *ctx, const uint32_t *len, uint64_t n, uint64_t *word_off) {
    LAUNCH(ctx, words_of_lengths_kernel, div_up(n + 1, 256), 256, 0, len, n, word_off);
    exclusive_scan<uint64_t>(ctx, word_off, n + 1, nullptr);
}

}  // namespace sb200

// ---- one host process, several GPUs: sb200_multi --------------------------------------------------------------------------------------
// SURVEY.md 8(b) sketches the boundary as sb200_create(n_gpus, device_ids): the caller of the reference's interfaces is ONE host thread
// (pipeline/stage.cpp:143-204), so the multi-GPU form of sb200_construct takes host reads, splits them by index over its GPUs, drives one
// rank per GPU from its own host threads (NCCL between distinct devices, the local communicator when a device id repeats — virtual
// ranks, for one-GPU boxes) and hands back ONE graph in the layout of the single-GPU call: the shards concatenated in rank order.
namespace sb200 {
struct MultiBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, waiting = 0;
    uint64_t gen = 0;
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g; });
    }
};

}  // namespace sb200

struct sb200_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<sb200_ctx *> ctx;
    std::vector<sb200_comm *> comm;
    std::string last_error;
    ~sb200_multi() {
        for (auto c : comm) delete c;
        for (auto c : ctx) if (c) sb200_destroy(c);
    }
};

namespace sb200 {

static void multi_create(sb200_multi *mg) {
    const int G = mg->n;
    bool repeats = false;
    for (int a = 0; a < G; ++a)
        for (int b = a + 1; b < G; ++b) repeats = repeats || mg->devices[a] == mg->devices[b];
    mg->ctx.assign((size_t) G, nullptr);
    mg->comm.assign((size_t) G, nullptr);
    for (int r = 0; r < G; ++r)
        if (sb200_create(mg->devices[r], &mg->ctx[r]) != 0) throw sb200_error(4, sb200_last_error(nullptr));
    if (repeats || G == 1) {
        comm_create_local(G, mg->comm.data());
        return;
    }
    uint8_t id[128];
    nccl_unique_id(id);
    std::vector<std::thread> th;
    std::vector<std::string> err((size_t) G);
    for (int r = 0; r < G; ++r)
        th.emplace_back([&, r] {
            try {
                mg->comm[r] = comm_create_nccl(mg->ctx[r], r, G, id);
            } catch (const std::exception &e) {
                err[r] = e.what();
            }
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < G; ++r)
        if (!err[r].empty()) throw sb200_error(5, err[r]);
}

static sb200_graph *multi_construct(sb200_multi *mg, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                                    const sb200_construct_params *p) {
    const int G = mg->n;
    std::vector<sb200_shard *> shard((size_t) G, nullptr);
    std::vector<std::string> err((size_t) G);
    MultiBarrier bar;
    bar.n = G;
    sb200_graph *g = nullptr;
    std::vector<uint64_t> kp_off((size_t) G + 1, 0), km_off((size_t) G + 1, 0);
    bool failed = false;
    auto rank_main = [&](int r) {
        sb200_ctx *ctx = mg->ctx[r];
        sb200_reads *rd = nullptr;
        bool ok = true;
        try {
            CUDA_CHECK(cudaSetDevice(ctx->device));
            const uint64_t lo = n_reads * (uint64_t) r / G, hi = n_reads * (uint64_t) (r + 1) / G;
            std::vector<uint64_t> off(hi - lo + 1);
            for (uint64_t i = lo; i <= hi; ++i) off[i - lo] = word_off[i] - word_off[lo];
            static const uint32_t zero32 = 0;
            if (sb200_reads_upload(ctx, words + word_off[lo], off.data(), hi > lo ? len + lo : &zero32, hi - lo, &rd) != 0) throw sb200_error(2, ctx->last_error);
            shard[r] = construct_sharded(ctx, mg->comm[r], rd, p, 0);
        } catch (const std::exception &e) {
            err[r] = e.what();
            mg->comm[r]->fail();
            ok = false;
        }
        if (rd) sb200_reads_free(rd);
        {
            std::lock_guard<std::mutex> lk(bar.m);
            failed = failed || !ok;
        }
        bar.wait();   // every rank has its shard (or has failed)
        if (failed) return;
        if (r == 0) {   // sizes -> offsets, host buffers from rank 0's pinned pool
            try {
                for (int q = 0; q < G; ++q) {
                    kp_off[(size_t) q + 1] = kp_off[(size_t) q] + shard[q]->kpomers->size;
                    km_off[(size_t) q + 1] = km_off[(size_t) q] + shard[q]->kmers->size;
                }
                g = new sb200_graph();
                g->ctx = ctx;
                memset(&g->view, 0, sizeof g->view);
                sb200_graph_view &v = g->view;
                const sb200_shard *s0 = shard[0];
                const unsigned W1 = s0->kpomers->words, W0 = s0->kmers->words, B = p->num_buckets;
                v.n_kpomers = kp_off[(size_t) G]; v.n_kmers = km_off[(size_t) G]; v.kpomer_instances = s0->total_instances; v.clipped = s0->clipped;
                if (p->fetch_kmers) {
                    v.kpomers = g->pin<uint64_t>(v.n_kpomers * W1 + 1);
                    v.kpomer_counts = g->pin<uint32_t>(v.n_kpomers + 1);
                    v.kmers = g->pin<uint64_t>(v.n_kmers * W0 + 1);
                    g->kp_starts.assign((size_t) B + 1, 0);
                    g->km_starts.assign((size_t) B + 1, 0);
                    const unsigned n_owned = B / (unsigned) G;
                    for (int q = 0; q < G; ++q)
                        for (unsigned b = 0; b < n_owned; ++b) {
                            const size_t gb = (size_t) q * n_owned + b;
                            g->kp_starts[gb] = kp_off[(size_t) q] + shard[q]->kpomers->bucket_starts_host[gb];
                            g->km_starts[gb] = km_off[(size_t) q] + shard[q]->kmers->bucket_starts_host[gb];
                        }
                    g->kp_starts[B] = v.n_kpomers; g->km_starts[B] = v.n_kmers;
                    v.kpomer_bucket_starts = g->kp_starts.data(); v.kmer_bucket_starts = g->km_starts.data();
                    v.d2h_bytes += v.n_kpomers * (W1 * 8 + 4) + v.n_kmers * W0 * 8;
                }
            } catch (const std::exception &e) {
                err[0] = e.what();
                std::lock_guard<std::mutex> lk(bar.m);
                failed = true;
            }
        }
        bar.wait();   // the host buffers exist
        if (failed) return;
        try {
            sb200_graph_view &v = g->view;
            const sb200_shard *s = shard[r];
            cudaStream_t st = ctx->stream;
            if (p->fetch_kmers) {   // every GPU brings its shard home over its own PCIe link
                const unsigned W1 = s->kpomers->words, W0 = s->kmers->words;
                CUDA_CHECK(cudaMemcpyAsync(const_cast<uint64_t *>(v.kpomers) + kp_off[r] * W1, s->kpomers->data.p, s->kpomers->size * W1 * 8, cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(cudaMemcpyAsync(const_cast<uint32_t *>(v.kpomer_counts) + kp_off[r], s->kpomers->counts.p, s->kpomers->size * 4, cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(cudaMemcpyAsync(const_cast<uint64_t *>(v.kmers) + km_off[r] * W0, s->kmers->data.p, s->kmers->size * W0 * 8, cudaMemcpyDeviceToHost, st));
            }
            if (r == 0) {   // masks, index bytes and the gathered unitigs live on rank 0
                uint8_t *masks = g->pin<uint8_t>(v.n_kmers + 1);
                CUDA_CHECK(cudaMemcpyAsync(masks, s->ext->masks.p, v.n_kmers, cudaMemcpyDeviceToHost, st));
                v.masks = masks;
                uint64_t mphf_serialize(const sb200_mphf *m, uint8_t *out);
                const uint64_t isz = mphf_serialize(s->mphf, nullptr);
                uint8_t *ib = g->pin<uint8_t>(isz + 8);
                mphf_serialize(s->mphf, ib);
                v.index_bytes = ib; v.index_size = isz;
                const sb200_unitigs *u = s->unitigs;
                v.n_unitigs = u->count; v.n_loops = u->n_loops; v.unitig_bases = u->total_bases; v.n_unitig_words = u->total_words;
                uint64_t *uw = g->pin<uint64_t>(u->total_words + 1), *uo = g->pin<uint64_t>(u->count + 1);
                uint32_t *ul = g->pin<uint32_t>(u->count + 1);
                CUDA_CHECK(cudaMemcpyAsync(uw, u->words.p, u->total_words * 8, cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(cudaMemcpyAsync(uo, u->word_off.p, (u->count + 1) * 8, cudaMemcpyDeviceToHost, st));
                CUDA_CHECK(cudaMemcpyAsync(ul, u->len.p, u->count * 4, cudaMemcpyDeviceToHost, st));
                v.unitig_words = uw; v.unitig_word_off = uo; v.unitig_len = ul;
                v.d2h_bytes += v.n_kmers + isz + u->total_words * 8 + (u->count + 1) * 8 + u->count * 4;
            }
            CUDA_CHECK(cudaStreamSynchronize(st));
        } catch (const std::exception &e) {
            err[r] = e.what();
            std::lock_guard<std::mutex> lk(bar.m);
            failed = true;
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < G; ++r) th.emplace_back(rank_main, r);
    for (auto &t : th) t.join();
    for (auto s : shard) if (s) { cudaSetDevice(s->ctx->device); delete s; }
    if (failed) {
        if (g) { cudaSetDevice(g->ctx->device); delete g; }
        std::string msg;
        for (int r = 0; r < G; ++r)
            if (!err[r].empty() && err[r].find("another rank of the local communicator failed") == std::string::npos) { msg = err[r]; break; }
        if (msg.empty())
            for (int r = 0; r < G; ++r) if (!err[r].empty()) { msg = err[r]; break; }
        throw sb200_error(1, msg.empty() ? std::string("sb200: multi-GPU construction failed") : msg);
    }
    return g;
}

}  // namespace sb200

template<class F>
static int guarded(sb200_ctx *ctx, sb200_comm *cm, F &&f) {
    try {
        CUDA_CHECK(cudaSetDevice(ctx->device));
        f();
        return 0;
    } catch (const sb200_error &e) {
        ctx->last_error = e.what();
        cudaGetLastError();
        if (cm) cm->fail();
        return e.code;
    } catch (const std::exception &e) {
        ctx->last_error = e.what();
        if (cm) cm->fail();
        return 3;
    }
}

static std::string g_comm_error;

extern "C" {

int sb200_comm_unique_id(uint8_t *id128) {
    try {
        sb200::nccl_unique_id(id128);
        return 0;
    } catch (const std::exception &e) {
        g_comm_error = e.what();
        return 5;
    }
}
const char *sb200_comm_last_error(void) { return g_comm_error.c_str(); }

int sb200_comm_create_nccl(sb200_ctx *ctx, int rank, int world, const uint8_t *id128, sb200_comm **out) {
    *out = nullptr;
    return guarded(ctx, nullptr, [&] { *out = sb200::comm_create_nccl(ctx, rank, world, id128); });
}
int sb200_comm_create_local(int world, sb200_comm **out) {
    try {
        sb200::comm_create_local(world, out);
        return 0;
    } catch (const std::exception &e) {
        g_comm_error = e.what();
        return 1;
    }
}
int sb200_comm_rank(const sb200_comm *c) { return c->rank; }
int sb200_comm_size(const sb200_comm *c) { return c->size; }
void sb200_comm_free(sb200_comm *c) { delete c; }

int sb200_construct_sharded(sb200_ctx *ctx, sb200_comm *comm, const sb200_reads *reads, const sb200_construct_params *params, int gather_to,
                            sb200_shard **out) {
    *out = nullptr;
    return guarded(ctx, comm, [&] {
        SB200_REQUIRE(reads && reads->ctx == ctx, "reads belong to another context");
        *out = sb200::construct_sharded(ctx, comm, reads, params, gather_to);
    });
}
const sb200_kmers *sb200_shard_kpomers(const sb200_shard *s) { return s->kpomers; }
const sb200_kmers *sb200_shard_kmers(const sb200_shard *s) { return s->kmers; }
const sb200_mphf *sb200_shard_mphf(const sb200_shard *s) { return s->mphf; }
const sb200_ext *sb200_shard_ext(const sb200_shard *s) { return s->ext; }
const sb200_unitigs *sb200_shard_unitigs(const sb200_shard *s) { return s->unitigs; }
int sb200_shard_info(const sb200_shard *s, sb200_shard_info_t *info) {
    memset(info, 0, sizeof *info);
    info->rank = s->rank; info->size = s->size;
    info->total_kpomers = s->total_kpomers; info->total_kmers = s->total_kmers; info->total_instances = s->total_instances;
    info->total_unitigs = s->total_unitigs; info->total_unitig_bases = s->total_unitig_bases; info->n_loops = s->n_loops; info->clipped = s->clipped;
    info->whole_set_fallback = s->whole_set_fallback ? 1 : 0; info->gathered = s->gathered ? 1 : 0;
    info->bytes_sent = s->bytes_sent; info->record_bytes = s->record_bytes; info->exchange_ms = s->exchange_ms;
    for (int i = 0; i < 8; ++i) info->stage_ms[i] = s->stage_ms[i];
    return 0;
}
int sb200_shard_walk_stats(const sb200_shard *s, uint64_t *out /* size x 6 */) {
    memcpy(out, s->walk_stats.data(), s->walk_stats.size() * 8);
    return 0;
}
void sb200_shard_free(sb200_shard *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    delete s;
}

static std::string g_multi_error;
int sb200_multi_create(int n_gpus, const int *device_ids, sb200_multi **out) {
    *out = nullptr;
    sb200_multi *mg = new sb200_multi();
    try {
        SB200_REQUIRE(n_gpus >= 1 && n_gpus <= 64 && device_ids, "number of GPUs out of range [1,64]");
        mg->n = n_gpus;
        mg->devices.assign(device_ids, device_ids + n_gpus);
        sb200::multi_create(mg);
        *out = mg;
        return 0;
    } catch (const sb200_error &e) {
        g_multi_error = e.what();
        delete mg;
        return e.code;
    } catch (const std::exception &e) {
        g_multi_error = e.what();
        delete mg;
        return 3;
    }
}
void sb200_multi_destroy(sb200_multi *mg) { delete mg; }
const char *sb200_multi_last_error(const sb200_multi *mg) { return mg ? mg->last_error.c_str() : g_multi_error.c_str(); }
int sb200_multi_size(const sb200_multi *mg) { return mg->n; }
sb200_ctx *sb200_multi_context(sb200_multi *mg, int rank) { return (rank >= 0 && rank < mg->n) ? mg->ctx[(size_t) rank] : nullptr; }
int sb200_multi_construct(sb200_multi *mg, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                          const sb200_construct_params *params, sb200_graph **out) {
    *out = nullptr;
    try {
        SB200_REQUIRE(params && word_off && len && (words || n_reads == 0), "null read buffers / parameters");
        *out = sb200::multi_construct(mg, words, word_off, len, n_reads, params);
        return 0;
    } catch (const sb200_error &e) {
        mg->last_error = e.what();
        return e.code;
    } catch (const std::exception &e) {
        mg->last_error = e.what();
        return 3;
    }
}

}  // extern "C"

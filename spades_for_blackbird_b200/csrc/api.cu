// api.cu — the C ABI of libspades_b200.so (include/sb200.h): argument checking, error mapping, host<->device copies.
#include <string.h>

#include <algorithm>
#include <memory>
#include <mutex>

#include "../../include/sb200.h"
#include "common.cuh"
#include "kmer_set.cuh"

namespace sb200 {
sb200_kmers *count_reads(sb200_ctx *ctx, const sb200_reads *rd, unsigned K, int canonical_only, int add_rc, unsigned B);
sb200_kmers *derive_kmers(sb200_ctx *ctx, const sb200_kmers *kp, unsigned B);
sb200_mphf *mphf_build(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes);
void mphf_lookup_device(sb200_ctx *ctx, const sb200_mphf *m, const uint64_t *recs_dev, uint64_t n, uint64_t *out_dev);
uint64_t mphf_serialize(const sb200_mphf *m, uint8_t *out);
uint64_t mphf_seq_idx_host(sb200_mphf *m, const uint64_t *rec);
sb200_ext *build_ext(sb200_ctx *ctx, const sb200_kmers *kpomers, const sb200_kmers *kmers, const sb200_mphf *mphf);
uint64_t tipclip(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, uint64_t bound);
sb200_unitigs *extract_unitigs(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops);
}  // namespace sb200

static std::string g_create_error;
static std::mutex g_create_mutex;

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

#include "graph.cuh"

extern "C" {

int sb200_create(int device, sb200_ctx **out) {
    *out = nullptr;
    std::lock_guard<std::mutex> lock(g_create_mutex);
    try {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw sb200_error(4, std::string("sb200: no CUDA device available (") + cudaGetErrorString(e) +
                                     "); this library has no CPU fallback");
        SB200_REQUIRE(device >= 0 && device < n, "device ordinal out of range");
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            throw sb200_error(4, std::string("sb200: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                     std::to_string(prop.minor) + "; libspades_b200 is built for sm_100a only (no fallback)");
        CUDA_CHECK(cudaSetDevice(device));
        sb200_ctx *ctx = new sb200_ctx();
        ctx->device = device;
        ctx->num_sms = prop.multiProcessorCount;
        ctx->trace = getenv("SB200_TRACE") != nullptr;
        ctx->force_jump_path = getenv("SB200_FORCE_JUMP") != nullptr;
        ctx->links = getenv("SB200_LINKS") != nullptr;
        ctx->no_walk_blocks = getenv("SB200_NO_WALK_BLOCKS") != nullptr;
        ctx->no_peer_stores = getenv("SB200_NO_PEER_STORES") != nullptr;
        if (const char *v = getenv("SB200_WALK_CAPTURE_WORDS")) ctx->walk_capture_words = (size_t) atoi(v);
        ctx->no_mask_payload = getenv("SB200_NO_MASK_PAYLOAD") != nullptr;
        ctx->no_place = getenv("SB200_NO_PLACE") != nullptr;
        ctx->no_fused_partition = getenv("SB200_NO_FUSED_PARTITION") != nullptr;
        ctx->atomic_partition = getenv("SB200_ATOMIC_PARTITION") != nullptr;
        ctx->counting_passes = getenv("SB200_COUNTING_PASSES") != nullptr;
        ctx->mphf_state_per_key = getenv("SB200_MPHF_STATE_PER_KEY") != nullptr;
        ctx->group_chunk = getenv("SB200_GROUP_KERNEL") && !strcmp(getenv("SB200_GROUP_KERNEL"), "chunk");
        ctx->trace_t0 = sb200_ctx::now_s();
        CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaHostAlloc((void **) &ctx->mailbox_host, sb200_ctx::MAILBOX_BYTES, cudaHostAllocMapped));
        CUDA_CHECK(cudaHostGetDevicePointer((void **) &ctx->mailbox_dev, ctx->mailbox_host, 0));
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t threshold = UINT64_MAX;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        *out = ctx;
        return 0;
    } catch (const sb200_error &e) {
        g_create_error = e.what();
        cudaGetLastError();
        return e.code;
    }
}

void sb200_destroy(sb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    ctx->dev_trim();
    ctx->pinned_free_all();
    if (ctx->mailbox_host) { cudaFreeHost(ctx->mailbox_host); ctx->mailbox_host = nullptr; ctx->mailbox_dev = nullptr; }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    ctx->event_pool.clear();
    // Handles created from this context may outlive it (garbage-collected wrappers): their device blocks are still
    // registered, so the small context record is kept alive for them and only the streams go away when nothing is left.
    if (ctx->dev_block_size.empty()) {
        cudaStreamDestroy(ctx->stream);
        cudaStreamDestroy(ctx->copy_stream);
        delete ctx;
    } else {
        ctx->destroyed = true;   // every later dev_free returns its block to the driver; the last one deletes the context (common.cuh)
    }
}

const char *sb200_last_error(const sb200_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int sb200_synchronize(sb200_ctx *ctx) {
    return guarded(ctx, [&] { CUDA_CHECK(cudaStreamSynchronize(ctx->stream)); });
}
void *sb200_stream(sb200_ctx *ctx) { return (void *) ctx->stream; }
uint64_t sb200_kernel_launches(sb200_ctx *ctx, int reset) {
    uint64_t v = ctx->kernel_launches;
    if (reset) ctx->kernel_launches = 0;
    return v;
}

int sb200_profile(sb200_ctx *ctx, int enable) {
    return guarded(ctx, [&] {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (auto &r : ctx->prof) { ctx->event_pool.push_back(r.a); ctx->event_pool.push_back(r.b); }
        ctx->prof.clear();
        ctx->profiling = enable != 0;
    });
}

// "name\tlaunches\ttotal_ms\n" per kernel, sorted by total time; returns the number of bytes needed
int sb200_profile_report(sb200_ctx *ctx, char *out, uint64_t cap, uint64_t *needed) {
    return guarded(ctx, [&] {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        std::vector<std::pair<std::string, std::pair<uint64_t, double>>> agg;
        for (auto &r : ctx->prof) {
            float ms = 0;
            CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
            bool found = false;
            for (auto &a : agg)
                if (a.first == r.name) { a.second.first++; a.second.second += ms; found = true; break; }
            if (!found) agg.push_back({r.name, {1, ms}});
        }
        std::sort(agg.begin(), agg.end(), [](const auto &x, const auto &y) { return x.second.second > y.second.second; });
        std::string s;
        for (auto &a : agg) {
            char line[256];
            snprintf(line, sizeof line, "%s\t%llu\t%.6f\n", a.first.c_str(), (unsigned long long) a.second.first, a.second.second);
            s += line;
        }
        *needed = s.size() + 1;
        if (out && cap) {
            size_t n = std::min<size_t>(cap - 1, s.size());
            memcpy(out, s.data(), n);
            out[n] = 0;
        }
    });
}

// ---- reads ---------------------------------------------------------------------------------------------------------------
static void reads_stats(sb200_reads *r, const uint32_t *len_host, uint64_t n) {
    uint64_t bases = 0; uint32_t mx = 0;
    for (uint64_t i = 0; i < n; ++i) { bases += len_host[i]; if (len_host[i] > mx) mx = len_host[i]; }
    r->n_bases = bases; r->max_len = mx;
}

int sb200_reads_upload(sb200_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                       sb200_reads **out) {
    *out = nullptr;
    return guarded(ctx, [&] {
        SB200_REQUIRE(word_off && len && (words || n_reads == 0), "null read buffers");
        uint64_t n_words = n_reads ? word_off[n_reads] : 0;
        std::unique_ptr<sb200_reads> r(new sb200_reads());   // released if an allocation or copy below throws
        r->ctx = ctx; r->n_reads = n_reads; r->n_words = n_words;
        r->words.alloc(ctx, n_words + 2);
        r->word_off.alloc(ctx, n_reads + 1);
        r->len.alloc(ctx, n_reads + 1);
        CUDA_CHECK(cudaMemcpyAsync(r->words.p, words, n_words * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(r->word_off.p, word_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(r->len.p, len, n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
        reads_stats(r.get(), len, n_reads);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *out = r.release();
    });
}

int sb200_reads_wrap_device(sb200_ctx *ctx, const uint64_t *d_words, const uint64_t *d_word_off, const uint32_t *d_len, uint64_t n_reads,
                            uint64_t n_words, sb200_reads **out) {
    *out = nullptr;
    return guarded(ctx, [&] {
        std::unique_ptr<sb200_reads> r(new sb200_reads());
        r->ctx = ctx; r->n_reads = n_reads; r->n_words = n_words;
        r->words.alloc(ctx, n_words + 2);
        r->word_off.alloc(ctx, n_reads + 1);
        r->len.alloc(ctx, n_reads + 1);
        CUDA_CHECK(cudaMemcpyAsync(r->words.p, d_words, n_words * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(r->word_off.p, d_word_off, (n_reads + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(r->len.p, d_len, n_reads * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        *out = r.release();
    });
}

void sb200_reads_free(sb200_reads *r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    delete r;
}

// ---- k-mer sets ---------------------------------------------------------------------------------------------------------
int sb200_count(sb200_ctx *ctx, const sb200_reads *reads, unsigned K, int canonical_only, int add_rc, unsigned num_buckets, sb200_kmers **out) {
    *out = nullptr;
    return guarded(ctx, [&] {
        SB200_REQUIRE(reads && reads->ctx == ctx, "reads belong to another context");
        SB200_REQUIRE(reads->n_reads > 0, "No kmers were extracted from reads. Check the read lengths and k-mer length settings");
        *out = sb200::count_reads(ctx, reads, K, canonical_only, add_rc, num_buckets);
    });
}

int sb200_derive_kmers(sb200_ctx *ctx, const sb200_kmers *kpomers, unsigned num_buckets, sb200_kmers **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::derive_kmers(ctx, kpomers, num_buckets); });
}

unsigned sb200_kmers_k(const sb200_kmers *s) { return s->k; }
unsigned sb200_kmers_words(const sb200_kmers *s) { return s->words; }
unsigned sb200_kmers_num_buckets(const sb200_kmers *s) { return s->num_buckets; }
uint64_t sb200_kmers_size(const sb200_kmers *s) { return s->size; }
uint64_t sb200_kmers_instances(const sb200_kmers *s) { return s->instances; }
int sb200_kmers_bucket_starts(const sb200_kmers *s, uint64_t *out) {
    memcpy(out, s->bucket_starts_host.data(), ((size_t) s->num_buckets + 1) * 8);
    return 0;
}
int sb200_kmers_download(const sb200_kmers *s, uint64_t first, uint64_t count, uint64_t *records_out) {
    return guarded(s->ctx, [&] {
        SB200_REQUIRE(first + count <= s->size, "record range out of bounds");
        CUDA_CHECK(cudaMemcpyAsync(records_out, s->data.p + first * s->words, count * s->words * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(s->ctx->stream));
    });
}
int sb200_kmers_counts_download(const sb200_kmers *s, uint64_t first, uint64_t count, uint32_t *counts_out) {
    return guarded(s->ctx, [&] {
        SB200_REQUIRE(s->counts.p != nullptr, "this k-mer set carries no multiplicities");
        SB200_REQUIRE(first + count <= s->size, "record range out of bounds");
        CUDA_CHECK(cudaMemcpyAsync(counts_out, s->counts.p + first, count * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(s->ctx->stream));
    });
}
const uint64_t *sb200_kmers_device_records(const sb200_kmers *s) { return s->data.p; }
const uint32_t *sb200_kmers_device_counts(const sb200_kmers *s) { return s->counts.p; }
void sb200_kmers_free(sb200_kmers *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    delete s;
}

// ---- MPHF -----------------------------------------------------------------------------------------------------------------
int sb200_mphf_build(sb200_ctx *ctx, const sb200_kmers *kmers, sb200_mphf **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::mphf_build(ctx, kmers, nullptr); });
}
uint64_t sb200_mphf_size(const sb200_mphf *m) { return m->total; }
uint64_t sb200_mphf_mem_size(const sb200_mphf *m) {
    // KMerIndex::mem_size -> sum over buckets of mphf::mem_size = bitSize/8 where bitSize = nchar*64 + ranks.capacity()*64
    // (BooPHF.h:212,489-500); capacity is the reserve(2 + size/512) of build_ranks.
    uint64_t bits = 0;
    for (size_t t = 0; t < m->domain_host.size(); ++t) {
        uint64_t d = m->domain_host[t];
        if (!d) continue;
        bits += (1 + d / 64) * 64 + (2 + d / 512) * 64;
    }
    return bits / 8;
}
int sb200_mphf_lookup(sb200_ctx *ctx, const sb200_mphf *m, const uint64_t *records, uint64_t n, uint64_t *idx_out) {
    return guarded(ctx, [&] {
        DevBuf<uint64_t> recs(ctx, n * m->words), idx(ctx, n);
        CUDA_CHECK(cudaMemcpyAsync(recs.p, records, n * m->words * 8, cudaMemcpyHostToDevice, ctx->stream));
        sb200::mphf_lookup_device(ctx, m, recs.p, n, idx.p);
        CUDA_CHECK(cudaMemcpyAsync(idx_out, idx.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    });
}
int sb200_mphf_seq_idx(const sb200_mphf *m, const uint64_t *record, uint64_t *idx_out) {
    return guarded(m->ctx, [&] { *idx_out = sb200::mphf_seq_idx_host(const_cast<sb200_mphf *>(m), record); });
}
int sb200_mphf_serialize(const sb200_mphf *m, uint8_t *out, uint64_t *size) {
    return guarded(m->ctx, [&] { *size = sb200::mphf_serialize(m, out); });
}
void sb200_mphf_free(sb200_mphf *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    delete m;
}

// ---- extension index ------------------------------------------------------------------------------------------------------
int sb200_ext_build(sb200_ctx *ctx, const sb200_kmers *kpomers, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::build_ext(ctx, kpomers, kmers, mphf); });
}
int sb200_ext_masks_download(const sb200_ext *e, uint8_t *masks_out) {
    return guarded(e->ctx, [&] {
        CUDA_CHECK(cudaMemcpyAsync(masks_out, e->masks.p, e->size, cudaMemcpyDeviceToHost, e->ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->ctx->stream));
    });
}
int sb200_ext_idx_download(const sb200_ext *e, uint32_t *idx_out) {
    return guarded(e->ctx, [&] {
        CUDA_CHECK(cudaMemcpyAsync(idx_out, e->idx.p, e->n_local * 4, cudaMemcpyDeviceToHost, e->ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->ctx->stream));
    });
}
void sb200_ext_free(sb200_ext *e) {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    delete e;
}

int sb200_tipclip(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, uint64_t length_bound, uint64_t *removed) {
    return guarded(ctx, [&] { *removed = sb200::tipclip(ctx, kmers, mphf, ext, length_bound); });
}

// ---- unitigs ---------------------------------------------------------------------------------------------------------------
int sb200_unitigs_extract(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops,
                          sb200_unitigs **out) {
    *out = nullptr;
    return guarded(ctx, [&] { *out = sb200::extract_unitigs(ctx, kmers, mphf, ext, with_loops); });
}
uint64_t sb200_unitigs_count(const sb200_unitigs *u) { return u->count; }
uint64_t sb200_unitigs_loops(const sb200_unitigs *u) { return u->n_loops; }
uint64_t sb200_unitigs_total_bases(const sb200_unitigs *u) { return u->total_bases; }
uint64_t sb200_unitigs_total_words(const sb200_unitigs *u) { return u->total_words; }
int sb200_unitigs_download(const sb200_unitigs *u, uint64_t *words_out, uint64_t *word_off_out, uint32_t *len_out) {
    return guarded(u->ctx, [&] {
        cudaStream_t s = u->ctx->stream;
        CUDA_CHECK(cudaMemcpyAsync(words_out, u->words.p, u->total_words * 8, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(word_off_out, u->word_off.p, (u->count + 1) * 8, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(len_out, u->len.p, u->count * 4, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    });
}
void sb200_unitigs_free(sb200_unitigs *u) {
    if (!u) return;
    cudaSetDevice(u->ctx->device);
    delete u;
}

// ---- whole path: see construct.cu ------------------------------------------------------------------------------------------

int sb200_graph_get(const sb200_graph *g, sb200_graph_view *view) {
    *view = g->view;
    return 0;
}
void sb200_graph_free(sb200_graph *g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    delete g;
}

}  // extern "C"

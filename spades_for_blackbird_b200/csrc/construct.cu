// construct.cu — sb200_construct: the whole path with HOST buffers on both sides.
//
// What spades-gbuilder does between read conversion and output (A/projects/gbuilder/main.cpp:165-181) and what
// spades-core's Construction stage does (C/stages/construction.cpp:469-483): count (k+1)-mers, derive k-mers, build
// the MPHF, fill the extension masks, [clip tips], extract unitigs.  Results leave the GPU on a second stream while the
// next stage computes: the (k+1)-mer table travels during k-mer counting, the k-mer table during MPHF/mask building,
// masks and index during unitig extraction.  Device buffers whose copy may still be in flight are kept alive until
// both streams have drained; host buffers come from the context's pinned pool.
#include <stdlib.h>
#include <string.h>

#include "../../include/sb200.h"
#include "common.cuh"
#include "graph.cuh"
#include "kmer_set.cuh"

namespace sb200 {
uint64_t mphf_serialize_host(const sb200_mphf *m, const uint64_t *bits_host, const uint64_t *ranks_host, uint8_t *out);
void mphf_serialize_device(sb200_ctx *ctx, const sb200_mphf *m, uint8_t *out_dev);
}

extern "C" int sb200_construct(sb200_ctx *ctx, const uint64_t *words, const uint64_t *word_off, const uint32_t *len, uint64_t n_reads,
                               const sb200_construct_params *p, sb200_graph **out) {
    *out = nullptr;
    sb200_reads *reads = nullptr;
    sb200_kmers *kp = nullptr, *km = nullptr;
    sb200_mphf *mp = nullptr;
    sb200_ext *ext = nullptr;
    sb200_unitigs *un = nullptr;
    sb200_graph *g = nullptr;
    cudaEvent_t ev = nullptr;
    int rc = 0;
    auto fail = [&](int e) { throw sb200_error(e, ctx->last_error); };
    try {
        CUDA_CHECK(cudaSetDevice(ctx->device));
        SB200_REQUIRE(p && (p->k & 1) && p->k >= 1 && p->k < 128, "k must be odd and in [1,128)");
        cudaStream_t s = ctx->stream, c = ctx->copy_stream;
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        auto copy_after_compute = [&]() {   // copy stream picks up after everything issued so far on the compute stream
            CUDA_CHECK(cudaEventRecord(ev, s));
            CUDA_CHECK(cudaStreamWaitEvent(c, ev, 0));
        };
        const bool timeline = getenv("SB200_TIMELINE") != nullptr;   // host wall clock at the stage boundaries (stages are blocking)
        const double t_start = sb200_ctx::now_s();
        auto mark = [&](const char *label) {
            if (timeline) fprintf(stderr, "[sb200_construct] %-28s %9.3f ms\n", label, (sb200_ctx::now_s() - t_start) * 1e3);
        };
        g = new sb200_graph();
        g->ctx = ctx;
        memset(&g->view, 0, sizeof g->view);
        sb200_graph_view &v = g->view;
        int e;
        if ((e = sb200_reads_upload(ctx, words, word_off, len, n_reads, &reads))) fail(e);
        v.h2d_bytes = reads->n_words * 8 + (n_reads + 1) * 8 + n_reads * 4;
        mark("reads uploaded");
        if ((e = sb200_count(ctx, reads, p->k + 1, 1, 1, p->num_buckets, &kp))) fail(e);
        v.n_kpomers = kp->size; v.kpomer_instances = kp->instances;
        mark("(k+1)-mers counted");
        if (p->fetch_kmers) {
            uint64_t *h = g->pin<uint64_t>(kp->size * kp->words);
            uint32_t *cn = g->pin<uint32_t>(kp->size);
            copy_after_compute();
            CUDA_CHECK(cudaMemcpyAsync(h, kp->data.p, kp->size * kp->words * 8, cudaMemcpyDeviceToHost, c));
            CUDA_CHECK(cudaMemcpyAsync(cn, kp->counts.p, kp->size * 4, cudaMemcpyDeviceToHost, c));
            v.kpomers = h; v.kpomer_counts = cn;
            g->kp_starts = kp->bucket_starts_host; v.kpomer_bucket_starts = g->kp_starts.data();
            v.d2h_bytes += kp->size * kp->words * 8 + kp->size * 4;
        }
        if ((e = sb200_derive_kmers(ctx, kp, p->num_buckets, &km))) fail(e);
        v.n_kmers = km->size;
        mark("k-mers derived");
        if (p->fetch_kmers) {
            uint64_t *h = g->pin<uint64_t>(km->size * km->words);
            copy_after_compute();
            CUDA_CHECK(cudaMemcpyAsync(h, km->data.p, km->size * km->words * 8, cudaMemcpyDeviceToHost, c));
            v.kmers = h;
            g->km_starts = km->bucket_starts_host; v.kmer_bucket_starts = g->km_starts.data();
            v.d2h_bytes += km->size * km->words * 8;
        }
        if ((e = sb200_mphf_build(ctx, km, &mp))) fail(e);
        mark("MPHF built");
        // KMerIndex::serialize: the byte stream is assembled on the device and travels as one copy; the host only fills in the
        // small fields between the bit-vectors once it has landed
        const uint64_t isz = sb200::mphf_serialize_host(mp, nullptr, nullptr, nullptr);
        DevBuf<uint8_t> index_dev(ctx, isz + 8);
        sb200::mphf_serialize_device(ctx, mp, index_dev.p);
        uint8_t *index_h = g->pin<uint8_t>(isz + 8);
        copy_after_compute();
        CUDA_CHECK(cudaMemcpyAsync(index_h, index_dev.p, isz, cudaMemcpyDeviceToHost, c));
        v.d2h_bytes += isz;
        if ((e = sb200_ext_build(ctx, kp, km, mp, &ext))) fail(e);
        if (p->tip_clip) {
            uint64_t removed = 0;
            if ((e = sb200_tipclip(ctx, km, mp, ext, p->tip_length_bound, &removed))) fail(e);
            v.clipped = removed;
        }
        mark("masks filled");
        uint8_t *masks = g->pin<uint8_t>(km->size);
        copy_after_compute();
        CUDA_CHECK(cudaMemcpyAsync(masks, ext->masks.p, km->size, cudaMemcpyDeviceToHost, c));
        v.masks = masks; v.d2h_bytes += km->size;
        if ((e = sb200_unitigs_extract(ctx, km, mp, ext, p->with_loops, &un))) fail(e);
        mark("unitigs extracted");
        v.n_unitigs = un->count; v.n_loops = un->n_loops; v.unitig_bases = un->total_bases; v.n_unitig_words = un->total_words;
        uint64_t *uw = g->pin<uint64_t>(un->total_words);
        uint64_t *uo = g->pin<uint64_t>(un->count + 1);
        uint32_t *ul = g->pin<uint32_t>(un->count);
        copy_after_compute();
        CUDA_CHECK(cudaMemcpyAsync(uw, un->words.p, un->total_words * 8, cudaMemcpyDeviceToHost, c));
        CUDA_CHECK(cudaMemcpyAsync(uo, un->word_off.p, (un->count + 1) * 8, cudaMemcpyDeviceToHost, c));
        CUDA_CHECK(cudaMemcpyAsync(ul, un->len.p, un->count * 4, cudaMemcpyDeviceToHost, c));
        v.unitig_words = uw; v.unitig_word_off = uo; v.unitig_len = ul;
        v.d2h_bytes += un->total_words * 8 + (un->count + 1) * 8 + un->count * 4;
        mark("copies issued");
        CUDA_CHECK(cudaStreamSynchronize(c));
        CUDA_CHECK(cudaStreamSynchronize(s));
        mark("copies drained");
        sb200::mphf_serialize_host(mp, nullptr, nullptr, index_h);   // small fields only
        v.index_bytes = index_h; v.index_size = isz;
        mark("index header fields");
        *out = g;
    } catch (const sb200_error &err) {
        ctx->last_error = err.what();
        cudaGetLastError();
        rc = err.code;
    } catch (const std::exception &err) {
        ctx->last_error = err.what();
        rc = 3;
    }
    if (rc) {   // make sure nothing is still reading the device buffers we are about to release
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
    }
    if (ev) cudaEventDestroy(ev);
    if (reads) sb200_reads_free(reads);
    if (kp) sb200_kmers_free(kp);
    if (km) sb200_kmers_free(km);
    if (mp) sb200_mphf_free(mp);
    if (ext) sb200_ext_free(ext);
    if (un) sb200_unitigs_free(un);
    if (rc && g) delete g;
    return rc;
}

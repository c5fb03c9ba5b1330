// common.cuh — context, error handling and stream-ordered device buffers for libspades_b200.so.
//
// One sb200_ctx per GPU: a non-blocking compute stream, a copy stream for overlapped D2H, and the CUDA
// stream-ordered allocator (cudaMallocAsync) with an unbounded release threshold, so that the per-stage
// temporaries of back-to-back runs are recycled without going to the driver.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <map>
#include <string>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#define SB200_MAX_WORDS 4

struct sb200_error : std::runtime_error {
    int code;
    sb200_error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                             \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            throw sb200_error(2, _b);                                                                \
        }                                                                                            \
    } while (0)

#define SB200_REQUIRE(cond, msg)                                          \
    do {                                                                  \
        if (!(cond)) throw sb200_error(1, std::string("sb200: ") + (msg)); \
    } while (0)

struct sb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int num_sms = 148;
    std::string last_error;
    uint64_t kernel_launches = 0;   // launches of OUR kernels since the last reset (bench.py "gpu_launches")

    // Optional per-kernel device timing (bench.py's live roofline numbers): a CUDA event pair around every launch on
    // the launching stream, aggregated by kernel name in sb200_profile_report().
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; double bytes; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t get_event() {
        cudaEvent_t e;
        if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
    void prof_begin(const char *name) {
        ProfRec r{name, get_event(), get_event(), 0.0};
        cudaEventRecord(r.a, stream);
        prof.push_back(r);
    }
    void prof_end() { cudaEventRecord(prof.back().b, stream); }

    // SB200_TRACE=1: host-side wall time between trace points (each point drains the stream first), to stderr
    bool trace = false;
    double trace_t0 = 0;
    bool group_chunk = false;       // SB200_GROUP_KERNEL=chunk: the sorting group kernel (segsort.cuh) instead of the hashing one (grouphash.cuh)
    bool counting_passes = false;      // SB200_COUNTING_PASSES=1: round-1 grouping (extract / derive, then two histogram + scatter passes) instead of the staged producer-fused partition (A/B, cross-check)
    bool atomic_partition = false;     // SB200_ATOMIC_PARTITION=1: one-pass partition through L2 atomics (partition.cuh) instead of extract / derive + counting passes (A/B, cross-check)
    bool no_fused_partition = false;   // SB200_NO_FUSED_PARTITION=1: sharded path extracts first and partitions afterwards (cross-check)
    bool no_place = false;          // SB200_NO_PLACE=1: k-mer indices by MPHF lookups even when the build recorded the placements (cross-check)
    bool no_mask_payload = false;   // SB200_NO_MASK_PAYLOAD=1: masks by MPHF lookups (fill_masks_kernel) even when the k-mer sort could carry them
    bool mphf_state_per_key = false;  // SB200_MPHF_STATE_PER_KEY=1: level 0 writes a 24-byte state per key for level 1 (round 1) instead of level 1 re-reading the keys (cross-check)
    bool no_peer_stores = false;    // SB200_NO_PEER_STORES=1: the sharded path's record exchanges through send buffers + all-to-all instead of peer stores from pass 1
    bool no_walk_blocks = false;    // SB200_NO_WALK_BLOCKS=1: lookup walks read bit-vector, rank and mask separately (cross-check)
    size_t walk_capture_words = 0;  // SB200_WALK_CAPTURE_WORDS=n: cap of the measuring walks' capture buffer per start edge (tests: force the re-walk)
    bool links = false;             // SB200_LINKS=1: direct walks by pointer chasing over a link table (one lookup per vertex up front) instead of a lookup
                                    // through the walk blocks per step — the round-1 scheme, slower since the blocks exist (tests cover both)
    bool force_jump_path = false;   // SB200_FORCE_JUMP=1: always extract unitigs by pointer jumping (tests cover both paths)
    static double now_s() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
    }
    void trace_point(const char *label) {
        if (!trace) return;
        cudaStreamSynchronize(stream);
        double t = now_s();
        fprintf(stderr, "[sb200] %-34s %9.3f ms\n", label, (t - trace_t0) * 1e3);
        trace_t0 = now_s();
    }

    // Caching device allocator.  All kernels of a context run on one stream, so a block released by the host can be handed
    // out again immediately: work that still uses it was enqueued earlier on the same stream.  (sb200_construct, which also
    // copies on a second stream, keeps its buffers until both streams have drained.)  Requests are rounded up to 256 B and
    // served from the smallest cached block that is large enough but not more than 25 % larger; otherwise cudaMalloc.
    std::multimap<size_t, void *> dev_free_blocks;
    std::unordered_map<void *, size_t> dev_block_size;
    size_t dev_bytes_total = 0;
    void *dev_alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t) 255;
        auto it = dev_free_blocks.lower_bound(bytes);
        if (it != dev_free_blocks.end() && it->first <= bytes + bytes / 4 + 4096) {
            void *p = it->second;
            dev_free_blocks.erase(it);
            return p;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {   // give cached blocks back to the driver and retry once
            cudaGetLastError();
            dev_trim();
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) {
            char b[160];
            snprintf(b, sizeof b, "sb200: out of device memory allocating %zu bytes (%zu bytes held): %s", bytes, dev_bytes_total,
                     cudaGetErrorString(e));
            cudaGetLastError();
            throw sb200_error(2, b);
        }
        dev_block_size[p] = bytes;
        dev_bytes_total += bytes;
        return p;
    }
    bool destroyed = false;   // sb200_destroy ran while handles were still alive: blocks go straight back to the driver
    void dev_free(void *p) {
        auto it = dev_block_size.find(p);
        if (it == dev_block_size.end()) return;
        if (destroyed) {
            cudaFree(p);
            dev_bytes_total -= it->second;
            dev_block_size.erase(it);
            if (dev_block_size.empty()) {   // the last handle of a destroyed context: the record and its streams go too
                cudaStreamDestroy(stream);
                cudaStreamDestroy(copy_stream);
                delete this;
            }
            return;
        }
        dev_free_blocks.insert({it->second, p});
    }
    void dev_trim() {   // release every cached (currently unused) block
        cudaStreamSynchronize(stream);
        cudaStreamSynchronize(copy_stream);
        for (auto &kv : dev_free_blocks) {
            cudaFree(kv.second);
            dev_bytes_total -= kv.first;
            dev_block_size.erase(kv.second);
        }
        dev_free_blocks.clear();
    }

    // Pinned host buffers for results that go back to the caller (sb200_construct): cudaMallocHost of gigabytes costs
    // hundreds of milliseconds, so blocks are recycled across calls (grow-only, freed with the context).
    struct PinnedBlock { void *p; size_t cap; bool used; };
    std::vector<PinnedBlock> pinned_pool;
    void *pinned_get(size_t bytes) {
        if (bytes == 0) bytes = 8;
        PinnedBlock *best = nullptr;
        for (auto &b : pinned_pool)
            if (!b.used && b.cap >= bytes && (!best || b.cap < best->cap)) best = &b;
        if (best) { best->used = true; return best->p; }
        void *p = nullptr;
        size_t cap = bytes + bytes / 16;   // a little slack so that run-to-run size jitter still hits the pool
        if (cudaMallocHost(&p, cap) != cudaSuccess) throw sb200_error(2, "cudaMallocHost failed for result buffer");
        pinned_pool.push_back({p, cap, true});
        return p;
    }
    void pinned_put(void *p) {
        for (auto &b : pinned_pool)
            if (b.p == p) { b.used = false; return; }
    }
    void pinned_free_all() {
        for (auto &b : pinned_pool) cudaFreeHost(b.p);
        pinned_pool.clear();
    }

    // Small device -> host read-backs (counts, totals, flags the host needs to size the next launch) do NOT use the copy
    // engine: while sb200_construct streams a gigabyte-sized result to the host on the copy stream, a 4-byte
    // cudaMemcpyAsync queues behind it on the same D2H engine and every stage stalls (measured: the k-mer stage took
    // 33.9 ms instead of 14.1 ms).  A one-block kernel stores the bytes into a host-mapped mailbox instead.
    static constexpr size_t MAILBOX_BYTES = 16384;
    uint8_t *mailbox_host = nullptr, *mailbox_dev = nullptr;
    void fetch(void *dst, const void *dev_src, size_t bytes);   // blocking: enqueue, synchronise the compute stream, copy out
};

// Device array from the context's caching allocator (sb200_ctx::dev_alloc): blocks are recycled in stream order on the
// context's compute stream, so a steady-state run of the path makes no driver allocation calls at all.
template<class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    sb200_ctx *c = nullptr;
    DevBuf() {}
    DevBuf(sb200_ctx *ctx, size_t count) { alloc(ctx, count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), c(o.c), own(o.own) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; c = o.c; own = o.own; o.p = nullptr; o.n = 0; }
        return *this;
    }
    bool own = true;   // false: a view of memory somebody else keeps (a peer-visible receive buffer of the communicator)
    void borrow(sb200_ctx *ctx, T *ptr, size_t count) { release(); c = ctx; p = ptr; n = count; own = false; }
    ~DevBuf() { release(); }
    void alloc(sb200_ctx *ctx, size_t count) {
        release();
        c = ctx;
        n = count;
        p = (T *) ctx->dev_alloc((count ? count : 1) * sizeof(T));
    }
    void release() {
        if (p && own) c->dev_free(p);
        p = nullptr; n = 0; own = true;
    }
    void zero() { CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), c->stream)); }
    size_t bytes() const { return n * sizeof(T); }
};

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned) ((a + b - 1) / b); }

static __global__ void mailbox_copy_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t bytes) {
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                    \
        if ((ctx)->profiling) (ctx)->prof_begin(#kernel);                   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);    \
        if ((ctx)->profiling) (ctx)->prof_end();                            \
        (ctx)->kernel_launches++;                                           \
        CUDA_CHECK(cudaGetLastError());                                     \
    } while (0)

inline void sb200_ctx::fetch(void *dst, const void *dev_src, size_t bytes) {
    if (bytes == 0) return;
    if (!mailbox_host || bytes > MAILBOX_BYTES) {
        CUDA_CHECK(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        return;
    }
    LAUNCH(this, mailbox_copy_kernel, 1, 256, 0, (const uint8_t *) dev_src, mailbox_dev, (uint32_t) bytes);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    memcpy(dst, mailbox_host, bytes);
}

// common.cuh — context, error handling and stream-ordered device buffers for libspades_b200.so.
//
// One sb200_ctx per GPU: a non-blocking compute stream, a copy stream for overlapped D2H, and the CUDA
// stream-ordered allocator (cudaMallocAsync) with an unbounded release threshold, so that the per-stage
// temporaries of back-to-back runs are recycled without going to the driver.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
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
};

// Stream-ordered device array.
template<class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(sb200_ctx *ctx, size_t count) { alloc(ctx, count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(sb200_ctx *ctx, size_t count) {
        release();
        s = ctx->stream;
        n = count;
        CUDA_CHECK(cudaMallocAsync((void **) &p, (count ? count : 1) * sizeof(T), s));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr; n = 0;
    }
    void zero() { CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    size_t bytes() const { return n * sizeof(T); }
};

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned) ((a + b - 1) / b); }

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                    \
        if ((ctx)->profiling) (ctx)->prof_begin(#kernel);                   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);    \
        if ((ctx)->profiling) (ctx)->prof_end();                            \
        (ctx)->kernel_launches++;                                           \
        CUDA_CHECK(cudaGetLastError());                                     \
    } while (0)

// graph.cuh — host-side result of sb200_construct: pinned buffers owned through the context's pinned pool.
#pragma once
#include "../../include/sb200.h"
#include "common.cuh"

struct sb200_graph {
    sb200_ctx *ctx = nullptr;
    sb200_graph_view view;
    std::vector<void *> pinned;
    std::vector<uint64_t> kp_starts, km_starts;
    template<class T>
    T *pin(size_t n) {
        void *p = ctx->pinned_get(n * sizeof(T));
        pinned.push_back(p);
        return (T *) p;
    }
    ~sb200_graph() {
        for (void *p : pinned) ctx->pinned_put(p);
    }
};

// comm.cu — NCCL (dlopen'ed) and local (virtual ranks) implementations of the exchange steps of the hash-sharded path (comm.cuh).
#include <dlfcn.h>
#include <nccl.h>

#include "comm.cuh"

namespace sb200 {

// ---- NCCL entry points, resolved at run time -----------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // RTLD_NOLOAD first: inside a PyTorch process the NCCL torch brought along is already mapped and must be the one we use
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.handle = h;
#define SB200_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
        SB200_NCCL_SYM(GetUniqueId); SB200_NCCL_SYM(CommInitRank); SB200_NCCL_SYM(CommDestroy); SB200_NCCL_SYM(GroupStart);
        SB200_NCCL_SYM(GroupEnd); SB200_NCCL_SYM(Send); SB200_NCCL_SYM(Recv); SB200_NCCL_SYM(AllGather); SB200_NCCL_SYM(AllReduce);
        SB200_NCCL_SYM(Broadcast); SB200_NCCL_SYM(GetErrorString);
#undef SB200_NCCL_SYM
    });
    SB200_REQUIRE(api.handle && api.GetUniqueId && api.CommInitRank && api.Send && api.Recv && api.AllGather && api.Broadcast && api.GroupStart,
                  "libnccl.so.2 not found (or too old): the multi-GPU path needs NCCL");
    return api;
}

#define NCCL_CHECK(expr)                                                                                                \
    do {                                                                                                                \
        ncclResult_t r_ = (expr);                                                                                       \
        if (r_ != ncclSuccess) {                                                                                        \
            char b_[384];                                                                                               \
            snprintf(b_, sizeof b_, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl_api().GetErrorString(r_));      \
            throw sb200_error(5, b_);                                                                                   \
        }                                                                                                               \
    } while (0)

void nccl_unique_id(uint8_t id[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    NCCL_CHECK(nccl_api().GetUniqueId(&u));
    memcpy(id, &u, 128);
}

sb200_comm *comm_create_nccl(sb200_ctx *ctx, int rank, int world, const uint8_t id[128]) {
    SB200_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world size out of range");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t c = nullptr;
    NCCL_CHECK(nccl_api().CommInitRank(&c, world, u, rank));
    sb200_comm *cm = new sb200_comm();
    cm->rank = rank; cm->size = world; cm->nccl = c;
    return cm;
}

void comm_create_local(int world, sb200_comm **out) {
    SB200_REQUIRE(world >= 1 && world <= 64, "world size out of range [1,64]");
    auto sh = std::make_shared<LocalShared>(world);
    for (int r = 0; r < world; ++r) {
        out[r] = new sb200_comm();
        out[r]->rank = r; out[r]->size = world; out[r]->local = sh;
    }
}

void LocalShared::barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (aborted) throw sb200_error(6, "sb200: another rank of the local communicator failed");
    const uint64_t gen = generation;
    if (++waiting == size) {
        waiting = 0;
        ++generation;
        cv.notify_all();
        return;
    }
    cv.wait(lk, [&] { return generation != gen || aborted; });
    if (generation == gen && aborted) throw sb200_error(6, "sb200: another rank of the local communicator failed");
}

void LocalShared::abort() {
    std::lock_guard<std::mutex> lk(m);
    aborted = true;
    cv.notify_all();
}

__global__ void or_bytes_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint64_t n_words) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) dst[i] |= src[i];
}

}  // namespace sb200

using namespace sb200;

sb200_comm::~sb200_comm() {
    for (PeerSlot &ps : peer) {
        for (size_t r = 0; r < ps.mapped.size(); ++r)
            if (ps.ipc[r] && ps.mapped[r]) cudaIpcCloseMemHandle(ps.mapped[r]);
        if (ps.mine) cudaFree(ps.mine);
    }
    if (nccl) nccl_api().CommDestroy((ncclComm_t) nccl);
}

void sb200_comm::fail() {
    if (local) local->abort();
}

void sb200_comm::barrier(sb200_ctx *ctx) {
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (local) { local->barrier(); return; }
    uint64_t v = 0, out[64];
    SB200_REQUIRE(size <= 64, "world size beyond 64");
    all_gather_host(ctx, &v, 1, out);
}

void sb200_comm::all_gather_host(sb200_ctx *ctx, const uint64_t *in, size_t n, uint64_t *out) {
    if (size == 1) { memcpy(out, in, n * 8); return; }
    if (local) {
        local->vals[(size_t) rank].assign(in, in + n);
        local->barrier();
        for (int r = 0; r < size; ++r) memcpy(out + (size_t) r * n, local->vals[(size_t) r].data(), n * 8);
        local->barrier();   // nobody overwrites its vector before everybody has read it
        return;
    }
    DevBuf<uint64_t> tmp(ctx, n * ((size_t) size + 1));
    CUDA_CHECK(cudaMemcpyAsync(tmp.p + (size_t) size * n, in, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(nccl_api().AllGather(tmp.p + (size_t) size * n, tmp.p, n, ncclUint64, (ncclComm_t) nccl, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, (size_t) size * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void sb200_comm::all_to_all_v(sb200_ctx *ctx, const void *send, const uint64_t *send_off, void *recv, const uint64_t *recv_off) {
    const uint8_t *s = (const uint8_t *) send;
    uint8_t *d = (uint8_t *) recv;
    bytes_sent += (send_off[size] - send_off[0]) - (send_off[rank + 1] - send_off[rank]);
    record_bytes += (send_off[size] - send_off[0]) - (send_off[rank + 1] - send_off[rank]);
    cudaEvent_t e0 = ctx->get_event(), e1 = ctx->get_event();
    if (local) {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // my send buffer is complete
        local->ptr[(size_t) rank] = send;
        local->vals[(size_t) rank].assign(send_off, send_off + size + 1);
        local->barrier();
        CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
        for (int r = 0; r < size; ++r) {
            const uint64_t bytes = recv_off[r + 1] - recv_off[r];
            const uint64_t so = local->vals[(size_t) r][(size_t) rank];
            SB200_REQUIRE(local->vals[(size_t) r][(size_t) rank + 1] - so == bytes, "all-to-all: send and receive counts disagree");
            if (bytes)
                CUDA_CHECK(cudaMemcpyAsync(d + recv_off[r], (const uint8_t *) local->ptr[(size_t) r] + so, bytes, cudaMemcpyDefault, ctx->stream));
        }
        CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
        local->barrier();   // senders may release their buffers now
    } else {
        CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
        NCCL_CHECK(nccl_api().GroupStart());
        for (int r = 0; r < size; ++r) {
            const uint64_t sb = send_off[r + 1] - send_off[r], rb = recv_off[r + 1] - recv_off[r];
            if (r == rank) continue;
            if (sb) NCCL_CHECK(nccl_api().Send(s + send_off[r], sb, ncclUint8, r, (ncclComm_t) nccl, ctx->stream));
            if (rb) NCCL_CHECK(nccl_api().Recv(d + recv_off[r], rb, ncclUint8, r, (ncclComm_t) nccl, ctx->stream));
        }
        NCCL_CHECK(nccl_api().GroupEnd());
        const uint64_t own = send_off[rank + 1] - send_off[rank];   // my own part never touches the network
        SB200_REQUIRE(own == recv_off[rank + 1] - recv_off[rank], "all-to-all: send and receive counts disagree");
        if (own) CUDA_CHECK(cudaMemcpyAsync(d + recv_off[rank], s + send_off[rank], own, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    exchange_ms += ms;
    ctx->event_pool.push_back(e0);
    ctx->event_pool.push_back(e1);
}

bool sb200_comm::peer_buffers(sb200_ctx *ctx, int slot, const uint64_t *need, void **ptrs) {
    if (peer_failed) return false;
    PeerSlot &ps = peer[slot];
    if (ps.cap.empty()) { ps.cap.assign((size_t) size, 0); ps.mapped.assign((size_t) size, nullptr); ps.ipc.assign((size_t) size, false); }
    bool grow = false;
    for (int r = 0; r < size; ++r) grow = grow || need[r] > ps.cap[(size_t) r];
    if (grow) {
        // nobody may still address a buffer that is about to be replaced
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < size; ++r) {
            if (need[r] <= ps.cap[(size_t) r]) continue;
            if (ps.ipc[(size_t) r] && ps.mapped[(size_t) r]) CUDA_CHECK(cudaIpcCloseMemHandle(ps.mapped[(size_t) r]));
            ps.mapped[(size_t) r] = nullptr; ps.ipc[(size_t) r] = false;
        }
        barrier(ctx);
        std::vector<bool> grew((size_t) size, false);
        for (int r = 0; r < size; ++r) {
            if (need[r] <= ps.cap[(size_t) r]) continue;
            grew[(size_t) r] = true;
            ps.cap[(size_t) r] = need[r] + need[r] / 4 + (1u << 20);
            if (r == rank) {
                if (ps.mine) CUDA_CHECK(cudaFree(ps.mine));
                ps.mine = nullptr;
                CUDA_CHECK(cudaSetDevice(ctx->device));
                CUDA_CHECK(cudaMalloc(&ps.mine, ps.cap[(size_t) r]));
                ps.mapped[(size_t) r] = ps.mine;
            }
        }
        uint64_t ok = 1;
        if (local) {
            CUDA_CHECK(cudaSetDevice(ctx->device));   // peer access is a property of the calling thread's current device
            local->ptr[(size_t) rank] = ps.mine;
            local->barrier();
            for (int r = 0; r < size; ++r) {
                if (r == rank || !grew[(size_t) r]) continue;
                void *q = const_cast<void *>(local->ptr[(size_t) r]);
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); ok = 0; continue; }
                if (at.device != ctx->device) {
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, ctx->device, at.device);
                    if (!can) { ok = 0; continue; }
                    const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                    cudaGetLastError();
                }
                ps.mapped[(size_t) r] = q;
            }
            local->barrier();
        } else {
            // one CUDA IPC handle (64 bytes) per rank goes round; only the buffers that were replaced are (re)opened
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
            uint64_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            std::vector<uint64_t> all((size_t) size * 8);
            if (grew[(size_t) rank]) {
                cudaIpcMemHandle_t hd;
                CUDA_CHECK(cudaIpcGetMemHandle(&hd, ps.mine));
                memcpy(h, &hd, 64);
            }
            all_gather_host(ctx, h, 8, all.data());
            for (int r = 0; r < size; ++r) {
                if (r == rank || !grew[(size_t) r]) continue;
                cudaIpcMemHandle_t hd;
                memcpy(&hd, all.data() + (size_t) r * 8, 64);
                void *q = nullptr;
                if (cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; continue; }
                ps.mapped[(size_t) r] = q; ps.ipc[(size_t) r] = true;
            }
        }
        // all or nothing
        std::vector<uint64_t> oks((size_t) size);
        all_gather_host(ctx, &ok, 1, oks.data());
        for (int r = 0; r < size; ++r)
            if (!oks[(size_t) r]) peer_failed = true;
        if (peer_failed) return false;
    }
    for (int r = 0; r < size; ++r) ptrs[r] = ps.mapped[(size_t) r];
    return true;
}

void sb200_comm::all_gather_v_inplace(sb200_ctx *ctx, void *buf, const uint64_t *off) {
    void *bufs[1] = {buf};
    const uint64_t *offs[1] = {off};
    all_gather_v_inplace_multi(ctx, 1, bufs, offs);
}

void sb200_comm::all_gather_v_inplace_multi(sb200_ctx *ctx, int n_bufs, void *const *bufs, const uint64_t *const *offs) {
    if (size == 1) return;
    for (int i = 0; i < n_bufs; ++i) bytes_sent += (offs[i][rank + 1] - offs[i][rank]) * (uint64_t) (size - 1);
    if (local) {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n_bufs; ++i) {
            uint8_t *b = (uint8_t *) bufs[i];
            const uint64_t *off = offs[i];
            local->ptr[(size_t) rank] = b;
            local->barrier();
            for (int r = 0; r < size; ++r) {
                const uint64_t bytes = off[r + 1] - off[r];
                if (r != rank && bytes)
                    CUDA_CHECK(cudaMemcpyAsync(b + off[r], (const uint8_t *) local->ptr[(size_t) r] + off[r], bytes, cudaMemcpyDefault, ctx->stream));
            }
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            local->barrier();
        }
        return;
    }
    // point-to-point in one group: every rank hands its slices to every peer and takes theirs — NVSwitch carries all pairs at once
    NCCL_CHECK(nccl_api().GroupStart());
    for (int i = 0; i < n_bufs; ++i) {
        uint8_t *b = (uint8_t *) bufs[i];
        const uint64_t *off = offs[i];
        const uint64_t mine = off[rank + 1] - off[rank];
        for (int d = 1; d < size; ++d) {
            const int to = (rank + d) % size, from = (rank - d + size) % size;
            const uint64_t nb = off[from + 1] - off[from];
            if (mine) NCCL_CHECK(nccl_api().Send(b + off[rank], mine, ncclUint8, to, (ncclComm_t) nccl, ctx->stream));
            if (nb) NCCL_CHECK(nccl_api().Recv(b + off[from], nb, ncclUint8, from, (ncclComm_t) nccl, ctx->stream));
        }
    }
    NCCL_CHECK(nccl_api().GroupEnd());
}

void sb200_comm::gather_v(sb200_ctx *ctx, const void *send, uint64_t bytes, void *recv, const uint64_t *recv_off, int root) {
    uint8_t *d = (uint8_t *) recv;
    if (rank != root) bytes_sent += bytes;
    if (local) {
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        local->ptr[(size_t) rank] = send;
        local->barrier();
        if (rank == root) {
            for (int r = 0; r < size; ++r) {
                const uint64_t nb = recv_off[r + 1] - recv_off[r];
                if (nb) CUDA_CHECK(cudaMemcpyAsync(d + recv_off[r], local->ptr[(size_t) r], nb, cudaMemcpyDefault, ctx->stream));
            }
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
        local->barrier();
        return;
    }
    NCCL_CHECK(nccl_api().GroupStart());
    if (rank == root) {
        for (int r = 0; r < size; ++r) {
            const uint64_t nb = recv_off[r + 1] - recv_off[r];
            if (r != root && nb) NCCL_CHECK(nccl_api().Recv(d + recv_off[r], nb, ncclUint8, r, (ncclComm_t) nccl, ctx->stream));
        }
    } else if (bytes) {
        NCCL_CHECK(nccl_api().Send(send, bytes, ncclUint8, root, (ncclComm_t) nccl, ctx->stream));
    }
    NCCL_CHECK(nccl_api().GroupEnd());
    if (rank == root && bytes) CUDA_CHECK(cudaMemcpyAsync(d + recv_off[root], send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
}

void sb200_comm::all_reduce_or_bytes(sb200_ctx *ctx, uint8_t *buf, uint64_t n, bool flags01) {
    if (size == 1 || n == 0) return;
    bytes_sent += n;   // order of magnitude (ring / tree details are NCCL's)
    if (local) {
        // every rank ORs the others' ORIGINAL arrays into a private sum, then replaces its own array
        const uint64_t n_words = (n + 3) / 4;   // the arrays are padded to a multiple of 4 bytes by their owners (ext.cu, shard.cu)
        DevBuf<uint32_t> acc(ctx, n_words), tmp(ctx, n_words);
        CUDA_CHECK(cudaMemcpyAsync(acc.p, buf, n_words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        local->ptr[(size_t) rank] = buf;
        local->barrier();
        for (int r = 0; r < size; ++r) {
            if (r == rank) continue;
            CUDA_CHECK(cudaMemcpyAsync(tmp.p, local->ptr[(size_t) r], n_words * 4, cudaMemcpyDefault, ctx->stream));
            LAUNCH(ctx, or_bytes_kernel, div_up(n_words, 256), 256, 0, acc.p, tmp.p, n_words);
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        local->barrier();   // everybody has read everybody's original
        CUDA_CHECK(cudaMemcpyAsync(buf, acc.p, n_words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        return;
    }
    // bit masks with disjoint supports: sum == or; 0/1 flags several ranks may raise: max == or
    NCCL_CHECK(nccl_api().AllReduce(buf, buf, n, ncclUint8, flags01 ? ncclMax : ncclSum, (ncclComm_t) nccl, ctx->stream));
}

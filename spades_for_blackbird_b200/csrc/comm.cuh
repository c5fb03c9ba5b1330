// comm.cuh — the exchange steps of the hash-sharded path behind one small interface with two implementations:
//   * NCCL over NVLink / NVSwitch (one rank per GPU; libnccl.so.2 is resolved at run time with dlopen, so that the library itself
//     loads on a box without NCCL and, inside a PyTorch process, shares the NCCL build torch has already mapped);
//   * "local": G virtual ranks = G host threads of one process, each with its own context (on the same GPU or on different ones),
//     exchanging through device-to-device copies between host barriers — the same orchestration code on a one-GPU box
//     (tests/test_gpu_sharded.py; B200_PROFILING.md rules out emulating ranks with kernels that wait on one another).
// The reference has no counterpart: its shuffle is fwrite(..., "ab") into kmers_raw<i> files under `omp critical`
// (C/utils/kmer_mph/kmer_splitter.hpp:140-161).  All device work is enqueued on the context's own stream, so the stages before and
// after an exchange need no cross-stream synchronisation.
#pragma once
#include <condition_variable>
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace sb200 {

struct LocalShared {   // the meeting point of the virtual ranks of one local communicator
    int size = 0;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    bool aborted = false;   // a rank failed: every barrier throws from now on instead of waiting for it
    std::vector<const void *> ptr;                 // one published device pointer per rank
    std::vector<std::vector<uint64_t>> vals;       // one published host vector per rank
    explicit LocalShared(int n) : size(n), ptr((size_t) n, nullptr), vals((size_t) n) {}
    void barrier();
    void abort();
};

}  // namespace sb200

struct sb200_comm {
    int rank = 0, size = 1;
    void *nccl = nullptr;                          // ncclComm_t, or nullptr for a local communicator
    std::shared_ptr<sb200::LocalShared> local;     // virtual ranks
    uint64_t bytes_sent = 0;                       // bytes this rank handed to other ranks since the last reset (NVLink roofline of the bench)
    uint64_t record_bytes = 0;                     // ... of which by the record exchanges (what exchange_ms timed)
    double exchange_ms = 0;                        // device time of the record exchanges (peer-store pass 1, or all-to-all) since the last reset
    struct PeerSlot {
        void *mine = nullptr;
        std::vector<uint64_t> cap;                 // capacity of every rank's buffer (the same bookkeeping on every rank)
        std::vector<void *> mapped;                // how this rank addresses every rank's buffer
        std::vector<bool> ipc;                     // mapped[r] came from cudaIpcOpenMemHandle
    };
    PeerSlot peer[2];
    bool peer_failed = false;                      // some pair cannot map: the exchanges go through NCCL / copies
    ~sb200_comm();

    // every rank contributes n host values; out[r * n + i] = value i of rank r
    void all_gather_host(sb200_ctx *ctx, const uint64_t *in, size_t n, uint64_t *out);
    // send[send_off[r], send_off[r + 1]) goes to rank r; recv[recv_off[s], recv_off[s + 1]) arrives from rank s (byte offsets)
    void all_to_all_v(sb200_ctx *ctx, const void *send, const uint64_t *send_off, void *recv, const uint64_t *recv_off);
    // Peer-visible receive buffers, one per slot and rank, kept for the life of the communicator: ptrs[r] = rank r's buffer of `slot`
    // as THIS rank addresses it (its own allocation; a CUDA IPC mapping when r is another process; the plain pointer when r is a
    // virtual rank of this process, with peer access enabled between the devices).  need[r] = bytes rank r's buffer must hold — every rank
    // passes the same array, so all of them know without talking when somebody has to grow and the handles have to go round again
    // (steady state: no communication at all).  Returns false on every rank when some pair of GPUs cannot map each other's memory.
    bool peer_buffers(sb200_ctx *ctx, int slot, const uint64_t *need, void **ptrs);
    // every rank holds the same layout; rank r's slice buf[off[r], off[r + 1]) is valid on r and is copied to everybody else (bytes)
    void all_gather_v_inplace(sb200_ctx *ctx, void *buf, const uint64_t *off);
    // the same for several buffers in ONE exchange (one NCCL group: one kernel, one handshake)
    void all_gather_v_inplace_multi(sb200_ctx *ctx, int n_bufs, void *const *bufs, const uint64_t *const *offs);
    // rank r's `bytes` bytes land at recv[recv_off[r] ...) on `root` (recv / recv_off only read on root)
    void gather_v(sb200_ctx *ctx, const void *send, uint64_t bytes, void *recv, const uint64_t *recv_off, int root);
    // element-wise OR of byte arrays over all ranks, in place (buffers are readable up to the next multiple of 4 bytes).  NCCL has no OR:
    // bit masks whose set bits are disjoint between ranks are summed, 0/1 flags (flags01) are max-ed
    void all_reduce_or_bytes(sb200_ctx *ctx, uint8_t *buf, uint64_t n, bool flags01);
    void barrier(sb200_ctx *ctx);
    void fail();   // called by a rank that is about to leave the collective sequence with an error
};

namespace sb200 {
void nccl_unique_id(uint8_t id[128]);
sb200_comm *comm_create_nccl(sb200_ctx *ctx, int rank, int world, const uint8_t id[128]);
void comm_create_local(int world, sb200_comm **out);
}  // namespace sb200

// kmer_set.cuh — device-resident sorted-unique k-mer sets (the GPU counterpart of kmers::KMerDiskStorage,
// C/utils/kmer_mph/kmer_index_builder.hpp:48-191) and the handles the C ABI passes around.
#pragma once
#include <memory>
#include <vector>
#include "common.cuh"

struct sb200_reads {
    sb200_ctx *ctx = nullptr;
    uint64_t n_reads = 0, n_words = 0, n_bases = 0;
    uint32_t max_len = 0;
    DevBuf<uint64_t> words;      // packed bases, every read word-aligned
    DevBuf<uint64_t> word_off;   // n_reads + 1
    DevBuf<uint32_t> len;        // n_reads
};

struct sb200_kmers {
    sb200_ctx *ctx = nullptr;
    unsigned k = 0, words = 0, num_buckets = 0;
    uint64_t size = 0;
    uint64_t instances = 0;              // window instances that were counted (0 for derived sets)
    DevBuf<uint64_t> data;               // size x words, file order: bucket, then array_less
    DevBuf<uint32_t> counts;             // multiplicity per record (empty for derived sets)
    DevBuf<uint8_t> masks_file;          // derived sets: InOutMask byte per record in FILE order, OR-ed while the set was sorted (may be empty)
    DevBuf<uint64_t> bucket_starts;      // num_buckets + 1 (device)
    std::vector<uint64_t> bucket_starts_host;
};

struct sb200_records {   // unsorted k-mer instances / candidates on their way through the hash shuffle
    sb200_ctx *ctx = nullptr;
    unsigned k = 0, words = 0;
    uint64_t n = 0;
    bool double_palindromes = false;   // canonical fwd+RC counting mode: a self-reverse-complement record counts twice
    bool marker = false;               // canonical forward-only mode: all-ones records are "filtered out" markers
    bool mask_payload = false;         // derived k-mer candidates: 3 padding bits above the k-mer carry the InOutMask bit (count.cu)
    DevBuf<uint64_t> data;
    // staged sharded exchange (count.cu shard_send_* / shard_receive): the records are grouped by (owner, coarse bin) and travel with the
    // sizes of those runs; a k-mer that fills its last word keeps its mask bit in a byte beside the record
    DevBuf<uint8_t> pay;
    DevBuf<uint32_t> coarse_counts;    // [n_owners x n_co] on the sender, [n_sources x n_co] on the receiver
};

struct sb200_mphf {
    sb200_ctx *ctx = nullptr;
    unsigned num_buckets = 0, words = 0;
    uint64_t total = 0;
    // Per (bucket, level) geometry, BooPHF layout: level l of bucket b has domain[b*25+l] bits and
    // nchar = 1 + domain/64 words starting at word_off[b*25+l]; its rank samples (one per 8 words) start at
    // rank_off[b*25+l].  A bucket's levels are contiguous, so a running popcount over the bucket's words is
    // BooPHF's cumulative rank offset.
    std::vector<uint64_t> domain_host, word_off_host, rank_off_host, bucket_word_start_host;
    std::vector<uint64_t> segment_starts_host;   // num_buckets + 1, with the reference's last-entry quirk
    std::vector<uint64_t> lastbitsetrank_host, bucket_sizes_host;
    uint64_t total_words = 0, total_ranks = 0;
    DevBuf<uint64_t> domain, word_off, rank_off, segment_starts;   // device copies of the tables
    DevBuf<uint64_t> bits;    // all bit-vectors
    DevBuf<uint64_t> ranks;   // all rank samples
    // Whole-table builds on one GPU also keep where every key landed: place[i] = global bit position (word << 6 | bit) of key i of the
    // set the index was built over, and pc_scan[w] = set bits before word w.  The index of key i is then pc_scan[w] + popc(bits below) —
    // no hashing, no level probing (ext.cu index_from_place_kernel).
    DevBuf<uint32_t> place, pc_scan;
    uint64_t final_level_keys = 0;
    // host copy of bits / ranks for single-key lookups from host code (sb200_mphf_seq_idx), made on first use
    std::vector<uint64_t> bits_host, ranks_host;
    bool host_copy = false;
};

struct sb200_ext {   // DeBruijnExtensionIndex payload: masks in MPHF-index order (+ successor links for the walks)
    sb200_ctx *ctx = nullptr;
    unsigned k = 0;
    uint64_t size = 0;          // entries of the mask array = keys of the whole index
    uint64_t n_local = 0;       // k-mers of the table this GPU holds (== size unless the table is sharded)
    DevBuf<uint8_t> masks;      // size bytes (padded to a multiple of 4), PerfectHashMap::data_
    DevBuf<uint32_t> idx;       // MPHF index of every k-mer in file order
    DevBuf<uint32_t> inv;       // file position of every MPHF index
    DevBuf<uint32_t> succ;      // successor of every oriented vertex with one outgoing edge (2 * size entries)
    bool masks_edited = false;  // tip clipping changed the masks: the file-order copy kept by the k-mer set is stale
    bool succ_valid = false;    // false once the masks were edited (tip clipping): links are then recomputed by lookup
};

struct sb200_unitigs {
    sb200_ctx *ctx = nullptr;
    unsigned k = 0;
    uint64_t count = 0, n_loops = 0, total_bases = 0, total_words = 0;
    DevBuf<uint64_t> word_off;   // count + 1: unitig i occupies words [word_off[i], word_off[i+1])
    DevBuf<uint32_t> len;        // count
    DevBuf<uint64_t> words;      // packed 2-bit, same layout as reads
};

namespace sb200 {
// Grouping plan of the hash-sharded path on the staged kernels (count.cu shard_plan): the same on every rank
struct ShardPlan {
    int p = 0, s = 0;              // value-prefix bits of the fine group key; fine groups per coarse bin = 2^s
    uint32_t n_go = 0, n_co = 0;   // fine groups / coarse bins per owner
};
// Hash-sharded sender (PEER instances of the pass-1 kernels): the runs of owner o's coarse bins are stored straight into o's receive
// buffer — peer memory over NVLink (CUDA IPC mapping, or a plain pointer when the owner is this GPU) — instead of into a local send buffer
// that an all-to-all would then move: the exchange IS the pass-1 kernel's output.  rec[o] / pay[o] point at the place of THIS rank's run
// in owner o's buffer (the cursors are relative to it); bins are owner-major, n_co per owner.
struct SpPeers {
    uint64_t *rec[64];
    uint8_t *pay[64];
    uint32_t n_co;
};

// Sender state of a record exchange through peer stores (count.cu shard_peer_*): after the count pass every size is known, the
// owners' buffers are addressed, and pass 1 writes into them.
struct ShardPeerSend {
    unsigned k = 0, words = 0;
    bool double_palindromes = false, mask_payload = false, side_pay = false;
    int pshift = -1;
    ShardPlan pl;
    DevBuf<uint32_t> raw;     // [G][n_co] records per (owner, coarse bin)
    DevBuf<uint32_t> cur1;    // [G][n_co] pass-1 cursors, relative to this rank's run in the owner's buffer
};
// EarlyTipClipper in three steps (ext.cu): the kill list and the masks span the WHOLE index, the k-mers may be one GPU's shard
struct TipClipState {
    DevBuf<uint8_t> kill;      // ext->size + 4: k-mers (MPHF index) to isolate
    DevBuf<uint8_t> tipped;    // 2 * kmers->size + 4: oriented local junctions that lost a tip
    DevBuf<unsigned long long> removed;
};
}  // namespace sb200

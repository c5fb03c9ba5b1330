// unitigs_walk.cuh — unbranching paths by direct walks, one thread per start edge (included by unitigs.cu).
//
// When no chain is longer than WALK_LIMIT vertices (every read set with sequencing errors: a branch every few bases),
// list ranking is overkill.  Every outgoing edge of every oriented junction is a work item, listed in the reference's
// discovery order (file order of the canonical k-mer, forward strand then reverse strand, A,C,G,T).  A thread follows its
// edge with one MPHF lookup per step — bit-vectors, rank samples and masks together are ~130 MB and stay L2-resident —
// exactly like the reference's ConstructSequenceWithEdge (debruijn_graph_constructor.hpp:236-245), but for all ~10^7
// start edges at once.  The first walk measures (length, end k-mer -> `!(s < !s)`); after two scans a second walk over the
// kept edges re-traces the path and emits the 2-bit packed sequence directly, 32 bases per stored word.  Only the masks,
// the MPHF and the junction's own k-mer are needed: nothing indexed by "all vertices" is built, which is also what lets
// several GPUs extract disjoint file ranges of junctions independently (tests/test_gpu_sharded.py).  A chain longer than
// WALK_LIMIT, or vertices no walk reached (perfect loops), send the whole extraction to the pointer-jumping path.
#pragma once

namespace sb200 {

constexpr uint32_t WALK_LIMIT = 1024;

template<int W>
__device__ __forceinline__ uint32_t walk_mask(const MphfDev &m, const uint8_t *__restrict__ masks, const uint64_t *y, int k) {
    bool minimal;
    uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
    uint32_t raw = __ldg(masks + idy);
    return minimal ? raw : mask_conj(raw);
}

// Work list of start edges, in the reference's discovery order: every outgoing edge of every oriented junction, t = 2 * (file index -
// first) + strand, elist[e] = (t << 2) | nucleotide.  Two passes over tiles of 1024 oriented vertices — count, scan of the tile counts,
// emit with a block scan inside the tile — so nothing of the size of the vertex set is written (the first version stored and scanned a
// degree per oriented vertex: 2.3 GB of traffic for a 29 MB list).  `file_masks` (masks in file order, a coalesced read) replaces the
// random masks[idx[.]] reads when the index-order array has not been edited by the tip clipper.
constexpr int EL_ITEMS = 4;
constexpr int EL_TILE = 256 * EL_ITEMS;

__device__ __forceinline__ uint32_t junction_out_mask(uint64_t t, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ masks,
                                                      const uint8_t *__restrict__ file_masks) {
    const uint32_t raw = file_masks ? (uint32_t) __ldg(file_masks + (t >> 1)) : (uint32_t) __ldg(masks + __ldg(idx + (t >> 1)));
    if (!mask_is_junction(raw)) return 0u;
    return ((t & 1) ? mask_conj(raw) : raw) & 15u;
}

__global__ void __launch_bounds__(256) junction_degree_kernel(uint64_t n, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ masks,
                                                             const uint8_t *__restrict__ file_masks, uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t sm[256 / 32 + 1];
    const uint64_t t0 = (uint64_t) blockIdx.x * EL_TILE + (uint64_t) threadIdx.x * EL_ITEMS;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < EL_ITEMS; ++i)
        if (t0 + i < 2 * n) c += (uint32_t) __popc(junction_out_mask(t0 + i, idx, masks, file_masks));
    uint32_t total;
    block_exclusive_scan<uint32_t, 256>(c, &total, sm);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) edge_list_kernel(uint64_t n, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ masks,
                                                       const uint8_t *__restrict__ file_masks, const uint32_t *__restrict__ tile_off,
                                                       uint32_t *__restrict__ elist) {
    __shared__ uint32_t sm[256 / 32 + 1];
    const uint64_t t0 = (uint64_t) blockIdx.x * EL_TILE + (uint64_t) threadIdx.x * EL_ITEMS;
    uint32_t mk[EL_ITEMS], c = 0;
#pragma unroll
    for (int i = 0; i < EL_ITEMS; ++i) {
        mk[i] = (t0 + i < 2 * n) ? junction_out_mask(t0 + i, idx, masks, file_masks) : 0u;
        c += (uint32_t) __popc(mk[i]);
    }
    uint32_t total;
    uint32_t o = tile_off[blockIdx.x] + block_exclusive_scan<uint32_t, 256>(c, &total, sm);
#pragma unroll
    for (int i = 0; i < EL_ITEMS; ++i) {
        uint32_t m = mk[i];
        while (m) {
            const uint32_t nuc = (uint32_t) __ffs((int) m) - 1u;
            m &= m - 1u;
            elist[o++] = ((uint32_t) (t0 + i) << 2) | nuc;
        }
    }
}

// Walks have very different lengths (1 .. 1000 steps): with one walk per thread a warp runs until its longest walk ends and
// most lanes idle (ncu, profiles/r1a: 1850 thread slots per lookup step).  All walk kernels are therefore PERSISTENT: a lane
// that finishes claims the next work item at once (one warp-aggregated atomic per refill), so every lane keeps a walk — and,
// for the link walks, a DRAM read — in flight.  Results are indexed by the work item, so the output order is unchanged.
constexpr uint32_t NO_WORK = 0xFFFFFFFFu;
constexpr uint32_t WORK_CHUNK = 256;   // items a warp takes from the global queue at a time
// Hands the next work items to the lanes that need one.  A warp draws chunks of WORK_CHUNK items from the global counter (one
// atomic per chunk: a per-refill atomic on a single address serialised the link walks, 3.6 -> 9.6 ms) and serves its lanes from
// the chunk; [cnext, cend) is the warp-uniform rest of the current chunk.
__device__ __forceinline__ uint32_t claim_work(bool need, uint32_t *__restrict__ counter, uint32_t n, uint32_t &cnext, uint32_t &cend) {
    const uint32_t m = __ballot_sync(0xffffffffu, need);
    if (!m) return NO_WORK;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (uint32_t) __popc(m & ((1u << lane) - 1u)), cnt = (uint32_t) __popc(m);
    const uint32_t avail = cend - cnext;
    uint32_t item = NO_WORK;
    if (rank < avail) item = cnext + rank;
    if (cnt > avail) {   // warp-uniform: the chunk runs out
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, WORK_CHUNK);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (rank >= avail) item = base + (rank - avail);
        cnext = base + (cnt - avail);
        cend = base + WORK_CHUNK;
    } else {
        cnext += cnt;
    }
    return (need && item < n) ? item : NO_WORK;
}

template<int W>
__global__ void __launch_bounds__(256) walk_measure_kernel(MphfDev m, const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ elist,
                                                          uint32_t n_e, const uint8_t *__restrict__ masks, uint32_t *__restrict__ elen,
                                                          uint32_t *__restrict__ kflag, unsigned long long *__restrict__ ewords,
                                                          unsigned long long *__restrict__ totals /* [0] chain vertices seen, [1] long chains, [4] kept bases */,
                                                          uint32_t *__restrict__ work) {
    unsigned long long chain_nodes = 0, kept_bases = 0;
    uint32_t too_long = 0;
    bool active = false, exhausted = false;
    uint32_t cnext = 0, cend = 0;
    uint32_t e = 0, c = 0, nn = 0, prev_first = 0;
    uint64_t x[W], y[W];
#pragma unroll
    for (int q = 0; q < W; ++q) { x[q] = 0; y[q] = 0; }
    while (true) {
        const bool need = !active && !exhausted;
        const uint32_t item = claim_work(need, work, n_e, cnext, cend);
        if (need) {
            if (item == NO_WORK) exhausted = true;
            else {
                e = item;
                const uint32_t code = elist[e];
                const uint32_t t = code >> 2;
                c = code & 3u;
                oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x);
                kmer_shl<W>(x, k, c, y);
                prev_first = kmer_base(x, 0);   // first base of the vertex before the current one
                nn = 1;
                active = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            const uint32_t mk = walk_mask<W>(m, masks, y, k);
            const bool junction = mask_is_junction(mk);
            if (junction || nn > WALK_LIMIT) {
                uint32_t len = 0;
                if (!junction) too_long += 1;
                else {
                    chain_nodes += nn - 1;
                    uint64_t rcn[W];
                    kmer_rc<W>(y, k, rcn);
                    const int cmp = kmer_lex_cmp<W>(x, rcn);
                    const bool keep = cmp > 0 || (cmp == 0 && c >= 3u - prev_first);   // tie: first edge nucleotide of the reverse path
                    if (keep) { len = nn; kept_bases += (unsigned long long) k + nn; }
                }
                elen[e] = len;
                kflag[e] = len ? 1u : 0u;
                ewords[e] = len ? (((unsigned long long) k + len + 31) >> 5) : 0ULL;
                active = false;
            } else {
                uint64_t z[W];
                prev_first = kmer_base(y, 0);
                kmer_shl<W>(y, k, nib_next(mk & 15u), z);
#pragma unroll
                for (int q = 0; q < W; ++q) y[q] = z[q];
                ++nn;
            }
        }
    }
    // one atomic per warp and counter
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        chain_nodes += __shfl_down_sync(0xffffffffu, chain_nodes, d);
        kept_bases += __shfl_down_sync(0xffffffffu, kept_bases, d);
        too_long += __shfl_down_sync(0xffffffffu, too_long, d);
    }
    // one set of global atomics per CTA (per-warp atomics on three addresses serialise in L2)
    __shared__ unsigned long long s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        if (chain_nodes) atomicAdd(&s_tot[0], chain_nodes);
        if (too_long) atomicAdd(&s_tot[1], (unsigned long long) too_long);
        if (kept_bases) atomicAdd(&s_tot[2], kept_bases);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_tot[0]) atomicAdd(&totals[0], s_tot[0]);
        if (s_tot[1]) atomicAdd(&totals[1], s_tot[1]);
        if (s_tot[2]) atomicAdd(&totals[4], s_tot[2]);
    }
}

__global__ void __launch_bounds__(256) kept_list_kernel(const uint32_t *__restrict__ kflag_scan, const uint32_t *__restrict__ elen, uint32_t n_e,
                                                       uint32_t *__restrict__ klist) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_e && elen[e]) klist[kflag_scan[e]] = e;
}

// ---- link table: the walks as pointer chasing ----------------------------------------------------------------------------
// With the whole k-mer table on this GPU every chain vertex (in = out = 1) gets one 32-bit word, indexed by its oriented
// MPHF position v = 2 * MPHF index + strand (strand 1 = the reverse complement of the stored k-mer):
//     bits 0-27  oriented MPHF position of its successor      bits 28-29  nucleotide appended by the step
//     bits 30-31 first base of the vertex itself (the reverse path's tie-breaker)        junctions hold LINK_JUNCTION.
// Building it costs one MPHF lookup per chain vertex in a fully parallel kernel; the walks then cost one 4-byte read per step
// instead of a canonicalisation, two XXH3 hashes, the level probes and a rank per step (ncu, profiles/r1a: the lookup walks
// executed 13.5 G warp instructions and pulled 25 GB through DRAM for 210 M steps).  The table lives in MPHF order, not file
// order, so that the successor's position IS the lookup result: a table in file order needed inv[lookup] per vertex, a random
// 4-byte read that DRAM serves as a 128-byte fetch (17 of the 25 GB links_kernel read, profiles/r1b); the price is that every
// vertex stores its word at a random place instead (the two strands of a k-mer are neighbours: one 8-byte store per k-mer).
constexpr uint32_t LINK_JUNCTION = 0xFFFFFFFFu;
constexpr uint32_t LINK_POS_MASK = 0x0FFFFFFFu;

template<int W>
__global__ void __launch_bounds__(256) links_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k, const uint32_t *__restrict__ idx,
                                                   const uint8_t *__restrict__ masks,
                                                   const uint8_t *__restrict__ file_masks /* masks in file order, or nullptr */,
                                                   uint32_t *__restrict__ link) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    const uint32_t me = __ldg(idx + (t >> 1));
    const uint32_t raw = file_masks ? (uint32_t) __ldg(file_masks + (t >> 1)) : (uint32_t) __ldg(masks + me);
    const int strand = (int) (t & 1);
    uint32_t out = LINK_JUNCTION;
    if (!mask_is_junction(raw)) {
        const uint32_t c = nib_next((strand ? mask_conj(raw) : raw) & 15u);
        uint64_t x[W], y[W];
        oriented_kmer<W>(kmers, t >> 1, strand, k, x);
        kmer_shl<W>(x, k, c, y);
        bool minimal;
        const uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
        out = (2u * idy + (minimal ? 0u : 1u)) | (c << 28) | (kmer_base(x, 0) << 30);
    }
    link[2u * me + (uint32_t) strand] = out;
}

// (the link walks stay one thread per start edge: they are bound by random DRAM reads, and the persistent form's extra
// registers and refill logic cost more than the idle lanes it removes: 3.6 + 1.5 ms against 4.6 + 1.8 ms)
template<int W>
__global__ void __launch_bounds__(256) walk_measure_links_kernel(MphfDev m, const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ elist,
                                                                uint32_t n_e, const uint32_t *__restrict__ inv, const uint32_t *__restrict__ link,
                                                                uint32_t *__restrict__ elen, uint32_t *__restrict__ efirst,
                                                                unsigned long long *__restrict__ ebases, uint32_t *__restrict__ kflag,
                                                                unsigned long long *__restrict__ ewords, unsigned long long *__restrict__ totals) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long chain_nodes = 0, kept_bases = 0;
    uint32_t too_long = 0;
    if (e < n_e) {
        uint32_t code = elist[e];
        uint32_t t = code >> 2, c = code & 3u;
        uint64_t x[W], y[W];
        oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x);
        kmer_shl<W>(x, k, c, y);
        bool minimal;
        const uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
        uint32_t v = 2u * idy + (minimal ? 0u : 1u);   // oriented MPHF position: the index space of the link table
        efirst[e] = v;
        uint32_t prev_first = kmer_base(x, 0);   // first base of the vertex before the current one
        uint32_t nn = 1;
        uint32_t L = __ldg(link + v);
        unsigned long long bases = 0;   // the first 32 nucleotides the chain appends after c: the emitting walk of a short path needs no link reads
        while (L != LINK_JUNCTION) {
            if (nn > WALK_LIMIT) { too_long = 1; break; }
            if (nn <= 32) bases |= (unsigned long long) ((L >> 28) & 3u) << (2 * (nn - 1));
            prev_first = L >> 30;
            v = L & LINK_POS_MASK;
            L = __ldg(link + v);
            ++nn;
        }
        uint32_t len = 0;
        if (!too_long) {
            chain_nodes = nn - 1;
            uint64_t rcn[W];
            oriented_kmer<W>(kmers, __ldg(inv + (v >> 1)), (v & 1) ? 0 : 1, k, rcn);   // rc of the end vertex (file position of its index)
            int cmp = kmer_lex_cmp<W>(x, rcn);
            bool keep = cmp > 0 || (cmp == 0 && c >= 3u - prev_first);   // tie: first edge nucleotide of the reverse path
            if (keep) { len = nn; kept_bases = (unsigned long long) k + nn; }
        }
        elen[e] = len;
        if (len) ebases[e] = bases;
        kflag[e] = len ? 1u : 0u;
        ewords[e] = len ? (((unsigned long long) k + len + 31) >> 5) : 0ULL;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        chain_nodes += __shfl_down_sync(0xffffffffu, chain_nodes, d);
        kept_bases += __shfl_down_sync(0xffffffffu, kept_bases, d);
        too_long += __shfl_down_sync(0xffffffffu, too_long, d);
    }
    // one set of global atomics per CTA (per-warp atomics on three addresses serialise in L2)
    __shared__ unsigned long long s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        if (chain_nodes) atomicAdd(&s_tot[0], chain_nodes);
        if (too_long) atomicAdd(&s_tot[1], (unsigned long long) too_long);
        if (kept_bases) atomicAdd(&s_tot[2], kept_bases);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_tot[0]) atomicAdd(&totals[0], s_tot[0]);
        if (s_tot[1]) atomicAdd(&totals[1], s_tot[1]);
        if (s_tot[2]) atomicAdd(&totals[4], s_tot[2]);
    }
}

template<int W>
__global__ void __launch_bounds__(256) walk_emit_links_kernel(const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ elist,
                                                             const uint32_t *__restrict__ klist, uint32_t n_kept, const uint32_t *__restrict__ link,
                                                             const uint32_t *__restrict__ elen, const uint32_t *__restrict__ efirst,
                                                             const unsigned long long *__restrict__ ebases,
                                                             const unsigned long long *__restrict__ ewords_scan, uint32_t *__restrict__ seq_len,
                                                             uint64_t *__restrict__ seq_word_off, uint64_t *__restrict__ out_words) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_kept) return;
    uint32_t e = klist[q];
    uint32_t code = elist[e];
    uint32_t t = code >> 2, c = code & 3u;
    const uint32_t nn = elen[e];
    const unsigned long long woff = ewords_scan[e];
    const uint32_t L = (uint32_t) k + nn;
    seq_len[q] = L;
    seq_word_off[q] = woff;
    uint64_t x[W];
    oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x);
    uint32_t pos = 0;
    uint64_t acc = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if ((w + 1) * 32 <= k) { out_words[woff + w] = x[w]; pos = (w + 1) * 32; }
    }
    if (pos < (uint32_t) k) { acc = x[W - 1]; pos = (uint32_t) k; }   // padding bits of x are zero
    auto push = [&](uint32_t base) {
        acc |= (uint64_t) base << (2 * (pos & 31));
        ++pos;
        if ((pos & 31) == 0) { out_words[woff + (pos >> 5) - 1] = acc; acc = 0; }
    };
    push(c);
    if (nn <= 33) {   // short path: the measuring walk left its nucleotides
        unsigned long long bases = ebases[e];
        for (uint32_t s = 1; s < nn; ++s) { push((uint32_t) bases & 3u); bases >>= 2; }
    } else {
        uint32_t v = efirst[e];
        for (uint32_t s = 1; s < nn; ++s) {
            const uint32_t lw = __ldg(link + v);
            push((lw >> 28) & 3u);
            v = lw & LINK_POS_MASK;
        }
    }
    if (pos & 31) out_words[woff + (pos >> 5)] = acc;
}

// second walk, kept edges only: packed output.  Bases are accumulated 32 to a word in a register and stored once per word.
template<int W>
__global__ void __launch_bounds__(256) walk_emit_kernel(MphfDev m, const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ elist,
                                                       const uint32_t *__restrict__ klist, uint32_t n_kept, const uint8_t *__restrict__ masks,
                                                       const uint32_t *__restrict__ elen, const unsigned long long *__restrict__ ewords_scan,
                                                       uint32_t *__restrict__ seq_len, uint64_t *__restrict__ seq_word_off,
                                                       uint64_t *__restrict__ out_words, uint32_t *__restrict__ work) {
    bool active = false, exhausted = false;
    uint32_t cnext = 0, cend = 0;
    uint32_t left = 0, pos = 0;   // left = lookup steps still to take
    uint64_t acc = 0;
    unsigned long long woff = 0;
    uint64_t y[W];
#pragma unroll
    for (int w = 0; w < W; ++w) y[w] = 0;
    while (true) {
        const bool need = !active && !exhausted;
        const uint32_t q = claim_work(need, work, n_kept, cnext, cend);
        if (need) {
            if (q == NO_WORK) exhausted = true;
            else {
                const uint32_t e = klist[q];
                const uint32_t code = elist[e];
                const uint32_t t = code >> 2, c = code & 3u;
                const uint32_t nn = elen[e];
                woff = ewords_scan[e];
                seq_len[q] = (uint32_t) k + nn;
                seq_word_off[q] = woff;
                uint64_t x[W];
                oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x);
                pos = 0;
                acc = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    if ((w + 1) * 32 <= k) { out_words[woff + w] = x[w]; pos = (w + 1) * 32; }
                }
                if (pos < (uint32_t) k) { acc = x[W - 1]; pos = (uint32_t) k; }   // padding bits of x are zero
                acc |= (uint64_t) c << (2 * (pos & 31));
                ++pos;
                if ((pos & 31) == 0) { out_words[woff + (pos >> 5) - 1] = acc; acc = 0; }
                kmer_shl<W>(x, k, c, y);
                left = nn - 1;
                active = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            if (left == 0) {
                if (pos & 31) out_words[woff + (pos >> 5)] = acc;
                active = false;
            } else {
                const uint32_t mk = walk_mask<W>(m, masks, y, k);
                const uint32_t b = nib_next(mk & 15u);
                acc |= (uint64_t) b << (2 * (pos & 31));
                ++pos;
                if ((pos & 31) == 0) { out_words[woff + (pos >> 5) - 1] = acc; acc = 0; }
                uint64_t z[W];
                kmer_shl<W>(y, k, b, z);
#pragma unroll
                for (int w = 0; w < W; ++w) y[w] = z[w];
                --left;
            }
        }
    }
}

__global__ void __launch_bounds__(256) count_nonjunction_kernel(const uint8_t *__restrict__ masks, uint64_t n, unsigned long long *__restrict__ total) {
    __shared__ uint32_t sm[256 / 32 + 1];
    uint32_t c = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        c += mask_is_junction(masks[i]) ? 0u : 1u;
    uint32_t block_total;
    block_exclusive_scan<uint32_t, 256>(c, &block_total, sm);
    if (threadIdx.x == 0 && block_total) atomicAdd(total, (unsigned long long) block_total);   // one atomic per CTA
}

// Result of the measuring walk over the junctions of file range [first, last)
struct WalkStats {
    unsigned long long chain_vertices = 0, long_chains = 0, words = 0, kept_bases = 0;
    uint32_t n_edges = 0, n_kept = 0;
};

// Returns nullptr when the direct walks do not apply (a chain longer than WALK_LIMIT, or — when checking the whole set —
// perfect loops to collect): the caller then runs the pointer-jumping path.  [first, last) restricts the junctions to a
// file range of k-mers; `stats_out` reports what was seen either way.
template<int W>
static sb200_unitigs *unitigs_walk_w(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int check_loops,
                                     uint64_t first, uint64_t last, WalkStats *stats_out) {
    uint64_t n = kmers->size;
    int k = (int) kmers->k;
    MphfDev m = mphf_dev(mphf);
    uint64_t n_range = last - first;
    if (2 * n_range >= (1ull << 30)) {   // the start-edge codes (t << 2 | nucleotide) are 32-bit
        if (!stats_out) return nullptr;   // whole set: the caller falls through to the pointer-jumping path
        SB200_REQUIRE(false, "more than 2^29 k-mers in one extraction range: shard the input");
    }
    WalkStats st;
    uint64_t nt = 2 * n_range;
    const uint32_t n_tiles = (uint32_t) div_up(std::max<uint64_t>(nt, 1), EL_TILE);
    DevBuf<uint32_t> deg(ctx, (uint64_t) n_tiles + 1);   // start edges per tile of EL_TILE oriented vertices, then their exclusive scan
    // masks in file order (a coalesced read) are valid as long as tip clipping has not edited the index-order array
    const uint8_t *fm = (kmers->masks_file.p && !ext->masks_edited) ? kmers->masks_file.p + first : nullptr;
    DevBuf<uint32_t> tot32(ctx, 2);
    DevBuf<unsigned long long> totals(ctx, 6);   // [0] chain vertices seen [1] long chains [2] total words [3] non-junction k-mers [4] bases
    totals.zero();
    DevBuf<uint32_t> work(ctx, 4);               // work-queue heads of the persistent walk kernels: [0] measure, [1] emit
    work.zero();
    const unsigned walk_grid = (unsigned) ctx->num_sms * 6;
    if (nt) LAUNCH(ctx, junction_degree_kernel, n_tiles, 256, 0, n_range, ext->idx.p + first, ext->masks.p, fm, deg.p);
    else deg.zero();
    exclusive_scan<uint32_t>(ctx, deg.p, n_tiles, tot32.p);
    ctx->fetch(&st.n_edges, tot32.p, 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    uint32_t n_e = st.n_edges;
    DevBuf<uint32_t> elist(ctx, (uint64_t) n_e + 1), elen(ctx, (uint64_t) n_e + 1), kflag(ctx, (uint64_t) n_e + 1);
    DevBuf<unsigned long long> ewords(ctx, (uint64_t) n_e + 1);
    const uint64_t *kbase = kmers->data.p + first * W;
    // pointer-chasing walks need the whole k-mer table and the inverse permutation on this GPU; a shard (or a table too
    // large for 28-bit positions) walks by MPHF lookups instead
    const bool use_links = !ctx->no_links && first == 0 && last == n && ext->n_local == ext->size && ext->inv.p != nullptr && 2 * n <= LINK_POS_MASK;
    DevBuf<uint32_t> link, efirst;
    DevBuf<unsigned long long> ebases;
    if (n_e) {
        LAUNCH(ctx, edge_list_kernel, n_tiles, 256, 0, n_range, ext->idx.p + first, ext->masks.p, fm, deg.p, elist.p);
        if (use_links) {
            link.alloc(ctx, 2 * n);
            efirst.alloc(ctx, (uint64_t) n_e + 1);
            ebases.alloc(ctx, (uint64_t) n_e + 1);
            LAUNCH(ctx, links_kernel<W>, div_up(2 * n, 256), 256, 0, m, kmers->data.p, n, k, ext->idx.p, ext->masks.p, fm, link.p);
            LAUNCH(ctx, walk_measure_links_kernel<W>, div_up(n_e, 256), 256, 0, m, kbase, k, elist.p, n_e, ext->inv.p, link.p, elen.p, efirst.p,
                   ebases.p, kflag.p, ewords.p, totals.p);
        } else {
            LAUNCH(ctx, walk_measure_kernel<W>, walk_grid, 256, 0, m, kbase, k, elist.p, n_e, ext->masks.p, elen.p, kflag.p, ewords.p, totals.p,
                   work.p);
        }
    }
    if (check_loops) LAUNCH(ctx, count_nonjunction_kernel, (unsigned) ctx->num_sms * 8, 256, 0, ext->masks.p, n, totals.p + 3);
    exclusive_scan<uint32_t>(ctx, kflag.p, n_e, tot32.p + 1);
    exclusive_scan<unsigned long long>(ctx, ewords.p, n_e, totals.p + 2);
    unsigned long long th[6];
    ctx->fetch(th, totals.p, 48);
    ctx->fetch(&st.n_kept, tot32.p + 1, 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    st.chain_vertices = th[0]; st.long_chains = th[1]; st.words = th[2]; st.kept_bases = th[4];
    if (stats_out) *stats_out = st;
    if (st.long_chains != 0) return nullptr;                          // some chain is too long for a sequential walk
    if (check_loops && th[0] != 2 * th[3]) return nullptr;            // vertices no walk reached: perfect loops exist
    sb200_unitigs *out = new sb200_unitigs();
    out->ctx = ctx; out->k = (unsigned) k; out->count = st.n_kept; out->n_loops = 0; out->total_words = st.words;
    out->total_bases = st.kept_bases;
    out->len.alloc(ctx, (uint64_t) st.n_kept + 1);
    out->word_off.alloc(ctx, (uint64_t) st.n_kept + 1);
    out->words.alloc(ctx, st.words + 1);
    if (st.n_kept) {
        DevBuf<uint32_t> klist(ctx, st.n_kept);
        LAUNCH(ctx, kept_list_kernel, div_up(n_e, 256), 256, 0, kflag.p, elen.p, n_e, klist.p);
        if (use_links)
            LAUNCH(ctx, walk_emit_links_kernel<W>, div_up(st.n_kept, 256), 256, 0, kbase, k, elist.p, klist.p, st.n_kept, link.p, elen.p, efirst.p,
                   ebases.p, ewords.p, out->len.p, out->word_off.p, out->words.p);
        else
            LAUNCH(ctx, walk_emit_kernel<W>, walk_grid, 256, 0, m, kbase, k, elist.p, klist.p, st.n_kept, ext->masks.p, elen.p, ewords.p,
                   out->len.p, out->word_off.p, out->words.p, work.p + 1);
    }
    uint64_t tw = st.words;
    CUDA_CHECK(cudaMemcpyAsync(out->word_off.p + st.n_kept, &tw, 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return out;
}

}  // namespace sb200

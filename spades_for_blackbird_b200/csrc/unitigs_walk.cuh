// unitigs_walk.cuh — unbranching paths by direct walks, one thread per start edge (included by unitigs.cu).
//
// When no chain is longer than WALK_LIMIT vertices (every read set with sequencing errors: a branch every few bases),
// list ranking is overkill.  Every outgoing edge of every oriented junction is a work item, listed in the reference's
// discovery order (file order of the canonical k-mer, forward strand then reverse strand, A,C,G,T).  A thread follows its
// edge with one MPHF lookup per step — through the walk blocks (mphf.cuh): index bits, rank and masks of a step in one 128-byte line —
// exactly like the reference's ConstructSequenceWithEdge (debruijn_graph_constructor.hpp:236-245), but for all ~10^7
// start edges at once.  The first walk measures (length, end k-mer -> `!(s < !s)`) and captures the path's nucleotides; after two scans
// the kept edges emit the 2-bit packed sequence from the capture (a second walk only for paths beyond the capture).  Only the masks,
// the MPHF and the junction's own k-mer are needed: nothing indexed by "all vertices" is built, which is also what lets
// several GPUs extract disjoint file ranges of junctions independently (tests/test_gpu_sharded.py).  A chain longer than
// WALK_LIMIT, or vertices no walk reached (perfect loops), send the whole extraction to the pointer-jumping path.
#pragma once

namespace sb200 {

constexpr uint32_t WALK_LIMIT = 1024;

template<int W>
__device__ __forceinline__ uint32_t walk_mask(const MphfDev &m, const uint8_t *__restrict__ masks, const uint64_t *y, int k) {
    uint64_t c[W];
    const bool minimal = kmer_canonical<W>(y, k, c);
    const uint32_t raw = m.wblk ? mphf_lookup_mask<W>(m, masks, c) : (uint32_t) __ldg(masks + (uint32_t) mphf_lookup<W>(m, c));
    return minimal ? raw : mask_conj(raw);
}

// Builds the walk blocks (mphf.cuh) from the complete index and the masks in index order: one thread per 16-byte piece of a line.
__global__ void __launch_bounds__(256) walk_blocks_kernel(const uint64_t *__restrict__ bits, const uint32_t *__restrict__ pc_scan, uint64_t words,
                                                         const uint8_t *__restrict__ masks, uint64_t n_masks, uint64_t n_lines, uint4 *__restrict__ out) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t line = t >> 3;
    if (line >= n_lines) return;
    const uint32_t piece = (uint32_t) t & 7u;
    const uint64_t w0 = line * WB_WORDS;
    const uint32_t rank = __ldg(pc_scan + w0);
    const uint64_t wend = w0 + WB_WORDS < words ? w0 + WB_WORDS : words;
    const uint32_t cnt = __ldg(pc_scan + wend) - rank;   // pc_scan has words + 1 entries
    const bool spill = cnt > WB_MASKS;
    union { uint4 v; uint64_t q[2]; uint32_t d[4]; uint8_t b[16]; } u;
    u.v = make_uint4(0u, 0u, 0u, 0u);
    int mfirst = 0;   // first byte of this piece that holds a mask
    if (piece == 0) {
        u.q[0] = w0 < words ? __ldg(bits + w0) : 0ULL;
        u.q[1] = w0 + 1 < words ? __ldg(bits + w0 + 1) : 0ULL;
        mfirst = 16;
    } else if (piece == 1) {
        u.q[0] = w0 + 2 < words ? __ldg(bits + w0 + 2) : 0ULL;
        u.d[2] = rank | (spill ? 0x80000000u : 0u);
        mfirst = 12;
    }
    const int j0 = (int) piece * 16 - (int) (8 * WB_WORDS + 4);   // mask number of byte 0 of this piece
    if (!spill && j0 + 16 > 0 && j0 + mfirst < (int) cnt) {
        // bytes [o, o + 16) of the mask array through aligned 32-bit loads and funnel shifts (o may be negative in piece 1: those bytes
        // are overwritten below), then cut to the masks of this line
        const int64_t o = (int64_t) rank + j0;
        const int64_t a = o & ~(int64_t) 3;
        const uint32_t sh = (uint32_t) (o & 3) * 8;
        const uint32_t *mw = reinterpret_cast<const uint32_t *>(masks);
        const int64_t last = (int64_t) ((n_masks + 3) / 4);   // the array is padded to a multiple of 4 bytes
        uint32_t w[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int64_t wi = a / 4 + q;
            w[q] = (wi >= 0 && wi < last) ? __ldg(mw + wi) : 0u;
        }
        union { uint4 v; uint32_t d[4]; uint8_t b[16]; } mk;
#pragma unroll
        for (int q = 0; q < 4; ++q) mk.d[q] = __funnelshift_r(w[q], w[q + 1], sh);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i >= mfirst && (uint32_t) (j0 + i) < cnt) u.b[i] = mk.b[i];
    }
    out[t] = u.v;
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

// A lane's walk ends at a junction; what follows (orientation filter against the start k-mer, result stores) and the start of its next
// walk (work claim, k-mer load, first shift) are long code paths that few lanes need at a time: they run in SERVICE ROUNDS, once at
// least WALK_SERVICE_MIN lanes of the warp wait for one, so that the step code — the part every walk spends its time in — runs with most
// lanes active (ncu r2l: 17.9 of 32 threads active per instruction with a service check in every iteration).
constexpr int WALK_SERVICE_MIN = 8;

template<int W>
__global__ void __launch_bounds__(256, 5) walk_measure_kernel(MphfDev m, const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ elist,
                                                          uint32_t n_e, const uint8_t *__restrict__ masks, uint32_t *__restrict__ elen,
                                                          uint32_t *__restrict__ kflag, unsigned long long *__restrict__ ewords,
                                                          unsigned long long *__restrict__ totals /* [0] chain vertices seen, [1] long chains, [4] kept bases, [5] kept paths longer than the capture */,
                                                          uint32_t *__restrict__ work, unsigned long long *__restrict__ escr, uint32_t cw) {
    uint32_t chain_nodes = 0, kept_bases = 0;   // per lane: far below 2^32
    uint32_t too_long = 0, uncaptured = 0;
    unsigned long long cap = 0;
    bool active = false, pending = false, exhausted = false;   // pending: the walk ended, its result is not written yet
    uint32_t cnext = 0, cend = 0;
    uint32_t e = 0, nn = 0, prev_first = 0;
    uint64_t y[W];
#pragma unroll
    for (int q = 0; q < W; ++q) y[q] = 0;
    while (true) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (__popc(idle) >= WALK_SERVICE_MIN && __any_sync(0xffffffffu, pending || (!active && !exhausted))) {
            if (pending) {
                uint32_t len = 0;
                if (nn == 0) too_long += 1;
                else {
                    chain_nodes += nn - 1;
                    const uint32_t code = elist[e];
                    const uint32_t t = code >> 2, c = code & 3u;
                    uint64_t x[W], rcn[W];
                    oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x);
                    kmer_rc<W>(y, k, rcn);
                    const int cmp = kmer_lex_cmp<W>(x, rcn);
                    const bool keep = cmp > 0 || (cmp == 0 && c >= 3u - prev_first);   // tie: first edge nucleotide of the reverse path
                    if (keep) {
                        len = nn;
                        kept_bases += (uint32_t) k + nn;
                        const uint32_t s = nn - 1;   // nucleotides appended after c
                        if (s > 32u * cw) uncaptured = 1;
                        else if (s & 31u) escr[(uint64_t) e * cw + (s >> 5)] = cap;
                    }
                }
                elen[e] = len;
                kflag[e] = len ? 1u : 0u;
                ewords[e] = len ? (((unsigned long long) k + len + 31) >> 5) : 0ULL;
                pending = false;
            }
            const bool need = !active && !exhausted;
            const uint32_t item = claim_work(need, work, n_e, cnext, cend);
            if (need) {
                if (item == NO_WORK) exhausted = true;
                else {
                    e = item;
                    const uint32_t code = elist[e];
                    const uint32_t t = code >> 2;
                    uint64_t x[W];
                    oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x);
                    kmer_shl<W>(x, k, code & 3u, y);
                    prev_first = kmer_base(x, 0);   // first base of the vertex before the current one
                    nn = 1;
                    cap = 0;
                    active = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            const uint32_t mk = walk_mask<W>(m, masks, y, k);
            const bool junction = mask_is_junction(mk);
            if (junction || nn > WALK_LIMIT) {
                if (!junction) nn = 0;   // marks a chain beyond the limit (a finished walk has nn >= 1)
                active = false;
                pending = true;
            } else {
                uint64_t z[W];
                prev_first = kmer_base(y, 0);
                const uint32_t b = nib_next(mk & 15u);
                // the nucleotides the chain appends after c are left for the emitting pass: 32 to a word, `cw` words per start edge
                const uint32_t s = nn - 1;
                if ((s >> 5) < cw) {
                    cap |= (unsigned long long) b << (2 * (s & 31u));
                    if ((s & 31u) == 31u) { escr[(uint64_t) e * cw + (s >> 5)] = cap; cap = 0; }
                }
                kmer_shl<W>(y, k, b, z);
#pragma unroll
                for (int q = 0; q < W; ++q) y[q] = z[q];
                ++nn;
            }
        }
    }
    // one atomic per warp and counter
    unsigned long long chain_sum = chain_nodes, kept_sum = kept_bases;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        chain_sum += __shfl_down_sync(0xffffffffu, chain_sum, d);
        kept_sum += __shfl_down_sync(0xffffffffu, kept_sum, d);
        too_long += __shfl_down_sync(0xffffffffu, too_long, d);
    }
    if (__any_sync(0xffffffffu, uncaptured != 0) && (threadIdx.x & 31) == 0) atomicAdd(&totals[5], 1ULL);   // rare
    // one set of global atomics per CTA (per-warp atomics on three addresses serialise in L2)
    __shared__ unsigned long long s_tot[3];
    if (threadIdx.x < 3) s_tot[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        if (chain_sum) atomicAdd(&s_tot[0], chain_sum);
        if (too_long) atomicAdd(&s_tot[1], (unsigned long long) too_long);
        if (kept_sum) atomicAdd(&s_tot[2], kept_sum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_tot[0]) atomicAdd(&totals[0], s_tot[0]);
        if (s_tot[1]) atomicAdd(&totals[1], s_tot[1]);
        if (s_tot[2]) atomicAdd(&totals[4], s_tot[2]);
    }
}

// Emitting pass over the kept edges when the measuring walk captured every kept path: no lookups, no link reads — the start k-mer, the
// edge nucleotide and the captured words, re-aligned to the output's word grid.
template<int W>
__global__ void __launch_bounds__(256) walk_emit_captured_kernel(const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ elist,
                                                                const uint32_t *__restrict__ klist, uint32_t n_kept, const uint32_t *__restrict__ elen,
                                                                const unsigned long long *__restrict__ escr, uint32_t cw,
                                                                const unsigned long long *__restrict__ ewords_scan, uint32_t *__restrict__ seq_len,
                                                                uint64_t *__restrict__ seq_word_off, uint64_t *__restrict__ out_words) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_kept) return;
    const uint32_t e = klist[q];
    const uint32_t code = elist[e];
    const uint32_t t = code >> 2, c = code & 3u;
    const uint32_t nn = elen[e];
    const unsigned long long woff = ewords_scan[e];
    seq_len[q] = (uint32_t) k + nn;
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
    acc |= (uint64_t) c << (2 * (pos & 31));
    ++pos;
    if ((pos & 31) == 0) { out_words[woff + (pos >> 5) - 1] = acc; acc = 0; }
    // the captured nucleotides start at output position k + 1: shift every captured word by that offset
    const uint32_t n_cap = nn - 1, sh = 2 * (pos & 31);
    const unsigned long long *src = escr + (uint64_t) e * cw;
    for (uint32_t j = 0; 32 * j < n_cap; ++j) {
        uint64_t wv = src[j];
        const uint32_t cnt = n_cap - 32 * j < 32 ? n_cap - 32 * j : 32;   // nucleotides in this word
        if (cnt < 32) wv &= (1ULL << (2 * cnt)) - 1ULL;
        acc |= wv << sh;
        const uint32_t room = 32 - (pos & 31);
        if (cnt >= room) {
            out_words[woff + (pos >> 5)] = acc;
            acc = sh ? wv >> (64 - sh) : 0ULL;
        }
        pos += cnt;
    }
    if (pos & 31) out_words[woff + (pos >> 5)] = acc;
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
                                                                unsigned long long *__restrict__ escr, uint32_t cw, uint32_t *__restrict__ kflag,
                                                                unsigned long long *__restrict__ ewords, unsigned long long *__restrict__ totals) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long chain_nodes = 0, kept_bases = 0;
    uint32_t too_long = 0, uncaptured = 0;
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
        unsigned long long cap = 0;   // the nucleotides the chain appends after c, 32 to a word, `cw` words per start edge: the emitting pass reads no links
        while (L != LINK_JUNCTION) {
            if (nn > WALK_LIMIT) { too_long = 1; break; }
            const uint32_t s = nn - 1;
            if ((s >> 5) < cw) {
                cap |= (unsigned long long) ((L >> 28) & 3u) << (2 * (s & 31u));
                if ((s & 31u) == 31u) { escr[(uint64_t) e * cw + (s >> 5)] = cap; cap = 0; }
            }
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
            if (keep) {
                len = nn;
                kept_bases = (unsigned long long) k + nn;
                const uint32_t s = nn - 1;
                if (s > 32u * cw) uncaptured = 1;
                else if (s & 31u) escr[(uint64_t) e * cw + (s >> 5)] = cap;
            }
        }
        elen[e] = len;
        kflag[e] = len ? 1u : 0u;
        ewords[e] = len ? (((unsigned long long) k + len + 31) >> 5) : 0ULL;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        chain_nodes += __shfl_down_sync(0xffffffffu, chain_nodes, d);
        kept_bases += __shfl_down_sync(0xffffffffu, kept_bases, d);
        too_long += __shfl_down_sync(0xffffffffu, too_long, d);
    }
    if (__any_sync(0xffffffffu, uncaptured != 0) && (threadIdx.x & 31) == 0) atomicAdd(&totals[5], 1ULL);   // rare
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
                                                             const unsigned long long *__restrict__ escr, uint32_t cw,
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
        unsigned long long bases = escr[(uint64_t) e * cw];
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
    // 16 masks per load (cudaMalloc'ed arrays are 256-byte aligned), the tail byte by byte
    const uint64_t n16 = n / 16, tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t) gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(masks) & 15u) == 0;
    if (aligned) {
        for (uint64_t i = tid; i < n16; i += nth) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(masks) + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int b = 0; b < 4; ++b) c += mask_is_junction((w[q] >> (8 * b)) & 255u) ? 0u : 1u;
        }
    }
    for (uint64_t i = (aligned ? n16 * 16 : 0) + tid; i < n; i += nth) c += mask_is_junction(masks[i]) ? 0u : 1u;
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
    // persistent walk kernels: exactly the CTAs that are resident at once
    int per_sm_measure = 4, per_sm_emit = 4;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_measure, walk_measure_kernel<W>, 256, 0));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_emit, walk_emit_kernel<W>, 256, 0));
    const unsigned walk_grid = (unsigned) (ctx->num_sms * std::max(per_sm_measure, 1)), emit_grid = (unsigned) (ctx->num_sms * std::max(per_sm_emit, 1));
    if (nt) LAUNCH(ctx, junction_degree_kernel, n_tiles, 256, 0, n_range, ext->idx.p + first, ext->masks.p, fm, deg.p);
    else deg.zero();
    exclusive_scan<uint32_t>(ctx, deg.p, n_tiles, tot32.p);
    ctx->fetch(&st.n_edges, tot32.p, 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    uint32_t n_e = st.n_edges;
    DevBuf<uint32_t> elist(ctx, (uint64_t) n_e + 1), elen(ctx, (uint64_t) n_e + 1), kflag(ctx, (uint64_t) n_e + 1);
    DevBuf<unsigned long long> ewords(ctx, (uint64_t) n_e + 1);
    const uint64_t *kbase = kmers->data.p + first * W;
    // Default: one lookup through the walk blocks per step (5.5 ms at config 2).  The link-table walks (SB200_LINKS=1: 3.7 ms of lookups
    // up front + 4.1 ms of pointer chasing) need the whole k-mer table and the inverse permutation on this GPU.
    const bool use_links = ctx->links && first == 0 && last == n && ext->n_local == ext->size && ext->inv.p != nullptr && 2 * n <= LINK_POS_MASK;
    DevBuf<uint32_t> link, efirst;
    // capture buffer of the measuring walks: `cw` words (32 nucleotides each) per start edge, the whole WALK_LIMIT when that fits 16 GB
    uint32_t cw = WALK_LIMIT / 32;
    while (cw > 1 && (uint64_t) n_e * cw * 8 > (16ull << 30)) cw >>= 1;
    if (ctx->walk_capture_words) cw = (uint32_t) std::min<size_t>(cw, ctx->walk_capture_words);
    DevBuf<unsigned long long> escr(ctx, (uint64_t) n_e * cw + 1);
    DevBuf<uint4> wblk;
    if (n_e && !use_links && mphf->pc_scan.p && !ctx->no_place && !ctx->no_walk_blocks) {
        // lookup walks: index bits, rank and masks of a step in one 128-byte line (mphf.cuh)
        const uint64_t n_lines = div_up(mphf->total_words, WB_WORDS);
        wblk.alloc(ctx, n_lines * 8);
        LAUNCH(ctx, walk_blocks_kernel, div_up(n_lines * 8, 256), 256, 0, mphf->bits.p, mphf->pc_scan.p, mphf->total_words, ext->masks.p, ext->size, n_lines, wblk.p);
        m.wblk = reinterpret_cast<const uint64_t *>(wblk.p);
    }
    if (n_e) {
        LAUNCH(ctx, edge_list_kernel, n_tiles, 256, 0, n_range, ext->idx.p + first, ext->masks.p, fm, deg.p, elist.p);
        if (use_links) {
            link.alloc(ctx, 2 * n);
            efirst.alloc(ctx, (uint64_t) n_e + 1);
            LAUNCH(ctx, links_kernel<W>, div_up(2 * n, 256), 256, 0, m, kmers->data.p, n, k, ext->idx.p, ext->masks.p, fm, link.p);
            LAUNCH(ctx, walk_measure_links_kernel<W>, div_up(n_e, 256), 256, 0, m, kbase, k, elist.p, n_e, ext->inv.p, link.p, elen.p, efirst.p,
                   escr.p, cw, kflag.p, ewords.p, totals.p);
        } else {
            LAUNCH(ctx, walk_measure_kernel<W>, walk_grid, 256, 0, m, kbase, k, elist.p, n_e, ext->masks.p, elen.p, kflag.p, ewords.p, totals.p,
                   work.p, escr.p, cw);
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
        if (th[5] == 0)   // every kept path was captured by the measuring walk
            LAUNCH(ctx, walk_emit_captured_kernel<W>, div_up(st.n_kept, 256), 256, 0, kbase, k, elist.p, klist.p, st.n_kept, elen.p, escr.p, cw,
                   ewords.p, out->len.p, out->word_off.p, out->words.p);
        else if (use_links)
            LAUNCH(ctx, walk_emit_links_kernel<W>, div_up(st.n_kept, 256), 256, 0, kbase, k, elist.p, klist.p, st.n_kept, link.p, elen.p, efirst.p,
                   escr.p, cw, ewords.p, out->len.p, out->word_off.p, out->words.p);
        else
            LAUNCH(ctx, walk_emit_kernel<W>, emit_grid, 256, 0, m, kbase, k, elist.p, klist.p, st.n_kept, ext->masks.p, elen.p, ewords.p,
                   out->len.p, out->word_off.p, out->words.p, work.p + 1);
    }
    uint64_t tw = st.words;
    CUDA_CHECK(cudaMemcpyAsync(out->word_off.p + st.n_kept, &tw, 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return out;
}

}  // namespace sb200

// unitigs.cu — unbranching paths (the condensed de Bruijn graph's edges) by pointer-jumping list ranking.
//
// Replaces UnbranchingPathExtractor::ExtractUnbranchingPathsAndLoops
// (C/assembly_graph/construction/debruijn_graph_constructor.hpp:182-388): the reference walks from every outgoing
// edge of every junction one MPHF lookup at a time and appends one nucleotide per step; here
//   1. every oriented non-junction k-mer knows its successor vertex (recorded for free while the masks were filled,
//      ext.cu; recomputed with one lookup per vertex only after tip clipping edited the masks),
//   2. pointer jumping gives each of them the last vertex of its chain and the distance to it (jump_kernel; pointer,
//      distance and a "finished" flag share one 64-bit word, so in-place updates are always observed consistently),
//   3. the junctions with outgoing edges are compacted into a work list; every start edge, enumerated in the
//      reference's order (file order of the canonical k-mer, forward strand then reverse strand, A,C,G,T), gets its
//      path length and end junction from step 2 and decides the orientation filter `!(s < !s)` from the two end k-mers
//      (ties by the first edge nucleotides),
//   4. a scan assigns output slots, heads write k+1 bases, every chain vertex writes its one base in parallel.
// Output order is therefore the reference's own (paths in discovery order, then perfect loops).
// Perfect loops (cycles without a junction) are what pointer jumping cannot finish; they are rare and handled by a
// single-thread kernel that restates CollectLoops/ConstructLoopFromVertex/SplitLoop (:248-265,308-344) literally.
#include "common.cuh"
#include "kmer_ops.cuh"
#include "kmer_set.cuh"
#include "mphf.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

MphfDev mphf_dev(const sb200_mphf *m);

constexpr uint32_t NONE = 0xFFFFFFFFu;

__device__ __forceinline__ bool nib_unique(uint32_t nib) { return nib && !(nib & (nib - 1)); }
__device__ __forceinline__ bool mask_is_junction(uint32_t m) { return !nib_unique(m & 15u) || !nib_unique(m >> 4); }
__device__ __forceinline__ uint32_t nib_next(uint32_t nib) { return (uint32_t) __ffs((int) nib) - 1u; }
// jump state of an oriented vertex: bits 0-31 pointer, bits 32-62 distance to it, bit 63 = "pointer is the last vertex of
// my chain" (an explicit flag: on a perfect loop a pointer can come back to its owner, which must not look finished)
constexpr unsigned long long ST_DONE = 1ULL << 63;
__device__ __forceinline__ uint64_t pack_state(uint32_t ptr, uint32_t dist, bool done) {
    return ((uint64_t) (dist & 0x7FFFFFFFu) << 32) | ptr | (done ? ST_DONE : 0ULL);
}
__device__ __forceinline__ uint32_t st_ptr(unsigned long long s) { return (uint32_t) s; }
__device__ __forceinline__ uint32_t st_dist(unsigned long long s) { return (uint32_t) (s >> 32) & 0x7FFFFFFFu; }
__device__ __forceinline__ bool st_done(unsigned long long s) { return (s & ST_DONE) != 0; }

template<int W>
__device__ __forceinline__ void oriented_kmer(const uint64_t *__restrict__ kmers, uint64_t filepos, int strand, int k, uint64_t *x) {
    uint64_t y[W];
    load_rec<W>(kmers, filepos, y);
    if (strand) kmer_rc<W>(y, k, x);
    else {
#pragma unroll
        for (int j = 0; j < W; ++j) x[j] = y[j];
    }
}

// base-wise lexicographic comparison from base 0: -1, 0, 1
template<int W>
__device__ __forceinline__ int kmer_lex_cmp(const uint64_t *a, const uint64_t *b) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t p = rev2(a[j]), q = rev2(b[j]);
        if (p != q) return p < q ? -1 : 1;
    }
    return 0;
}

// 1a. successor links by lookup (only when the links recorded by fill_masks_kernel are stale)
template<int W>
__global__ void __launch_bounds__(256) succ_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k, const uint32_t *__restrict__ idx,
                                                  const uint8_t *__restrict__ masks, uint32_t *__restrict__ succ) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint32_t id = idx[i];
    uint32_t v = 2 * id + strand;
    uint32_t raw = masks[id];
    if (mask_is_junction(raw)) { succ[v] = NONE; return; }
    uint32_t mk = strand ? mask_conj(raw) : raw;
    uint64_t x[W], y[W];
    oriented_kmer<W>(kmers, i, strand, k, x);
    kmer_shl<W>(x, k, nib_next(mk & 15u), y);
    bool minimal;
    uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
    succ[v] = 2 * idy + (minimal ? 0u : 1u);
}

// 1b. initial jump state from the links: a vertex whose successor is a junction is the last of its chain
__global__ void __launch_bounds__(256) state_init_kernel(uint64_t n_nodes, const uint8_t *__restrict__ masks, const uint32_t *__restrict__ succ,
                                                        unsigned long long *__restrict__ state) {
    uint64_t v = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    if (mask_is_junction(masks[v >> 1])) { state[v] = pack_state((uint32_t) v, 0, true); return; }
    uint32_t s = succ[v];
    state[v] = mask_is_junction(masks[s >> 1]) ? pack_state((uint32_t) v, 0, true) : pack_state(s, 1, false);
}

// 2. up to JUMPS pointer jumps per vertex and launch; a vertex is finished when its pointer is the last vertex of its
//    chain (ST_DONE).  In-place: stale reads only make a jump shorter, never wrong.
constexpr int JUMPS = 4;
__global__ void __launch_bounds__(256) jump_kernel(unsigned long long *__restrict__ state, uint64_t n_nodes, uint32_t *__restrict__ changed) {
    uint64_t v = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    unsigned long long sv = state[v];
    if (st_done(sv)) return;
#pragma unroll 1
    for (int it = 0; it < JUMPS; ++it) {
        unsigned long long sp = state[st_ptr(sv)];
        sv = pack_state(st_ptr(sp), st_dist(sv) + st_dist(sp), st_done(sp));
        if (st_done(sv)) break;
    }
    state[v] = sv;
    if (!st_done(sv) && *changed == 0) *changed = 1;
}

// 3a. work list: oriented junctions (t = 2 * file index + strand) that have at least one outgoing edge
__global__ void __launch_bounds__(256) junction_flag_kernel(uint64_t n, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ masks,
                                                           uint32_t *__restrict__ flag) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint32_t raw = masks[idx[t >> 1]];
    uint32_t mk = (t & 1) ? mask_conj(raw) : raw;
    flag[t] = (mask_is_junction(raw) && (mk & 15u)) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) compact_kernel(const uint32_t *__restrict__ flag_scan, uint64_t n, uint32_t total, uint32_t *__restrict__ list) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = flag_scan[i];
    uint32_t b = (i + 1 < n) ? flag_scan[i + 1] : total;
    if (b != a) list[a] = (uint32_t) i;
}

struct EdgeRec {       // one start edge, stored at 4 * (work-list position) + nucleotide
    uint32_t v1;       // first vertex after the junction
    uint32_t n;        // number of appended bases: |s| = k + n; 0 = edge absent or sequence dropped (`s < !s`)
    uint32_t chain;    // v1 is not a junction (the path has inner vertices)
    uint32_t pad;
};

// 3b. everything the reference's ConstructSequenceWithEdge + `if (s < !s) continue` decide for the start edges of one
//     oriented junction; also the number of kept sequences and their total length
template<int W>
__global__ void __launch_bounds__(128) edge_eval_kernel(MphfDev m, const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ jlist,
                                                       uint32_t n_j, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ inv,
                                                       const uint8_t *__restrict__ masks, const uint32_t *__restrict__ succ,
                                                       const unsigned long long *__restrict__ state, EdgeRec *__restrict__ edges,
                                                       uint32_t *__restrict__ cnt, unsigned long long *__restrict__ bases) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_j) return;
    uint32_t t = jlist[j];
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint32_t raw = masks[idx[i]];
    uint32_t mk = strand ? mask_conj(raw) : raw;
    uint64_t x[W];
    oriented_kmer<W>(kmers, i, strand, k, x);
    uint32_t c_kept = 0;
    unsigned long long b = 0;
#pragma unroll 1
    for (uint32_t c = 0; c < 4; ++c) {
        EdgeRec e;
        e.v1 = 0; e.n = 0; e.chain = 0; e.pad = 0;
        if (mk & (1u << c)) {
            uint64_t y[W], rcn[W];
            kmer_shl<W>(x, k, c, y);
            bool minimal;
            uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
            e.v1 = 2 * idy + (minimal ? 0u : 1u);
            e.chain = mask_is_junction(masks[idy]) ? 0u : 1u;
            uint32_t nn, cprime;   // cprime: first edge nucleotide of the reverse-complement path, only breaks ties
            if (!e.chain) {
                nn = 1;
                kmer_rc<W>(y, k, rcn);
                cprime = 3u - kmer_base(x, 0);
            } else {
                unsigned long long st = state[e.v1];
                uint32_t last = st_ptr(st);
                nn = st_dist(st) + 2;
                uint32_t vn = succ[last];
                // rc(kmer(vn)): the stored record if vn is the reverse strand, else its reverse complement
                oriented_kmer<W>(kmers, inv[vn >> 1], (vn & 1) ? 0 : 1, k, rcn);
                uint64_t lk[W];
                oriented_kmer<W>(kmers, inv[last >> 1], (int) (last & 1), k, lk);
                cprime = 3u - kmer_base(lk, 0);
            }
            int cmp = kmer_lex_cmp<W>(x, rcn);
            bool keep = cmp > 0 || (cmp == 0 && c >= cprime);
            if (keep) { e.n = nn; ++c_kept; b += (unsigned long long) k + nn; }
        }
        edges[4ull * j + c] = e;
    }
    cnt[j] = c_kept;
    bases[j] = b;
}

// 4a. heads: sequence length, first k+1 bases, and the output slot of the chain (indexed by its first vertex)
template<int W>
__global__ void __launch_bounds__(128) edge_emit_kernel(const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ jlist, uint32_t n_j,
                                                       const EdgeRec *__restrict__ edges, const uint32_t *__restrict__ cnt_off,
                                                       const unsigned long long *__restrict__ base_off, uint32_t *__restrict__ seq_len,
                                                       unsigned long long *__restrict__ seq_off, uint8_t *__restrict__ chars,
                                                       uint32_t *__restrict__ slot) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_j) return;
    uint32_t t = jlist[j];
    uint32_t u = cnt_off[j];
    unsigned long long off = base_off[j];
    uint64_t x[W];
    bool have_x = false;
#pragma unroll 1
    for (uint32_t c = 0; c < 4; ++c) {
        EdgeRec e = edges[4ull * j + c];
        if (e.n == 0) continue;
        if (!have_x) { oriented_kmer<W>(kmers, t >> 1, (int) (t & 1), k, x); have_x = true; }
        seq_len[u] = (uint32_t) k + e.n;
        seq_off[u] = off;
        for (int q = 0; q < k; ++q) chars[off + q] = (uint8_t) kmer_base_w<W>(x, q);
        chars[off + k] = (uint8_t) c;
        if (e.chain) slot[e.v1] = u;
        ++u;
        off += (unsigned long long) k + e.n;
    }
}

// 4b. every chain vertex v_j (j >= 1) writes base k+j of its sequence: the nucleotide of its unique outgoing edge.
// Its chain head is found through the reverse strand: walking rc(v_j) forward ends at rc(v_1) after j-1 steps.
__global__ void __launch_bounds__(256) chain_emit_kernel(uint64_t n_nodes, int k, const uint8_t *__restrict__ masks,
                                                        const unsigned long long *__restrict__ state, const uint32_t *__restrict__ slot,
                                                        const unsigned long long *__restrict__ seq_off, uint8_t *__restrict__ chars) {
    uint64_t v = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    uint32_t raw = masks[v >> 1];
    if (mask_is_junction(raw)) return;
    unsigned long long ss = state[v ^ 1];
    if (!st_done(ss)) return;   // perfect loop: never resolves
    uint32_t head = st_ptr(ss) ^ 1u;
    uint32_t u = slot[head];
    if (u == NONE) return;
    uint32_t j = st_dist(ss) + 1;
    uint32_t mk = (v & 1) ? mask_conj(raw) : raw;
    chars[seq_off[u] + (unsigned long long) k + j] = (uint8_t) nib_next(mk & 15u);
}

// ---- perfect loops -----------------------------------------------------------------------------------------------------
// candidates: canonical-strand vertices that are neither junctions nor resolved
__global__ void loop_flag_kernel(const uint32_t *__restrict__ idx, uint64_t n, const uint8_t *__restrict__ masks,
                                 const unsigned long long *__restrict__ state, uint32_t *__restrict__ flag) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = idx[i];
    uint32_t f = 0;
    if (!mask_is_junction(masks[id])) f = st_done(state[2 * id]) ? 0u : 1u;
    flag[i] = f;
}

__device__ __forceinline__ bool chars_less_rc(const uint8_t *s, unsigned long long L) {   // s < !s
    for (unsigned long long j = 0; j < L; ++j) {
        uint32_t a = s[j], b = 3u - s[L - 1 - j];
        if (a != b) return a < b;
    }
    return false;
}
__device__ __forceinline__ void chars_rc_inplace(uint8_t *s, unsigned long long L) {
    unsigned long long a = 0, b = L - 1;
    for (; a < b; ++a, --b) {
        uint8_t ca = (uint8_t) (3u - s[a]), cb = (uint8_t) (3u - s[b]);
        s[a] = cb; s[b] = ca;
    }
    if (a == b) s[a] = (uint8_t) (3u - s[a]);
}

// Single thread, file order, exactly CollectLoops.  write == 0: only count sequences / bases.
template<int W>
__global__ void loops_kernel(const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ idx, const uint8_t *__restrict__ masks,
                             const uint32_t *__restrict__ succ, const uint32_t *__restrict__ list, uint32_t n_list,
                             uint8_t *__restrict__ visited, int write, unsigned long long *__restrict__ totals /* [0]=seqs [1]=bases */,
                             uint8_t *__restrict__ scratch, uint32_t seq_base, unsigned long long char_base,
                             uint32_t *__restrict__ seq_len, unsigned long long *__restrict__ seq_off, uint8_t *__restrict__ chars) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long nseq = 0, nbases = 0;
    for (uint32_t q = 0; q < n_list; ++q) {
        uint32_t i = list[q];
        uint32_t id = idx[i];
        if (visited[id] == (uint8_t) (write + 1)) continue;
        uint32_t v = 2 * id;
        // walk the cycle from the stored (canonical) orientation
        unsigned long long n = 0;
        long long split = -1;
        uint32_t cur = v;
        if (write) {
            uint64_t x[W];
            load_rec<W>(kmers, i, x);
            for (int j = 0; j < k; ++j) scratch[j] = (uint8_t) kmer_base_w<W>(x, j);
        }
        do {
            visited[cur >> 1] = (uint8_t) (write + 1);
            uint32_t nx = succ[cur];
            if (split < 0 && nx == (cur ^ 1u)) split = (long long) n;   // self-reverse-complement (k+1)-mer at offset n
            if (write) {
                uint32_t raw = masks[cur >> 1];
                uint32_t mk = (cur & 1) ? mask_conj(raw) : raw;
                scratch[k + n] = (uint8_t) nib_next(mk & 15u);
            }
            cur = nx;
            ++n;
        } while (cur != v);
        unsigned long long L = (unsigned long long) k + n;
        if (split < 0) {
            if (write) {
                uint8_t *dst = chars + char_base + nbases;
                for (unsigned long long j = 0; j < L; ++j) dst[j] = scratch[j];
                if (chars_less_rc(dst, L)) chars_rc_inplace(dst, L);
                seq_len[seq_base + nseq] = (uint32_t) L;
                seq_off[seq_base + nseq] = char_base + nbases;
            }
            nseq += 1; nbases += L;
        } else {   // SplitLoop(s, pos): { s[pos, pos+k+1), s[pos+1, |s|-k) + s[0, pos+k) }
            unsigned long long pos = (unsigned long long) split;
            unsigned long long L1 = (unsigned long long) k + 1, L2 = (L - k - (pos + 1)) + (pos + k);
            if (write) {
                uint8_t *d1 = chars + char_base + nbases;
                for (unsigned long long j = 0; j < L1; ++j) d1[j] = scratch[pos + j];
                if (chars_less_rc(d1, L1)) chars_rc_inplace(d1, L1);
                seq_len[seq_base + nseq] = (uint32_t) L1;
                seq_off[seq_base + nseq] = char_base + nbases;
                uint8_t *d2 = d1 + L1;
                unsigned long long w = 0;
                for (unsigned long long j = pos + 1; j < L - k; ++j) d2[w++] = scratch[j];
                for (unsigned long long j = 0; j < pos + k; ++j) d2[w++] = scratch[j];
                if (chars_less_rc(d2, L2)) chars_rc_inplace(d2, L2);
                seq_len[seq_base + nseq + 1] = (uint32_t) L2;
                seq_off[seq_base + nseq + 1] = char_base + nbases + L1;
            }
            nseq += 2; nbases += L1 + L2;
        }
    }
    totals[0] = nseq;
    totals[1] = nbases;
}

// ---- packing ---------------------------------------------------------------------------------------------------------
__global__ void seq_words_kernel(const uint32_t *__restrict__ seq_len, uint64_t n, unsigned long long *__restrict__ nw) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nw[i] = ((unsigned long long) seq_len[i] + 31) >> 5;
}

// one warp per sequence: 32 bases -> one word per lane iteration
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ chars, const unsigned long long *__restrict__ seq_off,
                                                  const uint32_t *__restrict__ seq_len, const unsigned long long *__restrict__ word_off,
                                                  uint64_t n, uint64_t *__restrict__ words) {
    uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= n) return;
    const uint8_t *src = chars + seq_off[s];
    uint32_t L = seq_len[s];
    uint32_t nw = (L + 31) >> 5;
    for (uint32_t w = lane; w < nw; w += 32) {
        uint64_t acc = 0;
        uint32_t b0 = w * 32;
        uint32_t cnt = min(32u, L - b0);
        for (uint32_t j = 0; j < cnt; ++j) acc |= (uint64_t) (src[b0 + j] & 3u) << (2 * j);
        words[word_off[s] + w] = acc;
    }
}

}  // namespace sb200
#include "unitigs_walk.cuh"
namespace sb200 {

template<int W>
static sb200_unitigs *unitigs_w(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops) {
    uint64_t n = kmers->size;
    if (!ctx->force_jump_path) {
        sb200_unitigs *fast = unitigs_walk_w<W>(ctx, kmers, mphf, ext, with_loops, 0, n, nullptr);
        if (fast) return fast;
    }
    SB200_REQUIRE(2 * n < 0xFFFFFFF0ull, "more than 2^31 k-mers on one GPU: shard the input");
    int k = (int) kmers->k;
    uint64_t n_nodes = 2 * n;
    MphfDev m = mphf_dev(mphf);
    // 1. successor links
    DevBuf<uint32_t> succ_own;
    const uint32_t *succ = ext->succ.p;
    if (!ext->succ_valid) {
        succ_own.alloc(ctx, n_nodes);
        LAUNCH(ctx, succ_kernel<W>, div_up(n_nodes, 256), 256, 0, m, kmers->data.p, n, k, ext->idx.p, ext->masks.p, succ_own.p);
        succ = succ_own.p;
    }
    DevBuf<unsigned long long> state(ctx, n_nodes);
    LAUNCH(ctx, state_init_kernel, div_up(n_nodes, 256), 256, 0, n_nodes, ext->masks.p, succ, state.p);
    // 2. pointer jumping: every jump of a launch adds at least the distance its target had reached in the previous launch,
    //    so the guaranteed reach grows by (1 + JUMPS)x per launch: 5^16 > 2^32 vertices.  Whatever is still unfinished after
    //    MAX_ROUNDS launches lies on a perfect loop.
    constexpr int MAX_ROUNDS = 16;
    DevBuf<uint32_t> changed(ctx, MAX_ROUNDS + 2);
    changed.zero();
    int rounds = 0;
    bool converged = false;
    for (; rounds < MAX_ROUNDS; ++rounds) {
        LAUNCH(ctx, jump_kernel, div_up(n_nodes, 256), 256, 0, state.p, n_nodes, changed.p + rounds);
        uint32_t ch = 0;
        ctx->fetch(&ch, changed.p + rounds, 4);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (!ch) { converged = true; break; }
    }
    // perfect loops present?
    uint32_t n_loop_nodes = 0;
    DevBuf<uint32_t> loop_list;
    unsigned long long loop_totals[2] = {0, 0};
    DevBuf<unsigned long long> totals_dev(ctx, 4);
    DevBuf<uint8_t> visited, scratch;
    if (with_loops && !converged) {
        DevBuf<uint32_t> flag(ctx, n + 1);
        DevBuf<uint32_t> tot(ctx, 1);
        LAUNCH(ctx, loop_flag_kernel, div_up(n, 256), 256, 0, ext->idx.p, n, ext->masks.p, state.p, flag.p);
        exclusive_scan<uint32_t>(ctx, flag.p, n, tot.p);
        ctx->fetch(&n_loop_nodes, tot.p, 4);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (n_loop_nodes) {
            loop_list.alloc(ctx, n_loop_nodes);
            LAUNCH(ctx, compact_kernel, div_up(n, 256), 256, 0, flag.p, n, n_loop_nodes, loop_list.p);
            visited.alloc(ctx, n + 1); visited.zero();
            scratch.alloc(ctx, (uint64_t) n_loop_nodes + k + 8);
            LAUNCH(ctx, loops_kernel<W>, 1, 1, 0, kmers->data.p, k, ext->idx.p, ext->masks.p, succ, loop_list.p, n_loop_nodes, visited.p, 0,
                   totals_dev.p, scratch.p, 0u, 0ull, (uint32_t *) nullptr, (unsigned long long *) nullptr, (uint8_t *) nullptr);
            ctx->fetch(loop_totals, totals_dev.p, 16);
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
    }
    // 3. start edges of the junctions
    uint32_t n_j = 0;
    DevBuf<uint32_t> jlist;
    {
        DevBuf<uint32_t> flag(ctx, n_nodes + 1);
        DevBuf<uint32_t> tot(ctx, 1);
        LAUNCH(ctx, junction_flag_kernel, div_up(n_nodes, 256), 256, 0, n, ext->idx.p, ext->masks.p, flag.p);
        exclusive_scan<uint32_t>(ctx, flag.p, n_nodes, tot.p);
        ctx->fetch(&n_j, tot.p, 4);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        jlist.alloc(ctx, (uint64_t) n_j + 1);
        if (n_j) LAUNCH(ctx, compact_kernel, div_up(n_nodes, 256), 256, 0, flag.p, n_nodes, n_j, jlist.p);
    }
    DevBuf<EdgeRec> edges(ctx, 4ull * n_j + 4);
    DevBuf<uint32_t> cnt(ctx, (uint64_t) n_j + 1);
    DevBuf<unsigned long long> bases(ctx, (uint64_t) n_j + 1);
    uint32_t n_paths = 0;
    unsigned long long path_bases = 0;
    if (n_j) {
        LAUNCH(ctx, edge_eval_kernel<W>, div_up(n_j, 128), 128, 0, m, kmers->data.p, k, jlist.p, n_j, ext->idx.p, ext->inv.p, ext->masks.p, succ,
               state.p, edges.p, cnt.p, bases.p);
        DevBuf<uint32_t> cnt_total(ctx, 1);
        exclusive_scan<uint32_t>(ctx, cnt.p, n_j, cnt_total.p);
        exclusive_scan<unsigned long long>(ctx, bases.p, n_j, totals_dev.p + 2);
        ctx->fetch(&n_paths, cnt_total.p, 4);
        ctx->fetch(&path_bases, totals_dev.p + 2, 8);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }

    uint64_t n_seqs = (uint64_t) n_paths + loop_totals[0];
    uint64_t n_bases = path_bases + loop_totals[1];
    sb200_unitigs *out = new sb200_unitigs();
    out->ctx = ctx; out->k = (unsigned) k; out->count = n_seqs; out->n_loops = loop_totals[0]; out->total_bases = n_bases;
    out->len.alloc(ctx, n_seqs + 1);
    out->word_off.alloc(ctx, n_seqs + 1);
    DevBuf<unsigned long long> seq_off(ctx, n_seqs + 1);
    DevBuf<uint8_t> chars(ctx, n_bases + 8);
    if (n_paths) {
        DevBuf<uint32_t> slot(ctx, n_nodes);
        CUDA_CHECK(cudaMemsetAsync(slot.p, 0xFF, n_nodes * 4, ctx->stream));
        LAUNCH(ctx, edge_emit_kernel<W>, div_up(n_j, 128), 128, 0, kmers->data.p, k, jlist.p, n_j, edges.p, cnt.p, bases.p, out->len.p, seq_off.p,
               chars.p, slot.p);
        LAUNCH(ctx, chain_emit_kernel, div_up(n_nodes, 256), 256, 0, n_nodes, k, ext->masks.p, state.p, slot.p, seq_off.p, chars.p);
    }
    if (loop_totals[0]) {
        LAUNCH(ctx, loops_kernel<W>, 1, 1, 0, kmers->data.p, k, ext->idx.p, ext->masks.p, succ, loop_list.p, n_loop_nodes, visited.p, 1,
               totals_dev.p, scratch.p, n_paths, path_bases, out->len.p, seq_off.p, chars.p);
    }
    // 5. pack to 2 bits per base, every sequence word-aligned (the layout of Sequence / the binary reads)
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    LAUNCH(ctx, seq_words_kernel, div_up(n_seqs ? n_seqs : 1, 256), 256, 0, out->len.p, n_seqs, (unsigned long long *) out->word_off.p);
    exclusive_scan<unsigned long long>(ctx, (unsigned long long *) out->word_off.p, n_seqs, (unsigned long long *) (out->word_off.p + n_seqs));
    uint64_t total_words = 0;
    ctx->fetch(&total_words, out->word_off.p + n_seqs, 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    out->total_words = total_words;
    out->words.alloc(ctx, total_words + 1);
    if (n_seqs)
        LAUNCH(ctx, pack_kernel, div_up(n_seqs * 32, 256), 256, 0, chars.p, seq_off.p, out->len.p, (const unsigned long long *) out->word_off.p, n_seqs,
               out->words.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // temporaries above are released stream-ordered; results are ready
    return out;
}

// One GPU's share of a sharded extraction: the junctions among the k-mers of its own table shard, walked against the
// all-reduced masks and MPHF.  stats: [0] chain vertices seen, [1] chains longer than the walk limit, [2] start edges,
// [3] kept sequences, [4] kept bases, [5] non-junction k-mers of the WHOLE index (for the perfect-loop check
// sum_g stats_g[0] == 2 * stats[5]).  Returns nullptr (with stats filled) when a chain was too long.
sb200_unitigs *extract_unitigs_local(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, uint64_t *stats) {
    SB200_REQUIRE(ext->n_local == kmers->size && ext->k == kmers->k && ext->size == mphf->total, "extension index does not belong to this shard");
    WalkStats st;
    sb200_unitigs *u = nullptr;
    switch (kmers->words) {
        case 1: u = unitigs_walk_w<1>(ctx, kmers, mphf, ext, 0, 0, kmers->size, &st); break;
        case 2: u = unitigs_walk_w<2>(ctx, kmers, mphf, ext, 0, 0, kmers->size, &st); break;
        case 3: u = unitigs_walk_w<3>(ctx, kmers, mphf, ext, 0, 0, kmers->size, &st); break;
        default: u = unitigs_walk_w<4>(ctx, kmers, mphf, ext, 0, 0, kmers->size, &st); break;
    }
    DevBuf<unsigned long long> nj(ctx, 1);
    nj.zero();
    LAUNCH(ctx, count_nonjunction_kernel, (unsigned) ctx->num_sms * 8, 256, 0, ext->masks.p, ext->size, nj.p);
    unsigned long long h = 0;
    ctx->fetch(&h, nj.p, 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    stats[0] = st.chain_vertices; stats[1] = st.long_chains; stats[2] = st.n_edges; stats[3] = st.n_kept; stats[4] = st.kept_bases; stats[5] = h;
    return u;
}

sb200_unitigs *extract_unitigs(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops) {
    SB200_REQUIRE(ext->size == kmers->size && ext->n_local == kmers->size && ext->k == kmers->k,
                  "extension index does not belong to this k-mer set (sharded tables go through sb200_unitigs_extract_local)");
    switch (kmers->words) {
        case 1: return unitigs_w<1>(ctx, kmers, mphf, ext, with_loops);
        case 2: return unitigs_w<2>(ctx, kmers, mphf, ext, with_loops);
        case 3: return unitigs_w<3>(ctx, kmers, mphf, ext, with_loops);
        default: return unitigs_w<4>(ctx, kmers, mphf, ext, with_loops);
    }
}

}  // namespace sb200

// unitigs.cu — unbranching paths (the condensed de Bruijn graph's edges) by pointer-jumping list ranking.
//
// Replaces UnbranchingPathExtractor::ExtractUnbranchingPathsAndLoops
// (C/assembly_graph/construction/debruijn_graph_constructor.hpp:182-388): the reference walks from every outgoing
// edge of every junction one MPHF lookup at a time and appends one nucleotide per step; here
//   1. every oriented non-junction k-mer gets its successor node with ONE lookup (succ_kernel),
//   2. pointer jumping gives each of them the last vertex of its chain and the distance to it (jump_kernel; a
//      (pointer, distance) pair is one 64-bit word so in-place updates are always observed consistently),
//   3. every junction enumerates its start edges in the reference's order (file order of the canonical k-mer,
//      forward strand then reverse strand, A,C,G,T), gets path length and end junction from step 2 and decides the
//      orientation filter `!(s < !s)` from the two end k-mers (ties by the first edge nucleotides),
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

// 1. successor links + initial jump state
template<int W>
__global__ void __launch_bounds__(256) succ_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k, const uint32_t *__restrict__ idx,
                                                  const uint8_t *__restrict__ masks, uint32_t *__restrict__ succ,
                                                  unsigned long long *__restrict__ state) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint32_t id = idx[i];
    uint32_t v = 2 * id + strand;
    uint32_t raw = masks[id];
    if (mask_is_junction(raw)) {
        succ[v] = NONE;
        state[v] = pack_state(v, 0, true);
        return;
    }
    uint32_t mk = strand ? mask_conj(raw) : raw;
    uint64_t x[W], y[W];
    oriented_kmer<W>(kmers, i, strand, k, x);
    kmer_shl<W>(x, k, nib_next(mk & 15u), y);
    bool minimal;
    uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
    uint32_t s = 2 * idy + (minimal ? 0u : 1u);
    succ[v] = s;
    bool next_is_junction = mask_is_junction(masks[idy]);
    state[v] = next_is_junction ? pack_state(v, 0, true) : pack_state(s, 1, false);
}

// 2. one round of pointer jumping; a vertex is finished when its pointer is the last vertex of its chain (ST_DONE)
__global__ void __launch_bounds__(256) jump_kernel(unsigned long long *__restrict__ state, uint64_t n_nodes, uint32_t *__restrict__ changed) {
    uint64_t v = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    unsigned long long sv = state[v];
    if (st_done(sv)) return;
    unsigned long long sp = state[st_ptr(sv)];
    state[v] = pack_state(st_ptr(sp), st_dist(sv) + st_dist(sp), st_done(sp));
    if (*changed == 0) *changed = 1;
}

struct EdgeInfo {
    uint32_t v1;       // first vertex after the junction
    uint32_t n;        // number of appended bases: |s| = k + n
    bool chain;        // v1 is not a junction (the path has inner vertices)
    bool keep;         // !(s < !s)
};

// Everything the reference's ConstructSequenceWithEdge + `if (s < !s) continue` decide for start edge (x, c)
template<int W>
__device__ __forceinline__ EdgeInfo eval_edge(const MphfDev &m, const uint64_t *__restrict__ kmers, int k, const uint32_t *__restrict__ inv,
                                              const uint8_t *__restrict__ masks, const uint32_t *__restrict__ succ,
                                              const unsigned long long *__restrict__ state, const uint64_t *x, uint32_t c) {
    EdgeInfo e;
    uint64_t y[W], rcn[W];
    kmer_shl<W>(x, k, c, y);
    bool minimal;
    uint32_t idy = (uint32_t) mphf_lookup_oriented<W>(m, y, k, &minimal);
    e.v1 = 2 * idy + (minimal ? 0u : 1u);
    e.chain = !mask_is_junction(masks[idy]);
    uint32_t cprime;   // first edge nucleotide of the reverse-complement path, used only to break ties
    if (!e.chain) {
        e.n = 1;
        kmer_rc<W>(y, k, rcn);
        cprime = 3u - kmer_base(x, 0);
    } else {
        unsigned long long st = state[e.v1];
        uint32_t last = st_ptr(st);
        e.n = st_dist(st) + 2;
        uint32_t vn = succ[last];
        // rc(kmer(vn)): stored record if vn is the reverse strand, else its reverse complement
        oriented_kmer<W>(kmers, inv[vn >> 1], (vn & 1) ? 0 : 1, k, rcn);
        uint64_t lk[W];
        oriented_kmer<W>(kmers, inv[last >> 1], (int) (last & 1), k, lk);
        cprime = 3u - kmer_base(lk, 0);
    }
    int cmp = kmer_lex_cmp<W>(x, rcn);
    e.keep = cmp > 0 || (cmp == 0 && c >= cprime);
    return e;
}

// 3. per oriented junction: number of kept sequences and their total length
template<int W>
__global__ void __launch_bounds__(128) edge_count_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k,
                                                        const uint32_t *__restrict__ idx, const uint32_t *__restrict__ inv,
                                                        const uint8_t *__restrict__ masks, const uint32_t *__restrict__ succ,
                                                        const unsigned long long *__restrict__ state, uint32_t *__restrict__ cnt,
                                                        unsigned long long *__restrict__ bases) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint32_t raw = masks[idx[i]];
    uint32_t c_kept = 0;
    unsigned long long b = 0;
    if (mask_is_junction(raw)) {
        uint32_t mk = strand ? mask_conj(raw) : raw;
        if (mk & 15u) {
            uint64_t x[W];
            oriented_kmer<W>(kmers, i, strand, k, x);
            for (uint32_t c = 0; c < 4; ++c) {
                if (!(mk & (1u << c))) continue;
                EdgeInfo e = eval_edge<W>(m, kmers, k, inv, masks, succ, state, x, c);
                if (e.keep) { ++c_kept; b += (unsigned long long) k + e.n; }
            }
        }
    }
    cnt[t] = c_kept;
    bases[t] = b;
}

// 4a. heads: sequence length, first k+1 bases, and the output slot of the chain (indexed by its first vertex)
template<int W>
__global__ void __launch_bounds__(128) edge_emit_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k,
                                                       const uint32_t *__restrict__ idx, const uint32_t *__restrict__ inv,
                                                       const uint8_t *__restrict__ masks, const uint32_t *__restrict__ succ,
                                                       const unsigned long long *__restrict__ state, const uint32_t *__restrict__ cnt_off,
                                                       const unsigned long long *__restrict__ base_off, uint32_t *__restrict__ seq_len,
                                                       unsigned long long *__restrict__ seq_off, uint8_t *__restrict__ chars,
                                                       uint32_t *__restrict__ slot) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint32_t raw = masks[idx[i]];
    if (!mask_is_junction(raw)) return;
    uint32_t mk = strand ? mask_conj(raw) : raw;
    if (!(mk & 15u)) return;
    uint64_t x[W];
    oriented_kmer<W>(kmers, i, strand, k, x);
    uint32_t u = cnt_off[t];
    unsigned long long off = base_off[t];
    for (uint32_t c = 0; c < 4; ++c) {
        if (!(mk & (1u << c))) continue;
        EdgeInfo e = eval_edge<W>(m, kmers, k, inv, masks, succ, state, x, c);
        if (!e.keep) continue;
        seq_len[u] = (uint32_t) k + e.n;
        seq_off[u] = off;
        for (int j = 0; j < k; ++j) chars[off + j] = (uint8_t) kmer_base(x, j);
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
    if (!mask_is_junction(masks[id])) {
        f = st_done(state[2 * id]) ? 0u : 1u;
    }
    flag[i] = f;
}

__global__ void loop_compact_kernel(const uint32_t *__restrict__ flag_scan, uint64_t n, uint32_t total, uint32_t *__restrict__ list) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = flag_scan[i];
    uint32_t b = (i + 1 < n) ? flag_scan[i + 1] : total;
    if (b != a) list[a] = (uint32_t) i;
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
            for (int j = 0; j < k; ++j) scratch[j] = (uint8_t) kmer_base(x, j);
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

template<int W>
static sb200_unitigs *unitigs_w(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops) {
    uint64_t n = kmers->size;
    SB200_REQUIRE(2 * n < 0xFFFFFFF0ull, "more than 2^31 k-mers on one GPU: shard the input");
    int k = (int) kmers->k;
    uint64_t n_nodes = 2 * n;
    MphfDev m = mphf_dev(mphf);
    DevBuf<uint32_t> succ(ctx, n_nodes);
    DevBuf<unsigned long long> state(ctx, n_nodes);
    LAUNCH(ctx, succ_kernel<W>, div_up(n_nodes, 256), 256, 0, m, kmers->data.p, n, k, ext->idx.p, ext->masks.p, succ.p, state.p);
    // 2. pointer jumping: chains finish in ceil(log2(longest chain)) rounds; only perfect loops never do
    DevBuf<uint32_t> changed(ctx, 40);
    changed.zero();
    int rounds = 0;
    for (; rounds < 34; ++rounds) {
        LAUNCH(ctx, jump_kernel, div_up(n_nodes, 256), 256, 0, state.p, n_nodes, changed.p + rounds);
        uint32_t ch = 0;
        CUDA_CHECK(cudaMemcpyAsync(&ch, changed.p + rounds, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (!ch) break;
    }
    // perfect loops present?
    uint32_t n_loop_nodes = 0;
    DevBuf<uint32_t> loop_list;
    unsigned long long loop_totals[2] = {0, 0};
    DevBuf<unsigned long long> totals_dev(ctx, 4);
    DevBuf<uint8_t> visited, scratch;
    if (with_loops && rounds >= 34) {
        DevBuf<uint32_t> flag(ctx, n + 1);
        DevBuf<uint32_t> tot(ctx, 1);
        LAUNCH(ctx, loop_flag_kernel, div_up(n, 256), 256, 0, ext->idx.p, n, ext->masks.p, state.p, flag.p);
        exclusive_scan<uint32_t>(ctx, flag.p, n, tot.p);
        CUDA_CHECK(cudaMemcpyAsync(&n_loop_nodes, tot.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (n_loop_nodes) {
            loop_list.alloc(ctx, n_loop_nodes);
            LAUNCH(ctx, loop_compact_kernel, div_up(n, 256), 256, 0, flag.p, n, n_loop_nodes, loop_list.p);
            visited.alloc(ctx, n + 1); visited.zero();
            scratch.alloc(ctx, (uint64_t) n_loop_nodes + k + 8);
            LAUNCH(ctx, loops_kernel<W>, 1, 1, 0, kmers->data.p, k, ext->idx.p, ext->masks.p, succ.p, loop_list.p, n_loop_nodes, visited.p, 0,
                   totals_dev.p, scratch.p, 0u, 0ull, (uint32_t *) nullptr, (unsigned long long *) nullptr, (uint8_t *) nullptr);
            CUDA_CHECK(cudaMemcpyAsync(loop_totals, totals_dev.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        }
    }
    // 3. start edges
    DevBuf<uint32_t> cnt(ctx, n_nodes + 1);
    DevBuf<unsigned long long> bases(ctx, n_nodes + 1);
    LAUNCH(ctx, edge_count_kernel<W>, div_up(n_nodes, 128), 128, 0, m, kmers->data.p, n, k, ext->idx.p, ext->inv.p, ext->masks.p, succ.p, state.p,
           cnt.p, bases.p);
    DevBuf<uint32_t> cnt_total(ctx, 1);
    exclusive_scan<uint32_t>(ctx, cnt.p, n_nodes, cnt_total.p);
    exclusive_scan<unsigned long long>(ctx, bases.p, n_nodes, totals_dev.p + 2);
    uint32_t n_paths = 0;
    unsigned long long path_bases = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_paths, cnt_total.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(&path_bases, totals_dev.p + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));

    uint64_t n_seqs = (uint64_t) n_paths + loop_totals[0];
    uint64_t n_bases = path_bases + loop_totals[1];
    sb200_unitigs *out = new sb200_unitigs();
    out->ctx = ctx; out->k = (unsigned) k; out->count = n_seqs; out->n_loops = loop_totals[0]; out->total_bases = n_bases;
    out->len.alloc(ctx, n_seqs + 1);
    out->word_off.alloc(ctx, n_seqs + 1);
    DevBuf<unsigned long long> seq_off(ctx, n_seqs + 1);
    DevBuf<uint8_t> chars(ctx, n_bases + 8);
    DevBuf<uint32_t> slot(ctx, n_nodes);
    CUDA_CHECK(cudaMemsetAsync(slot.p, 0xFF, n_nodes * 4, ctx->stream));
    if (n_paths) {
        LAUNCH(ctx, edge_emit_kernel<W>, div_up(n_nodes, 128), 128, 0, m, kmers->data.p, n, k, ext->idx.p, ext->inv.p, ext->masks.p, succ.p, state.p,
               cnt.p, bases.p, out->len.p, seq_off.p, chars.p, slot.p);
        LAUNCH(ctx, chain_emit_kernel, div_up(n_nodes, 256), 256, 0, n_nodes, k, ext->masks.p, state.p, slot.p, seq_off.p, chars.p);
    }
    if (loop_totals[0]) {
        LAUNCH(ctx, loops_kernel<W>, 1, 1, 0, kmers->data.p, k, ext->idx.p, ext->masks.p, succ.p, loop_list.p, n_loop_nodes, visited.p, 1,
               totals_dev.p, scratch.p, n_paths, path_bases, out->len.p, seq_off.p, chars.p);
    }
    // 5. pack to 2 bits per base, every sequence word-aligned (the layout of Sequence / the binary reads)
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    LAUNCH(ctx, seq_words_kernel, div_up(n_seqs ? n_seqs : 1, 256), 256, 0, out->len.p, n_seqs, (unsigned long long *) out->word_off.p);
    exclusive_scan<unsigned long long>(ctx, (unsigned long long *) out->word_off.p, n_seqs, (unsigned long long *) (out->word_off.p + n_seqs));
    uint64_t total_words = 0;
    CUDA_CHECK(cudaMemcpyAsync(&total_words, out->word_off.p + n_seqs, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    out->total_words = total_words;
    out->words.alloc(ctx, total_words + 1);
    if (n_seqs)
        LAUNCH(ctx, pack_kernel, div_up(n_seqs * 32, 256), 256, 0, chars.p, seq_off.p, out->len.p, (const unsigned long long *) out->word_off.p, n_seqs,
               out->words.p);
    return out;
}

sb200_unitigs *extract_unitigs(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const sb200_ext *ext, int with_loops) {
    SB200_REQUIRE(ext->size == kmers->size && ext->k == kmers->k, "extension index does not belong to this k-mer set");
    switch (kmers->words) {
        case 1: return unitigs_w<1>(ctx, kmers, mphf, ext, with_loops);
        case 2: return unitigs_w<2>(ctx, kmers, mphf, ext, with_loops);
        case 3: return unitigs_w<3>(ctx, kmers, mphf, ext, with_loops);
        default: return unitigs_w<4>(ctx, kmers, mphf, ext, with_loops);
    }
}

}  // namespace sb200

// ext.cu — DeBruijnExtensionIndex payload (one InOutMask byte per k-mer, in MPHF-index order) and EarlyTipClipper.
//
// Replaces (reference file:line):
//   DeBruijnExtensionIndexBuilder::FillExtensionsFromIndex   C/utils/extension_index/kmer_extension_index_builder.hpp:44-59
//   InOutMask::AddOutgoing / AddIncoming                     C/utils/extension_index/kmer_extension_index.hpp:92-106
//   InvertableKeyWithHash::CountIdx                          C/utils/ph_map/key_with_hash.hpp:119-127
//   EarlyTipClipperProcessor / RemoveInconsistentForwardLinks C/assembly_graph/construction/early_simplification.hpp:20-152
// One thread per stored (k+1)-mer: two canonicalisations, two MPHF lookups (bit-vectors and rank samples are a few
// tens of MB and stay L2-resident), two byte-wide ORs done as 32-bit atomicOr on the containing word.
#include "common.cuh"
#include "kmer_ops.cuh"
#include "kmer_set.cuh"
#include "mphf.cuh"
#include "radix_sort.cuh"

namespace sb200 {

MphfDev mphf_dev(const sb200_mphf *m);

__device__ __forceinline__ void mask_or(uint8_t *masks, uint64_t idx, uint32_t bit) {
    atomicOr(reinterpret_cast<unsigned int *>(masks + (idx & ~3ULL)), (1u << bit) << (8 * (idx & 3)));
}
__device__ __forceinline__ void mask_and_not(uint8_t *masks, uint64_t idx, uint32_t bits) {
    atomicAnd(reinterpret_cast<unsigned int *>(masks + (idx & ~3ULL)), ~(bits << (8 * (idx & 3))));
}

// file_masks != nullptr: the masks were OR-ed in file order while the k-mers were sorted (count.cu); they only move to their
// place in MPHF-index order here.
template<int W>
__global__ void __launch_bounds__(256) index_of_kmers_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, uint32_t *__restrict__ idx,
                                                            uint32_t *__restrict__ inv, const uint8_t *__restrict__ file_masks,
                                                            uint8_t *__restrict__ masks) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r[W];
    load_rec<W>(kmers, i, r);
    uint32_t id = (uint32_t) mphf_lookup<W>(m, r);
    idx[i] = id;
    if (inv) inv[id] = (uint32_t) i;
    if (file_masks) masks[id] = file_masks[i];
}

// The same from the build's own record of where every key landed (sb200_mphf::place): index = set bits before the key's bit.  Two reads
// from tables that stay in L2 instead of the k-mer, two XXH3 hashes and the level probes.
__global__ void __launch_bounds__(256) index_from_place_kernel(const uint32_t *__restrict__ place, const uint32_t *__restrict__ pc_scan,
                                                              const uint64_t *__restrict__ bits, uint64_t n, uint32_t *__restrict__ idx,
                                                              uint32_t *__restrict__ inv, const uint8_t *__restrict__ file_masks,
                                                              uint8_t *__restrict__ masks) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = place[i];
    const uint32_t w = g >> 6;
    const uint32_t id = __ldg(pc_scan + w) + (uint32_t) __popcll(__ldg(bits + w) & ((1ULL << (g & 63u)) - 1ULL));
    idx[i] = id;
    if (inv) inv[id] = (uint32_t) i;
    if (file_masks) masks[id] = file_masks[i];
}

// Besides the two mask bits, every (k+1)-mer x = y -> z also records the link itself: succ[y] = z and, for the other
// strand, succ[rc(z)] = rc(y) (oriented vertex = 2 * index + strand).  A vertex with a single outgoing edge receives
// exactly one such write, so the unitig stage gets its successor links without a second round of MPHF lookups;
// vertices with several outgoing edges receive several (racing) writes and never read the slot.
template<int WS, int W>
__global__ void __launch_bounds__(256) fill_masks_kernel(MphfDev m, const uint64_t *__restrict__ kpomers, uint64_t n, int k, uint8_t *__restrict__ masks,
                                                        uint32_t *__restrict__ succ) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x[WS], a[W];
    load_rec<WS>(kpomers, i, x);
    uint32_t pnucl = kmer_base(x, 0), nnucl = kmer_base_w<WS>(x, k);
    bool min_a, min_b;
    kmer_subwindow<WS, W>(x, 0, k, a);
    uint64_t ia = mphf_lookup_oriented<W>(m, a, k, &min_a);
    mask_or(masks, ia, min_a ? nnucl : 7u - nnucl);                 // AddOutgoing(nnucl, as_is)
    kmer_subwindow<WS, W>(x, 1, k, a);
    uint64_t ib = mphf_lookup_oriented<W>(m, a, k, &min_b);
    mask_or(masks, ib, min_b ? pnucl + 4u : 7u - (pnucl + 4u));    // AddIncoming(pnucl, as_is)
    if (succ) {   // optional: measured at +7 ms for 147 M scattered 4-byte stores, the same as recomputing the links by lookup
        uint32_t vy = 2u * (uint32_t) ia + (min_a ? 0u : 1u), vz = 2u * (uint32_t) ib + (min_b ? 0u : 1u);
        succ[vy] = vz;
        succ[vz ^ 1u] = vy ^ 1u;
    }
}

template<int WS, int W>
static sb200_ext *build_ext_w(sb200_ctx *ctx, const sb200_kmers *kpomers, const sb200_kmers *kmers, const sb200_mphf *mphf) {
    sb200_ext *e = new sb200_ext();
    // The mask array covers the whole index (mphf->total entries); `kmers` may be one GPU's shard of the k-mer table, in which
    // case idx[] covers only the shard, the masks hold only the bits of this GPU's (k+1)-mers (summed over the GPUs by the
    // caller: distinct (k+1)-mers set distinct bits) and the inverse permutation is not built.
    const bool sharded = kmers->size != mphf->total;
    e->ctx = ctx; e->k = kmers->k; e->size = mphf->total; e->n_local = kmers->size;
    uint64_t padded = (e->size + 3) & ~3ULL;
    e->masks.alloc(ctx, padded + 4);
    e->masks.zero();
    e->idx.alloc(ctx, kmers->size);
    if (!sharded) e->inv.alloc(ctx, kmers->size);
    SB200_REQUIRE(2 * e->size < 0xFFFFFFF0ull, "more than 2^31 k-mers in one index");
    e->succ_valid = false;   // the direct-walk extraction needs no links; the pointer-jumping path computes them on demand
    MphfDev m = mphf_dev(mphf);
    // (a shard's masks are complete too — every candidate of a k-mer met at its owner — and land in this GPU's slice of the
    // zeroed full-size array; the caller's sum over the GPUs assembles the rest)
    const bool have_masks = kmers->masks_file.p != nullptr;
    if (kmers->size == 0) {
        // nothing of the table lives here (a rank that owns no k-mer)
    } else if (mphf->place.p && mphf->pc_scan.p && mphf->place.n == kmers->size && !ctx->no_place)   // the index was built over exactly this table on this GPU
        LAUNCH(ctx, index_from_place_kernel, div_up(kmers->size, 256), 256, 0, mphf->place.p, mphf->pc_scan.p, mphf->bits.p, kmers->size, e->idx.p,
               e->inv.p, have_masks ? kmers->masks_file.p : (const uint8_t *) nullptr, e->masks.p);
    else
        LAUNCH(ctx, index_of_kmers_kernel<W>, div_up(kmers->size, 256), 256, 0, m, kmers->data.p, kmers->size, e->idx.p, e->inv.p,
               have_masks ? kmers->masks_file.p : (const uint8_t *) nullptr, e->masks.p);
    if (!have_masks && kpomers->size) {
        auto fill_masks_kernel_ = fill_masks_kernel<WS, W>;
        LAUNCH(ctx, fill_masks_kernel_, div_up(kpomers->size, 256), 256, 0, m, kpomers->data.p, kpomers->size, (int) kmers->k, e->masks.p,
               e->succ.p);
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // blocking like every entry point: callers all-reduce the masks next
    return e;
}

// Extension index of a WHOLE k-mer table whose masks (MPHF-index order) already exist on the device: idx / inverse permutation by
// lookups, masks copied.  Used by the sharded path when rank 0 takes over the whole-set unitig extraction (shard.cu).
template<int W>
static sb200_ext *build_ext_from_masks_w(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const uint8_t *masks_dev) {
    sb200_ext *e = new sb200_ext();
    e->ctx = ctx; e->k = kmers->k; e->size = mphf->total; e->n_local = kmers->size;
    const uint64_t padded = (e->size + 3) & ~3ULL;
    e->masks.alloc(ctx, padded + 4);
    e->masks.zero();
    CUDA_CHECK(cudaMemcpyAsync(e->masks.p, masks_dev, e->size, cudaMemcpyDeviceToDevice, ctx->stream));
    e->idx.alloc(ctx, kmers->size);
    e->inv.alloc(ctx, kmers->size);
    e->masks_edited = true;   // no file-order copy of the masks exists for this table
    MphfDev m = mphf_dev(mphf);
    if (kmers->size)
        LAUNCH(ctx, index_of_kmers_kernel<W>, div_up(kmers->size, 256), 256, 0, m, kmers->data.p, kmers->size, e->idx.p, e->inv.p,
               (const uint8_t *) nullptr, e->masks.p);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return e;
}

sb200_ext *build_ext_from_masks(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, const uint8_t *masks_dev) {
    SB200_REQUIRE(mphf->total == kmers->size && mphf->words == kmers->words, "MPHF was not built over this k-mer set");
    switch (kmers->words) {
        case 1: return build_ext_from_masks_w<1>(ctx, kmers, mphf, masks_dev);
        case 2: return build_ext_from_masks_w<2>(ctx, kmers, mphf, masks_dev);
        case 3: return build_ext_from_masks_w<3>(ctx, kmers, mphf, masks_dev);
        default: return build_ext_from_masks_w<4>(ctx, kmers, mphf, masks_dev);
    }
}

sb200_ext *build_ext(sb200_ctx *ctx, const sb200_kmers *kpomers, const sb200_kmers *kmers, const sb200_mphf *mphf) {
    SB200_REQUIRE(kpomers->k == kmers->k + 1, "kpomers.k() must equal index.k() + 1");
    SB200_REQUIRE(mphf->total >= kmers->size && mphf->words == kmers->words, "MPHF was not built over this k-mer set");
    int WS = (int) kpomers->words, W = (int) kmers->words;
    if (WS == 1) return build_ext_w<1, 1>(ctx, kpomers, kmers, mphf);
    if (WS == 2 && W == 1) return build_ext_w<2, 1>(ctx, kpomers, kmers, mphf);
    if (WS == 2) return build_ext_w<2, 2>(ctx, kpomers, kmers, mphf);
    if (WS == 3 && W == 2) return build_ext_w<3, 2>(ctx, kpomers, kmers, mphf);
    if (WS == 3) return build_ext_w<3, 3>(ctx, kpomers, kmers, mphf);
    if (WS == 4 && W == 3) return build_ext_w<4, 3>(ctx, kpomers, kmers, mphf);
    return build_ext_w<4, 4>(ctx, kpomers, kmers, mphf);
}

// ---- early tip clipper ---------------------------------------------------------------------------------------------
// The reference walks the k-mer file sequentially and mutates masks as it goes; a tip's vertices all have a unique
// predecessor, so they are reachable only from their own junction and no other walk can observe their removal.
// Deciding every junction on the untouched masks and isolating afterwards therefore gives the reference's
// single-thread result (tests/test_gpu_parity.py pins it against the -t 1 reference run).
__constant__ int8_t c_unique_next[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};

__device__ __forceinline__ bool mask_unique(uint32_t nib) { return nib && !(nib & (nib - 1)); }

template<int W>
__device__ __forceinline__ uint32_t oriented_mask(const MphfDev &m, const uint8_t *masks, const uint64_t *x, int k, uint64_t *id_out) {
    bool minimal;
    uint64_t id = mphf_lookup_oriented<W>(m, x, k, &minimal);
    *id_out = id;
    uint32_t v = masks[id];
    return minimal ? v : mask_conj(v);
}

// Phase 1: for every oriented k-mer with >= 2 outgoing edges walk each branch (<= bound vertices); tips shorter than
// the longest branch are removed.  Vertices to isolate are recorded in `kill` (one byte per k-mer index) and applied
// by tipclip_apply_kernel, so that every decision sees the same snapshot.
template<int W>
__global__ void __launch_bounds__(128) tipclip_find_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k,
                                                          const uint8_t *__restrict__ masks, uint32_t bound, uint8_t *__restrict__ kill,
                                                          uint8_t *__restrict__ tipped /* 2 bits per file index: strand 0/1 */,
                                                          unsigned long long *__restrict__ removed) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint64_t y[W], x[W];
    load_rec<W>(kmers, i, y);
    if (strand) kmer_rc<W>(y, k, x);
    else {
#pragma unroll
        for (int j = 0; j < W; ++j) x[j] = y[j];
    }
    uint64_t id0;
    uint32_t mask = oriented_mask<W>(m, masks, x, k, &id0);
    if (__popc(mask & 15u) < 2) return;
    // pass 1: branch lengths (0 = not a tip)
    uint32_t len[4] = {0, 0, 0, 0};
    uint32_t maxlen = 0;
    for (uint32_t c = 0; c < 4; ++c) {
        if (!(mask & (1u << c))) continue;
        uint64_t cur[W], nxt[W];
        kmer_shl<W>(x, k, c, cur);
        uint32_t sz = 0;
        uint64_t id;
        uint32_t mk = oriented_mask<W>(m, masks, cur, k, &id);
        while (sz < bound && mask_unique(mk >> 4) && mask_unique(mk & 15u)) {   // FindForward, :110-121
            ++sz;
            kmer_shl<W>(cur, k, (uint32_t) c_unique_next[mk & 15u], nxt);
#pragma unroll
            for (int j = 0; j < W; ++j) cur[j] = nxt[j];
            mk = oriented_mask<W>(m, masks, cur, k, &id);
        }
        ++sz;
        if (!mask_unique(mk >> 4) || (mk & 15u) != 0) sz = 0;
        len[c] = sz;
        uint32_t l = sz == 0 ? 0xFFFFFFFFu : sz;
        if (l > maxlen) maxlen = l;
    }
    // pass 2: isolate the tips shorter than the longest branch (RemoveTips, :131-139)
    uint32_t rem = 0;
    for (uint32_t c = 0; c < 4; ++c) {
        if (!(mask & (1u << c)) || len[c] == 0 || len[c] >= maxlen) continue;
        uint64_t cur[W], nxt[W];
        kmer_shl<W>(x, k, c, cur);
        for (uint32_t s = 0; s < len[c]; ++s) {
            uint64_t id;
            uint32_t mk = oriented_mask<W>(m, masks, cur, k, &id);
            kill[id] = 1;
            if (s + 1 < len[c]) {
                kmer_shl<W>(cur, k, (uint32_t) c_unique_next[mk & 15u], nxt);
#pragma unroll
                for (int j = 0; j < W; ++j) cur[j] = nxt[j];
            }
        }
        rem += len[c];
    }
    if (rem) {
        atomicAdd(removed, (unsigned long long) rem);
        tipped[t] = 1;
    }
}

__global__ void tipclip_apply_kernel(uint8_t *__restrict__ masks, const uint8_t *__restrict__ kill, uint64_t n) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && kill[i]) masks[i] = 0;
}

// Phase 2: RemoveInconsistentForwardLinks (:20-35) on the junctions that lost a tip.
template<int W>
__global__ void __launch_bounds__(128) tipclip_links_kernel(MphfDev m, const uint64_t *__restrict__ kmers, uint64_t n, int k, uint8_t *__restrict__ masks,
                                                           const uint8_t *__restrict__ tipped) {
    uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n || !tipped[t]) return;
    uint64_t i = t >> 1;
    int strand = (int) (t & 1);
    uint64_t y[W], x[W];
    load_rec<W>(kmers, i, y);
    if (strand) kmer_rc<W>(y, k, x);
    else {
#pragma unroll
        for (int j = 0; j < W; ++j) x[j] = y[j];
    }
    bool minimal;
    uint64_t id0 = mphf_lookup_oriented<W>(m, x, k, &minimal);
    uint32_t raw = masks[id0];
    uint32_t mask = minimal ? raw : mask_conj(raw);
    uint32_t first = kmer_base(x, 0);
    uint32_t del = 0;
    for (uint32_t c = 0; c < 4; ++c) {
        if (!(mask & (1u << c))) continue;
        uint64_t nx[W], id;
        kmer_shl<W>(x, k, c, nx);
        uint32_t nm = oriented_mask<W>(m, masks, nx, k, &id);
        if (!(nm & (1u << (4 + first)))) del |= 1u << (minimal ? c : 7u - c);   // DeleteOutgoing(kh, c)
    }
    if (del) mask_and_not(masks, id0, del);
}

// The clipper in three steps, so that the hash-sharded path (shard.cu) can run step 1 and 3 over every GPU's own k-mers and combine the
// kill lists / mask slices in between: the masks and the kill list span the WHOLE index (ext->size), `kmers` may be one GPU's shard.
template<int W>
static void tipclip_find_w(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, uint64_t bound, TipClipState &st) {
    uint64_t n = kmers->size;
    MphfDev m = mphf_dev(mphf);
    st.kill.alloc(ctx, ext->size + 4); st.tipped.alloc(ctx, 2 * n + 4); st.removed.alloc(ctx, 1);
    st.kill.zero(); st.tipped.zero(); st.removed.zero();
    uint32_t b32 = bound > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t) bound;
    if (n) LAUNCH(ctx, tipclip_find_kernel<W>, div_up(2 * n, 128), 128, 0, m, kmers->data.p, n, (int) kmers->k, ext->masks.p, b32, st.kill.p, st.tipped.p, st.removed.p);
}

void tipclip_find(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, uint64_t bound, TipClipState &st) {
    switch (kmers->words) {
        case 1: tipclip_find_w<1>(ctx, kmers, mphf, ext, bound, st); break;
        case 2: tipclip_find_w<2>(ctx, kmers, mphf, ext, bound, st); break;
        case 3: tipclip_find_w<3>(ctx, kmers, mphf, ext, bound, st); break;
        default: tipclip_find_w<4>(ctx, kmers, mphf, ext, bound, st); break;
    }
}

void tipclip_apply(sb200_ctx *ctx, sb200_ext *ext, TipClipState &st) {
    if (ext->size) LAUNCH(ctx, tipclip_apply_kernel, div_up(ext->size, 256), 256, 0, ext->masks.p, st.kill.p, ext->size);
}

template<int W>
static void tipclip_links_w(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, TipClipState &st) {
    uint64_t n = kmers->size;
    MphfDev m = mphf_dev(mphf);
    if (n) LAUNCH(ctx, tipclip_links_kernel<W>, div_up(2 * n, 128), 128, 0, m, kmers->data.p, n, (int) kmers->k, ext->masks.p, st.tipped.p);
}

// returns the number of k-mers this GPU's junctions removed
uint64_t tipclip_links(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, TipClipState &st) {
    switch (kmers->words) {
        case 1: tipclip_links_w<1>(ctx, kmers, mphf, ext, st); break;
        case 2: tipclip_links_w<2>(ctx, kmers, mphf, ext, st); break;
        case 3: tipclip_links_w<3>(ctx, kmers, mphf, ext, st); break;
        default: tipclip_links_w<4>(ctx, kmers, mphf, ext, st); break;
    }
    unsigned long long r = 0;
    ctx->fetch(&r, st.removed.p, 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return r;
}

uint64_t tipclip(sb200_ctx *ctx, const sb200_kmers *kmers, const sb200_mphf *mphf, sb200_ext *ext, uint64_t bound) {
    TipClipState st;
    tipclip_find(ctx, kmers, mphf, ext, bound, st);
    tipclip_apply(ctx, ext, st);
    const uint64_t r = tipclip_links(ctx, kmers, mphf, ext, st);
    if (r) { ext->succ_valid = false; ext->masks_edited = true; }   // a junction that lost a tip may now have a single successor the racing writes did not keep
    return r;
}

}  // namespace sb200

// mphf.cu — BooPHF-identical minimal perfect hash, one per hash bucket, built level-synchronously on the GPU.
//
// Replaces KMerIndexBuilder::BuildIndex(index, storage) (C/utils/kmer_mph/kmer_index_builder.hpp:383-433) and
// boomphf::mphf::build/processLevel/processHash/insertIntoLevel/build_ranks (E/boomphf/BooPHF.h:423-441,677-696,
// 634-675,626-632,289-301).  All buckets advance through a level together: one kernel per level drops the keys that
// were placed at the previous level (their bit survived collision clearing), hashes the rest with the level's
// function and sets bits with atomicOr; a second hit on a bit is recorded in a collision vector instead of being
// resolved in place, so "placed" is simply bit & ~collision and no pass over the bit-vectors is needed between
// levels.  The order in which keys reach a level does not influence any bit, hence the active list is compacted
// with a warp-aggregated atomic append instead of a scan.  Level sizes come from the host in double precision with
// the reference's own expression, so the bit-vectors, rank samples and indices are bit-exact.
#include <math.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "kmer_ops.cuh"
#include "kmer_set.cuh"
#include "mphf.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

struct ActiveKey {
    uint64_t s0, s1;   // BooPHF hash state (levels >= 2 advance it with xorshift128*)
    uint32_t bucket;
    uint32_t pad;      // index of the key in the set
};

// level 0: hash every key, set its level-0 bit, start the active list (all keys, file order)
template<int W>
__global__ void __launch_bounds__(256) mphf_level0_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t B,
                                                         const uint64_t *__restrict__ domain, const uint64_t *__restrict__ word_off,
                                                         unsigned long long *__restrict__ bits, unsigned long long *__restrict__ coll,
                                                         ActiveKey *__restrict__ active) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r[W];
    load_rec<W>(keys, i, r);
    ActiveKey a;
    a.bucket = kmer_bucket<W>(r, B);
    a.pad = (uint32_t) i;   // position in the key set: where mphf_level_kernel reports the key's final bit
    xxh3_128<W>(r, a.s0, a.s1);
    uint64_t t = (uint64_t) a.bucket * MPHF_LEVELS;
    uint64_t pos = __umul64hi(a.s0, domain[t]);
    uint64_t w = word_off[t] + (pos >> 6);
    unsigned long long bit = 1ULL << (pos & 63);
    unsigned long long old = atomicOr(&bits[w], bit);
    if (old & bit) atomicOr(&coll[w], bit);
    if (active) active[i] = a;   // nullptr: level 1 starts from the keys again (mphf_level1_kernel) instead of from a 24-byte state per key
}

// level 1 straight from the keys: the state of a key after level 0 is a function of the key, so instead of writing 24 bytes per key in
// level 0 and reading them back here (3.5 GB for 73 M keys), level 1 reads the 16-byte keys again and recomputes bucket + XXH3_128; only
// the survivors (22 %) get a state record.  Placement test, level-1 insertion and the block-aggregated append are those of
// mphf_level_kernel below.
constexpr int MPHF_L1_ITEMS = 4;
template<int W>
__global__ void __launch_bounds__(256) mphf_level1_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t B, ActiveKey *__restrict__ out,
                                                         uint32_t *__restrict__ n_out, const uint64_t *__restrict__ domain,
                                                         const uint64_t *__restrict__ word_off, unsigned long long *__restrict__ bits,
                                                         unsigned long long *__restrict__ coll, uint32_t *__restrict__ place /* or nullptr */) {
    constexpr int ITEMS = MPHF_L1_ITEMS;
    __shared__ uint32_t s_wcnt[8];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t) blockDim.x * ITEMS;
    for (uint64_t base = (uint64_t) blockIdx.x * tile; base < n; base += (uint64_t) gridDim.x * tile) {
        ActiveKey a[ITEMS];
        bool keep[ITEMS];
        uint32_t before[ITEMS];
        uint32_t wtotal = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint64_t i = base + (uint64_t) j * blockDim.x + threadIdx.x;
            keep[j] = false;
            if (i < n) {
                uint64_t r[W];
                load_rec<W>(keys, i, r);
                a[j].bucket = kmer_bucket<W>(r, B);
                a[j].pad = (uint32_t) i;
                xxh3_128<W>(r, a[j].s0, a[j].s1);
                const uint64_t t = (uint64_t) a[j].bucket * MPHF_LEVELS;
                const uint64_t pos = __umul64hi(a[j].s0, domain[t]);
                const uint64_t w = word_off[t] + (pos >> 6);
                const unsigned long long bit = 1ULL << (pos & 63);
                const bool placed = (bits[w] & bit) && !(coll[w] & bit);
                keep[j] = !placed;
                if (placed && place) place[i] = (uint32_t) ((w << 6) | (pos & 63));
            }
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t m = __ballot_sync(0xffffffffu, keep[j]);
            before[j] = wtotal + (uint32_t) __popc(m & ((1u << lane) - 1u));
            wtotal += (uint32_t) __popc(m);
        }
        if (lane == 0) s_wcnt[warp] = wtotal;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const uint32_t c = s_wcnt[w]; s_wcnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(n_out, tot) : 0u;
        }
        __syncthreads();
        const uint32_t wbase = s_base + s_wcnt[warp];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (!keep[j]) continue;
            const uint64_t t = (uint64_t) a[j].bucket * MPHF_LEVELS + 1;
            const uint64_t pos = __umul64hi(a[j].s1, domain[t]);
            const uint64_t w = word_off[t] + (pos >> 6);
            const unsigned long long bit = 1ULL << (pos & 63);
            const unsigned long long old = atomicOr(&bits[w], bit);
            if (old & bit) atomicOr(&coll[w], bit);
            out[wbase + before[j]] = a[j];
        }
        __syncthreads();
    }
}

// level l >= 1: keep the keys that were not placed at level l-1, insert them into level l (if l is a bit level).
// The survivors are appended to `out` with ONE global atomic per CTA pass of 1024 keys: a warp-aggregated atomic per 32 keys put
// 2.3 M atomics on a single address for level 1 (all 73 M keys are examined there), and those serialise in L2 — 3.8 ms for a
// kernel whose memory traffic is worth 0.4 ms (profiles/r1b launch list).  The order of the survivors does not matter: the bits a
// level sets do not depend on it.
constexpr int MPHF_LEVEL_ITEMS = 4;
__global__ void __launch_bounds__(256) mphf_level_kernel(int level, const ActiveKey *__restrict__ in, const uint32_t *__restrict__ n_in,
                                                        ActiveKey *__restrict__ out, uint32_t *__restrict__ n_out,
                                                        const uint64_t *__restrict__ domain, const uint64_t *__restrict__ word_off,
                                                        unsigned long long *__restrict__ bits, unsigned long long *__restrict__ coll,
                                                        uint32_t *__restrict__ place /* or nullptr */) {
    constexpr int ITEMS = MPHF_LEVEL_ITEMS;
    __shared__ uint32_t s_wcnt[8];
    __shared__ uint32_t s_base;
    const uint32_t n = *n_in;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t) blockDim.x * ITEMS;
    for (uint64_t base = (uint64_t) blockIdx.x * tile; base < n; base += (uint64_t) gridDim.x * tile) {
        ActiveKey a[ITEMS];
        bool keep[ITEMS];
        uint32_t before[ITEMS];   // survivors of this warp ahead of this lane's item j
        uint32_t wtotal = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint64_t i = base + (uint64_t) j * blockDim.x + threadIdx.x;
            keep[j] = false;
            if (i < n) {
                a[j] = in[i];
                // Position this key probed at level-1.  For levels >= 2 the hash was xs_next() = new s1 + old s1, and after
                // that call s0 holds the old s1, so it can be rebuilt from the saved state.
                uint64_t prev_hash;
                if (level - 1 == 0) prev_hash = a[j].s0;
                else if (level - 1 == 1) prev_hash = a[j].s1;
                else prev_hash = a[j].s1 + a[j].s0;   // xs_next returned (new s1 + old s1); after the call s0 == old s1
                const uint64_t t = (uint64_t) a[j].bucket * MPHF_LEVELS + (level - 1);
                const uint64_t pos = __umul64hi(prev_hash, domain[t]);
                const uint64_t w = word_off[t] + (pos >> 6);
                const unsigned long long bit = 1ULL << (pos & 63);
                const bool placed = (bits[w] & bit) && !(coll[w] & bit);
                keep[j] = !placed;
                if (placed && place) place[a[j].pad] = (uint32_t) ((w << 6) | (pos & 63));
            }
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const uint32_t m = __ballot_sync(0xffffffffu, keep[j]);
            before[j] = wtotal + (uint32_t) __popc(m & ((1u << lane) - 1u));
            wtotal += (uint32_t) __popc(m);
        }
        if (lane == 0) s_wcnt[warp] = wtotal;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const uint32_t c = s_wcnt[w]; s_wcnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(n_out, tot) : 0u;
        }
        __syncthreads();
        const uint32_t wbase = s_base + s_wcnt[warp];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (!keep[j]) continue;
            if (level < MPHF_LEVELS - 1) {
                uint64_t h = (level == 1) ? a[j].s1 : xs_next(a[j].s0, a[j].s1);
                const uint64_t t = (uint64_t) a[j].bucket * MPHF_LEVELS + level;
                const uint64_t pos = __umul64hi(h, domain[t]);
                const uint64_t w = word_off[t] + (pos >> 6);
                const unsigned long long bit = 1ULL << (pos & 63);
                const unsigned long long old = atomicOr(&bits[w], bit);
                if (old & bit) atomicOr(&coll[w], bit);
            }
            out[wbase + before[j]] = a[j];
        }
        __syncthreads();   // s_wcnt / s_base are rewritten by the next pass
    }
}

__global__ void mphf_clear_popc_kernel(unsigned long long *__restrict__ bits, const unsigned long long *__restrict__ coll, uint64_t nwords,
                                       uint32_t *__restrict__ pc) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    unsigned long long v = bits[i] & ~coll[i];
    bits[i] = v;
    pc[i] = (uint32_t) __popcll(v);
}

// one block per (bucket, level): rank sample j = set bits in the bucket's words before word 8j of this level
__global__ void mphf_ranks_kernel(const uint32_t *__restrict__ pc_scan, const uint64_t *__restrict__ word_off,
                                  const uint64_t *__restrict__ rank_off, const uint64_t *__restrict__ nchar, uint64_t t0, uint64_t n_levels_total,
                                  uint64_t *__restrict__ ranks) {
    uint64_t t = t0 + blockIdx.x;
    if (t >= n_levels_total) return;
    uint64_t nc = nchar[t];
    if (nc == 0) return;
    uint64_t bucket_first = word_off[(t / MPHF_LEVELS) * MPHF_LEVELS];
    uint32_t base = pc_scan[bucket_first];
    uint64_t ns = (nc + 7) / 8;
    for (uint64_t j = threadIdx.x; j < ns; j += blockDim.x)
        ranks[rank_off[t] + j] = (uint64_t) (pc_scan[word_off[t] + 8 * j] - base);
}

__global__ void mphf_bucket_ends_kernel(const uint32_t *__restrict__ pc_scan, const uint64_t *__restrict__ word_off, uint32_t B,
                                        uint64_t total_words, uint32_t *__restrict__ ends) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    ends[b] = pc_scan[b < B ? word_off[(uint64_t) b * MPHF_LEVELS] : total_words];
}

__global__ void add_u32_kernel(uint32_t *__restrict__ a, uint64_t n, uint32_t v) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += v;
}

// slice_only (sharded construction, shard.cu): only the words of the shard's own buckets are initialised and processed — every other
// rank's slice arrives with the all-gather — and pc_scan is valid over the own slice at once (local prefix + the keys of all buckets
// before it), so that the shard's k-mers get their MPHF indices BEFORE the exchange.  Without it the foreign slices are zero: the
// contract of sb200_mphf_build_sharded ("sum the arrays over the GPUs").
template<int W>
static sb200_mphf *mphf_build_w(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes, bool slice_only) {
    // global_sizes != nullptr: `ks` is one GPU's shard (whole buckets); level geometry, rank offsets and segment starts are laid
    // out for ALL buckets, only the shard's keys are inserted.  Bit-vectors and rank samples of different shards are then
    // disjoint, so summing the arrays over the GPUs (all-reduce) yields exactly the single-GPU index.
    sb200_mphf *m = new sb200_mphf();
    m->ctx = ctx;
    uint32_t B = ks->num_buckets;
    uint64_t total_keys = 0;
    for (uint32_t b = 0; b < B; ++b)
        total_keys += global_sizes ? global_sizes[b] : ks->bucket_starts_host[b + 1] - ks->bucket_starts_host[b];
    m->num_buckets = B; m->words = W; m->total = total_keys;
    SB200_REQUIRE(total_keys < (1ull << 32), "more than 2^32-1 k-mers in one index");
    size_t NL = (size_t) B * MPHF_LEVELS;
    m->domain_host.assign(NL, 0); m->word_off_host.assign(NL, 0); m->rank_off_host.assign(NL, 0);
    std::vector<uint64_t> nchar(NL, 0);
    m->segment_starts_host.assign((size_t) B + 1, 0);
    m->lastbitsetrank_host.assign(B, 0);
    m->bucket_sizes_host.assign(B, 0);
    uint64_t words = 0, ranks = 0;
    for (uint32_t b = 0; b < B; ++b) {
        uint64_t n = global_sizes ? global_sizes[b] : ks->bucket_starts_host[b + 1] - ks->bucket_starts_host[b];
        m->segment_starts_host[b + 1] = n;
        m->bucket_sizes_host[b] = n;
        if (n > 0) {
            // boomphf::mphf::init + setup, BooPHF.h:409-421,575-588 (gamma = 4.0, kmer_index_builder.hpp:401-404)
            double gamma = 4.0;
            uint64_t hash_domain = (uint64_t) (size_t) ceil(double(n) * gamma);
            double p = 1.0 - pow(((gamma * (double) n - 1) / (gamma * (double) n)), (double) (n - 1));
            for (int l = 0; l < MPHF_LEVELS; ++l) {
                uint64_t d = ((uint64_t(hash_domain * pow(p, l)) + 63) / 64) * 64;
                if (d == 0) d = 64;
                size_t t = (size_t) b * MPHF_LEVELS + l;
                m->domain_host[t] = d;
                nchar[t] = 1 + d / 64;                       // bitVector(n): _nchar = 1 + n/64, BooPHF.h:144-148
                m->word_off_host[t] = words;
                m->rank_off_host[t] = ranks;
                words += nchar[t];
                ranks += (nchar[t] + 7) / 8;                 // build_ranks: one sample per 512 bits, :289-301
            }
        } else {
            for (int l = 0; l < MPHF_LEVELS; ++l) {
                size_t t = (size_t) b * MPHF_LEVELS + l;
                m->word_off_host[t] = words;
                m->rank_off_host[t] = ranks;
            }
        }
    }
    // kmer_index_builder.hpp:427-428 — the prefix loop stops at i < segments: the last entry stays a bucket size
    for (uint32_t i = 1; i < B; ++i) m->segment_starts_host[i] += m->segment_starts_host[i - 1];
    m->total_words = words; m->total_ranks = ranks;

    m->domain.alloc(ctx, NL); m->word_off.alloc(ctx, NL); m->rank_off.alloc(ctx, NL);
    m->segment_starts.alloc(ctx, (size_t) B + 1);
    DevBuf<uint64_t> nchar_dev(ctx, NL);
    CUDA_CHECK(cudaMemcpyAsync(m->domain.p, m->domain_host.data(), NL * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(m->word_off.p, m->word_off_host.data(), NL * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(m->rank_off.p, m->rank_off_host.data(), NL * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(nchar_dev.p, nchar.data(), NL * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(m->segment_starts.p, m->segment_starts_host.data(), ((size_t) B + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    // own word range [w0, w1): the buckets this table holds keys of
    uint64_t w0 = 0, w1 = words;
    uint32_t b_lo = 0, b_hi = B;
    if (slice_only) {
        b_lo = B; b_hi = 0;
        for (uint32_t b = 0; b < B; ++b)
            if (ks->bucket_starts_host[b + 1] > ks->bucket_starts_host[b]) { b_lo = std::min(b_lo, b); b_hi = std::max(b_hi, b + 1); }
        if (b_lo >= b_hi) { b_lo = b_hi = 0; w0 = w1 = 0; }
        else {
            w0 = m->word_off_host[(size_t) b_lo * MPHF_LEVELS];
            w1 = b_hi < B ? m->word_off_host[(size_t) b_hi * MPHF_LEVELS] : words;
        }
    }
    const uint64_t own = w1 - w0;
    m->bits.alloc(ctx, words + 1);
    m->ranks.alloc(ctx, ranks + 1);
    DevBuf<uint64_t> coll(ctx, own + 1);
    coll.zero();
    if (slice_only) {
        CUDA_CHECK(cudaMemsetAsync(m->bits.p + w0, 0, own * 8, ctx->stream));
        CUDA_CHECK(cudaMemsetAsync(m->bits.p + words, 0, 8, ctx->stream));
    } else m->bits.zero();
    uint64_t *coll_base = coll.p - w0;   // indexed by global word like `bits` (only [w0, w1] is ever touched)

    uint64_t n = ks->size;
    // whole-table build with 32-bit bit positions: remember where every key lands (see sb200_mphf::place)
    // (a shard records them too: the bit positions are global, and sb200_mphf_complete builds pc_scan once the index is whole)
    const bool keep_place = (words + 1) * 64 < (1ull << 32) && n > 0;
    if (keep_place) m->place.alloc(ctx, n);
    // state records exist from level 1 on, for the keys level 0 did not place (~22 % of the keys at gamma = 4); the lists are sized for
    // all keys (address space of the caching allocator, never touched beyond the survivors)
    const uint64_t cap_a = n;
    uint32_t n32 = (uint32_t) n;
    DevBuf<uint32_t> counters(ctx, MPHF_LEVELS + 1);
    DevBuf<ActiveKey> act_a, act_b;
    ActiveKey *src = nullptr, *dst = nullptr;
    unsigned grid = (unsigned) std::min<uint64_t>(div_up(n, 256 * MPHF_LEVEL_ITEMS), (uint64_t) ctx->num_sms * 16);
    int first_level = 1;
    if (n && !ctx->mphf_state_per_key) {
        act_a.alloc(ctx, cap_a); act_b.alloc(ctx, cap_a);
        counters.zero();
        LAUNCH(ctx, mphf_level0_kernel<W>, div_up(n, 256), 256, 0, ks->data.p, n, B, m->domain.p, m->word_off.p,
               (unsigned long long *) m->bits.p, (unsigned long long *) coll_base, (ActiveKey *) nullptr);
        auto mphf_level1_kernel_ = mphf_level1_kernel<W>;
        LAUNCH(ctx, mphf_level1_kernel_, grid, 256, 0, ks->data.p, n, B, act_a.p, counters.p + 1, m->domain.p, m->word_off.p,
               (unsigned long long *) m->bits.p, (unsigned long long *) coll_base, m->place.p);
        src = act_a.p; dst = act_b.p;
        first_level = 2;
    }
    if (n && ctx->mphf_state_per_key) {
        act_a.alloc(ctx, n); act_b.alloc(ctx, n);
        counters.zero();
        CUDA_CHECK(cudaMemcpyAsync(counters.p, &n32, 4, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(ctx, mphf_level0_kernel<W>, div_up(n, 256), 256, 0, ks->data.p, n, B, m->domain.p, m->word_off.p,
               (unsigned long long *) m->bits.p, (unsigned long long *) coll_base, act_a.p);
        src = act_a.p; dst = act_b.p;
    }
    if (!n) counters.zero();
    for (int l = first_level; l < MPHF_LEVELS && n; ++l) {   // (a rank that owns no k-mer at all only contributes zeroed arrays)
        LAUNCH(ctx, mphf_level_kernel, grid, 256, 0, l, src, counters.p + (l - 1), dst, counters.p + l, m->domain.p, m->word_off.p,
               (unsigned long long *) m->bits.p, (unsigned long long *) coll_base, m->place.p);
        std::swap(src, dst);
    }
    uint32_t final_keys = 0;
    ctx->fetch(&final_keys, counters.p + (MPHF_LEVELS - 1), 4);

    DevBuf<uint32_t> pc;
    if (keep_place) m->pc_scan.alloc(ctx, words + 1);
    else pc.alloc(ctx, words + 1);
    uint32_t *pcp = keep_place ? m->pc_scan.p : pc.p;
    // (entry w1 is the scan's total: a zero input; for a slice it lies in the next rank's region, rebuilt by mphf_complete)
    LAUNCH(ctx, mphf_clear_popc_kernel, div_up(own + 1, 256), 256, 0, (unsigned long long *) m->bits.p + w0,
           (const unsigned long long *) coll.p, own + 1, pcp + w0);
    if (slice_only) CUDA_CHECK(cudaMemsetAsync(pcp + w1, 0, 4, ctx->stream));
    exclusive_scan<uint32_t>(ctx, pcp + w0, own + 1, nullptr);
    if (slice_only && own && m->segment_starts_host[b_lo])
        LAUNCH(ctx, add_u32_kernel, div_up(own + 1, 256), 256, 0, pcp + w0, own + 1, (uint32_t) m->segment_starts_host[b_lo]);
    CUDA_CHECK(cudaMemsetAsync(m->ranks.p, 0, (ranks + 1) * 8, ctx->stream));   // buckets without keys here (other GPUs' shards) leave no garbage
    if (b_hi > b_lo)
        LAUNCH(ctx, mphf_ranks_kernel, (unsigned) ((b_hi - b_lo) * MPHF_LEVELS), 128, 0, pcp, m->word_off.p, m->rank_off.p, nchar_dev.p,
               (uint64_t) b_lo * MPHF_LEVELS, (uint64_t) b_hi * MPHF_LEVELS, m->ranks.p);
    // _lastbitsetrank per bucket = set bits of the whole bucket (a slice: mphf_complete fills them in once the index is whole)
    if (!slice_only) {
        std::vector<uint32_t> ends(B + 1, 0);
        DevBuf<uint32_t> ends_dev(ctx, (size_t) B + 1);
        LAUNCH(ctx, mphf_bucket_ends_kernel, div_up((uint64_t) B + 1, 256), 256, 0, pcp, m->word_off.p, B, words, ends_dev.p);
        ctx->fetch(ends.data(), ends_dev.p, ((size_t) B + 1) * 4);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t b = 0; b < B; ++b) m->lastbitsetrank_host[b] = ends[b + 1] - ends[b];
    } else CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    m->final_level_keys = final_keys;
    if (final_keys != 0) {
        delete m;
        SB200_REQUIRE(false, "BooPHF exact-map level reached (probability ~1e-16 per key): not supported on the GPU path");
    }
    return m;
}

static sb200_mphf *mphf_build_any(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes, bool slice_only) {
    switch (ks->words) {
        case 1: return mphf_build_w<1>(ctx, ks, global_sizes, slice_only);
        case 2: return mphf_build_w<2>(ctx, ks, global_sizes, slice_only);
        case 3: return mphf_build_w<3>(ctx, ks, global_sizes, slice_only);
        default: return mphf_build_w<4>(ctx, ks, global_sizes, slice_only);
    }
}
sb200_mphf *mphf_build(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes) { return mphf_build_any(ctx, ks, global_sizes, false); }
sb200_mphf *mphf_build_slice(sb200_ctx *ctx, const sb200_kmers *ks, const uint64_t *global_sizes) { return mphf_build_any(ctx, ks, global_sizes, true); }

__global__ void mphf_popc_kernel(const uint64_t *__restrict__ bits, uint64_t nwords, uint32_t *__restrict__ pc) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) pc[i] = (uint32_t) __popcll(bits[i]);
}

// After the shards' bit-vectors have been summed (all-reduce), every GPU holds the complete index: the per-word prefix popcounts that
// let a lookup rank with one read (mphf_lookup) can be built for it too.
void mphf_complete(sb200_ctx *ctx, sb200_mphf *m) {
    const uint64_t words = m->total_words;
    if (m->total == 0) return;
    // per-word prefix popcounts of the summed bit-vectors: kept for one-read ranks when bit positions fit 32 bits, and in any case
    // the source of every bucket's _lastbitsetrank (a sharded build only saw its own buckets' bits: the serialised index of a
    // completed shard must carry the totals of ALL buckets, BooPHF.h:514-532)
    const bool keep = (words + 1) * 64 < (1ull << 32);
    DevBuf<uint32_t> tmp;
    uint32_t *pcp;
    if (keep) { m->pc_scan.alloc(ctx, words + 1); pcp = m->pc_scan.p; }
    else { tmp.alloc(ctx, words + 1); pcp = tmp.p; }
    LAUNCH(ctx, mphf_popc_kernel, div_up(words + 1, 256), 256, 0, m->bits.p, words + 1, pcp);
    exclusive_scan<uint32_t>(ctx, pcp, words + 1, nullptr);
    const uint32_t B = m->num_buckets;
    std::vector<uint32_t> ends((size_t) B + 1, 0);
    DevBuf<uint32_t> ends_dev(ctx, (size_t) B + 1);
    LAUNCH(ctx, mphf_bucket_ends_kernel, div_up((uint64_t) B + 1, 256), 256, 0, pcp, m->word_off.p, B, words, ends_dev.p);
    ctx->fetch(ends.data(), ends_dev.p, ((size_t) B + 1) * 4);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t b = 0; b < B; ++b) m->lastbitsetrank_host[b] = ends[b + 1] - ends[b];
}

MphfDev mphf_dev(const sb200_mphf *m) {
    MphfDev d;
    d.domain = m->domain.p; d.word_off = m->word_off.p; d.rank_off = m->rank_off.p; d.segment_starts = m->segment_starts.p;
    d.bits = m->bits.p; d.ranks = m->ranks.p; d.num_buckets = m->num_buckets;
    d.pc_scan = (m->pc_scan.p && !m->ctx->no_place) ? m->pc_scan.p : nullptr;
    d.wblk = nullptr;
    return d;
}

template<int W>
__global__ void mphf_lookup_kernel(MphfDev m, const uint64_t *__restrict__ recs, uint64_t n, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r[W];
    load_rec<W>(recs, i, r);
    out[i] = mphf_lookup<W>(m, r);
}

// KMerIndex::seq_idx for n records resident on the device
void mphf_lookup_device(sb200_ctx *ctx, const sb200_mphf *m, const uint64_t *recs_dev, uint64_t n, uint64_t *out_dev) {
    MphfDev d = mphf_dev(m);
    unsigned g = div_up(n, 256);
    if (n == 0) return;
    switch (m->words) {
        case 1: LAUNCH(ctx, mphf_lookup_kernel<1>, g, 256, 0, d, recs_dev, n, out_dev); break;
        case 2: LAUNCH(ctx, mphf_lookup_kernel<2>, g, 256, 0, d, recs_dev, n, out_dev); break;
        case 3: LAUNCH(ctx, mphf_lookup_kernel<3>, g, 256, 0, d, recs_dev, n, out_dev); break;
        default: LAUNCH(ctx, mphf_lookup_kernel<4>, g, 256, 0, d, recs_dev, n, out_dev); break;
    }
}

// KMerIndex::seq_idx(const Seq &) for ONE key from host code (kmer_index.hpp:85-90, boomphf::mphf::lookup BooPHF.h:465-487, bitVector::rank
// :303-314): the consumers of the index after the path (link records, edge index, read mapping) look keys up one at a time from CPU
// threads.  The 5.8 bits per key of bit-vectors + rank samples are copied to the host once; the lookup itself is the reference's
// arithmetic (same XXH3 specialisations as the kernels, kmer_ops.cuh compiled for the host).  Thread-safe after the first call.
template<int W>
static uint64_t mphf_lookup_host(const sb200_mphf *m, const uint64_t *rec) {
    const uint32_t b = kmer_bucket<W>(rec, m->num_buckets);
    uint64_t s0, s1;
    xxh3_128<W>(rec, s0, s1);
    for (int l = 0; l < MPHF_LEVELS - 1; ++l) {
        const size_t t = (size_t) b * MPHF_LEVELS + l;
        const uint64_t h = (l == 0) ? s0 : (l == 1) ? s1 : xs_next(s0, s1);
        const uint64_t d = m->domain_host[t];
        if (d == 0) return ~0ULL;
        const uint64_t pos = mulhi64(h, d);
        const uint64_t *bv = m->bits_host.data() + m->word_off_host[t];
        const uint64_t word = bv[pos >> 6];
        if ((word >> (pos & 63)) & 1ULL) {
            uint64_t r = m->ranks_host[m->rank_off_host[t] + (pos >> 9)];
            for (uint64_t w = (pos >> 9) << 3; w < (pos >> 6); ++w) r += (uint64_t) __builtin_popcountll(bv[w]);
            r += (uint64_t) __builtin_popcountll(word & ((1ULL << (pos & 63)) - 1ULL));
            return m->segment_starts_host[b] + r;
        }
    }
    return ~0ULL;
}

uint64_t mphf_seq_idx_host(sb200_mphf *m, const uint64_t *rec) {
    if (!m->host_copy) {
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        if (!m->host_copy) {
            sb200_ctx *ctx = m->ctx;
            CUDA_CHECK(cudaSetDevice(ctx->device));
            m->bits_host.resize(m->total_words + 1);
            m->ranks_host.resize(m->total_ranks + 1);
            CUDA_CHECK(cudaMemcpy(m->bits_host.data(), m->bits.p, m->total_words * 8, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(m->ranks_host.data(), m->ranks.p, m->total_ranks * 8, cudaMemcpyDeviceToHost));
            m->host_copy = true;
        }
    }
    switch (m->words) {
        case 1: return mphf_lookup_host<1>(m, rec);
        case 2: return mphf_lookup_host<2>(m, rec);
        case 3: return mphf_lookup_host<3>(m, rec);
        default: return mphf_lookup_host<4>(m, rec);
    }
}

// KMerIndex::serialize (kmer_index.hpp:99-105) of per-bucket mphf::save (BooPHF.h:514-532) of bitVector::save (:316-323)
uint64_t mphf_serialize_host(const sb200_mphf *m, const uint64_t *bits_host, const uint64_t *ranks_host, uint8_t *out);

void mphf_serialize_device(sb200_ctx *ctx, const sb200_mphf *m, uint8_t *out_dev);

// The byte stream is assembled on the device (every bit-vector and rank array copied to its place) and comes home as ONE copy straight
// into the caller's buffer (pinned or not); the host then fills in the few small fields between the arrays.
uint64_t mphf_serialize(const sb200_mphf *m, uint8_t *out) {
    sb200_ctx *ctx = m->ctx;
    const uint64_t size = mphf_serialize_host(m, nullptr, nullptr, nullptr);
    if (!out) return size;
    DevBuf<uint8_t> dev(ctx, size + 8);
    mphf_serialize_device(ctx, m, dev.p);
    CUDA_CHECK(cudaMemcpyAsync(out, dev.p, size, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return mphf_serialize_host(m, nullptr, nullptr, out);   // small fields only; the bulk areas are skipped over
}

// bits_host / ranks_host: host copies of the device arrays.  out == nullptr: size query.  out != nullptr with bits_host ==
// nullptr: only the small fields are written and the bit-vector / rank bytes are skipped over (they were put in place by
// mphf_serialize_device + one D2H copy).
uint64_t mphf_serialize_host(const sb200_mphf *m, const uint64_t *bits_host, const uint64_t *ranks_host, uint8_t *out) {
    uint64_t total = 0;
    uint8_t *p = out;
    auto put = [&](const void *src, size_t n) {
        if (p) { memcpy(p, src, n); p += n; }
        total += n;
    };
    auto put_bulk = [&](const uint64_t *base, uint64_t first_word, size_t n) {
        if (p) {
            if (base) memcpy(p, base + first_word, n);
            p += n;
        }
        total += n;
    };
    uint64_t nseg = m->num_buckets;
    put(&nseg, 8);
    for (uint32_t b = 0; b < m->num_buckets; ++b) {
        double gamma = 4.0;
        int nb_levels = MPHF_LEVELS;
        uint64_t nelem = m->bucket_sizes_host[b];
        put(&gamma, 8);
        put(&nb_levels, 4);
        put(&m->lastbitsetrank_host[b], 8);
        put(&nelem, 8);
        for (int l = 0; l < MPHF_LEVELS; ++l) {
            size_t t = (size_t) b * MPHF_LEVELS + l;
            uint64_t size = m->domain_host[t];
            if (size == 0) {
                // Empty bucket: the reference leaves default-constructed bit-vectors (size 0, _nchar never set: what ITS save() writes for
                // them is undefined, BooPHF.h:137-140,316-323).  We write what its load() expects to read back for size 0 — resize(0)
                // makes one word (BooPHF.h:204-208,325-335) — so that KMerIndex::deserialize of these bytes is well-defined.
                uint64_t one = 1, zero = 0;
                put(&size, 8);
                put(&one, 8);
                put(&zero, 8);
                put(&zero, 8);   // no rank samples
                continue;
            }
            uint64_t nchar = 1 + size / 64;
            uint64_t nr = (nchar + 7) / 8;
            put(&size, 8);
            put(&nchar, 8);
            put_bulk(bits_host, m->word_off_host[t], nchar * 8);
            put(&nr, 8);
            if (nr) put_bulk(ranks_host, m->rank_off_host[t], nr * 8);
        }
        uint64_t nf = 0;
        put(&nf, 8);
    }
    put(m->segment_starts_host.data(), ((size_t) m->num_buckets + 1) * 8);
    return total;
}

// The same byte stream assembled on the device: every bit-vector and rank array is copied to its place in `out_dev`
// (`size` bytes, from mphf_serialize_host(m, 0, 0, 0)); the few small fields in between are left for the host to fill after
// the single D2H copy (mphf_serialize_host(m, nullptr, nullptr, host_copy)).  Saves the host a 50 MB scatter-gather.
struct SerSeg {
    uint64_t dst_byte, src_word, nwords;
    uint32_t is_rank, pad;
};
__global__ void __launch_bounds__(256) mphf_serialize_kernel(const SerSeg *__restrict__ segs, const uint64_t *__restrict__ bits,
                                                            const uint64_t *__restrict__ ranks, uint8_t *__restrict__ out) {
    const SerSeg sg = segs[blockIdx.x];
    const uint32_t *src = reinterpret_cast<const uint32_t *>((sg.is_rank ? ranks : bits) + sg.src_word);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + sg.dst_byte);   // the stream keeps 4-byte alignment (BooPHF.h:514-532 field sizes)
    for (uint64_t i = threadIdx.x; i < 2 * sg.nwords; i += blockDim.x) dst[i] = src[i];
}

void mphf_serialize_device(sb200_ctx *ctx, const sb200_mphf *m, uint8_t *out_dev) {
    std::vector<SerSeg> segs;
    uint64_t off = 8;
    for (uint32_t b = 0; b < m->num_buckets; ++b) {
        off += 8 + 4 + 8 + 8;
        for (int l = 0; l < MPHF_LEVELS; ++l) {
            size_t t = (size_t) b * MPHF_LEVELS + l;
            uint64_t size = m->domain_host[t];
            if (size == 0) { off += 32; continue; }   // empty bucket: size, nchar = 1, one zero word, no rank samples (host-written)
            uint64_t nchar = 1 + size / 64, nr = (nchar + 7) / 8;
            off += 16;
            segs.push_back(SerSeg{off, m->word_off_host[t], nchar, 0u, 0u});
            off += nchar * 8 + 8;
            if (nr) segs.push_back(SerSeg{off, m->rank_off_host[t], nr, 1u, 0u});
            off += nr * 8;
        }
        off += 8;
    }
    if (segs.empty()) return;
    DevBuf<SerSeg> d(ctx, segs.size());
    CUDA_CHECK(cudaMemcpyAsync(d.p, segs.data(), segs.size() * sizeof(SerSeg), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, mphf_serialize_kernel, (unsigned) segs.size(), 256, 0, d.p, m->bits.p, m->ranks.p, out_dev);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // `segs` (pageable) must outlive the H2D copy
}

}  // namespace sb200

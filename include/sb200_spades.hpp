// sb200_spades.hpp — the drop-in AT THE REFERENCE'S OWN TYPES.  Compiles against the SPAdes 3.15.4 source tree (assembler/src/common on
// the include path) and include/sb200.h; nothing else.  Where sb200_adapters.hpp mirrors the reference's class names for a stand-alone
// host program, the classes here ARE the reference's classes filled from the GPU:
//
//   sb200_spades::GpuKMerCounter            a genuine subclass of kmers::KMerCounter<RtSeq> (kmer_mph/kmer_index_builder.hpp:194-217):
//                                           Count / CountAll run on the B200 and return a real kmers::KMerDiskStorage<RtSeq> whose
//                                           bucket files kmers<i> (and final_kmers after merge()) hold the GPU's records — every
//                                           consumer of the reference (KMerIndexBuilder::BuildIndex(index, counter),
//                                           CoverageHashMapBuilder, edge-index builders, hammer) takes it unchanged.
//   sb200_spades::GpuExtensionIndexBuilder  the call shape of utils::DeBruijnExtensionIndexBuilder
//                                           (extension_index/kmer_extension_index_builder.hpp:62-106): fills a real
//                                           utils::DeBruijnExtensionIndex<> — KMerIndex through its own deserialize() of the GPU's
//                                           KMerIndex::serialize bytes (kmer_index.hpp:108-121), data_ (ph_map/perfect_hash_map.hpp:166)
//                                           from the GPU's mask array, final_kmers as the index's k-mer file — and returns the
//                                           (k+1)-mer storage like the reference.  The reference's UnbranchingPathExtractor,
//                                           EarlyTipClipperProcessor, FastGraphFromSequencesConstructor and GFAWriter then run on it
//                                           as they are (oracle/ref_dropin.cpp does exactly that and compares with the GPU's unitigs).
//   sb200_spades::PackedReads::FromGraphEdges + GpuKMerCounter(..., canonical_only, add_rc = false)
//                                           the counter behind the EDGE INDEX (assembly_graph/index/edge_index_builders.hpp:20-150:
//                                           DeBruijnGraphKMerSplitter / DeBruijnEdgeKMerSplitter feed every edge of the graph — conjugates
//                                           are edges of their own — through FillBufferFromSequence with the index's KmerFilter): SURVEY 8(f)4.
//   sb200_spades::GpuUnbranchingPaths       std::vector<Sequence> of UnbranchingPathExtractor::ExtractUnbranchingPathsAndLoops
//                                           (assembly_graph/construction/debruijn_graph_constructor.hpp:377-384) from the GPU.
// Streams: pass the FORWARD streams (single_binary_readers_for_libs(..., followed_by_rc = false, ...)); the RC stream the reference
// wraps around them (read_converter.cpp:212-242) is folded into the kernels (add_rc = 1).
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "io/reads/read_stream_vector.hpp"
#include "io/reads/single_read.hpp"
#include "sequence/sequence.hpp"
#include "utils/extension_index/kmer_extension_index.hpp"
#include "utils/kmer_mph/kmer_index_builder.hpp"

#include "sb200.h"

namespace sb200_spades {

namespace detail {
// KeyIteratingMap::kmers_ is private and only its friend KeyIteratingIndexBuilder sets it (ph_map/kmer_maps.hpp:163-196,
// perfect_hash_map_builder.hpp:47-70).  An explicit template instantiation may name a private member: this is the standard-conforming
// way to hand the k-mer file to the index without editing the reference or colliding with its builder's name.
template<class Tag> struct Stolen { static typename Tag::type ptr; };
template<class Tag> typename Tag::type Stolen<Tag>::ptr;
template<class Tag, typename Tag::type P>
struct Rob {
    struct Filler { Filler() { Stolen<Tag>::ptr = P; } };
    static Filler filler;
};
template<class Tag, typename Tag::type P> typename Rob<Tag, P>::Filler Rob<Tag, P>::filler;

typedef utils::DeBruijnExtensionIndex<> ExtIndex;
typedef utils::slim_kmer_index_traits<RtSeq> ExtTraits;
typedef utils::KeyIteratingMap<RtSeq, utils::InOutMask, ExtTraits, utils::DefaultStoring> ExtIterating;
typedef utils::PerfectHashMap<RtSeq, utils::InOutMask, ExtTraits, utils::DefaultStoring> ExtMap;
typedef utils::IndexWrapper<RtSeq, ExtTraits> ExtWrapper;
typedef kmers::KMerIndex<ExtTraits> ExtKMerIndex;
struct KmersTag { typedef fs::TmpFile ExtIterating::*type; };                     // set by friend KeyIteratingIndexBuilder in the reference
struct DataTag { typedef std::vector<utils::InOutMask> ExtMap::*type; };          // sized by friend PerfectHashMapBuilder in the reference
struct IndexTag { typedef std::shared_ptr<ExtKMerIndex> ExtWrapper::*type; };     // protected; filled by friend KMerIndexBuilder in the reference
template struct Rob<KmersTag, &ExtIterating::kmers_>;
template struct Rob<DataTag, &ExtMap::data_>;
template struct Rob<IndexTag, &ExtWrapper::index_ptr_>;

inline void check(sb200_ctx *ctx, int rc) {
    if (rc != 0) throw std::runtime_error(std::string("sb200: ") + sb200_last_error(ctx));
}

}  // namespace detail

// 2-bit packed forward reads of a stream list in the layout of sb200_reads_upload (= the payload of the reference's .seq records)
struct PackedReads {
    std::vector<uint64_t> words, word_off;
    std::vector<uint32_t> len;
    PackedReads() : word_off(1, 0) {}
    void push_back(const Sequence &s) {
        const size_t n = s.size(), nw = (n + 31) / 32, base = words.size();
        words.resize(base + nw, 0);
        for (size_t i = 0; i < n; ++i) words[base + i / 32] |= (uint64_t) (unsigned char) s[i] << (2 * (i % 32));
        word_off.push_back(words.size());
        len.push_back((uint32_t) n);
    }
    // every edge of a graph in iteration order (conjugate edges included, as DeBruijnGraphKMerSplitter::Split walks them)
    template<class Graph>
    static PackedReads FromGraphEdges(const Graph &g) {
        PackedReads p;
        for (auto it = g.ConstEdgeBegin(); !it.IsEnd(); ++it) p.push_back(g.EdgeNucls(*it));
        return p;
    }
    template<class Graph, class Edges>
    static PackedReads FromEdges(const Graph &g, const Edges &edges) {   // DeBruijnEdgeKMerSplitter: an explicit edge list
        PackedReads p;
        for (auto e : edges) p.push_back(g.EdgeNucls(e));
        return p;
    }
    template<class Streams>
    static PackedReads FromStreams(Streams &streams) {
        PackedReads p;
        for (auto &stream : streams) {
            stream.reset();
            io::SingleReadSeq r;
            while (!stream.eof()) {
                stream >> r;
                p.push_back(r.sequence());
            }
        }
        return p;
    }
};

// writes the buckets of a device-resident k-mer set into a real KMerDiskStorage (the files the reference's own counter leaves)
inline kmers::KMerDiskStorage<RtSeq> ToDiskStorage(fs::TmpDir work_dir, const sb200_kmers *s) {
    const unsigned B = sb200_kmers_num_buckets(s), W = sb200_kmers_words(s);
    kmers::KMerDiskStorage<RtSeq> storage(work_dir, sb200_kmers_k(s), kmer::KMerSegmentPolicy<RtSeq>(B));
    std::vector<uint64_t> starts(B + 1), buf;
    sb200_kmers_bucket_starts(s, starts.data());
    for (unsigned b = 0; b < B; ++b) {
        const uint64_t n = starts[b + 1] - starts[b];
        buf.resize(n * W + 1);
        if (n && sb200_kmers_download(s, starts[b], n, buf.data())) throw std::runtime_error("sb200: bucket download failed");
        auto file = storage.create(b);
        std::ofstream os(*file, std::ios::binary);
        os.write(reinterpret_cast<const char *>(buf.data()), std::streamsize(n * W * 8));
    }
    return storage;
}

class GpuKMerCounter : public kmers::KMerCounter<RtSeq> {
  public:
    // canonical_only = StoringTypeFilter<InvertableStoring> (ph_map/storing_traits.hpp:88-101); add_rc = the reference's streams were RC-wrapped
    GpuKMerCounter(fs::TmpDir work_dir, sb200_ctx *ctx, unsigned K, const PackedReads &reads, bool canonical_only = true, bool add_rc = true)
            : kmers::KMerCounter<RtSeq>(K), work_dir_(work_dir), ctx_(ctx), reads_(&reads), canonical_only_(canonical_only), add_rc_(add_rc) {}
    ~GpuKMerCounter() override { sb200_kmers_free(last_); }

    size_t kmer_size() const override { return RtSeq::GetDataSize(this->k()) * sizeof(RtSeq::DataType); }

    kmers::KMerDiskStorage<RtSeq> Count(unsigned num_buckets, unsigned /*num_threads*/) override {
        sb200_reads *rd = nullptr;
        detail::check(ctx_, sb200_reads_upload(ctx_, reads_->words.data(), reads_->word_off.data(), reads_->len.data(), reads_->len.size(), &rd));
        sb200_kmers_free(last_);
        last_ = nullptr;
        const int rc = sb200_count(ctx_, rd, this->k(), canonical_only_ ? 1 : 0, add_rc_ ? 1 : 0, num_buckets, &last_);
        sb200_reads_free(rd);
        if (rc != 0) FATAL_ERROR(sb200_last_error(ctx_));   // "No kmers were extracted ..." as kmer_index_builder.hpp:261-264
        return ToDiskStorage(work_dir_, last_);
    }
    kmers::KMerDiskStorage<RtSeq> CountAll(unsigned num_buckets, unsigned num_threads, bool merge = true) override {
        auto storage = Count(num_buckets, num_threads);
        if (merge) storage.merge();
        return storage;
    }
    const sb200_kmers *device_set() const { return last_; }   // the same set, still resident in HBM (multiplicities: sb200_kmers_counts_download)

  private:
    fs::TmpDir work_dir_;
    sb200_ctx *ctx_;
    const PackedReads *reads_;
    bool canonical_only_, add_rc_;
    sb200_kmers *last_ = nullptr;
};

// everything sb200 holds for one extension index: kept so that later stages (tip clipper, unitigs) can stay on the GPU
struct GpuIndexHandles {
    sb200_kmers *kpomers = nullptr, *kmers = nullptr;
    sb200_mphf *mphf = nullptr;
    sb200_ext *ext = nullptr;
    ~GpuIndexHandles() {
        sb200_ext_free(ext); sb200_mphf_free(mphf); sb200_kmers_free(kmers); sb200_kmers_free(kpomers);
    }
};

struct GpuExtensionIndexBuilder {
    sb200_ctx *ctx;
    explicit GpuExtensionIndexBuilder(sb200_ctx *c) : ctx(c) {}

    // kmer_extension_index_builder.hpp:62-80; `streams` = forward streams, nthreads = streams.size(), 10 * nthreads buckets (:74)
    template<class Streams>
    kmers::KMerDiskStorage<RtSeq> BuildExtensionIndexFromStream(fs::TmpDir work_dir, detail::ExtIndex &index, Streams &streams,
                                                                GpuIndexHandles *keep = nullptr, unsigned num_buckets = 0) const {
        PackedReads reads = PackedReads::FromStreams(streams);
        return BuildExtensionIndexFromReads(work_dir, index, reads, num_buckets ? num_buckets : 10 * (unsigned) streams.size(), keep);
    }

    kmers::KMerDiskStorage<RtSeq> BuildExtensionIndexFromReads(fs::TmpDir work_dir, detail::ExtIndex &index, const PackedReads &reads,
                                                               unsigned num_buckets, GpuIndexHandles *keep = nullptr) const {
        GpuIndexHandles local, *h = keep ? keep : &local;
        const unsigned k = index.k();
        sb200_reads *rd = nullptr;
        detail::check(ctx, sb200_reads_upload(ctx, reads.words.data(), reads.word_off.data(), reads.len.data(), reads.len.size(), &rd));
        int rc = sb200_count(ctx, rd, k + 1, 1, 1, num_buckets, &h->kpomers);
        sb200_reads_free(rd);
        if (rc != 0) FATAL_ERROR(sb200_last_error(ctx));
        FillFromKPOMers(work_dir, index, h);
        return ToDiskStorage(work_dir, h->kpomers);
    }

    // kmer_extension_index_builder.hpp:82-106 with the (k+1)-mers already on the device
    void FillFromKPOMers(fs::TmpDir work_dir, detail::ExtIndex &index, GpuIndexHandles *h) const {
        const unsigned k = index.k();
        VERIFY(sb200_kmers_k(h->kpomers) == k + 1);
        detail::check(ctx, sb200_derive_kmers(ctx, h->kpomers, sb200_kmers_num_buckets(h->kpomers), &h->kmers));
        detail::check(ctx, sb200_mphf_build(ctx, h->kmers, &h->mphf));
        detail::check(ctx, sb200_ext_build(ctx, h->kpomers, h->kmers, h->mphf, &h->ext));
        const uint64_t n = sb200_kmers_size(h->kmers);
        // KMerIndex through the reference's own reader (KMerIndex::deserialize, i.e. boomphf::mphf::load per bucket + segment starts) ...
        std::vector<uint8_t> bytes;
        uint64_t nbytes = 0;
        detail::check(ctx, sb200_mphf_serialize(h->mphf, nullptr, &nbytes));
        bytes.resize(nbytes);
        detail::check(ctx, sb200_mphf_serialize(h->mphf, bytes.data(), &nbytes));
        std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
        ss.write(reinterpret_cast<const char *>(bytes.data()), std::streamsize(nbytes));
        auto &index_ptr = static_cast<detail::ExtWrapper &>(index).*detail::Stolen<detail::IndexTag>::ptr;
        index_ptr->deserialize(ss);
        VERIFY(index_ptr->size() == n);
        // ... and data_ = the GPU's mask array, one InOutMask byte per k-mer in MPHF-index order
        static_assert(sizeof(utils::InOutMask) == 1, "InOutMask is one byte");
        auto &data = static_cast<detail::ExtMap &>(index).*detail::Stolen<detail::DataTag>::ptr;
        data.resize(n);
        if (n) detail::check(ctx, sb200_ext_masks_download(h->ext, reinterpret_cast<uint8_t *>(data.data())));
        // final_kmers: the file KeyIteratingMap iterates (kmer_begin), in the reference written by KMerDiskStorage::merge()
        fs::TmpFile fk = work_dir->tmp_file("final_kmers");
        {
            const unsigned W = sb200_kmers_words(h->kmers);
            std::vector<uint64_t> buf(n * W + 1);
            if (n) detail::check(ctx, sb200_kmers_download(h->kmers, 0, n, buf.data()));
            std::ofstream os(*fk, std::ios::binary);
            os.write(reinterpret_cast<const char *>(buf.data()), std::streamsize(n * W * 8));
        }
        static_cast<detail::ExtIterating &>(index).*detail::Stolen<detail::KmersTag>::ptr = fk;
    }
};

// UnbranchingPathExtractor(index, k).ExtractUnbranchingPathsAndLoops(...) on the GPU, as the reference's std::vector<Sequence>
inline std::vector<Sequence> GpuUnbranchingPaths(sb200_ctx *ctx, const GpuIndexHandles &h, bool with_loops = true) {
    sb200_unitigs *u = nullptr;
    detail::check(ctx, sb200_unitigs_extract(ctx, h.kmers, h.mphf, h.ext, with_loops ? 1 : 0, &u));
    const uint64_t n = sb200_unitigs_count(u), nw = sb200_unitigs_total_words(u);
    std::vector<uint64_t> words(nw + 1), off(n + 1);
    std::vector<uint32_t> len(n + 1);
    detail::check(ctx, sb200_unitigs_download(u, words.data(), off.data(), len.data()));
    sb200_unitigs_free(u);
    std::vector<Sequence> out;
    out.reserve(n);
    std::string s;
    for (uint64_t i = 0; i < n; ++i) {
        s.resize(len[i]);
        for (uint32_t j = 0; j < len[i]; ++j) s[j] = "ACGT"[(words[off[i] + j / 32] >> (2 * (j % 32))) & 3];
        out.emplace_back(s);
    }
    return out;
}

}  // namespace sb200_spades

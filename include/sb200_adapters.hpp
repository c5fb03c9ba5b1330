// sb200_adapters.hpp — C++ host side above the C ABI (include/sb200.h): the reference's own interface names for the hot
// path, so that spades-gbuilder / spades-kmercount / spades-core's Construction stage read the same with the GPU
// library underneath (INTEGRATION.md shows the reference-side edits).  Header-only, C++14 like the reference, links
// against libspades_b200.so only — no CUDA or torch types appear here.
//
// Reference interface (paths relative to /root/reference/assembler/src/common)             -> class here
//   io::ReadStreamList<io::SingleReadSeq>            io/reads/read_stream_vector.hpp          ReadStreamList
//   io::BinaryFileStream (.seq/.off files)           io/reads/binary_streams.hpp:19-133       ReadStreamList::FromBinaryFiles
//   io::BinaryWriter::ToBinary                       io/reads/binary_converter.cpp:50-113     WriteBinaryReads
//   io::LongestValid                                 io/reads/longest_valid_wrapper.hpp:15-52 LongestValid
//   Sequence                                         sequence/sequence.hpp                    Sequence
//   kmers::KMerSplitter / DeBruijnReadKMerSplitter   utils/kmer_mph/kmer_splitter.hpp:23-52, kmer_splitters.hpp:96-133
//   kmers::DeBruijnKMerKMerSplitter                  utils/kmer_mph/kmer_splitters.hpp:135-204
//   kmers::KMerCounter / KMerDiskCounter             utils/kmer_mph/kmer_index_builder.hpp:194-365
//   kmers::KMerDiskStorage                           utils/kmer_mph/kmer_index_builder.hpp:48-191
//   kmers::KMerIndex / KMerIndexBuilder              utils/kmer_mph/kmer_index.hpp:25-147, kmer_index_builder.hpp:368-453
//   utils::DeBruijnExtensionIndex(+Builder)          utils/extension_index/kmer_extension_index.hpp:242-339, ..._builder.hpp:19-110
//   debruijn_graph::EarlyTipClipperProcessor         assembly_graph/construction/early_simplification.hpp:37-160
//   debruijn_graph::UnbranchingPathExtractor         assembly_graph/construction/debruijn_graph_constructor.hpp:182-388
//   utils::CoverageHashMapBuilder                    utils/ph_map/coverage_hash_map_builder.hpp:15-54
//   debruijn_graph::FastGraphFromSequencesConstructor assembly_graph/construction/debruijn_graph_constructor.hpp:392-517  CondensedGraph
//   gfa::GFAWriter                                   io/graph/gfa_writer.cpp:18-52                                          CondensedGraph::WriteGFA
//   io::FastgWriter                                  io/graph/fastg_writer.cpp:35-46                                        CondensedGraph::WriteFASTG
// Error behaviour: the reference aborts through FATAL_ERROR / VERIFY (utils/logger/logger.hpp:177-190); here every
// nonzero ABI return becomes sb200::Error carrying sb200_last_error(), which a SPAdes build maps back to FATAL_ERROR.
// There is no CPU fallback: Context's constructor throws when no sm_100 device is present.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <ostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sb200.h"

namespace sb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// ---- context --------------------------------------------------------------------------------------------------------
class Context {
public:
    explicit Context(int device = 0) {
        int rc = sb200_create(device, &h_);
        if (rc) throw Error(rc, std::string("sb200_create: ") + sb200_last_error(nullptr));
    }
    ~Context() { if (h_) sb200_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    sb200_ctx *get() const { return h_; }
    void check(int rc) const { if (rc) throw Error(rc, sb200_last_error(h_)); }
    uint64_t kernel_launches(bool reset = false) const { return sb200_kernel_launches(h_, reset ? 1 : 0); }
private:
    sb200_ctx *h_ = nullptr;
};

// ---- Sequence: 2 bits per base, base i at bits 2(i%32) of word i/32 (sequence/sequence.hpp:24-35,186-197) -------------
inline char nucl(unsigned c) { return "ACGT"[c & 3u]; }
inline int dignucl(char c) {   // sequence/nucl.hpp:36-48; -1 for anything that is not ACGT
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1; }
}

class Sequence {
public:
    Sequence() {}
    explicit Sequence(const std::string &s) : size_(s.size()), words_((s.size() + 31) / 32, 0) {
        for (size_t i = 0; i < s.size(); ++i) {
            int c = dignucl(s[i]);
            if (c < 0) throw Error(1, "Sequence: invalid nucleotide");
            words_[i >> 5] |= (uint64_t) c << (2 * (i & 31));
        }
    }
    Sequence(const uint64_t *words, size_t size) : size_(size), words_(words, words + (size + 31) / 32) {}
    size_t size() const { return size_; }
    unsigned operator[](size_t i) const { return (unsigned) (words_[i >> 5] >> (2 * (i & 31))) & 3u; }
    const uint64_t *data() const { return words_.data(); }
    size_t data_size() const { return words_.size(); }   // Sequence::DataSize, sequence.hpp:67-69
    std::string str() const {
        std::string s(size_, 'A');
        for (size_t i = 0; i < size_; ++i) s[i] = nucl((*this)[i]);
        return s;
    }
    Sequence operator!() const {   // reverse complement
        Sequence r;
        r.size_ = size_;
        r.words_.assign(words_.size(), 0);
        for (size_t i = 0; i < size_; ++i) r.words_[i >> 5] |= (uint64_t) (3u - (*this)[size_ - 1 - i]) << (2 * (i & 31));
        return r;
    }
    bool operator==(const Sequence &o) const { return size_ == o.size_ && words_ == o.words_; }
    bool operator<(const Sequence &o) const {   // base-wise, shorter prefix first (sequence.hpp:222-230)
        size_t n = std::min(size_, o.size_);
        for (size_t i = 0; i < n; ++i) if ((*this)[i] != o[i]) return (*this)[i] < o[i];
        return size_ < o.size_;
    }
private:
    size_t size_ = 0;
    std::vector<uint64_t> words_;
};

// io::LongestValid (longest_valid_wrapper.hpp:15-52): the first longest run of ACGT of a raw read; [begin, end)
inline std::pair<size_t, size_t> LongestValid(const std::string &s) {
    size_t best_b = 0, best_e = 0, b = 0;
    for (size_t i = 0; i <= s.size(); ++i) {
        if (i == s.size() || dignucl(s[i]) < 0) {
            if (i - b > best_e - best_b) { best_b = b; best_e = i; }
            b = i + 1;
        }
    }
    return {best_b, best_e};
}

// ---- reads ----------------------------------------------------------------------------------------------------------
// The packed read set on the host: read r = words[word_off[r] .. word_off[r+1]), len[r] bases.  This is the layout of
// the records of the reference's .seq files with the per-record headers and offsets stripped, and the layout
// sb200_reads_upload takes.  The reference splits the set into `nthreads` streams only to parallelise; the result of
// the path does not depend on that split, so one list stands for all of them.
class ReadStreamList {
public:
    std::vector<uint64_t> words;
    std::vector<uint64_t> word_off{0};
    std::vector<uint32_t> len;
    std::vector<uint16_t> left_offset, right_offset;   // bases LongestValid trimmed on either side (single_read.hpp:176-185)
    size_t max_len = 0;
    uint64_t total_len = 0;

    size_t size() const { return len.size(); }
    void push_back(const Sequence &s, uint16_t left = 0, uint16_t right = 0) {
        words.insert(words.end(), s.data(), s.data() + s.data_size());
        word_off.push_back(words.size());
        len.push_back((uint32_t) s.size());
        left_offset.push_back(left);
        right_offset.push_back(right);
        max_len = std::max(max_len, s.size());
        total_len += s.size();
    }
    void push_back_raw(const std::string &raw) {   // EasyStream(handle_Ns = true) + binary conversion
        auto r = LongestValid(raw);
        push_back(Sequence(raw.substr(r.first, r.second - r.first)), (uint16_t) r.first, (uint16_t) (raw.size() - r.second));
    }
    Sequence operator[](size_t r) const { return Sequence(words.data() + word_off[r], len[r]); }

    // io::BinaryFileStream(prefix, portion_count, portion_num) for every portion at once: all records of prefix.seq.
    // Record = size_t length, ceil(length/32) uint64 words, uint16 left offset, uint16 right offset
    // (single_read.hpp:279-299, sequence.hpp:399-441); header = ReadStreamStat, 3 x 8 bytes (read_stream.hpp:19-36).
    // Paired files hold first/second alternately (paired_read.hpp:91-104), the orientation already applied on write.
    static ReadStreamList FromBinaryFiles(const std::string &prefix) {
        std::ifstream in(prefix + ".seq", std::ios::binary);
        if (!in) throw Error(1, "cannot open " + prefix + ".seq");
        uint64_t stat[3];
        in.read((char *) stat, sizeof stat);
        if (!in) throw Error(1, prefix + ".seq: truncated header");
        ReadStreamList l;
        l.len.reserve(stat[0]);
        l.word_off.reserve(stat[0] + 1);
        std::vector<uint64_t> buf;
        while (true) {
            uint64_t n;
            in.read((char *) &n, 8);
            if (!in) break;
            size_t nw = (n + 31) / 32;
            buf.resize(nw);
            uint16_t off[2];
            in.read((char *) buf.data(), (std::streamsize) (nw * 8));
            in.read((char *) off, 4);
            if (!in) throw Error(1, prefix + ".seq: truncated record");
            l.words.insert(l.words.end(), buf.begin(), buf.end());
            l.word_off.push_back(l.words.size());
            l.len.push_back((uint32_t) n);
            l.left_offset.push_back(off[0]);
            l.right_offset.push_back(off[1]);
            l.max_len = std::max<size_t>(l.max_len, n);
            l.total_len += n;
        }
        return l;
    }
};

// io::BinaryWriter::ToBinary for single reads: prefix.seq + prefix.off (file offset of every CHUNK-th record,
// binary_converter.hpp:35 CHUNK = 100, binary_converter.cpp:73-77).  rc = write every read reverse-complemented
// (the second mate of an FR pair, orientation.hpp:15-26).
inline void WriteBinaryReads(const ReadStreamList &l, const std::string &prefix, bool rc = false) {
    std::ofstream seq(prefix + ".seq", std::ios::binary), off(prefix + ".off", std::ios::binary);
    uint64_t stat[3] = {l.size(), l.max_len, l.total_len};
    seq.write((const char *) stat, sizeof stat);
    size_t rest = 1;
    for (size_t r = 0; r < l.size(); ++r) {
        if (!--rest) {
            uint64_t o = (uint64_t) seq.tellp();
            off.write((const char *) &o, 8);
            rest = 100;
        }
        Sequence s = rc ? !l[r] : l[r];
        uint64_t n = s.size();
        uint16_t offs[2] = {rc ? l.right_offset[r] : l.left_offset[r], rc ? l.left_offset[r] : l.right_offset[r]};
        seq.write((const char *) &n, 8);
        seq.write((const char *) s.data(), (std::streamsize) (s.data_size() * 8));
        seq.write((const char *) offs, 4);
    }
    if (!seq || !off) throw Error(1, "cannot write " + prefix + ".seq/.off");
}

// The read set resident in HBM (uploaded once, reused by every K of a multi-K run).
class DeviceReads {
public:
    DeviceReads(const Context &ctx, const ReadStreamList &l) : ctx_(ctx) {
        static const uint64_t zero64 = 0;
        static const uint32_t zero32 = 0;
        ctx.check(sb200_reads_upload(ctx.get(), l.words.empty() ? &zero64 : l.words.data(), l.word_off.data(),
                                     l.len.empty() ? &zero32 : l.len.data(), l.size(), &h_));
    }
    ~DeviceReads() { if (h_) sb200_reads_free(h_); }
    DeviceReads(const DeviceReads &) = delete;
    DeviceReads &operator=(const DeviceReads &) = delete;
    const sb200_reads *get() const { return h_; }
    const Context &ctx() const { return ctx_; }
private:
    const Context &ctx_;
    sb200_reads *h_ = nullptr;
};

// ---- KMerDiskStorage ----------------------------------------------------------------------------------------------------
// Movable value type like the reference's; the "files" are ranges of one device array in file order.
class KMerDiskStorage {
public:
    KMerDiskStorage() {}
    KMerDiskStorage(const Context &ctx, sb200_kmers *h) : ctx_(&ctx), h_(h, sb200_kmers_free) {
        starts_.resize(num_buckets() + 1);
        ctx.check(sb200_kmers_bucket_starts(h, starts_.data()));
    }
    unsigned k() const { return sb200_kmers_k(h_.get()); }
    unsigned kmer_words() const { return sb200_kmers_words(h_.get()); }
    size_t num_buckets() const { return sb200_kmers_num_buckets(h_.get()); }
    size_t total_kmers() const { return sb200_kmers_size(h_.get()); }
    size_t bucket_size(size_t i) const { return (size_t) (starts_[i + 1] - starts_[i]); }
    uint64_t kmer_instances() const { return sb200_kmers_instances(h_.get()); }
    // records of bucket i = the contents of the reference's kmers<i> file (bucket_begin(i) .. bucket_end(i))
    std::vector<uint64_t> bucket(size_t i) const { return download(starts_[i], bucket_size(i)); }
    // bucket_begin(i) / bucket_end(i): input iterators over the records of kmers<i>, each yielding (const uint64_t *, bytes) like
    // KMerDiskStorage::kmer_iterator (kmer_index_builder.hpp:57-92); the bucket is downloaded once and shared by the iterator copies
    typedef std::pair<const uint64_t *, size_t> KMerRawData;
    class kmer_iterator {
    public:
        typedef std::input_iterator_tag iterator_category;
        typedef KMerRawData value_type;
        typedef ptrdiff_t difference_type;
        typedef const KMerRawData *pointer;
        typedef KMerRawData reference;
        kmer_iterator() {}
        kmer_iterator(std::shared_ptr<std::vector<uint64_t>> data, unsigned words) : data_(std::move(data)), words_(words) {}
        KMerRawData operator*() const { return KMerRawData(data_->data() + pos_ * words_, (size_t) words_ * 8); }
        kmer_iterator &operator++() { ++pos_; return *this; }
        void operator+=(size_t n) { pos_ += n; }
        bool operator==(const kmer_iterator &o) const { return at_end() == o.at_end() && (at_end() || (data_ == o.data_ && pos_ == o.pos_)); }
        bool operator!=(const kmer_iterator &o) const { return !(*this == o); }
    private:
        bool at_end() const { return !data_ || pos_ * words_ >= data_->size(); }
        std::shared_ptr<std::vector<uint64_t>> data_;
        unsigned words_ = 1;
        size_t pos_ = 0;
    };
    kmer_iterator bucket_begin(size_t i) const { return kmer_iterator(std::make_shared<std::vector<uint64_t>>(bucket(i)), kmer_words()); }
    kmer_iterator bucket_end(size_t) const { return kmer_iterator(); }
    // merge() + final_kmers(): all buckets concatenated.  The reference merges the bucket FILES into <work_dir>/final_kmers and hands
    // the path around (kmer_index_builder.hpp:168-181); here the buckets already are one device array: merge() only records that the
    // caller asked for it, final_kmers(path) writes the file where a consumer wants one.
    void merge() { merged_ = true; }
    bool merged() const { return merged_; }
    std::vector<uint64_t> final_kmers() const { return download(0, total_kmers()); }
    std::string final_kmers(const std::string &path) const { write_final_kmers(path); return path; }
    // multiplicities in file order (empty for derived sets): what CoverageHashMapBuilder::FillCoverageFromStream counts
    std::vector<uint32_t> counts() const {
        std::vector<uint32_t> c(total_kmers());
        if (!c.empty()) ctx_->check(sb200_kmers_counts_download(h_.get(), 0, c.size(), c.data()));
        return c;
    }
    void write_final_kmers(const std::string &path) const {
        auto v = final_kmers();
        std::ofstream os(path, std::ios::binary);
        os.write((const char *) v.data(), (std::streamsize) (v.size() * 8));
    }
    const sb200_kmers *get() const { return h_.get(); }
    const Context &ctx() const { return *ctx_; }
    explicit operator bool() const { return (bool) h_; }
private:
    std::vector<uint64_t> download(uint64_t first, uint64_t n) const {
        std::vector<uint64_t> v(n * kmer_words());
        if (n) ctx_->check(sb200_kmers_download(h_.get(), first, n, v.data()));
        return v;
    }
    const Context *ctx_ = nullptr;
    std::shared_ptr<sb200_kmers> h_;
    std::vector<uint64_t> starts_;
    bool merged_ = false;
};

// ---- splitters: on the GPU a splitter is only the description of what to count (the records never hit a disk) ---------
struct KMerSplitter {
    virtual ~KMerSplitter() {}
    virtual unsigned K() const = 0;
    virtual KMerDiskStorage CountOnDevice(unsigned num_buckets) const = 0;
};

// DeBruijnReadKMerSplitter<Read, KmerFilter> over RC-wrapped streams.  canonical_only = StoringTypeFilter<InvertableStoring>
class DeBruijnReadKMerSplitter : public KMerSplitter {
public:
    DeBruijnReadKMerSplitter(const DeviceReads &reads, unsigned K, bool canonical_only = true, bool add_rc = true)
        : reads_(reads), K_(K), canonical_only_(canonical_only), add_rc_(add_rc) {}
    unsigned K() const override { return K_; }
    KMerDiskStorage CountOnDevice(unsigned num_buckets) const override {
        sb200_kmers *h = nullptr;
        reads_.ctx().check(sb200_count(reads_.ctx().get(), reads_.get(), K_, canonical_only_, add_rc_, num_buckets, &h));
        return KMerDiskStorage(reads_.ctx(), h);
    }
private:
    const DeviceReads &reads_;
    unsigned K_;
    bool canonical_only_, add_rc_;
};

// DeBruijnKMerKMerSplitter(K_target = K_source - 1, add_rc = true)
class DeBruijnKMerKMerSplitter : public KMerSplitter {
public:
    DeBruijnKMerKMerSplitter(const KMerDiskStorage &kpomers, unsigned K) : src_(kpomers), K_(K) {
        if (K + 1 != kpomers.k()) throw Error(1, "DeBruijnKMerKMerSplitter: K must be K_source - 1");
    }
    unsigned K() const override { return K_; }
    KMerDiskStorage CountOnDevice(unsigned num_buckets) const override {
        sb200_kmers *h = nullptr;
        src_.ctx().check(sb200_derive_kmers(src_.ctx().get(), src_.get(), num_buckets, &h));
        return KMerDiskStorage(src_.ctx(), h);
    }
private:
    const KMerDiskStorage &src_;
    unsigned K_;
};

struct KMerCounter {
    virtual ~KMerCounter() {}
    virtual KMerDiskStorage Count(unsigned num_buckets, unsigned num_threads) = 0;
    virtual KMerDiskStorage CountAll(unsigned num_buckets, unsigned num_threads, bool merge = true) = 0;
};

class KMerDiskCounter : public KMerCounter {
public:
    explicit KMerDiskCounter(const KMerSplitter &splitter) : splitter_(splitter) {}
    // num_threads only sized the reference's OpenMP team; the GPU path ignores it
    KMerDiskStorage Count(unsigned num_buckets, unsigned /*num_threads*/ = 1) override { return splitter_.CountOnDevice(num_buckets); }
    KMerDiskStorage CountAll(unsigned num_buckets, unsigned num_threads, bool /*merge*/ = true) override { return Count(num_buckets, num_threads); }
private:
    const KMerSplitter &splitter_;
};

// ---- KMerIndex / KMerIndexBuilder ---------------------------------------------------------------------------------------
class KMerIndex {
public:
    KMerIndex() {}
    size_t size() const { return h_ ? sb200_mphf_size(h_.get()) : 0; }
    size_t mem_size() const { return h_ ? sb200_mphf_mem_size(h_.get()) : 0; }
    // seq_idx for a batch of records (W words each); ~0 = fell through every level
    std::vector<uint64_t> seq_idx(const std::vector<uint64_t> &records, unsigned words) const {
        std::vector<uint64_t> idx(records.size() / words);
        if (!idx.empty()) ctx_->check(sb200_mphf_lookup(ctx_->get(), h_.get(), records.data(), idx.size(), idx.data()));
        return idx;
    }
    // seq_idx(const Seq &): one key from host code (kmer_index.hpp:85-90); the lookup runs on the host over a copy of the bit-vectors
    size_t seq_idx(const Sequence &kmer) const {
        uint64_t idx = 0;
        ctx_->check(sb200_mphf_seq_idx(h_.get(), kmer.data(), &idx));
        return (size_t) idx;
    }
    void serialize(std::ostream &os) const {
        uint64_t n = 0;
        ctx_->check(sb200_mphf_serialize(h_.get(), nullptr, &n));
        std::vector<uint8_t> b(n);
        ctx_->check(sb200_mphf_serialize(h_.get(), b.data(), &n));
        os.write((const char *) b.data(), (std::streamsize) n);
    }
    const sb200_mphf *get() const { return h_.get(); }
private:
    friend class KMerIndexBuilder;
    const Context *ctx_ = nullptr;
    std::shared_ptr<sb200_mphf> h_;
};

class KMerIndexBuilder {
public:
    void BuildIndex(KMerIndex &index, const KMerDiskStorage &storage) const {
        sb200_mphf *h = nullptr;
        storage.ctx().check(sb200_mphf_build(storage.ctx().get(), storage.get(), &h));
        index.ctx_ = &storage.ctx();
        index.h_.reset(h, sb200_mphf_free);
    }
    KMerDiskStorage BuildIndex(KMerIndex &index, KMerCounter &counter, unsigned num_buckets, unsigned num_threads = 1) const {
        KMerDiskStorage st = counter.Count(num_buckets, num_threads);
        BuildIndex(index, st);
        return st;
    }
};

// ---- DeBruijnExtensionIndex ----------------------------------------------------------------------------------------------
class DeBruijnExtensionIndex {
public:
    DeBruijnExtensionIndex(const Context &ctx, unsigned k) : ctx_(ctx), k_(k) {}
    unsigned k() const { return k_; }
    size_t size() const { return index_.size(); }
    const KMerIndex &index() const { return index_; }
    const KMerDiskStorage &kmers() const { return kmers_; }
    // PerfectHashMap::data_: one InOutMask byte per k-mer in MPHF-index order
    std::vector<uint8_t> data() const {
        std::vector<uint8_t> m(size());
        if (!m.empty()) ctx_.check(sb200_ext_masks_download(ext_.get(), m.data()));
        return m;
    }
    // MPHF index of every k-mer in file order (ConstructKWH(kmer).idx())
    std::vector<uint32_t> idx() const {
        std::vector<uint32_t> v(size());
        if (!v.empty()) ctx_.check(sb200_ext_idx_download(ext_.get(), v.data()));
        return v;
    }
    const Context &ctx() const { return ctx_; }
    sb200_ext *ext() const { return ext_.get(); }
private:
    friend class DeBruijnExtensionIndexBuilder;
    const Context &ctx_;
    unsigned k_;
    KMerDiskStorage kmers_;
    KMerIndex index_;
    std::shared_ptr<sb200_ext> ext_;
};

class DeBruijnExtensionIndexBuilder {
public:
    // Returns the (k+1)-mer storage like the reference (the caller feeds it to the coverage builder)
    KMerDiskStorage BuildExtensionIndexFromStream(DeBruijnExtensionIndex &index, const DeviceReads &streams, unsigned nthreads) const {
        DeBruijnReadKMerSplitter splitter(streams, index.k() + 1, true, true);
        KMerDiskCounter counter(splitter);
        KMerDiskStorage kpomers = counter.Count(10 * nthreads, nthreads);   // kmer_extension_index_builder.hpp:72-74
        BuildExtensionIndexFromKPOMers(index, kpomers, nthreads);
        return kpomers;
    }
    void BuildExtensionIndexFromKPOMers(DeBruijnExtensionIndex &index, const KMerDiskStorage &kpomers, unsigned nthreads) const {
        if (kpomers.k() != index.k() + 1) throw Error(1, "kpomers.k() must equal index.k() + 1");
        DeBruijnKMerKMerSplitter splitter(kpomers, index.k());
        KMerDiskCounter counter(splitter);
        index.kmers_ = KMerIndexBuilder().BuildIndex(index.index_, counter, 10 * nthreads, nthreads);   // ibid. :88-97
        sb200_ext *e = nullptr;
        index.ctx_.check(sb200_ext_build(index.ctx_.get(), kpomers.get(), index.kmers_.get(), index.index_.get(), &e));
        index.ext_.reset(e, sb200_ext_free);
    }
};

class EarlyTipClipperProcessor {
public:
    EarlyTipClipperProcessor(DeBruijnExtensionIndex &index, size_t length_bound) : index_(index), bound_(length_bound) {}
    size_t ClipTips() {
        uint64_t removed = 0;
        index_.ctx().check(sb200_tipclip(index_.ctx().get(), index_.kmers().get(), index_.index().get(), index_.ext(), bound_, &removed));
        return (size_t) removed;
    }
private:
    DeBruijnExtensionIndex &index_;
    size_t bound_;
};

class UnbranchingPathExtractor {
public:
    UnbranchingPathExtractor(DeBruijnExtensionIndex &index, size_t k) : index_(index), k_(k) {}
    // nchunks only partitioned the reference's OpenMP loop; the output order (paths in file order, then loops) is the same
    std::vector<Sequence> ExtractUnbranchingPaths(unsigned /*nchunks*/ = 1) const { return run(0); }
    std::vector<Sequence> ExtractUnbranchingPathsAndLoops(unsigned /*nchunks*/ = 1) const { return run(1); }
private:
    std::vector<Sequence> run(int with_loops) const {
        sb200_unitigs *u = nullptr;
        const Context &c = index_.ctx();
        c.check(sb200_unitigs_extract(c.get(), index_.kmers().get(), index_.index().get(), index_.ext(), with_loops, &u));
        std::shared_ptr<sb200_unitigs> guard(u, sb200_unitigs_free);
        last_ = guard;   // stays resident for the coverage filler (CoverageHashMap::FillCoverageAndFlanking)
        uint64_t n = sb200_unitigs_count(u);
        std::vector<uint64_t> words(sb200_unitigs_total_words(u) + 1), off(n + 1);
        std::vector<uint32_t> len(n + 1);
        c.check(sb200_unitigs_download(u, words.data(), off.data(), len.data()));
        std::vector<Sequence> out;
        out.reserve(n);
        for (uint64_t i = 0; i < n; ++i) out.emplace_back(words.data() + off[i], len[i]);
        return out;
    }
    DeBruijnExtensionIndex &index_;
    size_t k_;
    mutable std::shared_ptr<sb200_unitigs> last_;
public:
    const sb200_unitigs *device_sequences() const { return last_.get(); }   // the last result, still in HBM
};

// utils::PerfectHashMap<RtSeq, uint32_t> of CoverageHashMapBuilder::BuildIndex (ph_map/coverage_hash_map_builder.hpp:39-54): the
// reference builds a second MPHF over the (k+1)-mers and re-streams the reads; on the GPU the multiplicities are the run lengths of
// the counting sort (self-reverse-complement (k+1)-mers count twice, :31-36), moved into the index order of the same BooPHF.
class CoverageHashMap {
public:
    CoverageHashMap(const Context &ctx, const KMerDiskStorage &kpomers) : ctx_(ctx) {
        sb200_covmap *h = nullptr;
        ctx.check(sb200_coverage_map_build(ctx.get(), kpomers.get(), &h));
        h_.reset(h, sb200_coverage_map_free);
    }
    size_t size() const { return (size_t) sb200_coverage_map_size(h_.get()); }
    std::vector<uint32_t> data() const {   // data_: one count per (k+1)-mer in index order
        std::vector<uint32_t> v(size());
        if (!v.empty()) ctx_.check(sb200_coverage_map_values_download(h_.get(), v.data()));
        return v;
    }
    // FillCoverageAndFlankingFromPHM (assembly_graph/graph_support/coverage_filling.hpp:89-95) over the sequences the extractor left in
    // HBM: kc[i] = raw coverage of edge i (GFA KC:i), flank[2i] / flank[2i + 1] = raw flanking coverage of edge i / of its conjugate
    void FillCoverageAndFlanking(const UnbranchingPathExtractor &ex, std::vector<uint64_t> &kc, std::vector<uint64_t> &flank,
                                 unsigned averaging_range = 50) const {
        const sb200_unitigs *u = ex.device_sequences();
        if (!u) throw Error(1, "run the extractor first");
        const size_t n = (size_t) sb200_unitigs_count(u);
        kc.assign(n, 0);
        flank.assign(2 * n, 0);
        if (n) ctx_.check(sb200_unitigs_coverage(ctx_.get(), h_.get(), u, averaging_range, kc.data(), flank.data()));
    }
private:
    const Context &ctx_;
    std::shared_ptr<sb200_covmap> h_;
};

struct CoverageHashMapBuilder {
    // values in (k+1)-mer FILE order (what coverage.u32 of the reference driver lists)
    std::vector<uint32_t> FillCoverage(const KMerDiskStorage &kpomers) const { return kpomers.counts(); }
    CoverageHashMap BuildIndex(const Context &ctx, const KMerDiskStorage &kpomers) const { return CoverageHashMap(ctx, kpomers); }
};

// ---- the step right after the path: graph from the unitigs, GFA out ------------------------------------------------------------
// FastGraphFromSequencesConstructor::ConstructGraph: edge i gets id min_id + 2i (its conjugate the next id, the same id when
// the sequence is its own reverse complement); every sequence leaves a start and an end LinkRecord keyed by the MPHF index of
// its canonical first / last k-mer (one batched KMerIndex::seq_idx on the GPU instead of one lookup per record); records are
// sorted and every distinct index becomes a vertex pair (v, conjugate v) with ids min_id + 2j, min_id + 2j + 1.  WriteGFA
// emits what gfa::GFAWriter::WriteSegmentsAndLinks emits for that graph (DP:f:0 KC:i:0 as spades-gbuilder without -c; with
// SetCoverage the values FillCoverageAndFlankingFromPHM leaves, as with -c).  Lines come in id order; the reference's order of L lines depends on its adjacency containers, so files are
// compared after sorting the lines.
class CondensedGraph {
public:
    static constexpr uint64_t ID_BIAS = 3;   // omnigraph::GraphCore::ID_BIAS (assembly_graph/core/graph_core.hpp)
    struct Vertex { std::vector<uint64_t> incoming, outgoing; };   // oriented edge ids at the canonical vertex of the pair

    CondensedGraph(const DeBruijnExtensionIndex &index, const std::vector<Sequence> &sequences) : k_(index.k()), edges_(sequences) {
        const unsigned W = (k_ + 31) / 32;
        const size_t n = edges_.size();
        std::vector<uint64_t> recs(2 * n * W, 0);
        std::vector<uint8_t> is_rc(2 * n, 0);
        self_conj_.assign(n, 0);
        for (size_t i = 0; i < n; ++i) {
            const Sequence &s = edges_[i];
            self_conj_[i] = (s == !s) ? 1 : 0;
            for (int end = 0; end < 2; ++end) {   // StartLink / EndLink: canonical form of the first / last k-mer
                std::string km = s.str().substr(end ? s.size() - k_ : 0, k_);
                Sequence fwd(km), rc = !fwd;
                const bool minimal = fwd < rc;
                const Sequence &c = minimal ? fwd : rc;
                is_rc[2 * i + end] = minimal ? 0 : 1;
                memcpy(&recs[(2 * i + end) * W], c.data(), W * 8);
            }
        }
        std::vector<uint64_t> idx = index.index().seq_idx(recs, W);
        struct Rec { uint64_t key; uint64_t edge; };
        std::vector<Rec> records;
        records.reserve(2 * n);
        for (size_t i = 0; i < n; ++i) {
            const uint64_t e = ID_BIAS + 2 * i;
            records.push_back(Rec{(idx[2 * i] << 2) | ((uint64_t) is_rc[2 * i] << 1) | 1u, e});
            if (!self_conj_[i]) records.push_back(Rec{(idx[2 * i + 1] << 2) | ((uint64_t) is_rc[2 * i + 1] << 1), e});
        }
        std::sort(records.begin(), records.end(), [](const Rec &a, const Rec &b) { return a.key != b.key ? a.key < b.key : a.edge < b.edge; });
        end_of_.assign(2 * n, End{0, false});
        for (size_t i = 0; i < records.size(); ++i) {
            if (i == 0 || (records[i].key >> 2) != (records[i - 1].key >> 2)) vertices_.emplace_back();
            Vertex &v = vertices_.back();
            const bool rc = records[i].key & 2, start = records[i].key & 1;
            const uint64_t e = records[i].edge, ce = conjugate(e);
            // where the edge (or, for a start record, its conjugate) ENDS: the vertex of the record or its conjugate
            if (start) end_of_[ce - ID_BIAS] = End{vertices_.size() - 1, !rc};
            else end_of_[e - ID_BIAS] = End{vertices_.size() - 1, rc};
            // LinkEdge: the record attaches `e` to v (or to conjugate(v) when rc); the conjugate edge mirrors it on the other vertex
            if (start) { if (!rc) v.outgoing.push_back(e); else v.incoming.push_back(ce); }
            else       { if (!rc) v.incoming.push_back(e); else v.outgoing.push_back(ce); }
        }
    }
    size_t k() const { return k_; }
    size_t edge_count() const { return edges_.size(); }
    size_t vertex_count() const { return vertices_.size(); }
    const std::vector<Vertex> &vertices() const { return vertices_; }
    uint64_t conjugate(uint64_t e) const { return self_conj_[(e - ID_BIAS) >> 1] ? e : (((e - ID_BIAS) ^ 1u) + ID_BIAS); }

    // per-edge raw coverage (CoverageHashMap::FillCoverageAndFlanking): spades-gbuilder -c
    void SetCoverage(const std::vector<uint64_t> &kc) { kc_ = kc; }
    double coverage(size_t i) const {   // CoverageIndex::coverage: raw coverage / length in (k+1)-mers (assembly_graph/core/coverage.hpp:58-60)
        return kc_.empty() ? 0.0 : (double) kc_[i] / (double) (edges_[i].size() - k_);
    }

    void WriteGFA(std::ostream &os) const {
        for (size_t i = 0; i < edges_.size(); ++i)   // gfa_writer.cpp:18-26: DP:f:<float(coverage)> KC:i:<raw coverage>
            os << "S\t" << (ID_BIAS + 2 * i) << '\t' << edges_[i].str() << "\tDP:f:" << float(coverage(i)) << "\tKC:i:" << (kc_.empty() ? 0 : kc_[i]) << '\n';
        for (const Vertex &v : vertices_)
            for (uint64_t inc : v.incoming)
                for (uint64_t out : v.outgoing)
                    os << "L\t" << name(inc) << '\t' << orient(inc) << '\t' << name(out) << '\t' << orient(out) << '\t' << k_ << "M\n";
    }
    // io::FastgWriter::WriteSegmentsAndLinks (io/graph/fastg_writer.cpp:35-46) with spades-gbuilder's default naming
    // (BasicNamingF, io/utils/edge_namer.hpp:29-35: EDGE_<id>_length_<bases>_cov_<coverage>, "'" for the conjugate): one FASTA record per
    // ORIENTED edge, the header lists the edges that leave its end vertex (a std::set of names: string order), sequence wrapped at 60.
    // Records come in id order; the reference's order follows its edge iterator, so files are compared record by record after sorting.
    void WriteFASTG(std::ostream &os) const {
        for (size_t i = 0; i < edges_.size(); ++i) {
            const uint64_t e = ID_BIAS + 2 * i;
            for (int o = 0; o < (self_conj_[i] ? 1 : 2); ++o) {
                const uint64_t x = e + o;
                const End &end = end_of_[x - ID_BIAS];
                const Vertex &v = vertices_[end.vertex];
                std::vector<std::string> next;
                if (!end.conj) for (uint64_t y : v.outgoing) next.push_back(fastg_name(y));
                else for (uint64_t y : v.incoming) next.push_back(fastg_name(conjugate(y)));
                std::sort(next.begin(), next.end());
                next.erase(std::unique(next.begin(), next.end()), next.end());
                os << '>' << fastg_name(x);
                for (size_t j = 0; j < next.size(); ++j) os << (j ? ',' : ':') << next[j];
                os << ";\n";
                const std::string seq = o ? (!edges_[i]).str() : edges_[i].str();
                for (size_t p = 0; p < seq.size(); p += 60) os << seq.substr(p, 60) << '\n';
            }
        }
    }
private:
    struct End { size_t vertex; bool conj; };
    std::string fastg_name(uint64_t e) const {
        const uint64_t c = name(e);
        return "EDGE_" + std::to_string(c) + "_length_" + std::to_string(edges_[(c - ID_BIAS) >> 1].size()) + "_cov_" + std::to_string(coverage((c - ID_BIAS) >> 1)) +
               (e == c ? "" : "'");
    }
    uint64_t name(uint64_t e) const { return std::min(e, conjugate(e)); }          // io::CanonicalEdgeHelper (io/utils/edge_namer.hpp:71-86)
    char orient(uint64_t e) const { return e <= conjugate(e) ? '+' : '-'; }
    size_t k_;
    std::vector<Sequence> edges_;
    std::vector<uint8_t> self_conj_;
    std::vector<Vertex> vertices_;
    std::vector<End> end_of_;   // by oriented edge (id - ID_BIAS)
    std::vector<uint64_t> kc_;  // raw coverage per edge (empty: no coverage, DP:f:0 KC:i:0 as spades-gbuilder without -c)
};

}  // namespace sb200

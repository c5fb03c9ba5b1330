// oracle/ref_dropin.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Proves the drop-in at the reference's OWN types (include/sb200_spades.hpp): built against the unmodified SPAdes 3.15.4 headers
// where they lie under /root/reference (oracle/Makefile `make dropin`, output oracle/_ref/ref_dropin) and linked with
// libspades_b200.so, it runs the reference's CPU path and the GPU path side by side on the same reads and checks
//   A  kmers::KMerCounter<RtSeq>::Count — the GPU subclass against KMerDiskCounter: every bucket file byte for byte
//   B  a real utils::DeBruijnExtensionIndex<> filled from the GPU (KMerIndex::deserialize of the GPU's bytes, data_, final_kmers):
//      the REFERENCE's own lookup (ConstructKWH(kmer).idx(), BooPHF.h:465-487) on the GPU-built bit-vectors gives the reference's
//      indices for every k-mer, the masks are equal, and KMerIndex::serialize of both indices is the same byte stream
//   C  the REFERENCE's UnbranchingPathExtractor on the GPU-filled index == the reference's on its own index == the GPU's unitigs
//      (sequence by sequence, order included)
//   D  the REFERENCE's FastGraphFromSequencesConstructor + gfa::GFAWriter over the GPU-filled index and the GPU's unitigs writes
//      the same GFA lines as over the reference's own
//   E  the edge-index counter: KMerDiskCounter over the reference's DeBruijnGraphKMerSplitter on the condensed graph, with
//      the keep-all and the keep-minimal filter, against the GPU counter fed with the same edges: bucket files byte for byte
// Exit code 0 iff everything matches; one "name OK|FAIL" line per check on stdout.
#include "utils/extension_index/kmer_extension_index_builder.hpp"
#include "assembly_graph/construction/debruijn_graph_constructor.hpp"
#include "io/reads/vector_reader.hpp"
#include "io/reads/read_stream_vector.hpp"
#include "io/reads/rc_reader_wrapper.hpp"
#include "io/reads/longest_valid_wrapper.hpp"
#include "assembly_graph/core/graph.hpp"
#include "io/graph/gfa_writer.hpp"
#include "assembly_graph/index/edge_index_builders.hpp"
#include "utils/logger/log_writers.hpp"
#include "utils/filesystem/temporary.hpp"

#include "sb200_spades.hpp"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>

namespace {

std::vector<io::SingleRead> load_reads(const std::string &path) {
    std::vector<io::SingleRead> reads;
    std::ifstream in(path);
    std::string s;
    while (std::getline(in, s)) {
        if (s.empty() || s[0] == '>' || s[0] == '@') continue;
        io::SingleRead r("", s);
        io::LongestValid(r);
        reads.push_back(r);
    }
    return reads;
}

io::ReadStreamList<io::SingleReadSeq> make_streams(const std::vector<io::SingleRead> &reads, unsigned T, bool rc) {
    io::ReadStreamList<io::SingleReadSeq> streams;
    size_t n = reads.size();
    for (unsigned t = 0; t < T; ++t) {
        size_t b = n * t / T, e = n * (t + 1) / T;
        std::vector<io::SingleReadSeq> part;
        for (size_t i = b; i < e; ++i) part.emplace_back(reads[i].sequence());
        if (rc) streams.push_back(io::RCWrap<io::SingleReadSeq>(io::VectorReadStream<io::SingleReadSeq>(part)));
        else streams.push_back(io::VectorReadStream<io::SingleReadSeq>(part));
    }
    return streams;
}

struct IndexPeek : utils::DeBruijnExtensionIndex<> {   // IndexWrapper::index_ptr_ is protected; this accessor adds no state
    std::string bytes() const {
        std::ostringstream os(std::ios::binary);
        this->index_ptr_->serialize(os);
        return os.str();
    }
};

template<class Storage>
std::string bucket_bytes(const Storage &st, size_t b) {
    std::string out;
    for (auto it = st.bucket_begin(b); it != st.bucket_end(b); ++it) {
        auto rec = *it;
        out.append(reinterpret_cast<const char *>(rec.first), rec.second);
    }
    return out;
}

std::vector<std::string> gfa_lines(unsigned k, utils::DeBruijnExtensionIndex<> &ext, const std::vector<Sequence> &seqs) {
    debruijn_graph::DeBruijnGraph g(k);
    debruijn_graph::FastGraphFromSequencesConstructor<debruijn_graph::DeBruijnGraph>(k, ext).ConstructGraph(g, seqs);
    std::ostringstream os;
    gfa::GFAWriter writer(g, os);
    writer.WriteSegmentsAndLinks();
    std::vector<std::string> lines;
    std::istringstream is(os.str());
    std::string l;
    while (std::getline(is, l)) lines.push_back(l);
    std::sort(lines.begin(), lines.end());
    return lines;
}

int failures = 0;
void report(const char *name, bool ok, const std::string &detail = "") {
    std::cout << name << (ok ? " OK" : " FAIL") << (detail.empty() ? "" : " " + detail) << std::endl;
    if (!ok) ++failures;
}

}  // namespace

int main(int argc, char **argv) {
    std::string reads_path, out = "/tmp/sb200_dropin";
    unsigned k = 21, T = 2, B = 0;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto next = [&]() { return std::string(argv[++i]); };
        if (s == "--reads") reads_path = next();
        else if (s == "--out") out = next();
        else if (s == "-k") k = unsigned(std::stoul(next()));
        else if (s == "-t") T = unsigned(std::stoul(next()));
        else if (s == "--buckets") B = unsigned(std::stoul(next()));
        else { std::cerr << "usage: ref_dropin --reads R.txt --out DIR -k K -t T [--buckets B]\n"; return 2; }
    }
    if (!B) B = 10 * T;
    {   // keep the reference quiet
        logging::logger *lg = logging::create_logger("", logging::L_ERROR);
        lg->add_writer(std::make_shared<logging::console_writer>());
        logging::attach_logger(lg);
    }
    omp_set_num_threads(int(T));
    fs::make_dirs(out);
    auto workdir = fs::tmp::make_temp_dir(out, "dropin");
    auto reads = load_reads(reads_path);
    auto rc_streams = make_streams(reads, T, true);     // what the reference consumes: every read followed by its RC
    auto fwd_streams = make_streams(reads, T, false);   // what the GPU consumes

    sb200_ctx *ctx = nullptr;
    if (sb200_create(0, &ctx) != 0) {
        std::cerr << sb200_last_error(nullptr) << "\n";
        return 3;
    }
    using Index = utils::DeBruijnExtensionIndex<>;
    using Splitter = utils::DeBruijnReadKMerSplitter<io::SingleReadSeq, utils::StoringTypeFilter<Index::storing_type>>;

    // ---- A: the counter ----------------------------------------------------------------------------------------------------------
    sb200_spades::PackedReads packed = sb200_spades::PackedReads::FromStreams(fwd_streams);
    {
        kmers::KMerDiskCounter<RtSeq> ref_counter(workdir, Splitter(workdir, k + 1, rc_streams, 0));
        kmers::KMerDiskStorage<RtSeq> ref_storage = ref_counter.Count(B, T);
        sb200_spades::GpuKMerCounter gpu_counter(workdir, ctx, k + 1, packed);
        kmers::KMerCounter<RtSeq> &as_reference_interface = gpu_counter;   // the reference's abstract interface
        kmers::KMerDiskStorage<RtSeq> gpu_storage = as_reference_interface.Count(B, T);
        bool ok = ref_storage.num_buckets() == gpu_storage.num_buckets() && ref_storage.total_kmers() == gpu_storage.total_kmers();
        for (size_t b = 0; ok && b < ref_storage.num_buckets(); ++b) ok = bucket_bytes(ref_storage, b) == bucket_bytes(gpu_storage, b);
        report("A_counter_bucket_files", ok, std::to_string(gpu_storage.total_kmers()) + " (k+1)-mers in " + std::to_string(B) + " buckets");
        // ... and the reference's own index builder on top of the GPU counter (KMerIndexBuilder::BuildIndex(index, counter))
        gpu_storage.merge();
        ref_storage.merge();
        std::ifstream fa(*gpu_storage.final_kmers(), std::ios::binary), fb(*ref_storage.final_kmers(), std::ios::binary);
        std::string sa((std::istreambuf_iterator<char>(fa)), std::istreambuf_iterator<char>()), sb((std::istreambuf_iterator<char>(fb)), std::istreambuf_iterator<char>());
        report("A_counter_final_kmers_after_merge", sa == sb && !sa.empty());
    }

    // ---- B: the extension index ------------------------------------------------------------------------------------------------------
    Index ext_ref(k), ext_gpu(k);
    rc_streams.reset();
    auto kp_ref = utils::DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(workdir, ext_ref, rc_streams);
    sb200_spades::GpuIndexHandles handles;
    auto kp_gpu = sb200_spades::GpuExtensionIndexBuilder(ctx).BuildExtensionIndexFromReads(workdir, ext_gpu, packed, unsigned(kp_ref.num_buckets()), &handles);
    {
        bool ok = kp_ref.num_buckets() == kp_gpu.num_buckets();
        for (size_t b = 0; ok && b < kp_ref.num_buckets(); ++b) ok = bucket_bytes(kp_ref, b) == bucket_bytes(kp_gpu, b);
        report("B_kpomer_storage", ok);
        report("B_index_size", ext_ref.size() == ext_gpu.size(), std::to_string(ext_gpu.size()) + " k-mers");
        size_t n = 0, bad_idx = 0, bad_mask = 0, bad_kmer = 0;
        auto ir = ext_ref.kmer_begin(1), ig = ext_gpu.kmer_begin(1);   // both iterate their final_kmers file
        for (auto &a = ir[0], &b = ig[0]; a.good() && b.good(); ++a, ++b, ++n) {
            RtSeq x(k, *a), y(k, *b);
            if (!(x == y)) ++bad_kmer;
            auto kr = ext_ref.ConstructKWH(x), kg = ext_gpu.ConstructKWH(x);   // the reference's lookup code on both indices
            if (kr.idx() != kg.idx()) ++bad_idx;
            if (ext_ref.get_raw_value_reference(kr).get_mask() != ext_gpu.get_raw_value_reference(kg).get_mask()) ++bad_mask;
        }
        report("B_final_kmers_file", bad_kmer == 0 && n == ext_ref.size());
        report("B_reference_lookup_on_gpu_index", bad_idx == 0, std::to_string(n) + " lookups");
        report("B_masks", bad_mask == 0);
        // KMerIndex::serialize of an index with an EMPTY bucket is undefined in the reference (a default-constructed bitVector's _nchar is
        // never set, BooPHF.h:137-140,316-323): compare the byte streams only when every bucket holds k-mers
        std::vector<uint64_t> starts(sb200_kmers_num_buckets(handles.kmers) + 1);
        sb200_kmers_bucket_starts(handles.kmers, starts.data());
        bool any_empty = false;
        for (size_t b = 0; b + 1 < starts.size(); ++b) any_empty = any_empty || starts[b] == starts[b + 1];
        if (any_empty) report("B_kmer_index_serialize", true, "(skipped: empty buckets)");
        else report("B_kmer_index_serialize", static_cast<IndexPeek &>(ext_ref).bytes() == static_cast<IndexPeek &>(ext_gpu).bytes());
    }

    // ---- C: unitigs ---------------------------------------------------------------------------------------------------------------------
    std::vector<Sequence> gpu_seqs = sb200_spades::GpuUnbranchingPaths(ctx, handles, true);
    std::vector<Sequence> ref_on_gpu_index = debruijn_graph::UnbranchingPathExtractor(ext_gpu, k).ExtractUnbranchingPathsAndLoops(16 * T);
    std::vector<Sequence> ref_seqs = debruijn_graph::UnbranchingPathExtractor(ext_ref, k).ExtractUnbranchingPathsAndLoops(16 * T);
    auto same = [](const std::vector<Sequence> &a, const std::vector<Sequence> &b) {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); ++i)
            if (!(a[i] == b[i])) return false;
        return true;
    };
    report("C_reference_extractor_on_gpu_index", same(ref_on_gpu_index, ref_seqs), std::to_string(ref_seqs.size()) + " sequences");
    report("C_gpu_unitigs", same(gpu_seqs, ref_seqs));

    // ---- D: graph + GFA --------------------------------------------------------------------------------------------------------------------
    report("D_gfa_from_gpu_index_and_unitigs", gfa_lines(k, ext_gpu, gpu_seqs) == gfa_lines(k, ext_ref, ref_seqs));

    // ---- E: the edge index's counter ------------------------------------------------------------------------------------------------------
    {
        debruijn_graph::DeBruijnGraph g(k);
        debruijn_graph::FastGraphFromSequencesConstructor<debruijn_graph::DeBruijnGraph>(k, ext_ref).ConstructGraph(g, ref_seqs);
        sb200_spades::PackedReads edges = sb200_spades::PackedReads::FromGraphEdges(g);
        {
            using EdgeSplitter = debruijn_graph::DeBruijnGraphKMerSplitter<debruijn_graph::DeBruijnGraph, utils::StoringTypeFilter<utils::SimpleStoring>>;
            kmers::KMerDiskCounter<RtSeq> ref_counter(workdir, EdgeSplitter(workdir, k + 1, g));
            auto ref_storage = ref_counter.Count(B, T);
            sb200_spades::GpuKMerCounter gpu_counter(workdir, ctx, k + 1, edges, /*canonical_only=*/false, /*add_rc=*/false);
            auto gpu_storage = gpu_counter.Count(B, T);
            bool ok = ref_storage.total_kmers() == gpu_storage.total_kmers();
            for (size_t b = 0; ok && b < ref_storage.num_buckets(); ++b) ok = bucket_bytes(ref_storage, b) == bucket_bytes(gpu_storage, b);
            report("E_edge_index_counter_all_kmers", ok, std::to_string(gpu_storage.total_kmers()) + " (k+1)-mers of " + std::to_string(edges.len.size()) + " edges");
        }
        {
            using EdgeSplitter = debruijn_graph::DeBruijnGraphKMerSplitter<debruijn_graph::DeBruijnGraph, utils::StoringTypeFilter<utils::InvertableStoring>>;
            kmers::KMerDiskCounter<RtSeq> ref_counter(workdir, EdgeSplitter(workdir, k + 1, g));
            auto ref_storage = ref_counter.Count(B, T);
            sb200_spades::GpuKMerCounter gpu_counter(workdir, ctx, k + 1, edges, /*canonical_only=*/true, /*add_rc=*/false);
            auto gpu_storage = gpu_counter.Count(B, T);
            bool ok = ref_storage.total_kmers() == gpu_storage.total_kmers();
            for (size_t b = 0; ok && b < ref_storage.num_buckets(); ++b) ok = bucket_bytes(ref_storage, b) == bucket_bytes(gpu_storage, b);
            report("E_edge_index_counter_minimal_kmers", ok, std::to_string(gpu_storage.total_kmers()) + " (k+1)-mers");
        }
    }

    std::cout << (failures ? "DROPIN FAIL" : "DROPIN OK") << std::endl;
    return failures ? 1 : 0;
}

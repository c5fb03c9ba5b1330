// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference implementation (SPAdes 3.15.4 headers + a few of its .cpp
// files, compiled where they lie under /root/reference by oracle/Makefile into oracle/_ref/)
// through the exact call sequence of the hot path, and dumps every intermediate artefact so
// the CPU restatement (oracle/sb200_oracle.c) and the CUDA path can be compared byte-for-byte.
//
// Call sequence mirrored (reference file:line):
//   gbuilder   A/projects/gbuilder/main.cpp:165-181  and  C/stages/construction.cpp:218-372
//     KMerDiskCounter<RtSeq>(DeBruijnReadKMerSplitter<.., StoringTypeFilter<InvertableStoring>>).Count(B, T)
//         C/utils/extension_index/kmer_extension_index_builder.hpp:62-80
//     DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromKPOMers(...)      ibid. :82-106
//     [EarlyTipClipperProcessor(ext, bound).ClipTips()]  C/assembly_graph/construction/early_simplification.hpp:37-160
//     UnbranchingPathExtractor(ext,k).ExtractUnbranchingPathsAndLoops(nchunks)
//         C/assembly_graph/construction/debruijn_graph_constructor.hpp:377-384
//     [CoverageHashMapBuilder().BuildIndex(cov, kpomers, streams)]  C/utils/ph_map/coverage_hash_map_builder.hpp:39-54
//   tobinary   C/io/reads/binary_converter.cpp:50-113 (BinaryWriter::ToBinary) -> lib.seq / lib.off
//   kmercount  A/projects/kmercount/main.cpp:186-228: every k-window of read and RC(read), no canonical
//     filter, CountAll(16, T, merge=true).  The tool's own splitter class lives inside main.cpp, so the
//     driver uses the reference's DeBruijnReadKMerSplitter with the always-true filter
//     (StoringTypeFilter<SimpleStoring>) over RC-wrapped streams, which pushes the same multiset through
//     the same KMerSortingSplitter/KMerDiskCounter code; equality with the real tool is pinned by the
//     md5 d47405a3a23aed21661194c705a67970 on assembler/test_dataset k=21.
//
// Input: text file, one read per line (ACGTN; '>' / '@' header lines and blank lines ignored when --fasta).
// Reads pass through the reference's LongestValid() (C/io/reads/longest_valid_wrapper.hpp:15-52) exactly
// as io::EasyStream(handle_Ns=true) would do, then are split round-robin... no: CONTIGUOUSLY into T streams.
//
// Output directory (all little-endian, raw):
//   kpomers.<b>        sorted-unique (k+1)-mer records of bucket b   (gbuilder)
//   final_kmers        k-mers (k for gbuilder / kmercount) in file order (buckets concatenated)
//   bucket_sizes.u64   B uint64 sizes of the k-mer buckets
//   index.bin          KMerIndex::serialize bytes of the k-mer MPHF (gbuilder)
//   idx.u64            MPHF index of every k-mer in final_kmers order
//   masks.u8           raw InOutMask byte of every k-mer in final_kmers order (before extraction; after tip clipping if any)
//   masks_idx.u8       the data_ array itself, i.e. masks in MPHF-index order
//   unitigs.txt        one sequence per line, in the reference's output order (paths, then loops)
//   coverage.u32       per (k+1)-mer multiplicity in kpomer file order (if --coverage)
//   graph_cov.gfa      graph.gfa after FillCoverageAndFlankingFromPHM (spades-gbuilder -c); flanking.txt: raw flanking coverage per edge / conjugate
//   timing.txt         phase wall times in seconds
#include "utils/extension_index/kmer_extension_index_builder.hpp"
#include "utils/ph_map/coverage_hash_map_builder.hpp"
#include "assembly_graph/construction/debruijn_graph_constructor.hpp"
#include "assembly_graph/construction/early_simplification.hpp"
#include "io/reads/vector_reader.hpp"
#include "io/reads/read_stream_vector.hpp"
#include "io/reads/rc_reader_wrapper.hpp"
#include "io/reads/longest_valid_wrapper.hpp"
#include "io/reads/converting_reader_wrapper.hpp"
#include "io/reads/binary_converter.hpp"
#include "assembly_graph/core/graph.hpp"
#include "assembly_graph/graph_support/coverage_filling.hpp"
#include "io/graph/gfa_writer.hpp"
#include "io/graph/fastg_writer.hpp"
#include "utils/logger/log_writers.hpp"
#include "utils/filesystem/temporary.hpp"

#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <cstring>
#include <omp.h>

namespace {

double now() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Args {
    std::string mode, reads, out;
    unsigned k = 21, threads = 1, buckets = 0, nchunks = 0;
    long tip_bound = -1;
    bool coverage = false, quiet = false, no_dump = false, no_graph = false;
};

void write_file(const std::string &path, const void *p, size_t n) {
    std::ofstream os(path, std::ios::binary);
    os.write(reinterpret_cast<const char *>(p), std::streamsize(n));
}

std::vector<io::SingleRead> load_reads(const std::string &path) {
    std::vector<io::SingleRead> reads;
    std::ifstream in(path);
    std::string s;
    while (std::getline(in, s)) {
        if (s.empty() || s[0] == '>' || s[0] == '@')
            continue;
        io::SingleRead r("", s);
        io::LongestValid(r);   // what EasyStream(handle_Ns=true) does before binary conversion
        reads.push_back(r);
    }
    return reads;
}

// Split the reads contiguously into T in-memory streams of packed reads, each followed by its RC
// (read_converter.cpp:218-242 builds the same shape from the binary files).
io::ReadStreamList<io::SingleReadSeq> make_streams(const std::vector<io::SingleRead> &reads, unsigned T) {
    io::ReadStreamList<io::SingleReadSeq> streams;
    size_t n = reads.size();
    for (unsigned t = 0; t < T; ++t) {
        size_t b = n * t / T, e = n * (t + 1) / T;
        std::vector<io::SingleReadSeq> part;
        part.reserve(e - b);
        for (size_t i = b; i < e; ++i)
            part.emplace_back(reads[i].sequence());
        streams.push_back(io::RCWrap<io::SingleReadSeq>(io::VectorReadStream<io::SingleReadSeq>(part)));
    }
    return streams;
}

template<class Storage>
void dump_buckets(const Storage &st, const std::string &prefix, const std::string &sizes_path) {
    std::vector<uint64_t> sizes;
    for (size_t b = 0; b < st.num_buckets(); ++b) {
        sizes.push_back(st.bucket_size(b));
        std::ofstream os(prefix + std::to_string(b), std::ios::binary);
        for (auto it = st.bucket_begin(b); it != st.bucket_end(b); ++it) {
            auto rec = *it;
            os.write(reinterpret_cast<const char *>(rec.first), std::streamsize(rec.second));
        }
    }
    write_file(sizes_path, sizes.data(), sizes.size() * 8);
}

// IndexWrapper::index_ptr_ is protected (perfect_hash_map.hpp:35); this accessor adds no state.
struct IndexPeek : utils::DeBruijnExtensionIndex<> {
    template<class W> void serialize_index(W &w) const { this->index_ptr_->serialize(w); }
};

}  // namespace

int main(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto next = [&]() { return std::string(argv[++i]); };
        if (s == "--mode") a.mode = next();
        else if (s == "--reads") a.reads = next();
        else if (s == "--out") a.out = next();
        else if (s == "-k") a.k = unsigned(std::stoul(next()));
        else if (s == "-t") a.threads = unsigned(std::stoul(next()));
        else if (s == "--buckets") a.buckets = unsigned(std::stoul(next()));
        else if (s == "--nchunks") a.nchunks = unsigned(std::stoul(next()));
        else if (s == "--tip-bound") a.tip_bound = std::stol(next());
        else if (s == "--coverage") a.coverage = true;
        else if (s == "--quiet") a.quiet = true;
        else if (s == "--no-dump") a.no_dump = true;
        else if (s == "--no-graph") a.no_graph = true;
        else { std::cerr << "unknown arg " << s << "\n"; return 2; }
    }
    if (a.mode.empty() || a.reads.empty() || a.out.empty()) {
        std::cerr << "usage: ref_driver --mode gbuilder|kmercount|tobinary --reads R.txt --out DIR -k K -t T "
                     "[--buckets B] [--nchunks N] [--tip-bound L] [--coverage] [--quiet] [--no-dump] [--no-graph]\n";
        return 2;
    }
    if (!a.quiet) {
        logging::logger *lg = logging::create_logger("");
        lg->add_writer(std::make_shared<logging::console_writer>());
        logging::attach_logger(lg);
    } else {
        logging::logger *lg = logging::create_logger("", logging::L_ERROR);
        lg->add_writer(std::make_shared<logging::console_writer>());
        logging::attach_logger(lg);
    }
    omp_set_num_threads(int(a.threads));
    fs::make_dir(a.out);
    std::ostringstream timing;

    double t0 = now();
    auto reads = load_reads(a.reads);
    auto streams = make_streams(reads, a.threads);
    size_t bases = 0;
    for (const auto &r : reads) bases += r.size();
    timing << "load " << now() - t0 << "\n" << "reads " << reads.size() << "\nbases " << bases << "\n";
    auto workdir = fs::tmp::make_temp_dir(a.out, "ref");

    if (a.mode == "tobinary") {
        // io::BinaryWriter::ToBinary (C/io/reads/binary_converter.cpp:50-113) on the LongestValid-trimmed reads:
        // <out>/lib.seq + <out>/lib.off, the files io::BinaryFileStream reads back (binary_streams.hpp:48-97)
        io::ReadStream<io::SingleRead> rs = io::VectorReadStream<io::SingleRead>(reads);
        io::BinaryWriter(a.out + "/lib").ToBinary(rs);
    } else if (a.mode == "kmercount") {
        unsigned B = a.buckets ? a.buckets : 16;   // kmercount/main.cpp:215
        using Splitter = utils::DeBruijnReadKMerSplitter<io::SingleReadSeq,
                                                         utils::StoringTypeFilter<utils::SimpleStoring>>;
        t0 = now();
        kmers::KMerDiskCounter<RtSeq> counter(workdir, Splitter(workdir, a.k, streams, 0));
        auto storage = counter.Count(B, a.threads);
        timing << "count " << now() - t0 << "\n";
        if (!a.no_dump) {
            std::vector<uint64_t> sizes;
            for (size_t b = 0; b < storage.num_buckets(); ++b) sizes.push_back(storage.bucket_size(b));
            write_file(a.out + "/bucket_sizes.u64", sizes.data(), sizes.size() * 8);
            storage.merge();
            std::ifstream src(*storage.final_kmers(), std::ios::binary);
            std::ofstream(a.out + "/final_kmers", std::ios::binary) << src.rdbuf();
        }
    } else if (a.mode == "gbuilder") {
        unsigned B = a.buckets ? a.buckets : 10 * a.threads;   // kmer_extension_index_builder.hpp:74
        unsigned k = a.k;
        using Index = utils::DeBruijnExtensionIndex<>;
        Index ext(k);
        using Splitter = utils::DeBruijnReadKMerSplitter<io::SingleReadSeq,
                                                         utils::StoringTypeFilter<Index::storing_type>>;
        t0 = now();
        kmers::KMerDiskCounter<RtSeq> counter(workdir, Splitter(workdir, k + 1, streams, 0));
        auto kpomers = counter.Count(B, a.threads);
        double t1 = now();
        double dump_secs = 0;   // writing the artefacts out is not part of the path: path_total excludes it
        timing << "count_kpomers " << t1 - t0 << "\n";
        if (!a.no_dump) {
            dump_buckets(kpomers, a.out + "/kpomers.", a.out + "/kpomer_bucket_sizes.u64");
            dump_secs += now() - t1;
        }

        t1 = now();
        utils::DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromKPOMers(workdir, ext, kpomers, a.threads, 0);
        double t2 = now();
        timing << "extension_index " << t2 - t1 << "\n";

        size_t clipped = 0;
        if (a.tip_bound >= 0) {
            clipped = debruijn_graph::EarlyTipClipperProcessor(ext, size_t(a.tip_bound)).ClipTips();
            timing << "tipclip " << now() - t2 << "\nclipped " << clipped << "\n";
        }

        const double t_dump2 = now();
        if (!a.no_dump) {
            // k-mers in final_kmers order with idx and raw mask.
            std::ofstream fk(a.out + "/final_kmers", std::ios::binary);
            std::vector<uint64_t> idx;
            std::vector<uint8_t> masks;
            auto its = ext.kmer_begin(1);
            size_t W = RtSeq::GetDataSize(k);
            for (auto &it = its[0]; it.good(); ++it) {
                RtSeq km(k, *it);
                fk.write(reinterpret_cast<const char *>(km.data()), std::streamsize(W * 8));
                auto kwh = ext.ConstructKWH(km);
                idx.push_back(kwh.idx());
                masks.push_back(ext.get_raw_value_reference(kwh).get_mask());
            }
            write_file(a.out + "/idx.u64", idx.data(), idx.size() * 8);
            write_file(a.out + "/masks.u8", masks.data(), masks.size());
            std::vector<uint8_t> by_idx(masks.size());
            for (size_t i = 0; i < idx.size(); ++i) by_idx[idx[i]] = masks[i];
            write_file(a.out + "/masks_idx.u8", by_idx.data(), by_idx.size());
            std::ofstream ib(a.out + "/index.bin", std::ios::binary);
            static_cast<IndexPeek &>(ext).serialize_index(ib);
            dump_secs += now() - t_dump2;
        }

        // Coverage needs the streams again and must run before extraction only because extraction
        // does not touch the (k+1)-mer storage; order is irrelevant to the result.
        using CoverageMap = utils::PerfectHashMap<RtSeq, uint32_t, utils::slim_kmer_index_traits<RtSeq>,
                                                  utils::DefaultStoring>;
        CoverageMap cov(k + 1);
        if (a.coverage) {
            double tc = now();
            utils::CoverageHashMapBuilder().BuildIndex(cov, kpomers, streams);
            timing << "coverage " << now() - tc << "\n";
            const double t_dump3 = now();
            if (!a.no_dump) {
                std::vector<uint32_t> c;
                for (size_t b = 0; b < kpomers.num_buckets(); ++b)
                    for (auto it = kpomers.bucket_begin(b); it != kpomers.bucket_end(b); ++it) {
                        RtSeq x(k + 1, (*it).first);
                        c.push_back(cov.get_raw_value_reference(cov.ConstructKWH(x)));
                    }
                write_file(a.out + "/coverage.u32", c.data(), c.size() * 4);
                dump_secs += now() - t_dump3;
            }
        }

        double t3 = now();
        unsigned nchunks = a.nchunks ? a.nchunks : 16 * a.threads;   // debruijn_graph_constructor.hpp:542
        auto seqs = debruijn_graph::UnbranchingPathExtractor(ext, k).ExtractUnbranchingPathsAndLoops(nchunks);
        timing << "unitigs " << now() - t3 << "\nn_unitigs " << seqs.size() << "\n";
        timing << "path_total " << now() - t0 - dump_secs << "\n" << "dump " << dump_secs << "\n";
        if (!a.no_dump) {
            std::ofstream us(a.out + "/unitigs.txt");
            for (const auto &s : seqs) us << s.str() << '\n';
        }
        if (!a.no_dump && !a.no_graph) {
            // spades-gbuilder --gfa (A/projects/gbuilder/main.cpp:196-219): graph from the unitigs, segments + links
            debruijn_graph::DeBruijnGraph g(k);
            debruijn_graph::FastGraphFromSequencesConstructor<debruijn_graph::DeBruijnGraph>(k, ext).ConstructGraph(g, seqs);
            std::ofstream gf(a.out + "/graph.gfa");
            gfa::GFAWriter gfa_writer(g, gf);
            gfa_writer.WriteSegmentsAndLinks();
            // spades-gbuilder --fastg (main.cpp:218-220)
            const std::string fastg_name = a.out + "/graph.fastg";
            io::FastgWriter fastg_writer(g, fastg_name);
            fastg_writer.WriteSegmentsAndLinks();
            if (a.coverage) {
                // spades-gbuilder -c (main.cpp:200-211): per-edge coverage from the (k+1)-mer multiplicities
                // (GraphCoverageFiller, assembly_graph/graph_support/coverage_filling.hpp:16-95), then the same GFA with DP:f / KC:i
                omnigraph::FlankingCoverage<debruijn_graph::DeBruijnGraph> flanking_cov(g, 50);
                debruijn_graph::FillCoverageAndFlankingFromPHM(cov, g, flanking_cov);
                std::ofstream gc(a.out + "/graph_cov.gfa");
                gfa::GFAWriter cov_writer(g, gc);
                cov_writer.WriteSegmentsAndLinks();
                // FlankingCoverage::GetInCov / GetOutCov of every canonical edge (detail_coverage.hpp), one line per edge in id order
                std::ofstream fl(a.out + "/flanking.txt");
                for (debruijn_graph::EdgeId e : g.canonical_edges())
                    fl << g.int_id(e) << '\t' << flanking_cov.RawCoverage(e) << '\t' << flanking_cov.RawCoverage(g.conjugate(e)) << '\n';
            }
        }
    } else {
        std::cerr << "bad mode\n";
        return 2;
    }
    std::ofstream(a.out + "/timing.txt") << timing.str();
    std::cout << timing.str();
    return 0;
}

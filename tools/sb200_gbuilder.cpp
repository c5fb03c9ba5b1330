// sb200_gbuilder — the reference's spades-gbuilder / spades-kmercount call sequences over the C++ adapters
// (include/sb200_adapters.hpp -> libspades_b200.so).  It exists to show, and test, that the path is a drop-in at the
// level of the reference's own interfaces:
//   gbuilder   A/projects/gbuilder/main.cpp:165-181 :  BuildExtensionIndexFromStream -> [EarlyTipClipper] ->
//              UnbranchingPathExtractor::ExtractUnbranchingPathsAndLoops, unitigs written one per line / FASTA
//   kmercount  A/projects/kmercount/main.cpp:186-228:  KMerDiskCounter(all windows of read and RC).CountAll(16) -> final_kmers
// Input: a text file with one read per line (N allowed: LongestValid applies), or --binary PREFIX for the reference's
// own PREFIX.seq/.off files.  Not a product CLI — spades-gbuilder itself stays the front end.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <fstream>
#include <iostream>
#include <string>

#include "sb200_adapters.hpp"

static double now() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    std::string mode = "gbuilder", reads_path, binary_prefix, out, write_binary;
    unsigned k = 21, threads = 8;
    long tip_bound = -1;
    bool coverage = false, self_check = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--mode") mode = next();
        else if (a == "--reads") reads_path = next();
        else if (a == "--binary") binary_prefix = next();
        else if (a == "--write-binary") write_binary = next();
        else if (a == "--out") out = next();
        else if (a == "-k") k = (unsigned) atoi(next().c_str());
        else if (a == "-t") threads = (unsigned) atoi(next().c_str());
        else if (a == "--tip-clip") tip_bound = atol(next().c_str());
        else if (a == "--coverage") coverage = true;
        else if (a == "--self-check") self_check = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (out.empty() || (reads_path.empty() && binary_prefix.empty())) {
        fprintf(stderr, "usage: sb200_gbuilder --mode gbuilder|kmercount (--reads FILE | --binary PREFIX) --out DIR -k K [-t T] [--tip-clip BOUND] [--coverage]\n");
        return 2;
    }
    try {
        sb200::ReadStreamList reads;
        if (!binary_prefix.empty()) {
            reads = sb200::ReadStreamList::FromBinaryFiles(binary_prefix);
        } else {
            std::ifstream in(reads_path);
            if (!in) throw sb200::Error(1, "cannot open " + reads_path);
            std::string s;
            while (std::getline(in, s)) {
                if (s.empty() || s[0] == '>' || s[0] == '@') continue;
                reads.push_back_raw(s);
            }
        }
        if (!write_binary.empty()) sb200::WriteBinaryReads(reads, write_binary);
        sb200::Context ctx(0);   // throws without an sm_100 GPU: there is no CPU path
        double t0 = now();
        sb200::DeviceReads streams(ctx, reads);
        if (mode == "kmercount") {
            sb200::DeBruijnReadKMerSplitter splitter(streams, k, /*canonical_only=*/false, /*add_rc=*/true);
            sb200::KMerDiskCounter counter(splitter);
            sb200::KMerDiskStorage st = counter.CountAll(16, threads, true);
            st.write_final_kmers(out + "/final_kmers");
            printf("%zu kmers in total\n", st.total_kmers());
            return 0;
        }
        sb200::DeBruijnExtensionIndex index(ctx, k);
        sb200::KMerDiskStorage kpomers = sb200::DeBruijnExtensionIndexBuilder().BuildExtensionIndexFromStream(index, streams, threads);
        if (tip_bound >= 0) {
            size_t removed = sb200::EarlyTipClipperProcessor(index, (size_t) tip_bound).ClipTips();
            printf("%zu %u-mers were removed by early tip clipper\n", removed, k + 1);
        }
        // written before the extraction, like the reference driver's dump (the reference extractor isolates consumed vertices)
        {
            auto m = index.data();
            std::ofstream os(out + "/masks_idx.u8", std::ios::binary);
            os.write((const char *) m.data(), (std::streamsize) m.size());
            std::ofstream ib(out + "/index.bin", std::ios::binary);
            index.index().serialize(ib);
        }
        sb200::UnbranchingPathExtractor extractor(index, k);
        std::vector<sb200::Sequence> edges = extractor.ExtractUnbranchingPathsAndLoops(threads * 16);
        double t1 = now();
        kpomers.write_final_kmers(out + "/kpomers");
        index.kmers().write_final_kmers(out + "/final_kmers");
        if (coverage) {
            auto c = sb200::CoverageHashMapBuilder().FillCoverage(kpomers);
            std::ofstream os(out + "/coverage.u32", std::ios::binary);
            os.write((const char *) c.data(), (std::streamsize) (c.size() * 4));
        }
        if (self_check) {
            // the per-key and per-bucket faces of the reference's interfaces: KMerDiskStorage::bucket_begin/end iterate the records of
            // kmers<i> (kmer_index_builder.hpp:149-159), KMerIndex::seq_idx(const Seq &) looks one key up from host code (kmer_index.hpp:85-90)
            const sb200::KMerDiskStorage &km = index.kmers();
            const unsigned W = km.kmer_words();
            std::vector<uint64_t> all = km.final_kmers();
            std::vector<uint32_t> idx = index.idx();
            size_t pos = 0, bad = 0;
            for (size_t b = 0; b < km.num_buckets(); ++b) {
                size_t in_bucket = 0;
                for (auto it = km.bucket_begin(b); it != km.bucket_end(b); ++it, ++pos, ++in_bucket) {
                    auto rec = *it;
                    if (rec.second != W * 8 || memcmp(rec.first, &all[pos * W], W * 8) != 0) ++bad;
                    if (pos % 7 == 0 && index.index().seq_idx(sb200::Sequence(rec.first, k)) != idx[pos]) ++bad;
                }
                if (in_bucket != km.bucket_size(b)) ++bad;
            }
            printf("self-check %s: %zu records through the bucket iterators, single-key seq_idx on every 7th\n", (bad == 0 && pos == km.total_kmers()) ? "OK" : "FAIL", pos);
            if (bad || pos != km.total_kmers()) return 1;
        }
        std::ofstream us(out + "/unitigs.txt");
        for (const auto &e : edges) us << e.str() << "\n";
        {   // spades-gbuilder --gfa: graph from the unitigs (link records keyed by the MPHF), segments + links
            sb200::CondensedGraph graph(index, edges);
            std::ofstream gf(out + "/graph.gfa");
            graph.WriteGFA(gf);
            std::ofstream fg(out + "/graph.fastg");   // spades-gbuilder --fastg
            graph.WriteFASTG(fg);
            if (coverage) {   // spades-gbuilder -c (main.cpp:200-211): coverage map over the (k+1)-mers, per-edge coverage, the same GFA with DP:f / KC:i
                sb200::CoverageHashMap cov = sb200::CoverageHashMapBuilder().BuildIndex(ctx, kpomers);
                std::vector<uint64_t> kc, flank;
                cov.FillCoverageAndFlanking(extractor, kc, flank, 50);
                graph.SetCoverage(kc);
                std::ofstream gc(out + "/graph_cov.gfa");
                graph.WriteGFA(gc);
                std::ofstream fl(out + "/flanking.txt");
                for (size_t i = 0; i < kc.size(); ++i)
                    fl << (sb200::CondensedGraph::ID_BIAS + 2 * i) << '\t' << flank[2 * i] << '\t' << flank[2 * i + 1] << '\n';
            }
        }
        printf("%zu (k+1)-mers, %zu k-mers, %zu unitigs, %.3f s on device incl. transfers, %llu kernel launches\n", kpomers.total_kmers(),
               index.size(), edges.size(), t1 - t0, (unsigned long long) ctx.kernel_launches());
    } catch (const sb200::Error &e) {
        fprintf(stderr, "FATAL: %s\n", e.what());   // the reference: FATAL_ERROR -> exit(errno ? errno : -1)
        return 255;
    }
    return 0;
}

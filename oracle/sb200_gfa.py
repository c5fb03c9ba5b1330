"""TEST INFRASTRUCTURE — CPU restatement of the step right after the hot path:
FastGraphFromSequencesConstructor::ConstructGraph + gfa::GFAWriter::WriteSegmentsAndLinks.  Pure Python, small cases only.
Only tests/ may import this file; the product (include/sb200_adapters.hpp CondensedGraph) never does.

Reference lines followed (A = /root/reference/assembler/src/common):
  LinkRecord            A/assembly_graph/construction/debruijn_graph_constructor.hpp:400-430   key = idx << 2 | rc << 1 | start
  StartLink / EndLink   :432-448      canonical form of the first / last k-mer, MPHF index of it
  CollectLinkRecords    :450-466      edge i gets id min_id + 2i, its conjugate the next id; a self-conjugate edge has no end record
  ConstructGraph        :483-517      sort records, one vertex (pair) per distinct idx, LinkEdge per record
  LinkEdge              :468-478      rc records attach to conjugate(v); start -> outgoing, end -> incoming
  GFAWriter             A/io/graph/gfa_writer.cpp:18-52   S lines for canonical edges, L lines = incoming x outgoing of canonical vertices
  CanonicalEdgeHelper   A/io/utils/edge_namer.hpp:71-86   name = min(e, conj e), '+' iff e <= conj e
  ID_BIAS = 3           A/assembly_graph/core/graph_core.hpp (first id handed out)
  FastgWriter           A/io/graph/fastg_writer.cpp:19-46   one record per oriented edge, header = name:sorted set of successors;
  BasicNamingF          A/io/utils/edge_namer.hpp:29-35     EDGE_<id>_length_<bases>_cov_<std::to_string(coverage)>, "'" for the conjugate
Pinned by tests/golden/*.npz["gfa"] / ["fastg"] (the unmodified reference's own graph.gfa / graph.fastg, lines / records sorted).
"""

ID_BIAS = 3
_COMP = str.maketrans("ACGT", "TGCA")


def _rc(s):
    return s[::-1].translate(_COMP)


def _rtseq_less(a, b):
    """operator<(RuntimeSeq, RuntimeSeq) is nucleotide-wise from position 0 (rtseq.hpp:733-741) — NOT the word order the k-mer
    files are sorted in; with A<C<G<T it is plain string order."""
    return a < b


def _graph(unitigs, k, seq_idx):
    """-> (conj, vertices, end): vertices = list of (incoming, outgoing) oriented edge ids of the canonical vertex of every pair;
    end[e] = (vertex index, at the conjugate vertex?) for every oriented edge e"""
    self_conj = [u == _rc(u) for u in unitigs]

    def conj(e):
        i = (e - ID_BIAS) >> 1
        return e if self_conj[i] else (((e - ID_BIAS) ^ 1) + ID_BIAS)

    records = []
    for i, u in enumerate(unitigs):
        e = ID_BIAS + 2 * i
        for start in (True, False):
            if not start and self_conj[i]:
                continue
            km = u[:k] if start else u[-k:]
            r = _rc(km)
            is_rc = not _rtseq_less(km, r)
            records.append(((seq_idx(r if is_rc else km) << 2) | (int(is_rc) << 1) | int(start), e))
    records.sort()
    vertices, end = [], {}
    pos = 0
    while pos < len(records):
        j = pos
        inc, out = [], []
        while j < len(records) and (records[j][0] >> 2) == (records[pos][0] >> 2):
            key, e = records[j]
            is_rc, start = bool(key & 2), bool(key & 1)
            if start:   # LinkOutgoingEdge(v or conj v, e): e starts there, so conj(e) ends at the conjugate of that vertex
                (inc if is_rc else out).append(conj(e) if is_rc else e)
                end[conj(e)] = (len(vertices), not is_rc)
            else:       # LinkIncomingEdge(v or conj v, e)
                (out if is_rc else inc).append(conj(e) if is_rc else e)
                end[e] = (len(vertices), is_rc)
            j += 1
        vertices.append((inc, out))
        pos = j
    return conj, vertices, end


def edge_coverage(unitigs, k, count_of, averaging_range=50):
    """GraphCoverageFiller::FillCoverageFromEdges (A/assembly_graph/graph_support/coverage_filling.hpp:45-63): per sequence the sum of
    the multiplicities of its (k+1)-mers (count_of(canonical (k+1)-mer string) -> CoverageHashMap value) and the same over its first /
    last `averaging_range` (k+1)-mers (FlankingCoverage of the edge / of its conjugate).  Returns (kc, [(start, end)])."""
    kc, flank = [], []
    for u in unitigs:
        vals = []
        for p in range(len(u) - k):
            x = u[p:p + k + 1]
            r = _rc(x)
            vals.append(count_of(x if x <= r else r))
        rng = min(averaging_range, len(vals))
        kc.append(sum(vals))
        flank.append((sum(vals[:rng]), sum(vals[len(vals) - rng:])))
    return kc, flank


def cxx_float(x):
    """what `os << float(x)` prints (io/graph/gfa_writer.cpp:18-26: DP:f): the value rounded to binary32, then %g with 6 significant digits"""
    import struct
    return "%g" % struct.unpack("f", struct.pack("f", x))[0]


def gfa_lines(unitigs, k, seq_idx, kc=None):
    """unitigs: list[str] in extractor order; seq_idx(str canonical k-mer) -> MPHF index; kc: per-sequence raw coverage (spades-gbuilder
    -c) or None.  Returns the GFA lines in id order."""
    conj, vertices, _ = _graph(unitigs, k, seq_idx)
    if kc is None:
        lines = ["S\t%d\t%s\tDP:f:0\tKC:i:0" % (ID_BIAS + 2 * i, u) for i, u in enumerate(unitigs)]
    else:   # coverage(e) = raw coverage / length(e), length = number of (k+1)-mers (A/assembly_graph/core/coverage.hpp:58-60)
        lines = ["S\t%d\t%s\tDP:f:%s\tKC:i:%d" % (ID_BIAS + 2 * i, u, cxx_float(float(kc[i]) / float(len(u) - k)), kc[i]) for i, u in enumerate(unitigs)]
    for inc, out in vertices:
        for a in inc:
            for b in out:
                lines.append("L\t%d\t%s\t%d\t%s\t%dM" % (min(a, conj(a)), "+" if a <= conj(a) else "-",
                                                         min(b, conj(b)), "+" if b <= conj(b) else "-", k))
    return lines


def fastg_records(unitigs, k, seq_idx):
    """FASTG records (header line + sequence wrapped at 60) of every oriented edge, in id order."""
    conj, vertices, end = _graph(unitigs, k, seq_idx)

    def name(e):
        c = min(e, conj(e))
        return "EDGE_%d_length_%d_cov_%s%s" % (c, len(unitigs[(c - ID_BIAS) >> 1]), "%.6f" % 0.0, "" if e == c else "'")

    recs = []
    for i, u in enumerate(unitigs):
        e = ID_BIAS + 2 * i
        for x in ([e] if conj(e) == e else [e, e + 1]):
            v, at_conj = end[x]
            inc, out = vertices[v]
            nxt = sorted({name(y) for y in out} if not at_conj else {name(conj(y)) for y in inc})
            seq = u if x == e else _rc(u)
            head = ">" + name(x) + (":" + ",".join(nxt) if nxt else "") + ";"
            recs.append("\n".join([head] + [seq[p:p + 60] for p in range(0, len(seq), 60)]))
    return recs

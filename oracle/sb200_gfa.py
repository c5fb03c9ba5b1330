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
Pinned by tests/golden/*.npz["gfa"] (the unmodified reference's own graph.gfa, lines sorted).
"""

ID_BIAS = 3
_COMP = str.maketrans("ACGT", "TGCA")


def _rc(s):
    return s[::-1].translate(_COMP)


def _rtseq_less(a, b):
    """operator<(RuntimeSeq, RuntimeSeq) is nucleotide-wise from position 0 (rtseq.hpp:733-741) — NOT the word order the k-mer
    files are sorted in; with A<C<G<T it is plain string order."""
    return a < b


def gfa_lines(unitigs, k, seq_idx):
    """unitigs: list[str] in extractor order; seq_idx(str canonical k-mer) -> MPHF index.  Returns the GFA lines in id order."""
    n = len(unitigs)
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
    lines = ["S\t%d\t%s\tDP:f:0\tKC:i:0" % (ID_BIAS + 2 * i, u) for i, u in enumerate(unitigs)]
    pos = 0
    while pos < len(records):
        end = pos
        inc, out = [], []
        while end < len(records) and (records[end][0] >> 2) == (records[pos][0] >> 2):
            key, e = records[end]
            is_rc, start = bool(key & 2), bool(key & 1)
            if start:
                (inc if is_rc else out).append(conj(e) if is_rc else e)
            else:
                (out if is_rc else inc).append(conj(e) if is_rc else e)
            end += 1
        for a in inc:
            for b in out:
                lines.append("L\t%d\t%s\t%d\t%s\t%dM" % (min(a, conj(a)), "+" if a <= conj(a) else "-",
                                                         min(b, conj(b)), "+" if b <= conj(b) else "-", k))
        pos = end
    return lines

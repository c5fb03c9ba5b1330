"""Deterministic synthetic read sets of the shapes BASELINE.json / SURVEY.md §8(d) name.

Pure numpy; used by bench.py and by the tests (never by the product kernels).  Reads are produced directly in the
reference's packed layout (base j of a read at bits 2(j%32) of word j/32, every read word-aligned:
C/sequence/sequence.hpp:71-122, single_read.hpp:279-299) or as ACGT strings for the small cases.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_genome(n, seed):
    return np.random.default_rng(seed).integers(0, 4, size=n, dtype=np.uint8)


def revcomp_codes(a):
    return (3 - a[..., ::-1]).astype(np.uint8)


def sample_pairs(genome, n_pairs, read_len, insert, err, seed):
    """FR pairs: mate 1 forward at p, mate 2 = reverse complement of the fragment's tail; i.i.d. substitutions.

    Returns codes[n_reads, read_len] (uint8, 0..3), mates interleaved (r1, r2, r1, r2, ...)."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    start = rng.integers(0, G - insert + 1, size=n_pairs)
    idx = np.arange(read_len)
    r1 = genome[start[:, None] + idx[None, :]]
    r2 = genome[(start + insert - read_len)[:, None] + idx[None, :]]
    r2 = revcomp_codes(r2)
    reads = np.empty((2 * n_pairs, read_len), dtype=np.uint8)
    reads[0::2] = r1
    reads[1::2] = r2
    if err > 0:
        mask = rng.random(reads.shape) < err
        sub = rng.integers(1, 4, size=int(mask.sum()), dtype=np.uint8)
        reads[mask] = (reads[mask] + sub) & 3
    return reads


def pack_codes(reads):
    """codes[n, L] -> (words u64[n*ceil(L/32)], word_off u64[n+1], len u32[n]) in the reference layout."""
    n, L = reads.shape
    wpr = (L + 31) // 32
    padded = np.zeros((n, wpr * 32), dtype=np.uint64)
    padded[:, :L] = reads
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, None, :]
    words = (padded.reshape(n, wpr, 32) << shifts).sum(axis=2, dtype=np.uint64)
    word_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(wpr))
    return words.reshape(-1), word_off, np.full(n, L, dtype=np.uint32)


def pack_codes_chunked(reads, chunk=1 << 18):
    n, L = reads.shape
    wpr = (L + 31) // 32
    out = np.empty(n * wpr, dtype=np.uint64)
    for s in range(0, n, chunk):
        w, _, _ = pack_codes(reads[s:s + chunk])
        out[s * wpr:(s + len(reads[s:s + chunk])) * wpr] = w
    return out, np.arange(n + 1, dtype=np.uint64) * np.uint64(wpr), np.full(n, L, dtype=np.uint32)


def codes_to_strings(reads):
    return [bytes(_ACGT[r]).decode() for r in reads]


def isolate_config(genome_len=4_600_000, read_len=150, insert=350, coverage=100.0, err=0.005, seed=42):
    """BASELINE config #2 shape (scaled by genome_len): returns packed reads."""
    n_pairs = int(genome_len * coverage / (2 * read_len))
    g = random_genome(genome_len, seed)
    codes = sample_pairs(g, n_pairs, read_len, insert, err, seed + 1000)
    return pack_codes_chunked(codes)


# ---- the BASELINE.json workloads ---------------------------------------------------------
# config ids follow BASELINE.json `configs` 1-based: 1 = test_dataset (tests/golden/ecoli1k_k21.npz holds its reads),
# 2 = isolate 2x150 @100x k=55, 3 = the same reads at k = 21, 33, 55, 77, 4 = 2x250 @80x k=127, 5 = metagenome mix k=55.
WORKLOADS = {
    2: dict(name="synthetic isolate 4.6 Mbp genome, 2x150 reads at 100x, k=55 (BASELINE configs[1])", ks=(55,), genome_len=4_600_000,
            read_len=150, insert=350, coverage=100.0, seed=42, scaling="weak"),
    3: dict(name="same synthetic isolate, multi-K 21/33/55/77 back to back on the resident reads (BASELINE configs[2])", ks=(21, 33, 55, 77),
            genome_len=4_600_000, read_len=150, insert=350, coverage=100.0, seed=42, scaling="weak"),
    4: dict(name="long-k path: synthetic 4.6 Mbp genome, 2x250 reads at 80x, k=127 (BASELINE configs[3])", ks=(127,), genome_len=4_600_000,
            read_len=250, insert=500, coverage=80.0, seed=43, scaling="weak"),
    5: dict(name="metagenome mix: 200 genomes log-uniform 1-8 Mbp, log-normal abundances, ~1 Gbp of 2x150 reads, k=55 (BASELINE configs[4])",
            ks=(55,), read_len=150, insert=350, total_bases=1_000_000_000, n_genomes=200, seed=44, scaling="strong"),
}
ERROR_RATE = 0.005
CHUNK_PAIRS = 200_000


def _isolate_chunks(w, genome_len, rank):
    """(genome, [(n_pairs, seed)]) of one rank's share of an isolate workload: the chunking and seeds bench.py has used since round 1"""
    n_pairs = int(w["genome_len"] * w["coverage"] / (2 * w["read_len"]))   # per rank: weak scaling grows the genome, not the per-rank reads
    g = random_genome(genome_len, w["seed"])
    return g, [(min(CHUNK_PAIRS, n_pairs - s), 1000 + w["seed"] + s + 7_000_003 * rank) for s in range(0, n_pairs, CHUNK_PAIRS)]


def _metagenome(w):
    """genomes (list of code arrays) and pairs per genome: lengths log-uniform 1-8 Mbp, abundances log-normal(sigma = 1) normalised
    so that the reads total ~ total_bases"""
    rng = np.random.default_rng(w["seed"])
    lens = np.exp(rng.uniform(np.log(1e6), np.log(8e6), size=w["n_genomes"])).astype(np.int64)
    abund = rng.lognormal(0.0, 1.0, size=w["n_genomes"])
    share = abund * lens
    share /= share.sum()
    pairs = np.maximum((share * w["total_bases"] / (2 * w["read_len"])).astype(np.int64), 1)
    return lens, pairs


def workload_chunks(config, rank=0, world=1):
    """Yields codes[n, read_len] chunks of one rank's reads of a BASELINE workload (deterministic; the union over the ranks of a
    strong-scaling workload is the world = 1 read set)."""
    w = WORKLOADS[config]
    if config in (2, 3, 4):
        g, chunks = _isolate_chunks(w, w["genome_len"] * world, rank)
        for n, seed in chunks:
            yield sample_pairs(g, n, w["read_len"], w["insert"], ERROR_RATE, seed)
        return
    lens, pairs = _metagenome(w)
    c = 0
    for gi in range(w["n_genomes"]):
        g = None
        for s in range(0, int(pairs[gi]), CHUNK_PAIRS):
            if c % world == rank:
                if g is None:
                    g = random_genome(int(lens[gi]), 100_000 * w["seed"] + gi)
                yield sample_pairs(g, min(CHUNK_PAIRS, int(pairs[gi]) - s), w["read_len"], w["insert"], ERROR_RATE,
                                   200_000 * w["seed"] + 1000 * gi + s // CHUNK_PAIRS)
            c += 1


def workload_reads(config, rank=0, world=1, max_reads=None):
    """(words, word_off, len) of one rank's packed reads; max_reads bounds the sample (the FIRST reads of the set)"""
    w = WORKLOADS[config]
    parts, n = [], 0
    for codes in workload_chunks(config, rank, world):
        if max_reads is not None and n + len(codes) > max_reads:
            codes = codes[:max_reads - n]
        parts.append(pack_codes(codes)[0])
        n += len(codes)
        if max_reads is not None and n >= max_reads:
            break
    wpr = (w["read_len"] + 31) // 32
    words = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint64)
    return words, np.arange(n + 1, dtype=np.uint64) * np.uint64(wpr), np.full(n, w["read_len"], dtype=np.uint32)


def pack_text_sequences(text):
    """b"ACGT...\\nACG...\\n" -> (words, word_off, len) in the packed layout of the reads / unitigs (vectorised; test infrastructure for
    comparing the reference's unitigs.txt with the device arrays)"""
    raw = np.frombuffer(text, dtype=np.uint8)
    nl = np.flatnonzero(raw == 10)
    if len(raw) and raw[-1] != 10:
        nl = np.append(nl, len(raw))
    starts = np.concatenate([[0], nl[:-1] + 1]).astype(np.int64)
    lens = (nl - starts).astype(np.int64)
    keep = lens > 0
    starts, lens = starts[keep], lens[keep]
    lut = np.zeros(256, dtype=np.uint8)
    lut[ord("C")] = 1; lut[ord("G")] = 2; lut[ord("T")] = 3
    nwords = (lens + 31) // 32
    word_off = np.concatenate([[0], np.cumsum(nwords)]).astype(np.uint64)
    words = np.zeros(int(word_off[-1]), dtype=np.uint64)
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, :]
    step = 200_000
    for s in range(0, len(lens), step):
        e = min(s + step, len(lens))
        w0, w1 = int(word_off[s]), int(word_off[e])
        buf = np.zeros((w1 - w0) * 32, dtype=np.uint8)
        seq_id = np.repeat(np.arange(s, e), lens[s:e])
        within = np.arange(int(lens[s:e].sum()), dtype=np.int64) - np.repeat(np.cumsum(lens[s:e]) - lens[s:e], lens[s:e])
        src = starts[seq_id] + within
        dst = (word_off[seq_id].astype(np.int64) - w0) * 32 + within
        buf[dst] = lut[raw[src]]
        words[w0:w1] = (buf.reshape(-1, 32).astype(np.uint64) << shifts).sum(axis=1, dtype=np.uint64)
    return words, word_off, lens.astype(np.uint32)

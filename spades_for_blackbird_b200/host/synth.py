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

"""Seeded synthetic genomes and reads for the BASELINE.json configurations.

TEST / BENCH INFRASTRUCTURE -- nothing here is on the product path.

The recipes follow SURVEY.md appendix A.5 / A.7 (the probe commands the
baseline numbers were taken with): iid ACGT contigs from
``numpy.random.default_rng(seed)``, reads that are with probability
``hit_frac`` a substring of the reference (either strand when ``both_strands``)
and otherwise iid random, so that k-mer hit rate ~= hit_frac.
"""
from __future__ import annotations

import numpy as np

LUT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def random_contigs(n_contigs: int, contig_len: int, seed: int) -> np.ndarray:
    """(n_contigs, contig_len) uint8 ASCII matrix of iid ACGT (config 2/3 reference)."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=n_contigs * contig_len, dtype=np.uint8)
    return LUT[codes].reshape(n_contigs, contig_len)


def pangenome(base_len: int, n_copies: int, sub_rate: float, seed: int) -> np.ndarray:
    """(n_copies, base_len) uint8 ASCII: independently mutated copies of one iid base genome
    (config 4/5 'bacterial-pangenome-like' reference)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=base_len, dtype=np.uint8)
    out = np.empty((n_copies, base_len), dtype=np.uint8)
    for i in range(n_copies):
        m = rng.random(base_len) < sub_rate
        shift = rng.integers(1, 4, size=base_len, dtype=np.uint8)
        out[i] = LUT[np.where(m, (base + shift) & 3, base)]
    return out


def revcomp(seqs: np.ndarray) -> np.ndarray:
    return _COMP[seqs[..., ::-1]]


def sample_reads(ref: np.ndarray, n_reads: int, read_len: int, hit_frac: float, seed: int,
                 both_strands: bool = False, chunk: int = 1 << 20) -> np.ndarray:
    """(n_reads, read_len) uint8 ASCII reads from a (n_contigs, contig_len) reference."""
    rng = np.random.default_rng(seed)
    n_contigs, contig_len = ref.shape
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    flat = ref.reshape(-1)
    for a in range(0, n_reads, chunk):
        b = min(n_reads, a + chunk)
        m = b - a
        is_hit = rng.random(m) < hit_frac
        contig = rng.integers(0, n_contigs, size=m)
        off = rng.integers(0, contig_len - read_len + 1, size=m)
        strand = rng.integers(0, 2, size=m) if both_strands else np.zeros(m, dtype=np.int64)
        block = LUT[rng.integers(0, 4, size=(m, read_len), dtype=np.uint8)]
        h = np.nonzero(is_hit)[0]
        if h.size:
            start = contig[h] * contig_len + off[h]
            sub = flat[start[:, None] + ar[None, :]]
            flip = strand[h] == 1
            if flip.any():
                sub[flip] = revcomp(sub[flip])
            block[h] = sub
        out[a:b] = block
    return out


def write_fasta(path: str, seqs, names=None, line_width: int | None = None) -> None:
    """Write sequences (iterable of uint8 arrays / bytes) as FASTA with '\\n' line ends."""
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            s = bytes(s) if not isinstance(s, (bytes, bytearray)) else s
            name = names[i] if names is not None else f"s{i}"
            f.write(b">" + name.encode() + b"\n")
            if line_width:
                for j in range(0, len(s), line_width):
                    f.write(s[j:j + line_width] + b"\n")
            else:
                f.write(s + b"\n")


def write_fasta_matrix(path: str, mat: np.ndarray, prefix: str = "c") -> None:
    """Fast FASTA writer for an (n, len) uint8 matrix: one sequence line per row."""
    with open(path, "wb") as f:
        for i in range(mat.shape[0]):
            f.write(b">" + prefix.encode() + str(i).encode() + b"\n")
            f.write(mat[i].tobytes())
            f.write(b"\n")


def write_fastq(path: str, seqs) -> None:
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            s = bytes(s)
            f.write(b"@r" + str(i).encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


def ragged_to_batch(seqs) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate reads -> (ascii uint8, offsets int64[n+1]) as the C-ABI takes them."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    ascii_ = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    return ascii_, offsets


def matrix_to_batch(mat: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    n, L = mat.shape
    return np.ascontiguousarray(mat).reshape(-1), np.arange(n + 1, dtype=np.int64) * L

"""The narrow layout with more than 2^31 columns (the range of BASELINE configs[3]: ~2.6 G columns): u32 table rows, u32
counts in sectors and csectors, int64 results above 2^31, int32 results refused. An index of that size cannot come from a
fixture, so the four bit vectors are synthetic (one edge per column: any four bit vectors define the walk's arithmetic,
SBWT.hh:423-437), the oracle is built from the same arrays (oracle from_arrays: rank directories, C array and p-mer table
byte-equal to reference-written files on the golden fixtures, tests/test_oracle.py), and reads are spelled by following
edges from random columns, so that most k-mers are found and a third of the answers lie above 2^31."""
import numpy as np
import pytest

import oracle
import sbwt_b200 as S
from sbwt_b200.testing import synth

pytestmark = pytest.mark.gpu

N_COLS = (1 << 31) + (1 << 30) + 98_765  # 3.22e9 columns: narrow layout (< 2^32 - 256), values above 2^31


def _planted_reads(orc, codes, n_cols, n_reads, length, rng):
    """Follow edges from random columns: read = labels of the path, every k-mer of it ends in a unique column (w.h.p.)."""
    reads = []
    for u in rng.integers(0, n_cols - 1, size=n_reads):
        u = int(u)
        s = bytearray()
        for _ in range(length):
            if u == n_cols - 1:  # the one column without an edge
                break
            c = "ACGT"[codes[u]]
            s.append(ord(c))
            u = orc.C_array[codes[u]] + orc.rank(u, c)
        reads.append(bytes(s))
    return reads


@pytest.fixture(scope="module")
def big():
    rng = np.random.default_rng(2031)
    codes = rng.integers(0, 4, size=N_COLS, dtype=np.uint8)
    nw = (N_COLS + 63) // 64
    bits = []
    for c in range(4):
        b = codes == c
        b[-1] = False  # every node but the root has one incoming edge: n_nodes - 1 ones in all
        packed = np.packbits(b, bitorder="little")
        del b
        w = np.zeros(nw * 8, dtype=np.uint8)
        w[: packed.size] = packed
        bits.append(w.view(np.uint64))
    sgs = np.full(nw, np.uint64(0xFFFFFFFFFFFFFFFF))  # every column starts a suffix group: the edge invariant holds
    if N_COLS % 64:
        sgs[-1] = np.uint64((1 << (N_COLS % 64)) - 1)
    arrays = {"bits": bits, "sgs": sgs, "n_nodes": N_COLS, "n_kmers": N_COLS - 1, "k": 31, "precalc_k": 8}
    orc = oracle.OracleIndex(arrays=arrays)
    # reads: planted paths (found), planted with substitutions / N (walk-backs, restarts, probes), random (absent)
    reads = _planted_reads(orc, codes, N_COLS, 1500, 150, rng) + _planted_reads(orc, codes, N_COLS, 40, 700, rng)
    mutated = []
    for r in reads[:400]:
        r = bytearray(r)
        for q in rng.integers(0, len(r), size=3):
            r[int(q)] = ord("ACGTN"[int(rng.integers(0, 5))])
        mutated.append(bytes(r))
    reads += mutated + [bytes(synth.LUT[rng.integers(0, 4, size=150, dtype=np.uint8)]) for _ in range(500)]
    del codes
    a, off = synth.ragged_to_batch(reads)
    want = orc.query_batch(a, off, streaming=True)
    assert np.array_equal(want, orc.query_batch(a, off, streaming=False))
    assert (want >= (1 << 31)).sum() > 20_000 and (want == -1).sum() > 20_000
    return arrays, orc, reads, a, off, want


@pytest.mark.parametrize("compact", ["1", "0"])
def test_narrow_layout_above_2_31_columns(compact, big, monkeypatch):
    monkeypatch.setenv("SBWT_B200_COMPACT", compact)  # csectors (absolute u32 counts) / classic sectors
    arrays, orc, reads, a, off, want = big
    rng = np.random.default_rng(7)
    Carr = orc.C_array
    assert Carr[3] < (1 << 31) + (1 << 29) < N_COLS  # T columns, and part of G's, lie above 2^31
    idx = S.Index(arrays=dict(arrays, C=Carr, precalc=None))
    assert idx.n_nodes == N_COLS and idx.edges_only_at_group_starts and idx.compact_layout[0] == (compact == "1")
    # rank at the ends of the range
    pos = np.concatenate([rng.integers(0, N_COLS + 1, 3000), [0, N_COLS, N_COLS - 1, 1 << 31, (1 << 31) - 1, (1 << 31) + 223]]).astype(np.int64)
    chars = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=pos.size))
    np.testing.assert_array_equal(idx.rank(pos, chars), [orc.rank(int(q), chr(c)) for q, c in zip(pos, chars)])
    ses = S.Session(idx, a.size, len(reads))
    for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
        np.testing.assert_array_equal(ses.query_host(a, off, mode), want)
    with pytest.raises(S.SbwtGpuError, match="fewer than 2\\^31"):
        ses.query_host_i32(a, off, S.MODE_STREAMING)
    # the batched interval queries on the same layout
    l, r, m = idx.partial_search(a, off)
    for i in list(range(0, len(reads), 97)):
        assert (int(l[i]), int(r[i]), int(m[i])) == orc.partial_search(reads[i])
    hits = want[want >= (1 << 31)][:50]
    assert idx.get_kmers(hits[:8]) == [orc.get_kmer(int(x)) for x in hits[:8]]
    ses.close()
    idx.close()

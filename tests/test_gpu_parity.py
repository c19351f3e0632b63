"""Parity of the CUDA path with the reference, through the C ABI, on a real GPU.

Expected values come from (a) the committed outputs of the reference's own classes
(tests/golden, MANIFEST.json) and (b) the C oracle on the same seeded inputs; at
BASELINE.json's full sizes, from size-independent properties (streaming == per-k-mer search,
hit counts of planted reads)."""
import json
import os

import numpy as np
import pytest

import oracle
import sbwt_b200 as S
from conftest import c1_expected, c1_reads, golden, parse_expected, read_fasta_reads
from sbwt_b200.testing import build_index, read_sbwt, strip_streaming_support, synth

pytestmark = pytest.mark.gpu
MAN = json.load(open(golden("MANIFEST.json")))


def run_both(index_path, reads, modes=(S.MODE_STREAMING, S.MODE_SEARCH), case_mode=S.CASE_UPPER, **kw):
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(index_path)
    ses = S.Session(idx, max(1, a.size), max(1, len(reads)))
    res = {m: ses.query_host(a, off, m, case_mode).copy() for m in modes}
    ses.close()
    idx.close()
    return res


@pytest.mark.parametrize("name", ["cli_k6", "small_k31", "small_k63_rc", "small_k8_p0"])
def test_golden_fixtures(name, tmp_path):
    if name == "cli_k6":
        expected = open(golden(name, "known_answer.txt"), "rb").read() + open(golden(name, "edge.expected.txt"), "rb").read()
        reads = read_fasta_reads(golden(name, "queries.fna")) + read_fasta_reads(golden(name, "edge.fna"))
    else:
        expected = open(golden(name, "expected.txt"), "rb").read()
        reads = read_fasta_reads(golden(name, "reads.fna"))
    vals, counts = parse_expected(expected)
    res = run_both(golden(name, "index.sbwt"), reads)
    np.testing.assert_array_equal(res[S.MODE_STREAMING], vals)
    np.testing.assert_array_equal(res[S.MODE_SEARCH], vals)
    assert oracle.format_lines(res[S.MODE_STREAMING], counts) == expected
    # the --no-streaming-support flavour: per-k-mer path only, streaming must refuse like the reference
    ns = str(tmp_path / "ns.sbwt")
    strip_streaming_support(golden(name, "index.sbwt"), ns)
    np.testing.assert_array_equal(run_both(ns, reads, modes=(S.MODE_SEARCH,))[S.MODE_SEARCH], vals)
    with pytest.raises(S.SbwtGpuError, match="streaming search support not built"):
        run_both(ns, reads, modes=(S.MODE_STREAMING,))


def test_config1_coli3():
    """BASELINE config 1 (reference output md5 bbb3a7a4...)."""
    expected = c1_expected()
    vals, counts = parse_expected(expected)
    res = run_both(golden("c1", "index.sbwt"), c1_reads())
    np.testing.assert_array_equal(res[S.MODE_STREAMING], vals)
    np.testing.assert_array_equal(res[S.MODE_SEARCH], vals)
    assert oracle.format_lines(res[S.MODE_STREAMING], counts) == expected


def test_index_accessors_and_rank():
    idx = S.Index(golden("c1", "index.sbwt"))
    orc = oracle.OracleIndex(golden("c1", "index.sbwt"))
    assert (idx.n_nodes, idx.n_kmers, idx.k, idx.precalc_k) == (10401756, 10335847, 30, 8)
    assert idx.C_array == [1, 2567588, 5206718, 7835964] == orc.C_array
    assert idx.has_streaming_support and idx.edges_only_at_group_starts
    n = idx.n_nodes
    rng = np.random.default_rng(3)
    pos = np.concatenate([rng.integers(0, n + 1, 5000), [0, 1, 223, 224, 225, 447, 448, n - 1, n]]).astype(np.int64)
    chars = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=pos.size))
    got = idx.rank(pos, chars)
    want = np.array([orc.rank(int(p), chr(c)) for p, c in zip(pos, chars)])
    np.testing.assert_array_equal(got, want)
    with pytest.raises(S.SbwtGpuError, match="out of range"):
        idx.rank(np.array([n + 1]), b"A")


@pytest.mark.parametrize("env", [{"SBWT_B200_FORCE_WIDE": "3"}, {"SBWT_B200_WINDOW": "7"}, {"SBWT_B200_WINDOW": "1", "SBWT_B200_FORCE_WIDE": "1"}])
def test_wide_layout_and_window_splitting(env, monkeypatch):
    """The > 2^32-column layout (superblock bases) and the splitting of reads into windows, forced on small data."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    for name in ("small_k31", "small_k63_rc"):
        vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
        res = run_both(golden(name, "index.sbwt"), read_fasta_reads(golden(name, "reads.fna")))
        np.testing.assert_array_equal(res[S.MODE_STREAMING], vals)
        np.testing.assert_array_equal(res[S.MODE_SEARCH], vals)


def test_chunked_host_pipeline_and_empty_inputs():
    """Host batches larger than the session capacity are chunked; ragged / empty batches."""
    name = "small_k31"
    vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden(name, "index.sbwt"))
    ses = S.Session(idx, max_bases=max(len(r) for r in reads) + 100, max_reads=7)
    for pinned in (False, True):
        if pinned:
            pa, po = S.pinned_empty(a.size, np.uint8), S.pinned_empty(off.size, np.int64)
            pa[:], po[:] = a, off
            out = S.pinned_empty(vals.size, np.int64)
            got = ses.query_host(pa, po, S.MODE_STREAMING, out=out)
        else:
            got = ses.query_host(a, off, S.MODE_STREAMING)
        np.testing.assert_array_equal(got, vals)
    # offsets that do not start at zero (a slice of a larger batch)
    cut = 100
    sub = ses.query_host(a, off[cut:].copy(), S.MODE_SEARCH)
    np.testing.assert_array_equal(sub, vals[ses.count_outputs(off[:cut + 1]):])
    # no reads, and reads that are all shorter than k
    assert ses.query_host(a, np.zeros(1, np.int64), S.MODE_STREAMING).size == 0
    short = [b"ACGT", b"", b"A" * 30]
    sa, so = synth.ragged_to_batch(short)
    assert ses.query_host(sa, so, S.MODE_STREAMING).size == 0
    with pytest.raises(S.SbwtGpuError, match="longer than the session capacity"):
        la, lo = synth.ragged_to_batch([b"A" * (ses.index.k + 5000 + max(len(r) for r in reads))])
        ses.query_host(la, lo, S.MODE_SEARCH)


def test_case_exact_mode_matches_search_api():
    """CASE_EXACT = SBWT::search() on the caller's raw bytes: lower case is a miss (SBWT.hh:427)."""
    name = "small_k31"
    raw = []
    cur = None
    for line in open(golden(name, "reads.fna"), "rb").read().split(b"\n"):
        if line.startswith(b">"):
            if cur is not None:
                raw.append(cur)
            cur = b""
        elif cur is not None:
            cur += line
    raw.append(cur)
    a, off = synth.ragged_to_batch(raw)
    orc = oracle.OracleIndex(golden(name, "index.sbwt"))
    want = orc.query_batch(a, off, streaming=False)
    got = run_both(golden(name, "index.sbwt"), raw, modes=(S.MODE_SEARCH,), case_mode=S.CASE_EXACT)[S.MODE_SEARCH]
    np.testing.assert_array_equal(got, want)
    assert (want != parse_expected(open(golden(name, "expected.txt"), "rb").read())[0]).any()  # the fixture has lower case


@pytest.mark.parametrize("name", ["small_k31", "small_k8_p0", "small_k63_rc"])
def test_case_api_mode_matches_the_reference_on_mixed_case(name):
    """CASE_API = the direct API on raw bytes: MODE_STREAMING reproduces SBWT::streaming_search(const char*, len) on
    mixed-case reads (a lower-case base is a miss in a from-scratch search, accepted as the new character of a streaming
    step), MODE_SEARCH the search() loop; expected vectors written by the reference's own methods (`sbwt_ref api`).
    Dense, int32 and hits-only results, whole-batch and chunked sessions."""
    reads = open(golden(name, "mixed_case.txt"), "rb").read().split(b"\n")[:-1]
    want = {}
    for kind in ("streaming", "search"):
        rows = open(golden(name, f"mixed_case.{kind}.txt"), "rb").read().split(b"\n")[:-1]
        want[kind] = np.array([int(x) for row in rows for x in row.split()], dtype=np.int64)
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden(name, "index.sbwt"))
    for max_bases, max_reads in ((a.size, len(reads)), (max(len(r) for r in reads) * 3, 7)):
        ses = S.Session(idx, max_bases, max_reads)
        np.testing.assert_array_equal(ses.query_host(a, off, S.MODE_STREAMING, S.CASE_API), want["streaming"])
        np.testing.assert_array_equal(ses.query_host(a, off, S.MODE_SEARCH, S.CASE_API), want["search"])
        np.testing.assert_array_equal(ses.query_host(a, off, S.MODE_SEARCH, S.CASE_EXACT), want["search"])
        np.testing.assert_array_equal(ses.query_host_i32(a, off, S.MODE_STREAMING, S.CASE_API).astype(np.int64), want["streaming"])
        mask, hits, n = ses.query_host_hits(a, off, S.MODE_STREAMING, S.CASE_API)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: want["streaming"].size].astype(bool)
        got = np.full(want["streaming"].size, -1, dtype=np.int64)
        got[bits] = hits
        np.testing.assert_array_equal(got, want["streaming"])
        ses.close()
    idx.close()


def test_index_create_from_arrays_and_violated_invariant(tmp_path):
    """sbwt_gpu_index_create (the SBWT(A,C,G,T,...) constructor path) and the literal walk-back:
    clearing suffix-group marks makes columns with edges non-starts, so the one-sector shortcut
    must be abandoned; the reference semantics (oracle) is the judge."""
    name = "small_k31"
    d = read_sbwt(golden(name, "index.sbwt"))
    n = d["n_nodes"]
    sgs = d["sgs"][1].copy()
    rng = np.random.default_rng(5)
    for col in rng.integers(1, n, 4000):
        sgs[col >> 6] &= ~np.uint64(1 << (int(col) & 63))
    arrays = dict(bits=[w for _, w in d["bits"]], sgs=sgs, C=d["C"], precalc=d["precalc"].reshape(-1), precalc_k=d["precalc_k"],
                  n_nodes=n, n_kmers=d["n_kmers"], k=d["k"])
    # write the modified index so the oracle can load it (same layout, rank supports unchanged)
    raw = np.fromfile(golden(name, "index.sbwt"), dtype=np.uint8)
    tail = 8 + 32 + 8 + d["precalc"].size * 8 + 32
    start = raw.size - tail - sgs.size * 8
    mod = raw.copy()
    mod[start:start + sgs.size * 8] = sgs.view(np.uint8)
    p = str(tmp_path / "mod.sbwt")
    mod.tofile(p)
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    want = oracle.OracleIndex(p).query_batch(a, off, streaming=True)
    idx = S.Index(arrays=arrays)
    assert not idx.edges_only_at_group_starts
    ses = S.Session(idx, a.size, len(reads))
    np.testing.assert_array_equal(ses.query_host(a, off, S.MODE_STREAMING), want)
    # reads far longer than a work-item window (256 k-mers): on such an index the streaming answers depend on the previous
    # k-mer across what would be a window boundary (SBWT.hh:556-576), so the batch is planned one item per read
    ref = np.frombuffer(b"".join(read_fasta_reads(golden(name, "input.fna"))), dtype=np.uint8)
    long_reads = []
    for i in range(12):
        o = int(rng.integers(0, ref.size - 2500))
        r = ref[o:o + int(rng.integers(600, 2400))].copy()
        for q in rng.integers(0, r.size, size=int(rng.integers(0, 5))):
            r[int(q)] = ord("ACGTN"[int(rng.integers(0, 5))])
        long_reads.append(bytes(r))
    a2, off2 = synth.ragged_to_batch(long_reads)
    want2 = oracle.OracleIndex(p).query_batch(a2, off2, streaming=True)
    ses2 = S.Session(idx, a2.size, len(long_reads))
    np.testing.assert_array_equal(ses2.query_host(a2, off2, S.MODE_STREAMING), want2)
    ses2.close()
    idx2 = S.Index(p)
    assert not idx2.edges_only_at_group_starts
    # an inconsistent C array is rejected
    bad = dict(arrays)
    bad["C"] = d["C"] + np.array([0, 1, 0, 0])
    with pytest.raises(S.SbwtGpuError, match="C array does not match"):
        S.Index(arrays=bad)


def test_device_buffers_and_counters():
    """sbwt_gpu_query_device on torch-owned device memory, plus the rank-op / sector accounting."""
    import torch
    name = "small_k31"
    vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden(name, "index.sbwt"))
    ses = S.Session(idx, a.size, len(reads))
    da, do = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
    out = torch.full((vals.size,), -7, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
        out.fill_(-7)
        ses.query_device(da.data_ptr(), do.data_ptr(), len(reads), a.size, mode, out.data_ptr(), vals.size, st)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), vals)
        stats = ses.query_device_counted(da.data_ptr(), do.data_ptr(), len(reads), a.size, mode, out.data_ptr(), vals.size, st)
        assert stats.lookups == vals.size and stats.hits == int((vals >= 0).sum())
        assert stats.rank_ops > 0 and stats.rank_ops % 2 == 0 and 0 < stats.index_sectors <= stats.rank_ops + stats.lookups
        np.testing.assert_array_equal(out.cpu().numpy(), vals)
    # per-k-mer search does at least as many rank ops as streaming
    s1 = ses.query_device_counted(da.data_ptr(), do.data_ptr(), len(reads), a.size, S.MODE_SEARCH, out.data_ptr(), vals.size, st)
    s2 = ses.query_device_counted(da.data_ptr(), do.data_ptr(), len(reads), a.size, S.MODE_STREAMING, out.data_ptr(), vals.size, st)
    assert s1.rank_ops >= s2.rank_ops


def test_random_differential_vs_oracle(tmp_path):
    """Fresh seeded cases (several k, p, +-RC): GPU == C oracle on hits, misses, Ns, ragged lengths."""
    for seed, (k, p, rc) in enumerate([(31, 8, False), (21, 5, True), (32, 8, False), (33, 6, True), (64, 8, False), (12, 12, True), (5, 2, False)]):
        ref = synth.random_contigs(3, 6000, seed=100 + seed)
        fa = str(tmp_path / f"r{seed}.fna")
        synth.write_fasta(fa, [ref[i] for i in range(3)])
        ix = str(tmp_path / f"i{seed}.sbwt")
        build_index(fa, ix, k=k, precalc=p, add_rc=rc)
        rng = np.random.default_rng(200 + seed)
        reads = []
        for i in range(700):
            L = int(rng.integers(1, 5 * k))
            if i % 3 == 0:
                s = synth.LUT[rng.integers(0, 4, size=L, dtype=np.uint8)]
            else:
                L = min(L, 6000)
                o = int(rng.integers(0, 6000 - L + 1))
                s = ref[int(rng.integers(0, 3)), o:o + L].copy()
                if i % 5 == 0 and L:
                    s[int(rng.integers(0, L))] = ord("N")
                if i % 7 == 0 and L:
                    s[int(rng.integers(0, L))] = synth.LUT[int(rng.integers(0, 4))]
            reads.append(bytes(s))
        a, off = synth.ragged_to_batch(reads)
        orc = oracle.OracleIndex(ix)
        want = orc.query_batch(a, off, streaming=True)
        np.testing.assert_array_equal(orc.query_batch(a, off, streaming=False), want)
        res = run_both(ix, reads)
        np.testing.assert_array_equal(res[S.MODE_STREAMING], want, err_msg=f"k={k} p={p}")
        np.testing.assert_array_equal(res[S.MODE_SEARCH], want, err_msg=f"k={k} p={p}")


def test_search_table_lengths(tmp_path):
    """The runtime search table (longer than the file's precalc table, built on the device) never
    changes a result; a file whose own table does not follow from its bit vectors is used verbatim."""
    for name in ("small_k31", "small_k8_p0", "small_k63_rc"):
        vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
        reads = read_fasta_reads(golden(name, "reads.fna"))
        a, off = synth.ragged_to_batch(reads)
        idx = S.Index(golden(name, "index.sbwt"))
        file_p = idx.precalc_k
        assert idx.table_length >= file_p
        np.testing.assert_array_equal(idx.precalc().reshape(-1), read_sbwt(golden(name, "index.sbwt"))["precalc"].reshape(-1))
        ses = S.Session(idx, a.size, len(reads))
        for tp in (0, 1, file_p, 5, 8, 9, 11, 12) + ((13, 14, 15, 16) if name == "small_k31" else ()):  # (16 characters: 34 GB of rows)
            idx.set_table_length(tp)
            assert idx.table_length == min(tp, idx.k)
            for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
                np.testing.assert_array_equal(ses.query_host(a, off, mode), vals, err_msg=f"{name} tp={tp}")
    # corrupt one row of the file's table: the reference would follow it, so must we (and no other length is allowed)
    d = read_sbwt(golden("small_k31", "index.sbwt"))
    raw = np.fromfile(golden("small_k31", "index.sbwt"), dtype=np.uint8)
    pre = d["precalc"].copy()
    reads = read_fasta_reads(golden("small_k31", "reads.fna"))
    vals, lens = parse_expected(open(golden("small_k31", "expected.txt"), "rb").read())
    # a row some from-scratch search really uses: the first 8 characters of a read whose first k-mer is found
    first = np.concatenate([[0], np.cumsum(lens)])[:-1]
    ri = next(i for i, r in enumerate(reads) if lens[i] > 0 and vals[first[i]] >= 0 and set(r[:8].upper()) <= set(b"ACGT"))
    row = sum(b"ACGT".index(ch) << (2 * j) for j, ch in enumerate(reads[ri][:8].upper()))
    assert pre[row, 0] >= 0
    pre[row] = (-1, -1)
    start = raw.size - 32 - pre.size * 8
    raw[start:start + pre.size * 8] = pre.reshape(-1).view(np.uint8)
    p = str(tmp_path / "modtable.sbwt")
    raw.tofile(p)
    a, off = synth.ragged_to_batch(reads)
    want = oracle.OracleIndex(p).query_batch(a, off, streaming=True)
    assert (want != vals).any()
    idx = S.Index(p)
    assert idx.table_length == 8
    np.testing.assert_array_equal(S.Session(idx, a.size, len(reads)).query_host(a, off, S.MODE_STREAMING), want)
    with pytest.raises(S.SbwtGpuError, match="does not follow from its bit vectors"):
        idx.set_table_length(10)


@pytest.mark.parametrize("name", ["long_k80", "long_k255_rc"])
@pytest.mark.parametrize("wide", [False, True])
def test_k_above_64(name, wide, tmp_path, monkeypatch):
    """k > 64 (the reference's search has no limit): indexes built and answered by the reference itself with k = 80 and
    k = 255 (tests/golden/make_golden.py --only-long-k), through the literal kernels of long_kmer_kernels.cuh: both modes,
    dense / int32 / hits-only results, the device formatter's text, a chunked session, the forced-wide layout, and the
    per-k-mer path alone on a file without streaming support."""
    if wide:
        monkeypatch.setenv("SBWT_B200_FORCE_WIDE", "3")
    expected = open(golden(name, "expected.txt"), "rb").read()
    vals, counts = parse_expected(expected)
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden(name, "index.sbwt"))
    assert idx.k == MAN[name]["k"] > 64
    for max_bases, max_reads in ((a.size, len(reads)), (max(len(r) for r in reads) * 2, 5)):
        ses = S.Session(idx, max_bases, max_reads)
        for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
            np.testing.assert_array_equal(ses.query_host(a, off, mode), vals)
        if not wide:
            np.testing.assert_array_equal(ses.query_host_i32(a, off, S.MODE_STREAMING).astype(np.int64), vals)
            mask, hits, n = ses.query_host_hits(a, off, S.MODE_SEARCH)
            bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: vals.size].astype(bool)
            assert n == int((vals >= 0).sum()) and np.array_equal(bits, vals >= 0) and np.array_equal(hits, vals[vals >= 0])
        text, n_lookups = ses.query_host_text(a, off, S.MODE_STREAMING)
        assert text == expected and n_lookups == vals.size
        ses.close()
    idx.close()
    ns = str(tmp_path / "ns.sbwt")
    strip_streaming_support(golden(name, "index.sbwt"), ns)
    np.testing.assert_array_equal(run_both(ns, reads, modes=(S.MODE_SEARCH,))[S.MODE_SEARCH], vals)
    with pytest.raises(S.SbwtGpuError, match="streaming search support not built"):
        run_both(ns, reads, modes=(S.MODE_STREAMING,))


def test_scale_properties_config2_like(tmp_path):
    """At a BASELINE-like shape (random reference, 150-bp reads, 50 % planted): streaming == per-k-mer
    search on the GPU, planted reads hit on all 120 k-mers, random reads miss, and a sample agrees with the oracle."""
    ref = synth.random_contigs(20, 200_000, seed=42)
    raw = str(tmp_path / "ref.txt")
    with open(raw, "wb") as f:
        for i in range(ref.shape[0]):
            f.write(ref[i].tobytes() + b"\n")
    ix = str(tmp_path / "c2.sbwt")
    info = build_index(raw, ix, k=31, precalc=8, raw=True)
    assert info["n_kmers"] == 20 * (200_000 - 30)
    reads = synth.sample_reads(ref, 200_000, 150, 0.5, seed=43)
    a, off = synth.matrix_to_batch(reads)
    idx = S.Index(ix)
    ses = S.Session(idx, a.size, reads.shape[0])
    s = ses.query_host(a, off, S.MODE_STREAMING)
    q = ses.query_host(a, off, S.MODE_SEARCH)
    np.testing.assert_array_equal(s, q)
    per_read = (s.reshape(-1, 120) >= 0).sum(axis=1)
    assert set(np.unique(per_read)) <= {0, 120}
    assert 0.49 < (per_read == 120).mean() < 0.51
    assert s[s >= 0].min() >= 1 and s.max() < idx.n_nodes
    orc = oracle.OracleIndex(ix)
    m = 3000
    np.testing.assert_array_equal(orc.query_batch(a[:m * 150], off[:m + 1], streaming=True), s[:m * 120])


def test_int32_results_match_int64():
    """sbwt_gpu_query_host_i32 / query_device_i32: the same values in half the bytes (n_nodes < 2^31)."""
    import torch
    for name in ("small_k31", "small_k63_rc"):
        vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
        reads = read_fasta_reads(golden(name, "reads.fna"))
        a, off = synth.ragged_to_batch(reads)
        idx = S.Index(golden(name, "index.sbwt"))
        ses = S.Session(idx, max_bases=5000, max_reads=50)  # several chunks through the pipeline
        for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
            got = ses.query_host_i32(a, off, mode)
            assert got.dtype == np.int32
            np.testing.assert_array_equal(got.astype(np.int64), vals)
        ses2 = S.Session(idx, a.size, len(reads))
        da, do = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
        out = torch.full((vals.size,), -7, dtype=torch.int32, device="cuda")
        ses2.query_device_i32(da.data_ptr(), do.data_ptr(), len(reads), a.size, S.MODE_STREAMING, out.data_ptr(), vals.size,
                              torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy().astype(np.int64), vals)


def test_int32_results_refused_on_wide_index(monkeypatch):
    monkeypatch.setenv("SBWT_B200_FORCE_WIDE", "3")
    reads = read_fasta_reads(golden("small_k31", "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden("small_k31", "index.sbwt"))
    ses = S.Session(idx, a.size, len(reads))
    with pytest.raises(S.SbwtGpuError, match="int32 results need"):
        ses.query_host_i32(a, off, S.MODE_SEARCH)


def test_long_reads_and_mixed_runs(tmp_path):
    """Reads far longer than a work-item window, with planted stretches, substitutions and Ns in between:
    every switch between streaming and from-scratch search, across window boundaries."""
    ref = synth.random_contigs(2, 50_000, seed=7)
    fa = str(tmp_path / "r.fna")
    synth.write_fasta(fa, [ref[i] for i in range(2)])
    ix = str(tmp_path / "i.sbwt")
    build_index(fa, ix, k=31, precalc=8, add_rc=True)
    rng = np.random.default_rng(11)
    reads = []
    for i in range(60):
        parts = []
        for _ in range(int(rng.integers(1, 12))):
            L = int(rng.integers(1, 900))
            if rng.random() < 0.6:
                o = int(rng.integers(0, 50_000 - L))
                seg = ref[int(rng.integers(0, 2)), o:o + L].copy()
                for _ in range(int(rng.integers(0, 3))):
                    seg[int(rng.integers(0, L))] = synth.LUT[int(rng.integers(0, 4))]
            else:
                seg = synth.LUT[rng.integers(0, 4, size=L, dtype=np.uint8)]
            if rng.random() < 0.2:
                seg[int(rng.integers(0, L))] = ord("N")
            parts.append(seg)
        reads.append(bytes(np.concatenate(parts)))
    a, off = synth.ragged_to_batch(reads)
    want = oracle.OracleIndex(ix).query_batch(a, off, streaming=True)
    assert 0.2 < (want >= 0).mean() < 0.8
    res = run_both(ix, reads)
    np.testing.assert_array_equal(res[S.MODE_STREAMING], want)
    np.testing.assert_array_equal(res[S.MODE_SEARCH], want)


@pytest.mark.parametrize("stride", ["0", "1", "2", "5", "9", "40"])
def test_probe_strides_give_identical_results(stride, tmp_path, monkeypatch):
    """The streaming walk answers runs of misses by probing every D-th k-mer (walk2_kernel PROBE). Any stride,
    and no probing at all (0), must give the reference's answers: golden fixtures, config 1, and long reads with
    planted stretches, substitutions every few bases and Ns (restarts, partial coverage, ranges of > 32 segments),
    with small work-item windows as well."""
    monkeypatch.setenv("SBWT_B200_PROBE", stride)
    for name in ("small_k31", "small_k63_rc", "cli_k6"):
        expected = open(golden(name, "known_answer.txt" if name == "cli_k6" else "expected.txt"), "rb").read()
        reads = read_fasta_reads(golden(name, "queries.fna" if name == "cli_k6" else "reads.fna"))
        vals, _ = parse_expected(expected)
        np.testing.assert_array_equal(run_both(golden(name, "index.sbwt"), reads, modes=(S.MODE_STREAMING,))[S.MODE_STREAMING], vals)
    vals, _ = parse_expected(c1_expected())
    np.testing.assert_array_equal(run_both(golden("c1", "index.sbwt"), c1_reads(), modes=(S.MODE_STREAMING,))[S.MODE_STREAMING], vals)

    ref = synth.random_contigs(2, 60_000, seed=17)
    fa = str(tmp_path / "r.fna")
    synth.write_fasta(fa, [ref[i] for i in range(2)])
    ix = str(tmp_path / "i.sbwt")
    build_index(fa, ix, k=31, precalc=8, add_rc=False)
    rng = np.random.default_rng(int(stride) + 100)
    reads = []
    for i in range(80):
        parts = []
        for _ in range(int(rng.integers(1, 8))):
            L = int(rng.integers(1, 1500))
            kind = rng.random()
            if kind < 0.55:  # a stretch of the reference with substitutions at a random density
                o = int(rng.integers(0, 60_000 - L))
                seg = ref[int(rng.integers(0, 2)), o:o + L].copy()
                gap = int(rng.integers(3, 120))
                for pos in range(int(rng.integers(0, gap)), L, gap):
                    seg[pos] = synth.LUT[(int(np.where(synth.LUT == seg[pos])[0][0]) + 1) & 3]
            else:
                seg = synth.LUT[rng.integers(0, 4, size=L, dtype=np.uint8)]
            if rng.random() < 0.25:
                seg[int(rng.integers(0, L))] = ord("N")
            parts.append(seg)
        reads.append(bytes(np.concatenate(parts)))
    a, off = synth.ragged_to_batch(reads)
    want = oracle.OracleIndex(ix).query_batch(a, off, streaming=True)
    assert 0.05 < (want >= 0).mean() < 0.9
    for window in (None, "33", "1"):
        if window:
            monkeypatch.setenv("SBWT_B200_WINDOW", window)
        got = run_both(ix, reads, modes=(S.MODE_STREAMING,))[S.MODE_STREAMING]
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("wire", ["dense", "sparse"])
@pytest.mark.parametrize("threads", ["0", "1", "5"])
def test_result_wire_formats_of_the_host_pipeline(threads, wire, monkeypatch):
    """sbwt_gpu_query_host returns int64 either copied as such (SBWT_B200_WIDEN_THREADS=0), or copied as int32 and
    sign-extended by host threads (dense), or as hit masks + hits only, rebuilt by host threads (sparse; host_widen.hpp).
    Same values every way, int64 and int32 API, pinned or pageable buffers, many chunks in flight, chunks that are all
    hits, all misses, empty, and destinations off the 32-byte grid."""
    monkeypatch.setenv("SBWT_B200_WIDEN_THREADS", threads)
    monkeypatch.setenv("SBWT_B200_WIRE", wire)
    for name in ("small_k31", "small_k63_rc"):
        vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
        reads = read_fasta_reads(golden(name, "reads.fna"))
        a, off = synth.ragged_to_batch(reads)
        idx = S.Index(golden(name, "index.sbwt"))
        ses = S.Session(idx, max_bases=3000, max_reads=11)  # dozens of chunks over the three pipeline slots
        for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
            np.testing.assert_array_equal(ses.query_host(a, off, mode), vals)
            np.testing.assert_array_equal(ses.query_host_i32(a, off, mode).astype(np.int64), vals)
            out = S.pinned_empty(vals.size + 1, np.int64)
            out[:] = -7
            got = ses.query_host(a, off, mode, out=out[1:])  # destination off the 32-byte grid
            np.testing.assert_array_equal(got, vals)
            assert out[0] == -7
            out32 = S.pinned_empty(vals.size + 3, np.int32)
            out32[:] = -7
            got = ses.query_host_i32(a, off, mode, out=out32[3:])
            np.testing.assert_array_equal(got.astype(np.int64), vals)
            assert (out32[:3] == -7).all()
        ses.close()
        # larger chunks: several 4096-result blocks per chunk, runs of found reads, absent reads and reads too short for a k-mer
        rng = np.random.default_rng(21)
        big = []
        for i in range(600):
            r = reads[int(rng.integers(0, len(reads)))]
            kind = rng.random()
            big.append(r if kind < 0.5 else (bytes(synth.LUT[rng.integers(0, 4, size=len(r), dtype=np.uint8)]) if kind < 0.9 else r[:5]))
        a2, off2 = synth.ragged_to_batch(big)
        want = oracle.OracleIndex(golden(name, "index.sbwt")).query_batch(a2, off2, streaming=True)
        ses = S.Session(idx, max_bases=40_000, max_reads=1000)
        for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
            np.testing.assert_array_equal(ses.query_host(a2, off2, mode), want)
            np.testing.assert_array_equal(ses.query_host_i32(a2, off2, mode).astype(np.int64), want)
        ses.close()
        idx.close()


@pytest.mark.parametrize("layout", ["c64", "c96"])
@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_compact_one_hot_layout_gives_identical_results(mode, layout, tmp_path, monkeypatch):
    """The compact layout (two bits per column, device_index.cuh) is a pure representation choice: off (0), automatic
    (1) and forced (2, every index, so that indexes with many flagged blocks such as config 1 mix csector answers and
    classic-sector fallbacks inside one warp) must all reproduce the reference's output; int32 results and the
    counted instantiation included."""
    if mode == "0" and layout == "c96":
        pytest.skip("no compact layout: covered by the c64 leg")
    monkeypatch.setenv("SBWT_B200_COMPACT", mode)
    monkeypatch.setenv("SBWT_B200_LAYOUT", layout)  # csector format: 64 columns with absolute counts / 96 with relative ones
    for name in ("small_k31", "small_k63_rc", "cli_k6", "small_k8_p0"):
        expected = open(golden(name, "known_answer.txt" if name == "cli_k6" else "expected.txt"), "rb").read()
        reads = read_fasta_reads(golden(name, "queries.fna" if name == "cli_k6" else "reads.fna"))
        vals, _ = parse_expected(expected)
        res = run_both(golden(name, "index.sbwt"), reads)
        np.testing.assert_array_equal(res[S.MODE_STREAMING], vals)
        np.testing.assert_array_equal(res[S.MODE_SEARCH], vals)
    idx = S.Index(golden("small_k31", "index.sbwt"))
    used, flagged = idx.compact_layout
    assert used == (mode != "0") and (flagged < 0 if mode == "0" else 0 <= flagged < 0.05)
    idx.close()
    idx = S.Index(golden("c1", "index.sbwt"))
    used, flagged = idx.compact_layout
    assert used == (mode == "2") and (mode == "0" or 0.2 < flagged < 0.6)  # 39 % of coli3's blocks hold a branching or dead-end column
    reads = c1_reads()
    a, off = synth.ragged_to_batch(reads)
    vals, _ = parse_expected(c1_expected())
    ses = S.Session(idx, a.size, len(reads))
    for m in (S.MODE_STREAMING, S.MODE_SEARCH):
        np.testing.assert_array_equal(ses.query_host(a, off, m), vals)
        np.testing.assert_array_equal(ses.query_host_i32(a, off, m).astype(np.int64), vals)
    import torch
    d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
    d_out = torch.empty(vals.size, dtype=torch.int64, device="cuda")
    st = ses.query_device_counted(d_a.data_ptr(), d_off.data_ptr(), len(reads), a.size, S.MODE_STREAMING, d_out.data_ptr(), vals.size)
    np.testing.assert_array_equal(d_out.cpu().numpy(), vals)
    assert st.lookups == vals.size and st.hits == int((vals >= 0).sum())
    ses.close()
    idx.close()
    # long reads with planted stretches, substitutions and Ns on a +RC index (restarts, walk-backs, TODO ranges)
    ref = synth.random_contigs(2, 50_000, seed=7)
    fa = str(tmp_path / "r.fna")
    synth.write_fasta(fa, [ref[i] for i in range(2)])
    ix = str(tmp_path / "i.sbwt")
    build_index(fa, ix, k=31, precalc=8, add_rc=True)
    rng = np.random.default_rng(12)
    reads = []
    for i in range(60):
        parts = []
        for _ in range(int(rng.integers(1, 10))):
            L = int(rng.integers(1, 900))
            if rng.random() < 0.6:
                o = int(rng.integers(0, 50_000 - L))
                seg = ref[int(rng.integers(0, 2)), o:o + L].copy()
                for _ in range(int(rng.integers(0, 3))):
                    seg[int(rng.integers(0, L))] = synth.LUT[int(rng.integers(0, 4))]
            else:
                seg = synth.LUT[rng.integers(0, 4, size=L, dtype=np.uint8)]
            if rng.random() < 0.2:
                seg[int(rng.integers(0, L))] = ord("N")
            parts.append(seg)
        reads.append(bytes(np.concatenate(parts)))
    a, off = synth.ragged_to_batch(reads)
    want = oracle.OracleIndex(ix).query_batch(a, off, streaming=True)
    res = run_both(ix, reads)
    np.testing.assert_array_equal(res[S.MODE_STREAMING], want)
    np.testing.assert_array_equal(res[S.MODE_SEARCH], want)


def test_sharded_query_over_several_sessions():
    """sbwt_gpu_query_host_sharded: one batch cut into contiguous ranges of equal base count, one host thread per session
    (on a multi-GPU box one session per device; here replicas on whatever devices exist, several per device if need be),
    results in order in the caller's array -- equal to one sbwt_gpu_query_host call, empty shards and tiny batches included."""
    n_dev = S.device_count()
    for name in ("small_k31", "small_k63_rc"):
        vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
        reads = read_fasta_reads(golden(name, "reads.fna"))
        a, off = synth.ragged_to_batch(reads)
        for n_ses in (1, 2, 3, 5):
            idxs = [S.Index(golden(name, "index.sbwt"), device=i % n_dev) for i in range(n_ses)]
            sess = [S.Session(ix, max_bases=20_000, max_reads=500) for ix in idxs]
            for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
                np.testing.assert_array_equal(S.query_host_sharded(sess, a, off, mode), vals)
            # fewer reads than sessions: some shards are empty
            a1, off1 = synth.ragged_to_batch(reads[:2])
            want1 = sess[0].query_host(a1, off1, S.MODE_STREAMING).copy()
            np.testing.assert_array_equal(S.query_host_sharded(sess, a1, off1, S.MODE_STREAMING), want1)
            for s_ in sess:
                s_.close()
            for ix in idxs:
                ix.close()
    idx = S.Index(golden("small_k31", "index.sbwt"))
    ses = S.Session(idx, 1000, 10)
    with pytest.raises(S.SbwtGpuError, match="listed twice"):
        S.query_host_sharded([ses, ses], a, off, S.MODE_STREAMING)
    other = S.Index(golden("small_k63_rc", "index.sbwt"))
    ses2 = S.Session(other, 1000, 10)
    with pytest.raises(S.SbwtGpuError, match="replicas of one index"):
        S.query_host_sharded([ses, ses2], a, off, S.MODE_STREAMING)


@pytest.mark.parametrize("name", ["cli_k6", "small_k31", "small_k63_rc", "small_k8_p0"])
def test_other_read_only_queries(name, monkeypatch):
    """The batched device forms of partial_search, update_sbwt_interval, forward, contains, get_kmer and
    ascii_export_sets (SBWT.hh:369-381, 423-437, 526-537, 701-773; SubsetMatrixRank.hh:39-48) against the reference's
    own answers (f4.json) and, on random inputs, the C oracle; narrow and forced-wide layouts."""
    f4 = json.load(open(golden(name, "f4.json")))
    orc = oracle.OracleIndex(golden(name, "index.sbwt"))
    reads = read_fasta_reads(golden(name, f4["queries"]))
    for wide in (False, True):
        if wide:
            monkeypatch.setenv("SBWT_B200_FORCE_WIDE", "2")
        idx = S.Index(golden(name, "index.sbwt"))
        a, off = synth.ragged_to_batch(reads)
        l, r, m = idx.partial_search(a, off)
        assert np.stack([l, r, m], axis=1).tolist() == f4["partial_search"]
        fw = f4["forward"]
        np.testing.assert_array_equal(idx.forward(fw["nodes"], fw["chars"].encode()), fw["out"])
        gk = f4["get_kmer"]
        assert [x.decode() for x in idx.get_kmers(gk["ranks"])] == gk["kmers"]
        text = idx.ascii_export_sets()
        assert text == orc.export_sets() and len(text) == f4["export_len"]
        # random inputs against the oracle
        rng = np.random.default_rng(5)
        n = orc.n_nodes
        pos = rng.integers(0, n, size=2000)
        chars = bytes(rng.choice(np.frombuffer(b"ACGTNa", np.uint8), size=pos.size))
        np.testing.assert_array_equal(idx.contains(pos, chars), [orc.contains(int(p), chr(c)) for p, c in zip(pos, chars)])
        np.testing.assert_array_equal(idx.forward(pos, chars), [orc.forward(int(p), chr(c)) for p, c in zip(pos, chars)])
        # update_sbwt_interval: from the full interval, from intervals reached by a prefix, and {-1,-1} passing through
        strs = [bytes(x) for x in reads[:200]] + [b"", b"ACGN", b"acgt"]
        a2, off2 = synth.ragged_to_batch(strs)
        l0 = np.zeros(len(strs), dtype=np.int64)
        r0 = np.full(len(strs), n - 1, dtype=np.int64)
        l0[5], r0[5] = -1, -1
        gl, gr = idx.update_interval(a2, off2, l0, r0)
        want = [orc.update_interval(s_, int(a_), int(b_)) for s_, a_, b_ in zip(strs, l0, r0)]
        assert list(zip(gl.tolist(), gr.tolist())) == want
        half = [s_[: len(s_) // 2] for s_ in strs]
        rest = [s_[len(s_) // 2:] for s_ in strs]
        a3, off3 = synth.ragged_to_batch(half)
        hl, hr = idx.update_interval(a3, off3, l0, r0)
        a4, off4 = synth.ragged_to_batch(rest)
        fl, fr = idx.update_interval(a4, off4, hl, hr)
        assert list(zip(fl.tolist(), fr.tolist())) == want
        with pytest.raises(S.SbwtGpuError, match="out of range"):
            idx.get_kmers([n])
        idx.close()


def test_forward_needs_streaming_support():
    idx = S.Index(golden("cli_k6", "index_nostream.sbwt"))
    with pytest.raises(S.SbwtGpuError, match="Streaming support required"):
        idx.forward([1], b"A")
    idx.close()


def test_hits_only_results():
    """sbwt_gpu_query_host_hits: membership bitmap over the whole batch + the found values in order reproduce the dense
    result array of the oracle; chunked sessions (chunk boundaries inside mask words), pinned and pageable buffers,
    bitmap alone, both modes, and the bitmap alone on an index that is forced wide."""
    vals, _ = parse_expected(c1_expected())
    reads = c1_reads()
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden("c1", "index.sbwt"))
    n_out = vals.size

    def expand(mask, hits):
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n_out].astype(bool)
        out = np.full(n_out, -1, dtype=np.int64)
        out[bits] = hits
        return out

    for max_bases, max_reads in ((a.size, len(reads)), (7001, 13), (100_000, 1000)):
        ses = S.Session(idx, max_bases, max_reads)
        for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
            mask, hits, n = ses.query_host_hits(a, off, mode)
            assert n == int((vals >= 0).sum())
            np.testing.assert_array_equal(expand(mask, hits), vals)
        mask2, none, n2 = ses.query_host_hits(a, off, S.MODE_STREAMING, want_hits=False)
        assert none is None and n2 == n and np.array_equal(mask2, mask)
        pm, ph = S.pinned_empty(mask.size, np.uint32), S.pinned_empty(n_out, np.int32)
        pm[:] = 0xFFFFFFFF
        mask3, hits3, n3 = ses.query_host_hits(a, off, S.MODE_STREAMING, mask=pm, hits=ph)
        np.testing.assert_array_equal(expand(mask3, hits3), vals)
        ses.close()
    idx.close()


def test_chunks_without_results_and_late_errors():
    """Host batches cut into chunks by a small session: a chunk made of reads shorter than k only (no results, and its
    predecessor ends inside a bitmap word), then reads again; and a read that does not fit the session in a LATER chunk:
    the call fails with the reference-style message, the session stays usable."""
    reads = c1_reads()
    long_reads = [r for r in reads if len(r) >= 60][:11]
    mixed = long_reads[:5] + [b"ACGTACGTAC"] * 5 + [b""] * 2 + long_reads[5:]
    a, off = synth.ragged_to_batch(mixed)
    want = oracle.OracleIndex(golden("c1", "index.sbwt")).query_batch(a, off, streaming=True)
    idx = S.Index(golden("c1", "index.sbwt"))
    ses = S.Session(idx, max(len(r) for r in mixed) * 5, 5)
    n_out = want.size
    assert ses.count_outputs(off[:6]) % 32 != 0
    for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
        np.testing.assert_array_equal(ses.query_host(a, off, mode), want)
        pm = S.pinned_empty((n_out + 31) // 32, np.uint32)
        pm[:] = 0xFFFFFFFF
        mask, hits, n = ses.query_host_hits(a, off, mode, mask=pm)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n_out].astype(bool)
        got = np.full(n_out, -1, dtype=np.int64)
        got[bits] = hits
        np.testing.assert_array_equal(got, want)
        assert n == int((want >= 0).sum())
    too_long = long_reads[:7] + [b"ACGT" * 2000] + long_reads[7:]
    ba, bo = synth.ragged_to_batch(too_long)
    for call in (lambda: ses.query_host(ba, bo, S.MODE_STREAMING), lambda: ses.query_host_hits(ba, bo, S.MODE_STREAMING)):
        with pytest.raises(S.SbwtGpuError, match=r"read 7 \(8000 bases\) is longer than the session capacity"):
            call()
    bad = off.copy()
    bad[8] = bad[7] - 3
    with pytest.raises(S.SbwtGpuError, match="non-decreasing"):
        ses.query_host(a, bad, S.MODE_STREAMING)
    np.testing.assert_array_equal(ses.query_host(a, off, S.MODE_STREAMING), want)
    ses.close()
    idx.close()


def test_hits_only_bitmap_on_wide_index(monkeypatch):
    monkeypatch.setenv("SBWT_B200_FORCE_WIDE", "3")
    name = "small_k31"
    vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden(name, "index.sbwt"))
    ses = S.Session(idx, 5000, 50)
    mask, _, n = ses.query_host_hits(a, off, S.MODE_STREAMING, want_hits=False)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: vals.size].astype(bool)
    np.testing.assert_array_equal(bits, vals >= 0)
    assert n == int((vals >= 0).sum())
    with pytest.raises(S.SbwtGpuError, match="fewer than 2\\^31"):
        ses.query_host_hits(a, off, S.MODE_STREAMING)
    ses.close()
    idx.close()

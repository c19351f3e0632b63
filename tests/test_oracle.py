"""The C oracle (oracle/sbwt_oracle.c) against the reference's golden vector and the
fixtures produced by the reference's own classes (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from conftest import c1_expected, c1_reads, golden, parse_expected, read_fasta_reads
from sbwt_b200.testing import read_sbwt, strip_streaming_support, synth

MAN = json.load(open(golden("MANIFEST.json")))


def test_known_answer_cli_k6(tmp_path):
    """tests/test_CLI.hh:90 -- the reference's only hard-coded known answer for this path."""
    known = open(golden("cli_k6", "known_answer.txt")).read()
    for ix in ("index.sbwt", "index_nostream.sbwt"):
        idx = oracle.OracleIndex(golden("cli_k6", ix))
        for q in ("queries.fna", "queries.fq"):
            out = tmp_path / "o.txt"
            assert idx.search_file(golden("cli_k6", q), str(out)) == 22 + 32 + 8
            assert out.read_text() == known


def test_edge_cases_cli_k6(tmp_path):
    expected = open(golden("cli_k6", "edge.expected.txt")).read()
    for ix in ("index.sbwt", "index_nostream.sbwt"):
        idx = oracle.OracleIndex(golden("cli_k6", ix))
        out = tmp_path / "o.txt"
        idx.search_file(golden("cli_k6", "edge.fna"), str(out))
        assert out.read_text() == expected
    lines = expected.split("\n")
    assert lines[2] == "" and lines[3] == "74 "  # len < k -> empty line; len == k -> one value


@pytest.mark.parametrize("name", ["small_k31", "small_k63_rc", "small_k8_p0"])
def test_small_fixtures(name, tmp_path):
    expected = open(golden(name, "expected.txt"), "rb").read()
    assert hashlib.md5(expected).hexdigest() == MAN[name]["expected_md5"]
    idx = oracle.OracleIndex(golden(name, "index.sbwt"))
    assert idx.k == MAN[name]["k"] and idx.precalc_k == MAN[name]["precalc_k"]
    out = tmp_path / "o.txt"
    idx.search_file(golden(name, "reads.fna"), str(out))
    assert out.read_bytes() == expected
    # batch API, both paths, against the parsed reference output
    vals, counts = parse_expected(expected)
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    assert counts == [max(0, len(r) - idx.k + 1) for r in reads]
    np.testing.assert_array_equal(idx.query_batch(a, off, streaming=True), vals)
    np.testing.assert_array_equal(idx.query_batch(a, off, streaming=False), vals)
    assert oracle.format_lines(vals, counts) == expected
    ns = str(tmp_path / "ns.sbwt")
    strip_streaming_support(golden(name, "index.sbwt"), ns)
    idx2 = oracle.OracleIndex(ns)
    assert not idx2.has_streaming_support
    np.testing.assert_array_equal(idx2.query_batch(a, off, streaming=False), vals)
    with pytest.raises(RuntimeError, match="streaming search support not built"):
        idx2.query_batch(a, off, streaming=True)


def test_config1_coli3():
    """BASELINE config 1: coli3 k=30 x queries.fastq; reference output md5 bbb3a7a4...."""
    expected = c1_expected()
    assert hashlib.md5(expected).hexdigest() == "bbb3a7a497a2617b2d6ac993ac3560f3" == MAN["c1"]["expected_md5"]
    vals, counts = parse_expected(expected)
    assert vals.size == 355000 and int((vals >= 0).sum()) == 277375
    idx = oracle.OracleIndex(golden("c1", "index.sbwt"))
    assert (idx.n_nodes, idx.n_kmers, idx.k, idx.precalc_k) == (10401756, 10335847, 30, 8)
    assert idx.C_array == [1, 2567588, 5206718, 7835964]
    a, off = synth.ragged_to_batch(c1_reads())
    np.testing.assert_array_equal(idx.query_batch(a, off, streaming=True), vals)
    np.testing.assert_array_equal(idx.query_batch(a, off, streaming=False), vals)


def test_rank_directory_equals_definition():
    """rank_support_v5 arithmetic over the file's own directory == popcount of [0,pos)."""
    idx = oracle.OracleIndex(golden("small_k31", "index.sbwt"))
    n = idx.n_nodes
    rng = np.random.default_rng(1)
    pos = np.unique(np.concatenate([rng.integers(0, n + 1, 3000), [0, 1, 63, 64, 65, 383, 384, 385, 2047, 2048, 2049, n - 1, n]]))
    for c in "ACGT":
        for p in pos:
            assert idx.rank(int(p), c) == idx.rank_naive(int(p), c)
    assert idx.rank(n, "N") == 0
    assert idx.C_array[1] - idx.C_array[0] == idx.rank(n, "A")


def test_structural_invariants():
    """SURVEY.md section 8(a) note 7: edges only at suffix-group starts, total ones = n-1, groups <= 4."""
    for name in ("small_k31", "small_k63_rc", "cli_k6"):
        d = read_sbwt(golden(name, "index.sbwt"))
        n = d["n_nodes"]
        bits = [np.unpackbits(w.view(np.uint8), bitorder="little")[:n] for _, w in d["bits"]]
        sgs = np.unpackbits(d["sgs"][1].view(np.uint8), bitorder="little")[:n]
        assert sgs[0] == 1
        anyedge = bits[0] | bits[1] | bits[2] | bits[3]
        assert not (anyedge & (1 - sgs)).any()
        assert sum(int(b.sum()) for b in bits) == n - 1
        starts = np.flatnonzero(sgs)
        assert np.diff(np.append(starts, n)).max() <= 4


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/sbwt_ref not built (needs /root/reference)")
def test_against_live_reference_binary(tmp_path):
    """Differential test vs the reference classes on a fresh random case (build container only)."""
    rng = np.random.default_rng(1234)
    ref = synth.random_contigs(3, 5000, seed=77)
    fa = str(tmp_path / "in.fna")
    synth.write_fasta(fa, [ref[i] for i in range(3)])
    from sbwt_b200.testing import build_index
    ix = str(tmp_path / "i.sbwt")
    build_index(fa, ix, k=21, precalc=5, add_rc=True)
    reads = synth.sample_reads(ref, 500, 80, 0.5, seed=5, both_strands=True)
    reads[::7, 13] = ord("N")
    q = str(tmp_path / "q.fna")
    synth.write_fasta(q, [reads[i] for i in range(reads.shape[0])])
    o1, o2 = str(tmp_path / "o1.txt"), str(tmp_path / "o2.txt")
    oracle.ref_run("search", "-i", ix, "-q", q, "-o", o1)
    oracle.OracleIndex(ix).search_file(q, o2)
    assert open(o1, "rb").read() == open(o2, "rb").read()
    # rank values at random positions
    idx = oracle.OracleIndex(ix)
    pos = rng.integers(0, idx.n_nodes + 1, 200)
    inp = "".join(f"{int(p)} {c}\n" for p in pos for c in "ACGT").encode()
    got = oracle.ref_run("ranks", "-i", ix, stdin=inp).stdout.split()
    want = [idx.rank(int(p), c) for p in pos for c in "ACGT"]
    assert [int(x) for x in got] == want


F4 = ["cli_k6", "small_k31", "small_k63_rc", "small_k8_p0"]


@pytest.mark.parametrize("name", F4)
def test_other_queries_against_the_reference(name):
    """partial_search / forward / get_kmer / ascii_export_sets (SBWT.hh:369-381, 526-537, 701-773): the C restatement
    against the answers of the reference's own methods (tests/golden/<name>/f4.json, written by `sbwt_ref`)."""
    f4 = json.load(open(golden(name, "f4.json")))
    idx = oracle.OracleIndex(golden(name, "index.sbwt"))
    reads = read_fasta_reads(golden(name, f4["queries"]))
    # (SeqIO upper-cases what `sbwt_ref partial` sees; partial_search upper-cases again, so the raw reads agree)
    assert [list(idx.partial_search(r)) for r in reads] == f4["partial_search"]
    fw = f4["forward"]
    assert [idx.forward(n, c) for n, c in zip(fw["nodes"], fw["chars"])] == fw["out"]
    gk = f4["get_kmer"]
    assert [idx.get_kmer(r).decode() for r in gk["ranks"]] == gk["kmers"]
    text = idx.export_sets()
    assert len(text) == f4["export_len"] and hashlib.md5(text).hexdigest() == f4["export_md5"]
    if "export_text" in f4:
        assert text.decode() == f4["export_text"]
    # contains() is the bit the export prints; update_interval composes to search()
    n = idx.n_nodes
    cols = text[:-1].decode()
    pos, at = 0, 0
    while at < len(cols) and pos < min(n, 500):
        j = at
        while cols[j].isupper():
            j += 1
        members = cols[at:j + 1].upper().replace("$", "")
        assert [idx.contains(pos, c) for c in "ACGT"] == [c in members for c in "ACGT"]
        pos, at = pos + 1, j + 1


def _mixed_case(name):
    reads = open(golden(name, "mixed_case.txt"), "rb").read().split(b"\n")[:-1]
    want = {}
    for kind in ("streaming", "search"):
        rows = open(golden(name, f"mixed_case.{kind}.txt"), "rb").read().split(b"\n")[:-1]
        assert len(rows) == len(reads)
        want[kind] = np.array([int(x) for row in rows for x in row.split()], dtype=np.int64)
    return reads, want


@pytest.mark.parametrize("name", ["small_k31", "small_k8_p0", "small_k63_rc"])
def test_mixed_case_direct_api_against_the_reference(name):
    """Raw mixed-case bytes through the direct API: the oracle's streaming_search upper-cases the new character of a
    streaming step only (SBWT.hh:565), its search() takes bytes as they are (SBWT.hh:427) -- against the vectors the
    reference's own methods returned (`sbwt_ref api`, tests/golden/make_golden.py --only-mixed)."""
    reads, want = _mixed_case(name)
    a, off = synth.ragged_to_batch(reads)
    orc = oracle.OracleIndex(golden(name, "index.sbwt"))
    np.testing.assert_array_equal(orc.query_batch(a, off, streaming=True), want["streaming"])
    np.testing.assert_array_equal(orc.query_batch(a, off, streaming=False), want["search"])
    assert (want["streaming"] != want["search"]).any()  # the fixture exercises the difference


@pytest.mark.parametrize("name", ["long_k80", "long_k255_rc"])
def test_k_above_64_against_the_reference(name):
    """k = 80 and k = 255: index built and queries answered by the reference's own classes (driver compiled with
    MAX_KMER_LENGTH=255); the oracle reproduces the text in both modes."""
    expected = open(golden(name, "expected.txt"), "rb").read()
    vals, counts = parse_expected(expected)
    reads = read_fasta_reads(golden(name, "reads.fna"))
    a, off = synth.ragged_to_batch(reads)
    orc = oracle.OracleIndex(golden(name, "index.sbwt"))
    assert orc.k > 64
    for streaming in (True, False):
        got = orc.query_batch(a, off, streaming=streaming)
        np.testing.assert_array_equal(got, vals)
    assert oracle.format_lines(vals, counts) == expected

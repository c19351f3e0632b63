"""The `sbwt search` command line (sbwt_b200/csrc/cli_search.cpp) and the C++ mirror of the
reference's SBWT<> surface (sbwt_b200/csrc/SBWT.hh): tests shaped like the reference's
tests/test_CLI.hh and api example."""
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle
import sbwt_b200 as S
from conftest import ROOT, c1_expected, c1_reads, golden
from sbwt_b200.testing import strip_streaming_support, synth

CLI = os.path.join(ROOT, "sbwt_b200", "csrc", "sbwt_search")


@pytest.fixture(scope="module", autouse=True)
def _build_cli():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "sbwt_b200", "csrc"), "all"], check=True, capture_output=True)


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


# ------------------------------------------------------------------ CPU: option handling / errors (sbwt_search.cpp:149-199, sbwt.cpp:42-57)

def test_help_and_usage_errors(tmp_path):
    r = run()
    assert r.returncode == 1 and "Query all k-mers of all input reads." in r.stderr
    r = run("search", "--help")
    assert r.returncode == 1 and "--gzip-output" in r.stderr
    r = run("search", "-i", str(tmp_path / "missing.sbwt"), "-q", golden("cli_k6", "queries.fna"), "-o", str(tmp_path / "o.txt"))
    assert r.returncode == 1 and "Error opening file" in r.stderr
    r = run("search", "-i", golden("cli_k6", "index.sbwt"), "-q", golden("cli_k6", "queries.fna"))
    assert r.returncode == 1 and "Option 'out-file' has no value" in r.stderr
    r = run("build", "-i", "x")
    assert r.returncode == 1
    bad = tmp_path / "bad.sbwt"
    bad.write_bytes(open(golden("cli_k6", "index.sbwt"), "rb").read().replace(b"plain-matrix", b"plain-mAtrix"))
    r = run("search", "-i", str(bad), "-q", golden("cli_k6", "queries.fna"), "-o", str(tmp_path / "o.txt"))
    assert r.returncode == 1 and "unrecognized variant" in r.stderr
    rrr = tmp_path / "rrr.sbwt"
    rrr.write_bytes(open(golden("cli_k6", "index.sbwt"), "rb").read().replace(b"\x0c\0\0\0\0\0\0\0plain-matrix", b"\x0a\0\0\0\0\0\0\0rrr-matrix", 1))
    r = run("search", "-i", str(rrr), "-q", golden("cli_k6", "queries.fna"), "-o", str(tmp_path / "o.txt"))
    assert r.returncode == 1 and "plain-matrix variant only" in r.stderr


@pytest.mark.skipif(S.device_count() > 0, reason="checks the behaviour WITHOUT a GPU")
def test_cli_has_no_cpu_fallback(tmp_path):
    r = run("search", "-i", golden("cli_k6", "index.sbwt"), "-q", golden("cli_k6", "queries.fna"), "-o", str(tmp_path / "o.txt"))
    assert r.returncode == 1 and "no CUDA device available" in r.stderr


# ------------------------------------------------------------------ GPU: tests/test_CLI.hh:20-113

@pytest.mark.gpu
def test_end_to_end_query_known_answer(tmp_path):
    known = open(golden("cli_k6", "known_answer.txt")).read()
    qs = {}
    for ext in (".fna", ".fq"):
        src = golden("cli_k6", "queries" + ext)
        qs[ext] = src
        gz = str(tmp_path / ("q" + ext + ".gz"))
        with gzip.open(gz, "wb") as f:
            f.write(open(src, "rb").read())
        qs[ext + ".gz"] = gz
    for ix in ("index.sbwt", "index_nostream.sbwt"):
        for ext, q in qs.items():
            o = str(tmp_path / "o.txt")
            r = run("search", "-o", o, "-i", golden("cli_k6", ix), "-q", q)
            assert r.returncode == 0, r.stderr
            assert open(o).read() == known, (ix, ext)
            assert "us/query" in r.stderr and "us/query end-to-end" in r.stderr
    # list mode + --gzip-output (test_CLI.hh:85-110)
    ins, outs = tmp_path / "in.txt", tmp_path / "out.txt"
    names = [str(tmp_path / f"o{i}.txt.gz") for i in range(4)]
    ins.write_text("\n".join(qs.values()) + "\n")
    outs.write_text("\n".join(names) + "\n")
    r = run("search", "-o", str(outs), "-i", golden("cli_k6", "index.sbwt"), "-q", str(ins), "--gzip-output")
    assert r.returncode == 0, r.stderr
    for n in names:
        assert gzip.open(n, "rt").read() == known
    outs.write_text(names[0] + "\n")
    r = run("search", "-o", str(outs), "-i", golden("cli_k6", "index.sbwt"), "-q", str(ins))
    assert r.returncode == 1 and "Number of input and output files does not match (4 vs 1)" in r.stderr


@pytest.mark.gpu
def test_edge_cases_and_reader_errors(tmp_path):
    o = str(tmp_path / "o.txt")
    r = run("search", "-o", o, "-i", golden("cli_k6", "index.sbwt"), "-q", golden("cli_k6", "edge.fna"), "--batch-bases", "40")
    assert r.returncode == 0, r.stderr
    assert open(o).read() == open(golden("cli_k6", "edge.expected.txt")).read()
    cases = {"nonl.fna": (b">a\nACGTACGT", "ended unexpectedly"), "empty_line.fna": (b">a\nACGT\n\nACGT\n", "Empty line"),
             "empty_seq.fna": (b">a\n>b\nACGT\n", "Empty sequence"), "nostart.fna": (b"ACGT\n", "does not start with '>'"),
             "nostart.fq": (b">a\nACGT\n+\nIIII\n", "does not start with '@'"), "trunc.fq": (b"@a\nACGT\n+\n", "ended unexpectedly"),
             "x.txt2": (b">a\nACGT\n", "Unknown file format")}
    for name, (data, msg) in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        r = run("search", "-o", o, "-i", golden("cli_k6", "index.sbwt"), "-q", str(p))
        assert r.returncode == 1 and msg in r.stderr, (name, r.stderr)
        # the oracle's reader fails on the same inputs
        if name != "x.txt2":
            with pytest.raises(RuntimeError):
                oracle.OracleIndex(golden("cli_k6", "index.sbwt")).search_file(str(p), o)


@pytest.mark.gpu
def test_config1_cli_md5(tmp_path):
    """BASELINE config 1 through the command line: output md5 bbb3a7a4... for both index flavours."""
    q = str(tmp_path / "queries.fastq")
    synth.write_fastq(q, c1_reads())
    ns = str(tmp_path / "ns.sbwt")
    strip_streaming_support(golden("c1", "index.sbwt"), ns)
    for ix in (golden("c1", "index.sbwt"), ns):
        o = str(tmp_path / "o.txt")
        r = run("search", "-o", o, "-i", ix, "-q", q, "--batch-bases", "100000")
        assert r.returncode == 0, r.stderr
        assert hashlib.md5(open(o, "rb").read()).hexdigest() == "bbb3a7a497a2617b2d6ac993ac3560f3"
    assert open(o, "rb").read() == c1_expected()


@pytest.mark.gpu
def test_cpp_mirror_surface(tmp_path):
    exe = str(tmp_path / "test_mirror")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "sbwt_b200"), "-lsbwt_b200", "-Wl,-rpath," + os.path.join(ROOT, "sbwt_b200")], check=True)
    name = "small_k31"
    ref = open(golden(name, "input.fna"), "rb").read().split(b"\n")
    genome = b"".join(x for x in ref[1:400] if not x.startswith(b">"))
    # (mixed case on purpose: the mirror's streaming_search / search take raw bytes like the reference's, SBWT.hh:427,565)
    read = genome[1000:1060] + genome[1060:1064].lower() + genome[1064:1100] + b"N" + genome[1101:1180]
    out = str(tmp_path / "re.sbwt")
    r = subprocess.run([exe, golden(name, "index.sbwt"), out, read.decode()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().split("\n"))
    orc = oracle.OracleIndex(golden(name, "index.sbwt"))
    a, off = synth.ragged_to_batch([read])
    want = orc.query_batch(a, off, streaming=True)
    want_search = orc.query_batch(a, off, streaming=False)
    assert (want != want_search).any() and (want[30:34] >= 0).all() and (want_search[30:61] == -1).all()
    assert [int(x) for x in lines["streaming"].split()] == want.tolist()
    assert [int(x) for x in lines["search"].split()] == want_search.tolist()
    assert [int(x) for x in lines["rebuilt_streaming"].split()] == want.tolist()
    assert lines["header"].split() == [str(orc.n_nodes), str(orc.n_kmers), "31", "8", "1"] + [str(c) for c in orc.C_array]
    assert lines["interval"].split() == [str(want[0]), str(want[0])] and want[0] >= 0
    pl, pr, plen = (int(x) for x in lines["partial"].split())
    assert plen == 100 and pl <= pr  # the walk from the full interval stops at the N
    fw = [int(x) for x in lines["forward"].split()]
    k = 31  # (forward() is case-sensitive like rank(): a lower-case character has no edge)
    assert fw[-1] == -1 and fw[:-1] == [(-1 if read[i + k:i + k + 1].islower() else int(want[i + 1])) if want[i] >= 0 else -2 for i in range(len(fw) - 1)]
    assert [int(x) for x in lines["rank"].split()] == [orc.rank(orc.n_nodes, "A"), orc.rank(orc.n_nodes // 2, "G"), 0]
    assert lines["rebuilt_same_C"] == "1" and lines["rebuilt_same_precalc"] == "1"
    assert "incompatible version of SBWT" in lines["version_error"]
    assert open(out, "rb").read() == open(golden(name, "index.sbwt"), "rb").read()  # serialize() reproduces the file bit for bit


@pytest.mark.gpu
def test_devices_option_splits_every_batch_over_index_replicas(tmp_path):
    """--devices a,b,...: one replica of the index per listed device, every batch cut into contiguous read ranges, one host
    thread per replica, text written in order -- byte-identical to the single-device run and to the reference's output
    (on a one-GPU box the replicas share device 0)."""
    g = lambda f: golden("small_k31", f)
    expected = open(g("expected.txt"), "rb").read()
    n = max(1, S.device_count())
    devs = ",".join(str(i % n) for i in range(3))
    o = str(tmp_path / "multi.txt")
    r = run("search", "-i", g("index.sbwt"), "-q", g("reads.fna"), "-o", o, "--batch-bases", "5000", "--devices", devs)
    assert r.returncode == 0, r.stderr
    assert open(o, "rb").read() == expected


# ------------------------------------------------------------------ GPU: the command line against the reference's own reader / loop

@pytest.mark.gpu
def test_cli_matches_the_reference_on_wellformed_and_malformed_files(tmp_path):
    """`sbwt_search` against `oracle/_ref/sbwt_ref search` (seq_io::Reader + the reference's classes) file by file: same
    exit code, same error message, same output bytes -- including what has been written before a malformed record is
    met (the reference answers read by read; the GPU pipeline hands a parse error on only after the reads in front of
    it have been answered). Corpus: tests/test_fastx.py (every malformation SeqIO singles out) + well-formed files."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/sbwt_ref not built")
    import numpy as np
    from test_fastx import BAD, fasta, fastq
    ix = golden("cli_k6", "index.sbwt")
    rng = np.random.default_rng(4)
    corpus = dict(BAD)
    corpus["ok_multi.fna"] = fasta(rng, 60, True, max_len=120)
    corpus["ok.fq"] = fastq(rng, 60, max_len=120)
    corpus["ok_crlf.fna"] = fasta(rng, 20, True, crlf=True, max_len=60)
    for name, data in sorted(corpus.items()):
        for prefix in (b"", (fastq(rng, 30, max_len=80) if name.endswith(".fq") else fasta(rng, 30, True, max_len=80))):
            if "wrong_start" in name or name.startswith("empty"):
                prefix = b""
            p = tmp_path / name
            p.write_bytes(prefix + data)
            o_ref, o_gpu = str(tmp_path / "ref.txt"), str(tmp_path / "gpu.txt")
            for f in (o_ref, o_gpu):
                if os.path.exists(f):
                    os.remove(f)
            ref = subprocess.run([oracle.REF_BIN, "search", "-i", ix, "-q", str(p), "-o", o_ref], capture_output=True, text=True)
            for extra in ([], ["--batch-bases", "100"]):
                gpu = run("search", "-i", ix, "-q", str(p), "-o", o_gpu, *extra)
                assert gpu.returncode == ref.returncode, (name, gpu.stderr, ref.stderr)
                if ref.returncode != 0:
                    want = [line for line in ref.stderr.splitlines() if "rror" in line][-1]
                    got = [line for line in gpu.stderr.splitlines() if "rror" in line][-1]
                    assert got == want, (name, got, want)
                ref_out = open(o_ref, "rb").read() if os.path.exists(o_ref) else b""
                gpu_out = open(o_gpu, "rb").read() if os.path.exists(o_gpu) else b""
                assert gpu_out == ref_out, (name, extra, len(gpu_out), len(ref_out))


@pytest.mark.gpu
def test_cxxopts_option_forms(tmp_path):
    """--name=value, -nvalue and grouped short flags, as cxxopts accepts them (sbwt_search.cpp:149-165)."""
    import gzip
    ix, q = golden("cli_k6", "index.sbwt"), golden("cli_k6", "queries.fna")
    known = open(golden("cli_k6", "known_answer.txt")).read()
    o = str(tmp_path / "o.txt")
    for args in (["--index-file=" + ix, "--query-file=" + q, "--out-file=" + o], ["-i" + ix, "-q" + q, "-o" + o],
                 ["-i", ix, "--query-file", q, "-o", o, "stray-positional"]):
        if os.path.exists(o):
            os.remove(o)
        r = run("search", *args)
        assert r.returncode == 0, r.stderr
        assert open(o).read() == known
    oz = str(tmp_path / "o.txt.gz")
    r = run("search", "-i", ix, "-q", q, "-zo", oz)
    assert r.returncode == 0, r.stderr
    assert gzip.open(oz).read().decode() == known
    r = run("search", "-i", ix, "-q", q, "-o")
    assert r.returncode == 1 and "is missing an argument" in r.stderr
    r = run("search", "-i", ix, "-q", q, "-o", o, "--no-such-option")
    assert r.returncode == 1 and "does not exist" in r.stderr
    r = run("search", "-i", ix, "-q", q, "-o", o, "--batch-bases", "0")
    assert r.returncode == 1 and "must be positive" in r.stderr

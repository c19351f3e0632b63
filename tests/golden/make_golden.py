#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ (run in the BUILD CONTAINER only).

Needs /root/reference and oracle/_ref/sbwt_ref (`make -C oracle ref`): every
expected output below is produced by the reference's own classes
(SBWT<SubsetMatrixRank<...>>::search / streaming_search, seq_io::Reader)
through that driver, never by code of this repository. Index files are
written by sbwt_b200/testing/build_plain_matrix and asserted byte-identical
to what the reference's constructors serialize:
  * cli_k6, small_k31, small_k63_rc: vs the in-memory constructor
    (NodeBOSSInMemoryConstructor.hh) through `sbwt_ref build-inmem`;
  * c1 (coli3 k=30): vs the KMC-based `sbwt build` binary when a cmake build of the
    reference is available at $SBWT_REF_BIN (one-off probe, recorded in MANIFEST.json).

usage: python tests/golden/make_golden.py
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from sbwt_b200.testing import build_index, strip_streaming_support, synth  # noqa: E402

REF = "/root/reference"
REF_SBWT_BIN = os.environ.get("SBWT_REF_BIN", "/tmp/sbwt_probe/bin/sbwt")


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def ref_search(index, queries, out):
    oracle.ref_run("search", "-i", index, "-q", queries, "-o", out)


def same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def case_cli_k6(man, tmp):
    """tests/test_CLI.hh:20-113: k=6, +RC, p=4; the reference's only hard-coded known answer."""
    d = os.path.join(HERE, "cli_k6")
    os.makedirs(d, exist_ok=True)
    seqs = [b"ACTAGTGTAGCTACAAA", b"ATGTGCTGATGCTAGCATTTTTTT", b"GTGTACTAGTGTGTAGTCGAT"]  # test_CLI.hh:21-22
    synth.write_fasta(os.path.join(d, "input.fna"), seqs)
    idx = os.path.join(d, "index.sbwt")
    build_index(os.path.join(d, "input.fna"), idx, k=6, precalc=4, add_rc=True)
    refidx = os.path.join(tmp, "k6.ref.sbwt")
    oracle.ref_run("build-inmem", "-i", os.path.join(d, "input.fna"), "-o", refidx, "-k", "6", "-p", "4", "--add-reverse-complements")
    assert same(idx, refidx), "builder differs from the reference in-memory constructor (k6)"
    strip_streaming_support(idx, os.path.join(d, "index_nostream.sbwt"))
    oracle.ref_run("build-inmem", "-i", os.path.join(d, "input.fna"), "-o", refidx, "-k", "6", "-p", "4",
                   "--add-reverse-complements", "--no-streaming-support")
    assert same(os.path.join(d, "index_nostream.sbwt"), refidx)

    queries = [b"GGAGAACTAGTGTAGCTACAAAGAGAG", b"AGTGTGTAGCAAAATGTGCTGATGCTAGCAAAAAAAA", b"CTCTACACACTTC"]  # test_CLI.hh:49
    known = ("-1 -1 -1 -1 -1 74 55 77 22 47 36 70 19 31 8 4 3 -1 -1 -1 -1 -1 \n"
             "57 78 23 47 36 -1 -1 -1 -1 -1 52 -1 -1 39 73 54 15 65 53 38 72 20 46 35 11 -1 -1 -1 -1 2 2 2 \n"
             "-1 -1 26 5 25 66 -1 -1 \n")  # test_CLI.hh:90
    synth.write_fasta(os.path.join(d, "queries.fna"), queries)
    synth.write_fastq(os.path.join(d, "queries.fq"), queries)
    with open(os.path.join(d, "known_answer.txt"), "w") as f:
        f.write(known)
    for q in ("queries.fna", "queries.fq"):
        for ix in ("index.sbwt", "index_nostream.sbwt"):
            out = os.path.join(tmp, "o.txt")
            ref_search(os.path.join(d, ix), os.path.join(d, q), out)
            assert open(out).read() == known, (q, ix)

    # edge cases of SURVEY.md section 8(c): lowercase, N, IUPAC, len<k, len==k, multi-line FASTA, CRLF
    edge = os.path.join(d, "edge.fna")
    with open(edge, "wb") as f:
        f.write(b">lower\nggagaactagtgtagctacaaagagag\n")
        f.write(b">N\nGGAGAACTAGTNTAGCTACAAAGAGAG\n")
        f.write(b">short\nACTAG\n")
        f.write(b">exact\nACTAGT\n")
        f.write(b">multiline\nGGAGAACTAGTG\nTAGCTACAAAGAGAG\n")
        f.write(b">iupac\nACTAGTRTAGCTAC\n")
        f.write(b">crlf\nACTAGTGTAG\r\n")
        f.write(b">allN\nNNNNNNNNNNNN\n")
        f.write(b">mixedcase\nActAgTgTaGcTaCaAa\n")
        f.write(b">one\nA\n")
    ref_search(idx, edge, os.path.join(d, "edge.expected.txt"))
    out = os.path.join(tmp, "o.txt")
    ref_search(os.path.join(d, "index_nostream.sbwt"), edge, out)
    assert same(out, os.path.join(d, "edge.expected.txt"))
    man["cli_k6"] = {"index_md5": md5(idx), "index_nostream_md5": md5(os.path.join(d, "index_nostream.sbwt")),
                     "builder_equals_reference_inmem": True}


def mutate(rng, reads, rate):
    m = rng.random(reads.shape) < rate
    sub = synth.LUT[rng.integers(0, 4, size=reads.shape, dtype=np.uint8)]
    return np.where(m, sub, reads)


def ragged_reads(rng, ref, n, k, both_strands):
    """Mixed bag: exact substrings, mutated substrings, random, with N / lowercase, lengths around k."""
    flat = ref.reshape(-1)
    reads = []
    for i in range(n):
        kind = i % 8
        L = int(rng.integers(1, 4 * k)) if kind == 0 else int(rng.integers(k, 6 * k))
        if kind == 1:
            L = k
        if kind == 2:
            L = k - 1
        if kind in (3,):
            s = synth.LUT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        else:
            c = int(rng.integers(0, ref.shape[0]))
            L = min(L, ref.shape[1])
            o = int(rng.integers(0, ref.shape[1] - L + 1))
            s = ref[c, o:o + L].copy()
            if both_strands and rng.random() < 0.5:
                s = synth.revcomp(s)
        if kind == 4:
            s = mutate(rng, s, 0.02)
        if kind == 5 and L > 0:
            s = s.copy()
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(0, L))] = ord("N")
        if kind == 6:
            s = np.frombuffer(bytes(s).lower(), dtype=np.uint8).copy()
            lo = rng.random(L) < 0.5
            s = np.where(lo, s, np.frombuffer(bytes(s).upper(), dtype=np.uint8))
        reads.append(bytes(s))
    return reads


def case_small(man, tmp, name, ref, k, p, add_rc, n_reads, seed):
    d = os.path.join(HERE, name)
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "input.fna")
    synth.write_fasta(fa, [ref[i] for i in range(ref.shape[0])], line_width=70)
    idx = os.path.join(d, "index.sbwt")
    build_index(fa, idx, k=k, precalc=p, add_rc=add_rc)
    refidx = os.path.join(tmp, name + ".ref.sbwt")
    args = ["build-inmem", "-i", fa, "-o", refidx, "-k", str(k), "-p", str(p)]
    if add_rc:
        args.append("--add-reverse-complements")
    oracle.ref_run(*args)
    assert same(idx, refidx), f"builder differs from the reference in-memory constructor ({name})"
    rng = np.random.default_rng(seed)
    reads = ragged_reads(rng, ref, n_reads, k, add_rc)
    synth.write_fasta(os.path.join(d, "reads.fna"), reads)
    ref_search(idx, os.path.join(d, "reads.fna"), os.path.join(d, "expected.txt"))
    ns = os.path.join(tmp, name + ".ns.sbwt")
    strip_streaming_support(idx, ns)
    out = os.path.join(tmp, "o.txt")
    ref_search(ns, os.path.join(d, "reads.fna"), out)
    assert same(out, os.path.join(d, "expected.txt")), "reference: streaming != per-k-mer search"
    man[name] = {"index_md5": md5(idx), "k": k, "precalc_k": p, "add_rc": add_rc, "n_reads": n_reads,
                 "expected_md5": md5(os.path.join(d, "expected.txt")), "builder_equals_reference_inmem": True}


def case_long_k(man, tmp, name, ref, k, p, add_rc, n_reads, seed):
    """k > 64: index built AND answered by the reference itself (the driver compiled with MAX_KMER_LENGTH=255,
    oracle/Makefile target ref255; the repo's test builder stops at k = 64)."""
    ref255 = os.path.join(os.path.dirname(oracle.REF_BIN), "sbwt_ref_k255")
    subprocess.run(["make", "-s", "-C", os.path.dirname(os.path.dirname(oracle.REF_BIN)), "ref255"], check=True)
    d = os.path.join(HERE, name)
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "input.fna")
    synth.write_fasta(fa, [ref[i] for i in range(ref.shape[0])], line_width=70)
    idx = os.path.join(d, "index.sbwt")
    args = [ref255, "build-inmem", "-i", fa, "-o", idx, "-k", str(k), "-p", str(p)]
    if add_rc:
        args.append("--add-reverse-complements")
    subprocess.run(args, check=True, capture_output=True)
    rng = np.random.default_rng(seed)
    reads = [r for r in ragged_reads(rng, ref, n_reads, k, add_rc) if len(r) > 0]  # (SeqIO rejects empty FASTA records)
    synth.write_fasta(os.path.join(d, "reads.fna"), reads)
    subprocess.run([ref255, "search", "-i", idx, "-q", os.path.join(d, "reads.fna"), "-o", os.path.join(d, "expected.txt")], check=True, capture_output=True)
    ns = os.path.join(tmp, name + ".ns.sbwt")
    strip_streaming_support(idx, ns)
    out = os.path.join(tmp, "o.txt")
    subprocess.run([ref255, "search", "-i", ns, "-q", os.path.join(d, "reads.fna"), "-o", out], check=True, capture_output=True)
    assert same(out, os.path.join(d, "expected.txt")), "reference: streaming != per-k-mer search"
    man[name] = {"index_md5": md5(idx), "k": k, "precalc_k": p, "add_rc": add_rc, "n_reads": len(reads),
                 "expected_md5": md5(os.path.join(d, "expected.txt")), "built_by": "the reference's in-memory constructor (sbwt_ref_k255)"}


LONG_K_CASES = [("long_k80", dict(k=80, p=8, add_rc=False, n_reads=160, seed=41), (3, 2500, 40)),
                ("long_k255_rc", dict(k=255, p=6, add_rc=True, n_reads=120, seed=43), (2, 2200, 42))]


def case_c1(man, tmp):
    """BASELINE config 1: example_data/coli3.fna k=30 (defaults: p=8, streaming support) x example_data/queries.fastq."""
    d = os.path.join(HERE, "c1")
    os.makedirs(d, exist_ok=True)
    idx = os.path.join(d, "index.sbwt")
    build_index(os.path.join(REF, "example_data/coli3.fna"), idx, k=30, precalc=8)
    entry = {"index_md5": md5(idx)}
    if os.path.exists(REF_SBWT_BIN):
        refidx = os.path.join(tmp, "c1.ref.sbwt")
        os.makedirs(os.path.join(tmp, "kmc"), exist_ok=True)
        subprocess.run([REF_SBWT_BIN, "build", "-i", os.path.join(REF, "example_data/coli3.fna"), "-o", refidx, "-k", "30",
                        "-t", "8", "-m", "4", "-d", os.path.join(tmp, "kmc")], check=True, capture_output=True)
        assert same(idx, refidx), "builder differs from `sbwt build` (KMC) on config 1"
        entry["builder_equals_reference_sbwt_build"] = True
        out = os.path.join(tmp, "c1.cli.txt")
        subprocess.run([REF_SBWT_BIN, "search", "-i", refidx, "-q", os.path.join(REF, "example_data/queries.fastq"), "-o", out],
                       check=True, capture_output=True)
        entry["reference_cli_output_md5"] = md5(out)
    # reads only (headers / qualities are irrelevant to the path), one per line, gzipped
    reads = []
    with open(os.path.join(REF, "example_data/queries.fastq"), "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(1, len(lines), 4):
        reads.append(lines[i])
    with gzip.GzipFile(os.path.join(d, "reads.txt.gz"), "wb", mtime=0) as f:
        f.write(b"\n".join(reads) + b"\n")
    out = os.path.join(tmp, "c1.txt")
    ref_search(idx, os.path.join(REF, "example_data/queries.fastq"), out)
    entry["expected_md5"] = md5(out)
    if "reference_cli_output_md5" in entry:
        assert entry["expected_md5"] == entry["reference_cli_output_md5"]
    with gzip.GzipFile(os.path.join(d, "expected.txt.gz"), "wb", mtime=0) as f:
        f.write(open(out, "rb").read())
    entry["n_reads"] = len(reads)
    man["c1"] = entry


def case_f4(man, name, queries, seed):
    """The other read-only queries (SBWT.hh:369-381, 526-537, 701-773) answered by the reference's own methods through
    `sbwt_ref partial / forward / getkmer / export`; written to <name>/f4.json."""
    d = os.path.join(HERE, name)
    idx = os.path.join(d, "index.sbwt")
    n_nodes = oracle.OracleIndex(idx).n_nodes
    rng = np.random.default_rng(seed)
    part = [[int(x) for x in line.split()] for line in oracle.ref_run("partial", "-i", idx, "-q", os.path.join(d, queries)).stdout.decode().splitlines()]
    nodes = np.concatenate([rng.integers(0, n_nodes, size=300), [0, 1, n_nodes - 1]]).astype(np.int64)
    chars = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=nodes.size))
    fwd = [int(x) for x in oracle.ref_run("forward", "-i", idx, stdin="".join(f"{n} {chr(c)}\n" for n, c in zip(nodes, chars)).encode()).stdout.split()]
    ranks = np.concatenate([rng.integers(0, n_nodes, size=120), [0, 1, n_nodes - 1]]).astype(np.int64)
    kmers = oracle.ref_run("getkmer", "-i", idx, stdin="".join(f"{r}\n" for r in ranks).encode()).stdout.decode().split()
    tmp = tempfile.mkdtemp(prefix="golden_f4_")
    try:
        exp = os.path.join(tmp, "export.txt")
        oracle.ref_run("export", "-i", idx, "-o", exp)
        export = open(exp, "rb").read()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out = {"queries": queries, "partial_search": part, "forward": {"nodes": nodes.tolist(), "chars": chars.decode(), "out": fwd},
           "get_kmer": {"ranks": ranks.tolist(), "kmers": kmers}, "export_md5": hashlib.md5(export).hexdigest(), "export_len": len(export)}
    if len(export) < 4000:
        out["export_text"] = export.decode()
    with open(os.path.join(d, "f4.json"), "w") as f:
        json.dump(out, f)
    man.setdefault(name, {})["f4"] = {"partial": len(part), "forward": len(fwd), "get_kmer": len(kmers), "export_md5": out["export_md5"]}


def case_mixed(man, name, queries, seed):
    """The direct API on raw mixed-case bytes: `sbwt_ref api` = SBWT::streaming_search(const char*, len) and the search()
    loop on reads in which bases and runs of bases were lower-cased (SBWT.hh:427 takes bytes as they are, SBWT.hh:565
    upper-cases the new character of a streaming step). Written to <name>/mixed_case.txt (one read per line),
    mixed_case.streaming.txt and mixed_case.search.txt (one result vector per line)."""
    d = os.path.join(HERE, name)
    idx = os.path.join(d, "index.sbwt")
    rng = np.random.default_rng(seed)
    reads = []
    cur = []
    for line in open(os.path.join(d, queries), "rb").read().split(b"\n"):
        if line.startswith(b">"):
            if cur:
                reads.append(b"".join(cur))
            cur = []
        elif line:
            cur.append(line)
    if cur:
        reads.append(b"".join(cur))
    out = []
    for r in reads[:200]:
        a = np.frombuffer(r, np.uint8).copy()
        mode = rng.integers(0, 4)
        if mode == 0:      # single bases
            m = rng.random(a.size) < 0.02
        elif mode == 1:    # one run
            m = np.zeros(a.size, bool)
            if a.size > 10:
                s0 = int(rng.integers(0, a.size - 5)); m[s0:s0 + int(rng.integers(1, 40))] = True
        elif mode == 2:    # the first base(s) only: the first k-mer misses, the rest streams on or restarts
            m = np.zeros(a.size, bool); m[: int(rng.integers(1, 3))] = True
        else:              # all lower case
            m = np.ones(a.size, bool)
        up = (a >= 65) & (a <= 90)
        a[m & up] += 32
        out.append(a.tobytes())
    text = b"\n".join(out) + b"\n"
    open(os.path.join(d, "mixed_case.txt"), "wb").write(text)
    st = oracle.ref_run("api", "-i", idx, stdin=text).stdout
    se = oracle.ref_run("api", "-i", idx, "--no-streaming", stdin=text).stdout
    open(os.path.join(d, "mixed_case.streaming.txt"), "wb").write(st)
    open(os.path.join(d, "mixed_case.search.txt"), "wb").write(se)
    man.setdefault(name, {})["mixed_case"] = {"reads": len(out), "streaming_md5": hashlib.md5(st).hexdigest(), "search_md5": hashlib.md5(se).hexdigest(),
                                              "differ": st != se}


MIXED_CASES = [("small_k31", "reads.fna", 31), ("small_k8_p0", "reads.fna", 32), ("small_k63_rc", "reads.fna", 33)]
F4_CASES = [("cli_k6", "queries.fna", 21), ("small_k31", "reads.fna", 22), ("small_k63_rc", "reads.fna", 23), ("small_k8_p0", "reads.fna", 24)]


def main():
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    oracle.build(with_ref=True)
    if "--only-long-k" in sys.argv:  # add the k > 64 fixtures to an existing set
        man = json.load(open(os.path.join(HERE, "MANIFEST.json")))
        tmp = tempfile.mkdtemp(prefix="golden_")
        try:
            for name, kw, (nc, L, sd) in LONG_K_CASES:
                case_long_k(man, tmp, name, synth.random_contigs(nc, L, seed=sd), **kw)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
            json.dump(man, f, indent=1, sort_keys=True)
        return
    if "--only-mixed" in sys.argv:  # add the mixed-case fixtures to an existing set
        man = json.load(open(os.path.join(HERE, "MANIFEST.json")))
        for name, q, seed in MIXED_CASES:
            case_mixed(man, name, q, seed)
        with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
            json.dump(man, f, indent=1, sort_keys=True)
        return
    if "--only-f4" in sys.argv:  # add the f4.json fixtures to an existing set
        man = json.load(open(os.path.join(HERE, "MANIFEST.json")))
        for name, q, seed in F4_CASES:
            case_f4(man, name, q, seed)
        with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
            json.dump(man, f, indent=1, sort_keys=True)
        return
    man = {"generator": "tests/golden/make_golden.py", "expected_outputs_from": "oracle/_ref/sbwt_ref (reference classes)"}
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        case_cli_k6(man, tmp)
        case_small(man, tmp, "small_k31", synth.random_contigs(4, 12000, seed=7), k=31, p=8, add_rc=False, n_reads=600, seed=8)
        case_small(man, tmp, "small_k63_rc", synth.pangenome(6000, 4, 0.05, seed=9), k=63, p=8, add_rc=True, n_reads=400, seed=10)
        case_small(man, tmp, "small_k8_p0", synth.random_contigs(2, 3000, seed=11), k=8, p=0, add_rc=True, n_reads=300, seed=12)
        case_c1(man, tmp)
        for name, q, seed in F4_CASES:
            case_f4(man, name, q, seed)
        for name, q, seed in MIXED_CASES:
            case_mixed(man, name, q, seed)
        for name, kw, (nc, L, sd) in LONG_K_CASES:
            case_long_k(man, tmp, name, synth.random_contigs(nc, L, seed=sd), **kw)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    print(json.dumps(man, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

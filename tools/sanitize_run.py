#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): the golden fixtures and config 1 through both query modes,
int64 and int32 results, the hits-only call and the batched interval queries, checked against the committed reference outputs.
The probe stride comes from SBWT_B200_PROBE (0 / 1 / 9 are run by tools/gpu_sanitize.sh)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbwt_b200 as S  # noqa: E402
from conftest import c1_expected, c1_reads, golden, parse_expected, read_fasta_reads  # noqa: E402
from sbwt_b200.testing import synth  # noqa: E402

cases = [("small_k31", "reads.fna", "expected.txt"), ("small_k63_rc", "reads.fna", "expected.txt"), ("cli_k6", "queries.fna", "known_answer.txt")]
if "--with-c1" in sys.argv:
    cases.append(("c1", None, None))
for name, q, e in cases:
    if name == "c1":
        reads, vals = c1_reads()[:1500], None
        full, _ = parse_expected(c1_expected())
        vals = full[: sum(max(0, len(r) - 30 + 1) for r in reads)]
    else:
        reads = read_fasta_reads(golden(name, q))
        vals, _ = parse_expected(open(golden(name, e), "rb").read())
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(golden(name, "index.sbwt"))
    for cap in ((a.size, len(reads)), (max(len(r) for r in reads) + 500, 37)):
        ses = S.Session(idx, *cap)
        for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
            assert np.array_equal(ses.query_host(a, off, mode), vals), (name, mode)
            assert np.array_equal(ses.query_host_i32(a, off, mode).astype(np.int64), vals), (name, mode, "i32")
            mask, hits, n = ses.query_host_hits(a, off, mode)
            bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[: vals.size].astype(bool)
            assert np.array_equal(bits, vals >= 0) and np.array_equal(hits, vals[vals >= 0]), (name, mode, "hits")
        ses.close()
    if name != "c1":
        f4 = json.load(open(golden(name, "f4.json")))
        l, r, m = idx.partial_search(a, off)
        assert np.stack([l, r, m], axis=1).tolist() == f4["partial_search"]
        assert [x.decode() for x in idx.get_kmers(f4["get_kmer"]["ranks"][:20])] == f4["get_kmer"]["kmers"][:20]
        assert np.array_equal(idx.forward(f4["forward"]["nodes"], f4["forward"]["chars"].encode()), f4["forward"]["out"])
    idx.close()
    print("ok", name, vals.size, "results", flush=True)

# the direct API's mixed-case mode (case_fixup_kernel, lower_mask_kernel) and k > 64 (long_kmer_kernels.cuh)
name = "small_k31"
reads = open(golden(name, "mixed_case.txt"), "rb").read().split(b"\n")[:-1]
want = np.array([int(x) for row in open(golden(name, "mixed_case.streaming.txt"), "rb").read().split(b"\n")[:-1] for x in row.split()], dtype=np.int64)
a, off = synth.ragged_to_batch(reads)
idx = S.Index(golden(name, "index.sbwt"))
ses = S.Session(idx, max(len(r) for r in reads) * 3, 7)
assert np.array_equal(ses.query_host(a, off, S.MODE_STREAMING, S.CASE_API), want)
assert np.array_equal(ses.query_host_i32(a, off, S.MODE_STREAMING, S.CASE_API).astype(np.int64), want)
ses.close()
idx.close()
print("ok mixed case", want.size, "results", flush=True)
name = "long_k80"
reads = read_fasta_reads(golden(name, "reads.fna"))
vals, _ = parse_expected(open(golden(name, "expected.txt"), "rb").read())
a, off = synth.ragged_to_batch(reads)
idx = S.Index(golden(name, "index.sbwt"))
ses = S.Session(idx, max(len(r) for r in reads) * 2, 5)
for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
    assert np.array_equal(ses.query_host(a, off, mode), vals), (name, mode)
ses.close()
idx.close()
print("ok", name, vals.size, "results", flush=True)

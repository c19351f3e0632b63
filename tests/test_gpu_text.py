"""print_vector on the device (sbwt_gpu_query_host_text / sbwt_gpu_format_device) against the reference's
text: the committed golden outputs and the oracle's restatement of sbwt_search.cpp:21-43."""
import numpy as np
import pytest

import oracle
import sbwt_b200 as S
from conftest import c1_expected, c1_reads, golden, parse_expected, read_fasta_reads
from sbwt_b200.testing import synth

pytestmark = pytest.mark.gpu


def text_of(index_path, reads, mode, max_bases=None, max_reads=None):
    a, off = synth.ragged_to_batch(reads)
    idx = S.Index(index_path)
    ses = S.Session(idx, max_bases or max(1, a.size), max_reads or max(1, len(reads)))
    text, n = ses.query_host_text(a, off, mode)
    ses.close()
    idx.close()
    return text, n


@pytest.mark.parametrize("name", ["cli_k6", "small_k31", "small_k63_rc", "small_k8_p0"])
def test_text_matches_golden(name):
    if name == "cli_k6":
        expected = open(golden(name, "known_answer.txt"), "rb").read() + open(golden(name, "edge.expected.txt"), "rb").read()
        reads = read_fasta_reads(golden(name, "queries.fna")) + read_fasta_reads(golden(name, "edge.fna"))
    else:
        expected = open(golden(name, "expected.txt"), "rb").read()
        reads = read_fasta_reads(golden(name, "reads.fna"))
    vals, _ = parse_expected(expected)
    for mode in (S.MODE_STREAMING, S.MODE_SEARCH):
        text, n = text_of(golden(name, "index.sbwt"), reads, mode)
        assert text == expected
        assert n == vals.size
    # chunked through a small session: many batches, pieces still arrive in order
    text, n = text_of(golden(name, "index.sbwt"), reads, S.MODE_STREAMING, max_bases=max(len(r) for r in reads) + 64, max_reads=5)
    assert text == expected and n == vals.size


def test_text_config1():
    expected = c1_expected()
    text, n = text_of(golden("c1", "index.sbwt"), c1_reads(), S.MODE_STREAMING)
    assert text == expected and n == 355000


def test_text_empty_and_short_reads():
    idx = S.Index(golden("small_k31", "index.sbwt"))
    ses = S.Session(idx, 4096, 64)
    a, off = synth.ragged_to_batch([b"ACGT", b"", b"A" * 30])
    text, n = ses.query_host_text(a, off, S.MODE_STREAMING)
    assert text == b"\n\n\n" and n == 0
    text, n = ses.query_host_text(a, np.zeros(1, np.int64), S.MODE_STREAMING)
    assert text == b"" and n == 0


@pytest.mark.parametrize("i32", [False, True])
def test_format_device_adversarial_values(i32):
    """The formatter alone on values of every digit count, zeros (an empty field in the reference), ragged
    reads, reads without k-mers; int64 and int32 inputs; text capacity overflow is reported, not written."""
    import torch
    rng = np.random.default_rng(11)
    idx = S.Index(golden("small_k31", "index.sbwt"))
    k = idx.k
    n_reads = 3001
    lens = rng.integers(0, 400, n_reads)
    lens[:40] = np.arange(40)            # reads shorter than, equal to and just above k
    lens[100] = 5000                     # one long read
    off = np.zeros(n_reads + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    counts = np.maximum(0, lens - k + 1)
    n_vals = int(counts.sum())
    top = 31 if i32 else 63
    digits = rng.integers(0, top, n_vals)
    vals = (rng.integers(0, 1 << 62, n_vals) >> (62 - digits)).astype(np.int64)   # all magnitudes, zeros included
    vals[rng.random(n_vals) < 0.3] = -1
    vals[:8] = [0, 1, 9, 10, 99, 100, (1 << top) - 1, -1]
    want = oracle.format_lines(vals, counts)
    ses = S.Session(idx, int(off[-1]) + 1, n_reads)
    d_vals = torch.from_numpy(vals.astype(np.int32) if i32 else vals).cuda()
    d_off = torch.from_numpy(off).cuda()
    cap = len(want) + 7
    d_text = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    for shift in (0, 1, 13):             # unaligned destination
        d_text.zero_()
        ses.format_device(d_vals.data_ptr(), i32, d_off.data_ptr(), n_reads, d_text.data_ptr() + shift, cap, d_n.data_ptr())
        torch.cuda.synchronize()
        assert int(d_n.item()) == len(want)
        got = d_text.cpu().numpy()
        assert got[shift:shift + len(want)].tobytes() == want
        assert not got[:shift].any() and not got[shift + len(want):].any()   # nothing outside the text
    d_text.zero_()
    ses.format_device(d_vals.data_ptr(), i32, d_off.data_ptr(), n_reads, d_text.data_ptr(), len(want) - 1, d_n.data_ptr())
    torch.cuda.synchronize()
    assert int(d_n.item()) == len(want) and not d_text.cpu().numpy().any()
    ses.close()
    idx.close()

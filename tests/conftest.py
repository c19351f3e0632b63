"""Shared fixtures. GPU tests are marked @pytest.mark.gpu; everything else runs on CPU."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Compile the C oracle and the index constructor once (test infrastructure, not the product)."""
    import oracle
    from sbwt_b200.testing import build_tools
    oracle.build()
    build_tools()


def golden(*parts):
    return os.path.join(GOLDEN, *parts)


def read_fasta_reads(path):
    """Sequences of a FASTA file the way seq_io::Reader hands them to the index: concatenated lines, upper-cased."""
    reads, cur = [], None
    with open(path, "rb") as f:
        for line in f.read().split(b"\n"):
            if line.startswith(b">"):
                if cur is not None:
                    reads.append(cur)
                cur = b""
            elif cur is not None:
                cur += line
    if cur is not None:
        reads.append(cur)
    return [r.upper() for r in reads]


def parse_expected(text):
    """Reference output text -> (int64 values, per-read counts)."""
    if isinstance(text, bytes):
        text = text.decode()
    lines = text.split("\n")
    assert lines[-1] == ""
    counts, vals = [], []
    for ln in lines[:-1]:
        toks = ln.split()
        counts.append(len(toks))
        vals.extend(int(t) for t in toks)
    return np.array(vals, dtype=np.int64), counts


def c1_reads():
    with gzip.open(golden("c1", "reads.txt.gz"), "rb") as f:
        return [r for r in f.read().split(b"\n") if r]


def c1_expected():
    with gzip.open(golden("c1", "expected.txt.gz"), "rb") as f:
        return f.read()

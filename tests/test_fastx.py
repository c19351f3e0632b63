"""ParallelFastxReader (sbwt_b200/csrc/fastx.hpp) against the serial FastxReader, which restates
seq_io::Reader::get_next_read_to_buffer (SeqIO/include/SeqIO/SeqIO.hh:255-360): same batches byte for byte, same
exceptions, over well-formed files (single- and multi-line FASTA, FASTQ, CRLF, long reads, tiny batches) and over
every malformation the reference singles out (empty lines, empty sequences, a last line without a newline, truncated
FASTQ records, an empty FASTQ header). CPU only."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fx") / "test_fastx")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", out, os.path.join(ROOT, "tests", "cpp", "test_fastx.cpp"), "-lz", "-lpthread"], check=True)
    return out


def both(exe, path, max_bases, max_reads, threads):
    r = subprocess.run([exe, path, str(max_bases), str(max_reads), str(threads)], capture_output=True, text=True, check=True)
    a, b = r.stdout.strip().split("\n")
    return a, b


def write_bgzf(path, data, block=60_000):
    """bgzip's container: gzip members of < 64 KiB whose header carries the member size in a 'BC' extra field, then the
    empty end-of-file member."""
    import struct
    import zlib
    with open(path, "wb") as f:
        for a in list(range(0, len(data), block)) + [None]:
            chunk = b"" if a is None else data[a:a + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            body = c.compress(chunk) + c.flush()
            bsize = 12 + 6 + len(body) + 8
            f.write(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1))
            f.write(body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def rand_seq(rng, n):
    return bytes(np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=n)])


def fasta(rng, n, multi_line, crlf=False, max_len=400):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        s = rand_seq(rng, int(rng.integers(1, max_len)))
        out.append(b">r%d some header > with @ signs" % i + nl)
        if multi_line:
            w = int(rng.integers(1, 90))
            out += [s[j:j + w] + nl for j in range(0, len(s), w)]
        else:
            out.append(s + nl)
    return b"".join(out)


def fastq(rng, n, crlf=False, max_len=400):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        s = rand_seq(rng, int(rng.integers(1, max_len)))
        q = bytes(rng.integers(33, 74, size=len(s), dtype=np.uint8))  # '@' and '+' occur as quality values
        if i % 7 == 0:
            q = b"@" + q[1:]
        out += [b"@r%d" % i + nl, s + nl, b"+" + nl, q + nl]
    return b"".join(out)


@pytest.mark.parametrize("kind", ["fa1", "fam", "fq", "fa_crlf", "fq_crlf"])
def test_wellformed_files_give_identical_batches(exe, tmp_path, kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    data = {"fa1": lambda: fasta(rng, 3000, False), "fam": lambda: fasta(rng, 3000, True), "fq": lambda: fastq(rng, 3000),
            "fa_crlf": lambda: fasta(rng, 500, True, crlf=True), "fq_crlf": lambda: fastq(rng, 500, crlf=True)}[kind]()
    path = str(tmp_path / ("x.fq" if kind.startswith("fq") else "x.fna"))
    open(path, "wb").write(data)
    with gzip.open(path + ".gz", "wb", compresslevel=1) as f:  # the same bytes through the inflate-ahead path
        f.write(data)
    bg = str(tmp_path / ("b.fq.gz" if kind.startswith("fq") else "b.fna.gz"))  # ... and through the member-parallel BGZF path
    write_bgzf(bg, data, block=7_000)
    assert gzip.open(bg).read() == data
    for max_bases, max_reads, threads in [(1 << 30, 1 << 30, 4), (5000, 1 << 30, 3), (1, 1 << 30, 2), (1 << 30, 7, 8), (100_000, 100, 1), (333, 5, 5)]:
        a, b = both(exe, path, max_bases, max_reads, threads)
        assert not a.startswith("ERROR"), a
        assert a == b, (kind, max_bases, max_reads, threads)
        az, bz = both(exe, path + ".gz", max_bases, max_reads, threads)
        assert az == a and bz == a, (kind, "gz", max_bases, max_reads, threads)
        ag, bg_ = both(exe, bg, max_bases, max_reads, threads)
        assert ag == a and bg_ == a, (kind, "bgzf", max_bases, max_reads, threads)


def test_random_batch_shapes_and_thread_counts(exe, tmp_path):
    """The records of a window are formed by slices of its lines in parallel: batch limits, windows and slice boundaries
    that fall anywhere inside records (multi-line FASTA of uneven line widths, FASTQ), 1 to 16 threads, also with a
    malformed record late in the file (the reads in front of it are still delivered, then the same error)."""
    rng = np.random.default_rng(77)
    recs = []
    for i in range(4000):
        s_ = rand_seq(rng, int(rng.integers(1, 900)))
        w = int(rng.integers(1, 120))
        recs.append(b">r%d some text\n" % i + b"".join(s_[j:j + w] + b"\n" for j in range(0, len(s_), w)))
    fa = b"".join(recs)
    files = {"m.fna": fa, "q.fq": fastq(rng, 4000, max_len=700), "bad_late.fna": fa + b">x\n>y\nACGT\n", "bad_mid.fna": b"".join(recs[:2500]) + b">e\n\nAC\n" + b"".join(recs[2500:])}
    for name, data in files.items():
        path = str(tmp_path / name)
        open(path, "wb").write(data)
        for _ in range(12):
            mb = int(rng.choice([1, 50, 997, 20_000, 333_333, 1 << 30]))
            mr = int(rng.choice([1, 3, 64, 1000, 1 << 30]))
            th = int(rng.choice([1, 2, 3, 7, 16]))
            a, b = both(exe, path, mb, mr, th)
            assert a == b, (name, mb, mr, th, a[:200], b[:200])
            assert a.startswith("ERROR") == name.startswith("bad"), (name, a[:200])


def test_long_reads_exceed_the_scan_window(exe, tmp_path):
    rng = np.random.default_rng(5)
    path = str(tmp_path / "long.fna")
    with open(path, "wb") as f:
        for i in range(6):
            s = rand_seq(rng, 3_000_000 + i)
            f.write(b">c%d\n" % i)
            for j in range(0, len(s), 70):
                f.write(s[j:j + 70] + b"\n")
    with open(path, "rb") as f, gzip.open(path + ".gz", "wb", compresslevel=1) as g:
        g.write(f.read())
    for mb in (10, 1_000_000, 1 << 30):
        a, b = both(exe, path, mb, 1 << 30, 4)
        assert not a.startswith("ERROR") and a == b
        az, bz = both(exe, path + ".gz", mb, 1 << 30, 4)
        assert az == a and bz == a


BAD = {
    "fa_empty_line_after_header.fna": b">a\nACGT\n>b\n\nACGT\n>c\nAC\n",
    "fa_empty_line_in_sequence.fna": b">a\nACGT\n>b\nAC\n\nGT\n",
    "fa_empty_sequence.fna": b">a\nACGT\n>b\n>c\nACGT\n",
    "fa_header_at_end.fna": b">a\nACGT\n>b\n",
    "fa_no_final_newline.fna": b">a\nACGT\n>b\nACGTT",
    "fa_trailing_empty_line.fna": b">a\nACGT\n>b\nACGTT\n\n",
    "fa_wrong_start.fna": b"ACGT\n>b\nACGTT\n",
    "fq_truncated.fq": b"@a\nACGT\n+\nIIII\n@b\nAC\n+\n",
    "fq_no_final_newline.fq": b"@a\nACGT\n+\nIIII\n@b\nAC\n+\nII",
    "fq_empty_sequence.fq": b"@a\nACGT\n+\nIIII\n@b\n\n+\n\n@c\nAC\n+\nII\n",
    "fq_empty_header.fq": b"@a\nACGT\n+\nIIII\n\nAC\n+\nII\n@c\nAC\n+\nII\n",
    "fq_wrong_start.fq": b"a\nACGT\n+\nIIII\n",
    "fq_extra_blank_lines.fq": b"@a\nACGT\n+\nIIII\n\n\n",
    "empty.fna": b"",
    "empty.fq": b"",
}


@pytest.mark.parametrize("name", sorted(BAD))
def test_malformed_files_fail_or_parse_exactly_like_the_serial_reader(exe, tmp_path, name):
    path = str(tmp_path / name)
    # a few good records in front, so that the anomaly is met in a later batch as well as in the first
    rng = np.random.default_rng(3)
    for prefix in (b"", (fastq(rng, 40) if name.endswith(".fq") else fasta(rng, 40, True))):
        if name.startswith("fa_wrong") or name.startswith("fq_wrong") or name.startswith("empty"):
            prefix = b""
        open(path, "wb").write(prefix + BAD[name])
        with gzip.open(path + ".gz", "wb") as f:
            f.write(prefix + BAD[name])
        for max_bases, max_reads, threads in [(1 << 30, 1 << 30, 4), (50, 1 << 30, 2), (1 << 30, 1, 3)]:
            a, b = both(exe, path, max_bases, max_reads, threads)
            assert a == b, (name, a, b)
            az, bz = both(exe, path + ".gz", max_bases, max_reads, threads)
            assert az == bz == a.replace(name, name + ".gz"), (name, "gz", az, bz, a)
            bgp = str(tmp_path / ("bg_" + name + ".gz"))
            write_bgzf(bgp, prefix + BAD[name], block=25)
            ag, bg_ = both(exe, bgp, max_bases, max_reads, threads)
            assert ag == bg_ == a.replace(name, "bg_" + name + ".gz"), (name, "bgzf", ag, bg_, a)


def test_gzip_input_large_enough_to_slide_the_buffer(exe, tmp_path):
    """more than the 64 MB after which the inflate buffer drops what was consumed, multi-line FASTA and FASTQ"""
    rng = np.random.default_rng(10)
    seq = rand_seq(rng, 5_000_000)
    path = str(tmp_path / "big.fna.gz")
    with gzip.open(path, "wb", compresslevel=1) as f:
        for i in range(40):
            f.write(b">c%d\n" % i)
            s_ = seq[i * 1000:i * 1000 + 4_000_000]
            f.write(b"\n".join(s_[j:j + 80] for j in range(0, len(s_), 80)) + b"\n")
    for mb, mr in ((1 << 30, 1 << 30), (3_000_000, 1 << 30), (20_000_000, 3)):
        a, b = both(exe, path, mb, mr, 4)
        assert not a.startswith("ERROR") and a == b and a.split()[1] == "40"


def test_gzip_input_small(exe, tmp_path):
    rng = np.random.default_rng(9)
    data = fastq(rng, 500)
    path = str(tmp_path / "x.fastq.gz")
    with gzip.open(path, "wb") as f:
        f.write(data)
    plain = str(tmp_path / "x.fastq")
    open(plain, "wb").write(data)
    a, b = both(exe, path, 10_000, 1 << 30, 4)
    a2, b2 = both(exe, plain, 10_000, 1 << 30, 4)
    assert not a.startswith("ERROR") and a == b == a2 == b2

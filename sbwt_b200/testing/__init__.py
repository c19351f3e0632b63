"""Test / bench infrastructure: synthetic data, an index constructor and a numpy
parser of the reference's .sbwt file. Nothing here is on the product path
(index construction is out of scope)."""
from __future__ import annotations

import json
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILDER_SRC = os.path.join(HERE, "build_plain_matrix.cpp")
BUILDER_BIN = os.path.join(HERE, "build_plain_matrix")


def build_tools(force: bool = False) -> None:
    """Compile the index constructor (g++ -O3 -fopenmp)."""
    if not force and os.path.exists(BUILDER_BIN) and os.path.getmtime(BUILDER_BIN) >= os.path.getmtime(BUILDER_SRC):
        return
    subprocess.run(["g++", "-O3", "-fopenmp", "-std=c++17", "-o", BUILDER_BIN, BUILDER_SRC,
                    "-static-libstdc++", "-static-libgcc"], check=True)


def build_index(inputs, out_path: str, k: int, precalc: int = 8, streaming: bool = True,
                add_rc: bool = False, raw: bool = False, threads: int = 0) -> dict:
    """Build a plain-matrix .sbwt file, byte-compatible with `sbwt build` (tests/test_builder.py)."""
    build_tools()
    if isinstance(inputs, str):
        inputs = [inputs]
    cmd = [BUILDER_BIN, "-q", "-i", ",".join(inputs), "-o", out_path, "-k", str(k), "-p", str(precalc)]
    if not streaming:
        cmd.append("--no-streaming-support")
    if add_rc:
        cmd.append("--add-reverse-complements")
    if raw:
        cmd.append("--raw")
    if threads:
        cmd += ["-t", str(threads)]
    res = subprocess.run(cmd, capture_output=True, check=True)
    return json.loads(res.stdout.decode().strip().splitlines()[-1])


def read_sbwt(path: str) -> dict:
    """Parse the serialized plain-matrix layout into numpy arrays."""
    data = np.fromfile(path, dtype=np.uint8)
    pos = 0

    def take(n):
        nonlocal pos
        b = data[pos:pos + n]
        if b.size != n:
            raise ValueError("truncated .sbwt file")
        pos += n
        return b

    def i64():
        return struct.unpack("<q", take(8).tobytes())[0]

    def string():
        return take(i64()).tobytes().decode()

    def bitvec():
        nbits = i64()
        return nbits, take(((nbits + 63) // 64) * 8).view(np.uint64).copy()

    out = {"variant": string(), "version": string()}
    out["bits"] = [bitvec() for _ in range(4)]
    out["rank_support"] = [bitvec()[1] for _ in range(4)]
    out["sgs"] = bitvec()
    assert i64() == 32
    out["C"] = take(32).view(np.int64).copy()
    nb = i64()
    out["precalc"] = take(nb).view(np.int64).copy().reshape(-1, 2)
    out["precalc_k"], out["n_nodes"], out["n_kmers"], out["k"] = i64(), i64(), i64(), i64()
    if pos != data.size:
        raise ValueError("trailing bytes in .sbwt file")
    return out


def strip_streaming_support(src: str, dst: str) -> None:
    """Rewrite an index without its suffix-group vector: exactly what
    `sbwt build --no-streaming-support` writes (verified byte-for-byte on config 1)."""
    d = read_sbwt(src)
    raw = np.fromfile(src, dtype=np.uint8)
    nb, words = d["sgs"]
    # locate the vector: it sits right before the C array block, whose position is fixed from the end
    tail = 8 + 32 + 8 + d["precalc"].size * 8 + 32
    start = raw.size - tail - (8 + words.size * 8)
    with open(dst, "wb") as f:
        f.write(raw[:start].tobytes())
        f.write(struct.pack("<q", 0))
        f.write(raw[raw.size - tail:].tobytes())

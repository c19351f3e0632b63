"""The index constructor (test infrastructure) reproduces the committed index files, which
make_golden.py asserted byte-identical to the reference constructors' output."""
import hashlib
import json

import pytest

from conftest import golden
from sbwt_b200.testing import build_index, read_sbwt, strip_streaming_support

MAN = json.load(open(golden("MANIFEST.json")))


def md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


@pytest.mark.parametrize("name,k,p,rc", [("cli_k6", 6, 4, True), ("small_k31", 31, 8, False),
                                         ("small_k63_rc", 63, 8, True), ("small_k8_p0", 8, 0, True)])
def test_rebuild_matches_manifest(name, k, p, rc, tmp_path):
    out = str(tmp_path / "i.sbwt")
    info = build_index(golden(name, "input.fna"), out, k=k, precalc=p, add_rc=rc)
    assert md5(out) == MAN[name]["index_md5"] == md5(golden(name, "index.sbwt"))
    d = read_sbwt(out)
    assert (d["n_nodes"], d["n_kmers"], d["k"], d["precalc_k"]) == (info["n_nodes"], info["n_kmers"], k, p)
    assert d["variant"] == "plain-matrix" and d["version"] == "v0.1"


def test_strip_streaming_support(tmp_path):
    out = str(tmp_path / "ns.sbwt")
    strip_streaming_support(golden("cli_k6", "index.sbwt"), out)
    assert md5(out) == MAN["cli_k6"]["index_nostream_md5"]
    out2 = str(tmp_path / "ns2.sbwt")
    build_index(golden("cli_k6", "input.fna"), out2, k=6, precalc=4, add_rc=True, streaming=False)
    assert md5(out2) == md5(out)

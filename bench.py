#!/usr/bin/env python
"""bench.py -- k-mer lookups/s of the plain-matrix SBWT query path on B200.

A "step" is one pass of the hot path (pack -> plan -> walk) over one batch of synthetic reads.
Default workload = BASELINE.json configs[1] ("c2"): 100 Mbp random-DNA reference, k=31
plain-matrix index with streaming support (p=8), 10 M x 150 bp reads at 50 % hit rate,
streaming_search. `--workload c3` is the same index without streaming support (per-k-mer
search), `--workload c4s` / `c5s` a scaled pangenome-like index (+RC, k = 31 / 63) that exceeds L2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4s|c5s|tiny] [--reads R]

Prints ONE JSON line (rank 0). `value` is measured with inputs resident in HBM, `e2e` through
sbwt_gpu_query_host with pinned HOST buffers (H2D + D2H inside the timed region). `roofline` carries the
counted algorithmic sector bytes of the walk kernel against the measured HBM copy peak, the DRAM bytes one
launch really moves (profiles/traffic.json, from ncu), the measured random-sector ceilings and the same kernel
timed with the shorter search table; `cpu_baseline` is the reference's own query code on the host cores.
`--impl reference` runs only that (the reference arm of the contract).
Multi-GPU (torchrun, one rank per GPU): the index is replicated, every rank answers its own
shard of reads, no collective on the data path (weak scaling); NCCL is used only for the
barrier and the max-over-ranks of the timing the contract asks for.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sbwt_b200.testing import build_index, synth  # noqa: E402

CACHE = os.path.join(ROOT, ".cache", "bench")

WORKLOADS = {
    # name: reference generator, k, streaming, add_rc, default reads
    "c2": dict(desc="configs[1]: 100 Mbp random DNA, k=31 plain-matrix (p=8, streaming support), 150 bp reads, 50% hit rate, streaming_search",
               ref=("contigs", 100, 1_000_000, 42), k=31, streaming=True, rc=False, reads=10_000_000),
    "c3": dict(desc="configs[2]: same reference built with --no-streaming-support, per-k-mer search (k-p dependent steps per k-mer)",
               ref=("contigs", 100, 1_000_000, 42), k=31, streaming=False, rc=False, reads=10_000_000),
    "c4s": dict(desc="configs[3] scaled: pangenome-like 40 x 5 Mbp mutated copies (5% subst.), k=31 +RC, index > L2, streaming_search",
                ref=("pangenome", 40, 5_000_000, 44), k=31, streaming=True, rc=True, reads=10_000_000),
    "c5s": dict(desc="configs[4] scaled: the same pangenome-like reference, k=63 (+RC), index > L2, streaming_search, 88 lookups per 150-bp read",
                ref=("pangenome", 40, 5_000_000, 44), k=63, streaming=True, rc=True, reads=10_000_000),
    "c4": dict(desc="configs[3] at full size: pangenome-like 400 x 5 Mbp mutated copies (5% subst.) = 2 Gbp, k=31 +RC (~3.2 G columns: narrow layout with values >= 2^31), streaming_search",
               ref=("pangenome", 400, 5_000_000, 44), k=31, streaming=True, rc=True, reads=10_000_000),
    "c5": dict(desc="configs[4] at full size: the same 2 Gbp pangenome-like reference, k=63 (+RC), streaming_search, 88 lookups per 150-bp read",
               ref=("pangenome", 400, 5_000_000, 44), k=63, streaming=True, rc=True, reads=10_000_000),
    "c2q": dict(desc="experiment: a quarter-size configs[1] reference (25 Mbp) -- the index certainly fits in L2", ref=("contigs", 25, 1_000_000, 42),
                k=31, streaming=True, rc=False, reads=10_000_000),
    "tiny": dict(desc="smoke-sized: 2 Mbp random DNA, k=31, streaming", ref=("contigs", 2, 1_000_000, 42), k=31, streaming=True,
                 rc=False, reads=200_000),
}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def reference_matrix(spec):
    kind, n, length, seed = spec
    if kind == "contigs":
        return synth.random_contigs(n, length, seed)
    return synth.pangenome(length, n, 0.05, seed)


def ensure_index(name: str, w: dict) -> tuple[str, np.ndarray]:
    """Build (or reuse) the synthetic index with the in-repo constructor; returns (path, reference matrix)."""
    os.makedirs(CACHE, exist_ok=True)
    ref = reference_matrix(w["ref"])
    tag = "_".join(str(x) for x in w["ref"]) + f"_k{w['k']}" + ("_rc" if w["rc"] else "") + ("" if w["streaming"] else "_nostream")
    path = os.path.join(CACHE, tag + ".sbwt")
    if not os.path.exists(path):
        raw = os.path.join(CACHE, tag + ".raw")
        with open(raw, "wb") as f:
            for i in range(ref.shape[0]):
                f.write(ref[i].tobytes())
                f.write(b"\n")
        t0 = time.time()
        info = build_index(raw, path + ".tmp", k=w["k"], precalc=8, streaming=w["streaming"], add_rc=w["rc"], raw=True)
        os.replace(path + ".tmp", path)
        os.remove(raw)
        import resource
        log(f"built {path}: {info} in {time.time() - t0:.1f}s (builder peak RSS {resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1e6:.1f} GB)")
    return path, ref


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def write_sample_fasta(path: str, reads: np.ndarray) -> None:
    n, L = reads.shape
    m = np.empty((n, L + 3), dtype=np.uint8)
    m[:, 0], m[:, 1], m[:, 2:L + 2], m[:, L + 2] = ord(">"), ord("\n"), reads, ord("\n")
    m.tofile(path)


def cpu_reference(index_path: str, reads: np.ndarray, streaming: bool, threads: int | None = None, reps: int = 1) -> dict:
    """Time the reference's own query code (oracle/_ref/sbwt_ref, compiled from /root/reference) on the
    host cores over a bounded sample; falls back to the C port of the oracle when _ref is absent."""
    import oracle
    threads = threads or os.cpu_count() or 1
    n = int(min(reads.shape[0], 2_000_000))  # c2: ~3 s on 16 threads (~50 core-seconds); c3: ~20 s
    sample = reads[:n]
    if oracle.ref_available():
        q = os.path.join(CACHE, f"cpu_sample_{os.getpid()}.fna")
        write_sample_fasta(q, sample)
        try:
            res = oracle.ref_run("timed", "-i", index_path, "-q", q, "-t", str(threads), "-r", str(reps))
        finally:
            os.remove(q)
        r = json.loads(res.stdout.decode().strip().splitlines()[-1])
        return {"value": r["lookups"] / r["seconds"], "unit": "lookups/s", "cores": threads, "kind": "reference",
                "sample": f"first {n} reads ({r['lookups']} lookups, {r['hits']} hits) of the workload, {threads} threads over "
                          f"SBWT::{'streaming_search' if streaming else 'search'} compiled from the reference sources",
                "seconds": r["seconds"], "checksum": r["checksum"], "hits": r["hits"], "lookups": r["lookups"]}
    a, off = synth.matrix_to_batch(sample[: max(20_000, n // max(1, threads))])
    idx = oracle.OracleIndex(index_path)
    t0 = time.perf_counter()
    out = idx.query_batch(a, off, streaming=streaming)
    sec = time.perf_counter() - t0
    return {"value": out.size / sec, "unit": "lookups/s", "cores": 1, "kind": "port",
            "sample": f"first {off.size - 1} reads on 1 thread of the C oracle port (oracle/_ref not present)", "seconds": sec,
            "checksum": int(out.sum()), "hits": int((out >= 0).sum()), "lookups": int(out.size)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the workload's)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--e2e-chunk-bases", type=int, default=48_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    n_reads = args.reads or w["reads"]
    L = args.read_len
    k = w["k"]
    config = {"workload": f"{args.workload}: {w['desc']}", "reads_per_gpu": n_reads, "read_len": L, "k": k, "precalc_k": 8,
              "path": "streaming_search" if w["streaming"] else "search", "index": "replicated per GPU",
              "l2_policy": "inputs+outputs (>= 11 GB per step at full size) exceed L2; the index itself stays L2/HBM resident across steps as in production"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        path, ref = ensure_index(args.workload, w)
        threads = os.cpu_count() or 1
        # each step = one bounded sample; shrink it when K + W is large so the whole arm stays within minutes
        n_s = int(min(n_reads, 2_000_000, max(100_000, 24_000_000 // (max(1, args.warmup) + max(1, args.steps)))))
        reads = synth.sample_reads(ref, n_s, L, 0.5, seed=43, both_strands=w["rc"])
        best = None
        for i in range(max(1, args.warmup) + max(1, args.steps)):
            r = cpu_reference(path, reads, w["streaming"], threads)
            if i >= max(1, args.warmup) and (best is None or r["value"] > best["value"]):
                best = r
        line = {"impl": "reference", "metric": "kmer_lookups_per_s", "value": best["value"], "unit": "lookups/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
                "cpu_baseline": {k_: best[k_] for k_ in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": best["value"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch

    import sbwt_b200 as S

    if not torch.cuda.is_available() or S.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the SBWT GPU query path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    if rank == 0:
        path, ref = ensure_index(args.workload, w)
    barrier()
    if rank != 0:
        path, ref = ensure_index(args.workload, w)

    t0 = time.time()
    reads = synth.sample_reads(ref, n_reads, L, 0.5, seed=43 + rank, both_strands=w["rc"])
    a, off = synth.matrix_to_batch(reads)
    log(f"rank {rank}: {n_reads} reads generated in {time.time() - t0:.1f}s")
    mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
    idx = S.Index(path, device=local_rank)
    ses = S.Session(idx, a.size, n_reads)
    n_out = ses.count_outputs(off)
    d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
    d_out = torch.empty(n_out, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out, stream)

    stats = ses.query_device_counted(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out, stream)
    launches_per_step = stats.kernel_launches
    hits_gpu = int((d_out >= 0).sum().item())
    assert stats.lookups == n_out and stats.hits == hits_gpu
    # spot parity inside the bench: first reads against the oracle (checker only, outside any timed region)
    if rank == 0:
        import oracle
        m = min(2000, n_reads)
        want = oracle.OracleIndex(path).query_batch(a[: m * L], off[: m + 1], streaming=w["streaming"])
        got = d_out[: want.size].cpu().numpy()
        assert np.array_equal(got, want), "GPU output differs from the oracle"

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ses.set_timing(True)
    walk_ms, prep_ms = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    S.launch_count(reset=True)
    barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(args.steps):
            step()
            p_ms, w_ms = ses.last_timing()
            prep_ms.append(p_ms)
            walk_ms.append(w_ms)
        e1.record()
        torch.cuda.synchronize()
        barrier()
    gpu_launches = S.launch_count()
    ses.set_timing(False)
    ms_total = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * n_out / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ e2e through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        chunk_bases = min(args.e2e_chunk_bases, a.size)
        ses_h = S.Session(idx, chunk_bases, chunk_bases // max(1, L - 2) + 16)
        h_a, h_off, h_out = S.pinned_empty(a.size, np.uint8), S.pinned_empty(off.size, np.int64), S.pinned_empty(n_out, np.int64)
        h_a[:], h_off[:] = a, off
        ses_h.query_host(h_a, h_off, mode, out=h_out)  # warm-up (allocates the pipeline slots)
        assert np.array_equal(h_out[: 120 * 50], d_out[: 120 * 50].cpu().numpy())
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ses_h.query_host(h_a, h_off, mode, out=h_out)
        sec = (time.perf_counter() - t0) / n_e2e
        if dist is not None:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        e2e = {"value": world * n_out / sec, "unit": "lookups/s", "h2d_bytes_per_step": int(a.nbytes + off.nbytes),
               "d2h_bytes_per_step": int(n_out * 8), "ms_per_step": sec * 1e3, "steps": n_e2e,
               "api": "sbwt_gpu_query_host (pinned host buffers, int64 results: what SBWT::streaming_search returns)",
               "widen_threads": ses_h.widen_threads(), "host_threads": os.cpu_count()}
        sparse_wire = ses_h.widen_threads() > 0 and os.environ.get("SBWT_B200_WIRE", "sparse") != "dense"
        sparse_bytes = int(n_out // 8 + 4 * hits_gpu + n_out // 1024)  # hit masks + the hits + block bases
        if sparse_wire:
            e2e["result_wire_format"] = ("sparse: one hit bit per result + the hits only (int32) cross PCIe; host threads rebuild the caller's "
                                         "int64 array (host_widen.hpp; SBWT_B200_WIRE=dense sends every result as int32 instead)")
            e2e["d2h_bytes_per_step"] = sparse_bytes
        elif ses_h.widen_threads() > 0:
            e2e["result_wire_format"] = ("dense: int32 over PCIe, sign-extended into the caller's int64 array by host threads "
                                         "(host_widen.hpp; SBWT_B200_WIDEN_THREADS, default = hardware threads / visible GPUs, <= 8)")
            e2e["d2h_bytes_per_step"] = int(n_out * 4)
        else:
            e2e["result_wire_format"] = "int64 over PCIe"

        def timed_leg(session, fn_name, out_buf):
            getattr(session, fn_name)(h_a, h_off, mode, out=out_buf)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                getattr(session, fn_name)(h_a, h_off, mode, out=out_buf)
            s_ = (time.perf_counter() - t0) / n_e2e
            if dist is not None:
                t_ = torch.tensor([s_], device="cuda", dtype=torch.float64)
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                s_ = float(t_.item())
            return s_

        if ses_h.widen_threads() > 0 and world == 1:
            # the same call with the int64 values themselves crossing PCIe (no host threads involved)
            prev_env = os.environ.get("SBWT_B200_WIDEN_THREADS")
            os.environ["SBWT_B200_WIDEN_THREADS"] = "0"
            ses_d = S.Session(idx, chunk_bases, chunk_bases // max(1, L - 2) + 16)
            h_out[:4096] = -7
            sec_d = timed_leg(ses_d, "query_host", h_out)
            os.environ.pop("SBWT_B200_WIDEN_THREADS")
            if prev_env is not None:
                os.environ["SBWT_B200_WIDEN_THREADS"] = prev_env
            assert np.array_equal(h_out[: 120 * 50], d_out[: 120 * 50].cpu().numpy())
            e2e["int64_over_pcie"] = {"value": world * n_out / sec_d, "unit": "lookups/s", "d2h_bytes_per_step": int(n_out * 8),
                                      "ms_per_step": sec_d * 1e3, "api": "sbwt_gpu_query_host with SBWT_B200_WIDEN_THREADS=0"}
            ses_d.close()
        if idx.n_nodes < (1 << 31) and world == 1:
            # the same call with int32 results (same values; half the PCIe bytes of the result copy, which bounds e2e)
            h_out32 = S.pinned_empty(n_out, np.int32)
            ses_h.query_host_i32(h_a, h_off, mode, out=h_out32)
            assert np.array_equal(h_out32[: 120 * 50].astype(np.int64), h_out[: 120 * 50])
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                ses_h.query_host_i32(h_a, h_off, mode, out=h_out32)
            sec32 = (time.perf_counter() - t0) / n_e2e
            if dist is not None:
                t = torch.tensor([sec32], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec32 = float(t.item())
            e2e["int32_results"] = {"value": world * n_out / sec32, "unit": "lookups/s", "d2h_bytes_per_step": sparse_bytes if sparse_wire else int(n_out * 4),
                                    "ms_per_step": sec32 * 1e3, "api": "sbwt_gpu_query_host_i32"}
            del h_out32
        ses_h.close()

    # the same kernel with the search table of round r01a..r01m (10 characters): more dependent sector reads per lookup,
    # slower per lookup, but a higher sector rate -- reported next to the headline so the two can be told apart
    alt = None
    if world == 1 and idx.table_length > 10 and idx.precalc_k <= 10:
        tp_default = idx.table_length
        try:
            idx.set_table_length(10)
            st10 = ses.query_device_counted(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out, stream)
            ses.set_timing(True)
            t10 = []
            for _ in range(3 + min(args.steps, 10)):
                step()
                t10.append(ses.last_timing()[1])
            alt = {"table_length": 10, "kernel_ms": float(np.mean(t10[3:])), "algorithmic_sectors_per_step": int(st10.index_sectors),
                   "rank_ops_per_step": int(st10.rank_ops)}
        except Exception as e:  # informational leg: never let it take the bench line down
            log(f"shorter-table leg skipped: {e}")
            alt = None
        finally:
            ses.set_timing(False)
            idx.set_table_length(tp_default)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    w_ms = float(np.mean(walk_ms))
    sector_bytes = stats.index_sectors * 32
    achieved = sector_bytes / (w_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                # DRAM bytes actually moved per launch (ncu) / this run's kernel time / peak: how much of HBM the kernel uses
                "traffic_frac": (traffic / (w_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "kernel": "walk2_kernel",
                "kernel_ms": w_ms, "kernel_share_of_step": w_ms / ms_per_step, "prep_ms": float(np.mean(prep_ms)),
                "algorithmic_sectors_per_step": stats.index_sectors, "sectors_per_s": stats.index_sectors / (w_ms * 1e-3),
                "rank_ops_per_step": stats.rank_ops, "rank_ops_per_s": stats.rank_ops / (w_ms * 1e-3),
                "nominal_frac": stats.rank_ops * 32 / (w_ms * 1e-3) / 1e9 / peak,
                "table_length": int(idx.table_length),
                # sector bytes + the bytes the walk must also move: 8 B per result written, the packed reads and work items read
                "io_bytes_per_step": int(n_out * 8 + a.size * 3 // 8 + n_reads * 16),
                "frac_with_io": (sector_bytes + n_out * 8 + a.size * 3 // 8 + n_reads * 16) / (w_ms * 1e-3) / 1e9 / peak}
    if alt:
        alt["achieved"] = alt["algorithmic_sectors_per_step"] * 32 / (alt["kernel_ms"] * 1e-3) / 1e9
        alt["frac"] = alt["achieved"] / peak
        alt["lookups_per_s_kernel"] = n_out / (alt["kernel_ms"] * 1e-3)
        alt["note"] = ("same kernel, shorter search table: each from-scratch search does 3 more dependent interval steps (more sector reads, "
                       "higher sector rate, fewer lookups/s); the default table trades sector rate for lookups/s")
        roofline["shorter_table"] = alt
    if not args.no_probe:
        try:
            dram = S.sector_probe(local_rank, 8 << 30, 1 << 28, 32)
            l2 = S.sector_probe(local_rank, 48 << 20, 1 << 28, 32)
            # the structure the walk reads at random on every step: the rank sectors in use (a row of the search table is read
            # once per from-scratch search and is not part of this buffer)
            hot = 32 * (idx.n_nodes // 96 + 1) if (idx.compact_layout[0] and w["streaming"]) else 128 * (idx.n_nodes // 224 + 1)
            same = S.sector_probe(local_rank, max(1 << 20, hot), 1 << 28, 32)
            roofline["random_sector_ceiling"] = {"dram_sectors_per_s": dram, "l2_sectors_per_s": l2, "index_sized_buffer_sectors_per_s": same,
                                                 "frac_of_index_sized_ceiling": roofline["sectors_per_s"] / same,
                                                 "index_sized_buffer_bytes": int(hot),
                                                 "how": "sbwt_gpu_sector_probe: 2^28 independent uniformly random 32-byte loads over 8 GiB / 48 MiB / a buffer the size of the rank sectors in use",
                                                 "note": "an L2 miss moves a whole 128-byte line (4 sectors) from HBM whatever cudaLimitMaxL2FetchGranularity says, so "
                                                         "the DRAM figure x 128 B is ~94 % of the HBM copy peak; ~62 MB of L2 are usable for randomly read data "
                                                         "(profiles/r01g_l2_capacity.txt)"}
        except Exception as e:  # the probe is informational
            roofline["random_sector_ceiling"] = {"error": str(e)}

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_reference(path, reads, w["streaming"])
        cpu = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": "kmer_lookups_per_s", "value": value, "unit": "lookups/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": config, "clocks": clk.summary(), "e2e": e2e,
            "gpu_launches": int(gpu_launches), "launches_per_step": int(launches_per_step), "roofline": roofline, "cpu_baseline": cpu,
            "lookups_per_step_per_gpu": int(n_out), "hit_rate": hits_gpu / n_out, "index_device_bytes": int(idx.device_bytes),
            "n_nodes": int(idx.n_nodes),
            "index_layout": ({"kind": "compact one-hot csectors (96 columns x 4 characters per 32-byte sector) + classic sectors for flagged blocks",
                              "flagged_block_fraction": idx.compact_layout[1], "hot_bytes": int(32 * (idx.n_nodes // 96 + 1))}
                             if (idx.compact_layout[0] and w["streaming"]) else
                             {"kind": "classic sectors (224 columns x 1 character per 32-byte sector)",
                              "flagged_block_fraction": idx.compact_layout[1], "hot_bytes": int(128 * (idx.n_nodes // 224 + 1))})}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — BASELINE config 2 on B200: uniform-random batched incr building ~1.5 B nnz over
13 M rows, then 500 M point gets at a 50 % hit rate (SURVEY.md 8d "C2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A *step* is one batch of `--batch` (2^26) incr ops per GPU.
The timed steps are always the LAST K batches of the 2 B-op stream: with the default K = 30 that
is the whole build; with a smaller K the first batches are applied untimed ("prefill") so the
timed region still ends at ~1.5 B nnz — the largest, least cache-friendly table.

  value      incr Mops/s, all K batches resident in HBM before the timed region starts
  e2e        the same build through the C-ABI with HOST (pinned) buffers: the H2D copy of every
             batch and the D2H read of the control block are inside the timed region
  roofline   dominant kernel k_upsert: 108 algorithmic bytes/op / launch time
             (CUDA events on the library's stream) against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref) on the box's host cores, bounded prefix
  --impl reference   the reference arm: the same stream through the reference's own C API with
             pthreads (src/smatrix_benchmark.c:98-132 shape), each step a bounded sample
N > 1 (torchrun): rows are hash-partitioned by owner rank; every rank routes its slice of each
batch with an all-to-all over NCCL and updates its own shard (weak scaling: per-GPU work fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = 13_000_000
YCOLS = 256
TOTAL_OPS = 2_000_000_000
TOTAL_GETS = 500_000_000
SEED_BUILD, SEED_GET = 2, 3
INCR_BYTES, GET_BYTES = 108, 76          # algorithmic bytes per op, SURVEY.md 8(d)
INCR_SECTORS, GET_SECTORS = 3, 2         # random 32 B sectors per op


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=1 << 26, help="incr ops per step per GPU")
    p.add_argument("--rows", type=int, default=ROWS, help="rows per GPU")
    p.add_argument("--ycols", type=int, default=YCOLS)
    p.add_argument("--total-ops", type=int, default=TOTAL_OPS, help="build ops per GPU")
    p.add_argument("--gets", type=int, default=TOTAL_GETS, help="point gets per GPU")
    p.add_argument("--arena-gib", type=int, default=52,
                   help="slab memory each matrix reserves at smatrix_open (SMATRIX_ARENA_GIB); 0 = on demand")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-probes", action="store_true")
    p.add_argument("--phase-series", action="store_true", help="diagnostic: host phase split of every step on stderr")
    p.add_argument("--cpu-sample", type=int, default=16_000_000, help="ops in the CPU baseline sample")
    p.add_argument("--ref-step", type=int, default=2_000_000, help="ops per step of the reference arm")
    return p.parse_args()


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                 str(index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        sel = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or \
              [r for (_, r) in self.rows[-3:] if len(r) >= 7]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in sel)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in sel)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(sel[0][1]), "reasons": reasons,
                "samples": len(sel), "power_w_max": max(float(r[2]) for r in sel)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ reference arm
def cpu_reference_run(n_steps, warmup, step_ops, rows, ycols, threads_list):
    """The unmodified reference (oracle/_ref, else the C restatement) on the host cores: one
    matrix, cumulative prefix of the C2 stream, `step_ops` per step.  Warm-up steps double as
    the thread-count sweep (the reference scales negatively)."""
    from oracle import cpu
    cpu.build(ref=True)
    kind = "reference" if cpu.have_reference() else "port"
    m = cpu.CpuMatrix(kind)
    if kind == "port":
        threads_list = [1]           # the restatement has no locks
    first, sweep = 0, {}
    for w in range(max(warmup, len(threads_list))):
        t = threads_list[w % len(threads_list)]
        s = m.bench_c2_incr(t, SEED_BUILD, first, step_ops, rows, ycols)
        sweep.setdefault(t, []).append(step_ops / s / 1e6)
        first += step_ops
    best_t = max(sweep, key=lambda t: max(sweep[t]))
    secs = 0.0
    for _ in range(n_steps):
        secs += m.bench_c2_incr(best_t, SEED_BUILD, first, step_ops, rows, ycols)
        first += step_ops
    incr_mops = n_steps * step_ops / secs / 1e6
    gets = min(step_ops * 2, 4_000_000)
    gs = m.bench_c2_get(best_t, SEED_GET, SEED_BUILD, 0, gets, first, rows, ycols)
    m.close()
    return {"kind": kind, "cores": best_t, "host_cores": os.cpu_count(), "incr_mops": incr_mops,
            "get_mops": gets / gs / 1e6, "secs": secs,
            "sweep_mops": {str(t): round(max(v), 3) for t, v in sweep.items()},
            "sample": f"C2 stream prefix: {first} incr ops cumulative on one matrix "
                      f"({n_steps} timed steps x {step_ops} ops after {first - n_steps * step_ops} warm-up ops), "
                      f"then {gets} gets; rows={rows} ycols={ycols}"}


def thread_candidates():
    n = os.cpu_count() or 1
    c = [1, 2, 4, 8, 16, 32]
    return [t for t in c if t <= n] or [1]


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(a.steps, a.warmup, a.ref_step, a.rows * a.gpus, a.ycols, thread_candidates())
    line = {
        "impl": "reference", "metric": "incr_mops_c2", "value": r["incr_mops"], "unit": "Mops/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": r["secs"] / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "C2 uniform incr build (SURVEY.md 8d), reference CPU path, bounded sample",
                   "rows": a.rows * a.gpus, "ycols": a.ycols, "ops_per_step": a.ref_step},
        "get_mops": r["get_mops"],
        "cpu_baseline": {"value": r["incr_mops"], "unit": "Mops/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "host_cores": r["host_cores"], "sweep_mops": r["sweep_mops"]},
        "e2e": {"value": r["incr_mops"], "unit": "Mops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def main_ours(a):
    import torch
    import torch.distributed as dist
    from libsmatrix_b200 import SparseMatrix

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.pop("SMATRIX_ARENA_GIB", None)

    def with_arena(make):
        """Only matrices that hold the table reserve the slab arena (helper handles do not)."""
        def wrapped():
            if a.arena_gib:
                os.environ["SMATRIX_ARENA_GIB"] = str(a.arena_gib)
            try:
                return make()
            finally:
                os.environ.pop("SMATRIX_ARENA_GIB", None)
        return wrapped
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)   # all-to-all under the update
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B, K, W = a.batch, a.steps, a.warmup
    rows_total = a.rows * world
    n_batches = max(K, -(-a.total_ops // B))            # batches in the per-GPU stream
    prefill = n_batches - K

    if world > 1:
        from libsmatrix_b200.sharded import ShardedSparseMatrix
        class _Unordered(ShardedSparseMatrix):   # the C2 stream never writes column 0: order-free
            def incr_batch(self, xs, ys, vals=None):
                super().incr_batch(xs, ys, vals, ordered=False)
        mk = with_arena(lambda: _Unordered(rank, world, local))
    else:
        mk = with_arena(lambda: SparseMatrix(device=local))

    sampler = ClockSampler(local)
    ibuf = lambda n: torch.empty(n, dtype=torch.int32, device=dev)

    # ---- warm-up on a scratch matrix (W full-size batches from a different seed)
    gen = SparseMatrix(device=local) if world > 1 else None
    scratch = mk()
    g = gen or scratch
    wx, wy = ibuf(B), ibuf(B)
    for w in range(W):
        g.gen_c2_ops(99, (rank * W + w) * B, B, rows_total, a.ycols, wx.data_ptr(), wy.data_ptr())
        scratch.incr_batch(wx, wy, None)
    wq = scratch.get_batch(wx, wy)
    del wq
    scratch.close()

    m = mk()
    if world > 1:
        m.reserve_route(B)      # symmetric inboxes + IPC exchange: communicator-style set-up, untimed
    g = gen or m
    # global op index of (rank, batch k, offset i): ranks interleave batch-wise
    first_of = lambda k: (k * world + rank) * B

    # ---- prefill (untimed, measured separately)
    t_prefill = 0.0
    for k in range(prefill):
        g.gen_c2_ops(SEED_BUILD, first_of(k), B, rows_total, a.ycols, wx.data_ptr(), wy.data_ptr())
        barrier()
        t0 = time.perf_counter()
        m.incr_batch(wx, wy, None)
        barrier()
        t_prefill += time.perf_counter() - t0
    del wx, wy

    # ---- timed build: all K batches resident in HBM first
    xs = [ibuf(B) for _ in range(K)]
    ys = [ibuf(B) for _ in range(K)]
    for j in range(K):
        g.gen_c2_ops(SEED_BUILD, first_of(prefill + j), B, rows_total, a.ycols,
                     xs[j].data_ptr(), ys[j].data_ptr())
    m.set_kernel_timing(True)
    launches0, rounds0 = m.stat("launches"), m.stat("rounds")
    barrier()
    wall0 = time.time()
    m.timer_start()
    step_ms, kern_ms, prev_k = [], [], 0
    PHASES = ("partition", "upsert", "grow_plan", "slab", "migrate", "dir")
    prev_ph = [0.0] * len(PHASES) + [m.stat("rounds"), m.stat("launches")]
    for j in range(K):
        t_s = time.perf_counter()
        m.incr_batch(xs[j], ys[j], None)        # synchronous: returns when the device is done
        step_ms.append(round((time.perf_counter() - t_s) * 1e3, 2))
        k_now = m.stat("kernel_ns")
        kern_ms.append(round((k_now - prev_k) / 1e6, 2))
        prev_k = k_now
        if a.phase_series and rank == 0:
            now = [m.stat("ns_" + k) / 1e6 for k in PHASES] + [m.stat("rounds"), m.stat("launches")]
            print(f"step {j}: {step_ms[-1]} ms; " + ", ".join(f"{k} {now[i] - prev_ph[i]:.2f}" for i, k in enumerate(PHASES))
                + f"; rounds {now[-2] - prev_ph[-2]}, launches {now[-1] - prev_ph[-1]}", file=sys.stderr)
            prev_ph = now
    ms_build = m.timer_stop_ms()
    barrier()
    wall1 = time.time()
    ms_build = max_over_ranks(ms_build)
    launches = m.stat("launches") - launches0
    rounds = m.stat("rounds") - rounds0
    upsert_ns = m.stat("kernel_ns")
    phases = {k: round(m.stat("ns_" + k) / 1e6 / K, 3) for k in
              ("partition", "upsert", "grow_plan", "slab", "migrate", "dir")}
    m.set_kernel_timing(False)
    clocks = sampler.window(wall0, wall1)
    incr_mops = K * B * world / (ms_build * 1e-3) / 1e6
    del xs, ys
    nnz_local, rows_local = m.stat("nnz"), m.stat("rows")
    vsum_local = m.stat("value_sum")     # every op added 1: the table-wide sum must equal the op count
    stats = {k: m.stat(k) for k in ("dir_cap", "slab_bytes", "device_bytes", "row_grows", "dir_grows")}

    # ---- timed gets (50 % hits): all queries resident first
    n_build = n_batches * B * world
    G = a.gets
    gstep = B
    qx, qy, out = ibuf(G), ibuf(G), ibuf(G)
    for off in range(0, G, gstep):
        cnt = min(gstep, G - off)
        g.gen_c2_queries(SEED_GET, SEED_BUILD, rank * G + off, cnt, n_build, rows_total, a.ycols,
                         qx[off:].data_ptr(), qy[off:].data_ptr())
    m.set_kernel_timing(True)
    barrier()
    m.timer_start()
    for off in range(0, G, gstep):
        cnt = min(gstep, G - off)
        m.get_batch(qx[off:off + cnt], qy[off:off + cnt], out[off:off + cnt])
    ms_get = max_over_ranks(m.timer_stop_ms())
    barrier()
    get_ns = m.stat("kernel_ns")
    m.set_kernel_timing(False)
    get_mops = G * world / (ms_get * 1e-3) / 1e6
    hits = int((out != 0).sum().item())
    odd_hits = int((out[1::2] != 0).sum().item()) if (rank * G) % 2 == 0 else int((out[0::2] != 0).sum().item())
    del qx, qy, out

    # ---- read path on the same table (BASELINE config 4 shape): rowlen over all rows, getrow of 2 M rows
    reads = None
    if world == 1:
        ids = ((torch.arange(a.rows, device=dev, dtype=torch.int64) * 2654435761) & 0xFFFFFFFF)
        ids = torch.where(ids >= 2**31, ids - 2**32, ids).to(torch.int32)
        m.rowlen_batch(ids[:1 << 20])                                   # warm
        m.timer_start(); rl = m.rowlen_batch(ids); ms_rl = m.timer_stop_ms()
        sample = ids[: min(a.rows, 1 << 21)].contiguous()
        import ctypes as _C
        offs = torch.empty(sample.numel() + 1, dtype=torch.int64, device=dev)
        lib, h = m._lib, m._handle()
        total = int(lib.smatrix_getrow_batch(h, sample.data_ptr(), sample.numel(), offs.data_ptr(), None, 0))
        pairs = torch.empty(2 * total, dtype=torch.int32, device=dev)
        m.timer_start()
        got = int(lib.smatrix_getrow_batch(h, sample.data_ptr(), sample.numel(), offs.data_ptr(), pairs.data_ptr(), total))
        ms_gr = m.timer_stop_ms()
        reads = {"rowlen_mops": a.rows / (ms_rl * 1e-3) / 1e6, "rowlen_sum": int(rl.to(torch.int64).sum().item()),
                 "getrow_rows": sample.numel(), "getrow_pairs": got, "getrow_ms": ms_gr,
                 "getrow_gpairs_per_s": got / (ms_gr * 1e-3) / 1e9,
                 "getrow_algorithmic_gbs": (32 * sample.numel() + 16 * got) / (ms_gr * 1e-3) / 1e9,
                 "getrow_value_sum": int(pairs[1::2].to(torch.int64).sum().item())}
        del pairs, offs, ids, sample

    # ---- roofline probes in the same process (random 32 B sector reads / 4 B atomics, 32 GiB)
    probes = None
    if not a.no_probes and rank == 0:
        pm = gen or m
        foot = 32 << 30
        acc = 1 << 30
        probes = {"footprint_gib": 32,
                  "random_read_32B_per_s": pm.probe_random_read(foot, acc, 32),
                  "random_read_8B_per_s": pm.probe_random_read(foot, acc, 8),
                  "random_atomic_4B_per_s": pm.probe_random_atomic(foot, acc)}
    m.close()

    # ---- e2e: the same build through host (pinned) buffers
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, torch, dev, mk, gen, B, K, prefill, n_batches, rank, world, barrier, max_over_ranks)

    if world > 1:
        t = torch.tensor([nnz_local, rows_local, vsum_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        nnz_total, rows_seen, vsum_total = int(t[0].item()), int(t[1].item()), int(t[2].item())
    else:
        nnz_total, rows_seen, vsum_total = nnz_local, rows_local, vsum_local
    applied = n_batches * B * world
    checks = {"value_sum": vsum_total, "ops_applied": applied, "value_sum_ok": vsum_total == applied,
              "hit_fraction_exact": (odd_hits == 0 and hits == G - G // 2) if world == 1 else (odd_hits == 0),
              "rows_ok": rows_seen <= rows_total}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        r = cpu_reference_run(4, 0, a.cpu_sample // 8, a.rows, a.ycols, thread_candidates())
        cpu = {"value": r["incr_mops"], "unit": "Mops/s", "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"], "host_cores": r["host_cores"], "get_mops": r["get_mops"],
               "sweep_mops": r["sweep_mops"]}

    sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak_gbs()
    upsert_launches = max(rounds, 1)
    ops_per_launch = K * B / upsert_launches
    ach = INCR_BYTES * K * B / (upsert_ns * 1e-9) / 1e9 if upsert_ns else None
    traffic = None
    try:   # dram bytes per op of k_upsert from the committed ncu capture (profiles/), scaled to one launch
        with open(os.path.join(ROOT, "profiles", "traffic_r1.json")) as f:
            traffic = json.load(f)["upsert_dram_bytes_per_op"] * ops_per_launch
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_upsert<INCR>", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_op": INCR_BYTES, "ops_per_launch": ops_per_launch,
                "launches": upsert_launches, "avg_launch_ms": upsert_ns / upsert_launches / 1e6 if upsert_ns else None,
                "kernel_share_of_step": upsert_ns / 1e6 / ms_build if upsert_ns else None}
    get_ach = GET_BYTES * G / (get_ns * 1e-9) / 1e9 if get_ns else None
    roofline["get"] = {"kernel": "k_get", "achieved": get_ach, "frac": (get_ach / peak) if get_ach else None,
                       "algorithmic_bytes_per_op": GET_BYTES}
    if probes:
        r32 = probes["random_read_32B_per_s"]
        roofline["random_sector"] = {
            "R32_sectors_per_s": r32, "atomic_4B_per_s": probes["random_atomic_4B_per_s"],
            "read_8B_per_s": probes["random_read_8B_per_s"], "footprint_gib": probes["footprint_gib"],
            "incr_frac": INCR_SECTORS * (K * B / (upsert_ns * 1e-9)) / r32 if upsert_ns else None,
            "get_frac": GET_SECTORS * (G / (get_ns * 1e-9)) / r32 if get_ns else None,
            "note": "achieved = ops/s x algorithmic random sectors per op (3 incr, 2 get) / measured random 32 B read rate"}
    line = {
        "metric": "incr_mops_c2", "value": incr_mops, "unit": "Mops/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_build / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "C2: uniform-random batched incr, 2^26 ops/step/GPU, building ~1.5 B nnz over 13 M rows per GPU, "
                               "then point gets at 50 % hits (SURVEY.md 8d)",
                   "rows": rows_total, "ycols": a.ycols, "ops_per_step": B * world, "timed_ops": K * B * world,
                   "prefill_ops": prefill * B * world, "gets": G * world,
                   "l2": "inputs larger than L2 (512 MiB of keys per step, table >> 126 MB)",
                   "arena_gib": a.arena_gib, "arena_note": "slab arena reserved by smatrix_open (outside the timed region); "
                   "on-demand cudaMalloc is the fallback and costs 0.3-8 ms/step on this pool",
                   "chunk_ops": int(os.environ.get("SMATRIX_CHUNK", 1 << 25)),
                   "parallelism": f"row-hash shard x{world}" if world > 1 else "single GPU"},
        "get_mops": get_mops, "get_ms": ms_get, "get_hit_fraction": hits / G, "checks": checks,
        "nnz": nnz_total, "rows_present": rows_seen, "prefill_s": t_prefill,
        "table": stats, "clocks": clocks, "gpu_launches": launches, "upsert_rounds": rounds,
        "host_phase_ms_per_step": phases, "step_ms": step_ms, "step_upsert_kernel_ms": kern_ms,
        "roofline": roofline,
    }
    if reads:
        line["reads"] = reads
    if e2e:
        line["e2e"] = e2e
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(a, torch, dev, make_matrix, gen, B, K, prefill, n_batches, rank, world, barrier, max_over_ranks):
    """Same stream, HOST buffers: each timed step is one incr_batch(host arrays) call per rank —
    H2D of the batch, (N > 1: the route,) the update, and the D2H reads of the control block.
    Per step: barrier, wall clock around the call, max over ranks."""
    rows_total = a.rows * world
    first_of = lambda k: (k * world + rank) * B
    dx = torch.empty(B, dtype=torch.int32, device=dev)
    dy = torch.empty(B, dtype=torch.int32, device=dev)
    hx = torch.empty(B, dtype=torch.int32, pin_memory=True)
    hy = torch.empty(B, dtype=torch.int32, pin_memory=True)
    scratch = make_matrix()                      # warm-up of the host-pointer path (W steps, other seed)
    g = gen or scratch
    for w in range(a.warmup):
        g.gen_c2_ops(98, (rank * a.warmup + w) * B, B, rows_total, a.ycols, dx.data_ptr(), dy.data_ptr())
        hx.copy_(dx); hy.copy_(dy)
        torch.cuda.synchronize()
        scratch.incr_batch(hx, hy, None)
    scratch.close()
    m = make_matrix()
    if world > 1:
        m.reserve_route(B)
    g = gen or m
    for k in range(prefill):
        g.gen_c2_ops(SEED_BUILD, first_of(k), B, rows_total, a.ycols, dx.data_ptr(), dy.data_ptr())
        m.incr_batch(dx, dy, None)
    rounds0 = m.stat("rounds")
    secs, series = 0.0, []
    for j in range(K):
        g.gen_c2_ops(SEED_BUILD, first_of(prefill + j), B, rows_total, a.ycols, dx.data_ptr(), dy.data_ptr())
        hx.copy_(dx); hy.copy_(dy)
        barrier()
        t0 = time.perf_counter()
        m.incr_batch(hx, hy, None)               # returns after the device finished (synchronous API)
        dt = max_over_ranks(time.perf_counter() - t0)
        series.append(round(dt * 1e3, 2))
        secs += dt
    rounds = m.stat("rounds") - rounds0
    incr = K * B * world / secs / 1e6
    # gets through host buffers: queries up, values down
    G = min(a.gets, 4 * B)
    n_build = n_batches * B * world
    hq = torch.empty(B, dtype=torch.int32, pin_memory=True)
    hr = torch.empty(B, dtype=torch.int32, pin_memory=True)
    gsecs, done = 0.0, 0
    while done < G:
        cnt = min(B, G - done)
        g.gen_c2_queries(SEED_GET, SEED_BUILD, rank * G + done, cnt, n_build, rows_total, a.ycols,
                         dx.data_ptr(), dy.data_ptr())
        hx.copy_(dx); hq.copy_(dy)
        barrier()
        t0 = time.perf_counter()
        m.get_batch(hx[:cnt], hq[:cnt], hr[:cnt])
        gsecs += max_over_ranks(time.perf_counter() - t0)
        done += cnt
    m.close()
    return {"value": incr, "unit": "Mops/s", "h2d_bytes_per_step": 8 * B * world,
            "d2h_bytes_per_step": 1072 * max(1, rounds // max(K, 1)) * world, "ms_per_step": secs / K * 1e3, "step_ms": series,
            "get_mops": G * world / gsecs / 1e6, "get_h2d_bytes_per_step": 8 * B * world, "get_d2h_bytes_per_step": 4 * B * world,
            "note": "pinned host arrays through incr_batch / get_batch (N = 1: the C-ABI smatrix_incr_batch / smatrix_get_batch "
                    "with host pointers; N > 1: the sharded API stages each rank's slice, then routes); barrier, wall clock "
                    "around the call, max over ranks"}


def _json_only_stdout():
    """Everything that C libraries print to fd 1 (NCCL's version banner, ...) goes to stderr; the
    returned file object is the real stdout for the ONE JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    args = parse()
    _REAL_STDOUT = _json_only_stdout()
    _print = print

    def print(*a, **k):     # noqa: A001 - only the JSON line is printed with flush=True below
        k.setdefault("file", _REAL_STDOUT)
        _print(*a, **k)
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)

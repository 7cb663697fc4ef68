#!/usr/bin/env python
"""bench.py — feature rows/s of the feature_generation hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R]

One "step" = one pass of the whole hot path (19 raw features + 19 rolling z-scores for every snapshot row)
over one synthetic instrument-day resident in HBM.  Workload at N = 1: BASELINE.json configs[1], a single
instrument-day of 50 M snapshots x 10 levels (m = 10 ticks and ~2 trades per snapshot).  For N > 1 every
rank owns one such instrument-day (the reference shards by instrument-day, scripts/multidate_feature_to_csv.sh:24):
weak scaling, no data-path collective; the only collective is the NCCL all-reduce of the 19x3 global column
moments after the pass (and the timing barrier).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over ranks);
`e2e` goes through the C ABI with pinned HOST buffers (H2D + pass + D2H inside the timed region); the record also carries the
bytes and GB/s per direction, `ceiling` (the SAME pinned buffers and byte counts moved by plain copies, no kernels: what the
host's wire allows at this N), `frac_of_copy_ceiling`, and `norm_only` (the call with raw = NULL; not the headline).  For N > 1
the line also carries `timeshard`: ONE stream of N x 50 M rows split by time with the halo exchange inside the step.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified reference
compiled from /root/reference; falls back to the oracle port) on the host cores.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "feature_rows_per_sec"
UNIT = "rows/s"
LEVELS = 10
TICKS_PER_ROW = 10
FULL_ROWS = 50_000_000


def algorithmic_bytes_per_row(levels: int, m: int, q: float) -> float:
    """inputs read once + 19 raw + 19 normalised doubles written once."""
    return 340 + 24 * levels + 16 * m + q


def ncu_traffic_per_row():
    """DRAM bytes per row and kernel from the committed `ncu --set full` captures (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, m = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = max(mx, m)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(l.split(",")[1]) for _, l in self.lines[-3:] if len(l.split(",")) > 2] or [0.0]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_arm(args, rank: int, world: int):
    """The reference's own CPU implementation on the host cores (bounded sample of the same workload)."""
    if rank != 0:
        return
    import numpy as np

    import mrp_b200
    from mrp_b200 import synth
    from oracle import binding as ob

    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    rows_each = args.ref_rows
    # one independent instrument-day prefix per host thread: the reference is single-threaded per stream and
    # parallel across instrument-days (scripts/multidate_feature_to_csv.sh:24,52-66)
    specs = [synth.SynthSpec(n_rows=rows_each, levels=LEVELS, ticks_per_row=TICKS_PER_ROW, seed=0x5EED + i) for i in range(threads)]
    batch = synth.to_numpy(synth.make_batch(specs))
    hv = mrp_b200.HostView(batch)
    use_ref = ob.ref_available(LEVELS)
    kind = "reference" if use_ref else "port"

    def one():
        t0 = time.perf_counter()
        if use_ref:
            ob.run_reference(hv, None, use_book=False, nthreads=threads)
        else:
            ob.run_oracle(hv, nthreads=threads, want_dirs=False, want_counts=False)
        return time.perf_counter() - t0

    for _ in range(max(0, args.warmup)):
        one()
    times = [one() for _ in range(max(1, args.steps))]
    sec = sum(times) / len(times)
    rows = hv.n_rows
    value = rows / sec
    sample = f"{threads} streams x {rows_each} rows (prefix of the 50 M-row instrument-day shape), {threads} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": max(0, args.warmup), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"single instrument-day, {FULL_ROWS} snapshots x {LEVELS} levels (BASELINE.json configs[1])",
                   "levels": LEVELS, "ticks_per_row": TICKS_PER_ROW, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def cpu_baseline(levels: int, budget_rows: int):
    """Bounded CPU sample timed beside the GPU number (rank 0, N = 1 only)."""
    import mrp_b200
    from mrp_b200 import synth
    from oracle import binding as ob

    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    specs = [synth.SynthSpec(n_rows=budget_rows, levels=levels, ticks_per_row=TICKS_PER_ROW, seed=0x5EED + i) for i in range(threads)]
    hv = mrp_b200.HostView(synth.to_numpy(synth.make_batch(specs)))
    use_ref = ob.ref_available(levels)
    t0 = time.perf_counter()
    if use_ref:
        ob.run_reference(hv, None, use_book=False, nthreads=threads)
    else:
        ob.run_oracle(hv, nthreads=threads, want_dirs=False, want_counts=False)
    sec = time.perf_counter() - t0
    return {"value": hv.n_rows / sec, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": f"{threads} streams x {budget_rows} rows, one host thread per stream, {sec:.1f} s"}


def config0_leg(events: int = 600_000):
    """BASELINE.json configs[0]: the reference's own features_to_csv pipeline on a seeded synthetic MBO day, unmodified
    (oracle/_ref/libmrp_refpipe.so) and with the INTEGRATION.md §3 patch (libmrp_gpupipe.so: GpuFeatureProcessor -> C ABI -> CUDA).
    Wall times of both, whole program (event decode, book, features, CSV).  Returns None when the two builds are not here."""
    import tempfile

    libs = {k: os.path.join(ROOT, "oracle", "_ref", f"libmrp_{k}pipe.so") for k in ("ref", "gpu")}
    if not all(os.path.exists(p) for p in libs.values()):
        return None
    out = {"events_per_instrument": events, "what": "features_to_csv end to end: synthetic MBO events -> order book -> features -> four CSV files"}
    with tempfile.TemporaryDirectory() as tmp:
        for k, lib in libs.items():
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_pipe.py"), lib, os.path.join(tmp, k), "--events", str(events)],
                               capture_output=True, text=True)
            if r.returncode != 0:
                return None
            info = json.loads(r.stdout.strip().splitlines()[-1])
            if info["rc"] != 0:
                return None
            out["seconds_reference" if k == "ref" else "seconds_patched_gpu"] = info["seconds"]
    out["note"] = ("the event loop (decode, order book, FeatureEngine deques) is the reference's own host code in both; only the feature stage differs, "
                   "and at this size the GPU context creation is a visible part of the patched run")
    return out


def timeshard_leg(args, rank: int, world: int, dev, rows_per_gpu: int, steps: int, warmup: int):
    """BASELINE.json configs[4]: ONE stream of rows_per_gpu * world snapshots split by time.  Every rank generates only its own
    rows (range-restartable generator), binds them with headroom for the left neighbour's tail, and per step: packs its tail
    (mrp_fg_halo_export), ships it as ONE message over NVLink (NCCL send / recv), unpacks the neighbour's (mrp_fg_halo_import)
    and runs the pass.  Returns the sub-record (identical on every rank)."""
    import torch
    import torch.distributed as dist

    import mrp_b200
    from mrp_b200 import synth
    from mrp_b200 import timeshard as tsh

    total = rows_per_gpu * world
    spec = synth.SynthSpec(n_rows=total, levels=LEVELS, ticks_per_row=TICKS_PER_ROW, seed=0x5EED)
    r0, r1 = tsh.shard_bounds(total, world, rank)
    shard = synth.make_stream_range(spec, r0, r1, device=dev)
    shard.pop("trade_size", None)
    eng = mrp_b200.Engine(mrp_b200.default_params(), device=dev.index)
    ts = tsh.TimeShard(eng, shard, rank, world)
    del shard
    torch.cuda.empty_cache()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        ts.exchange()
        eng.run(stream)
        ts.carry_poison()
    barrier()
    launches0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    ev[0].record()
    sent = 0
    for i in range(steps):
        sent = ts.exchange()
        ev[2 * i + 1].record()
        eng.run(stream)
        ts.carry_poison()   # the normaliser's "poisoned for ever" state crosses shard boundaries too: one all-gather of 57 integers per rank
        ev[2 * i + 2].record()
    barrier()
    ms = ev[0].elapsed_time(ev[-1]) / steps
    ex_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)) / steps
    launches = eng.launch_count() - launches0
    t = torch.tensor([ms, ex_ms, float(sent)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ex_ms, sent_max = (float(x) for x in t.tolist())
    # the one collective of the path: global column moments of the z-scores over every rank's OWN rows
    own = ts.own_rows
    nm = torch.from_numpy(eng.norm_moments(stream, rows=own)).to(dev)
    dist.all_reduce(nm, op=dist.ReduceOp.SUM)
    rec = {
        "rows_total": total, "rows_per_gpu": rows_per_gpu, "ms_per_step": ms, "value": total / (ms * 1e-3), "unit": UNIT,
        "exchange_ms": ex_ms, "nvlink_bytes_per_rank_per_step": int(sent_max), "halo_rows": int(eng.halo_rows()),
        "halo_extra_work": eng.halo_rows() / rows_per_gpu, "gpu_launches": launches, "steps": steps, "warmup": warmup,
        "global_rows_in_moments": float(nm[0, 0].item()),
        "sharding": ("ONE stream split by time; per step each rank packs its last %d rows (+ one window of ticks / trades) with mrp_fg_halo_export, "
                     "ships them to its right neighbour as ONE NCCL message over NVLink, unpacks the left neighbour's with mrp_fg_halo_import and "
                     "recomputes them as halo; one all-reduce of the column moments of the own rows" % eng.halo_rows()),
    }
    eng.close()
    del ts
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=FULL_ROWS, help="snapshot rows per GPU (default: the 50 M-row config)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows for the host-buffer end-to-end leg (0 = same as --rows if host memory allows)")
    ap.add_argument("--ref-rows", type=int, default=200000, help="rows per stream of the CPU sample (about 15 s per host thread)")
    ap.add_argument("--timeshard", action="store_true", help="BASELINE.json configs[4]: ONE stream of rows*N snapshots split by time across the N ranks (halo over NVLink)")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="rows per pipelined chunk of the end-to-end leg (0 = library default)")
    ap.add_argument("--levels", type=int, default=LEVELS, help="book depth (50 = the deep-book config, BASELINE.json configs[3])")
    ap.add_argument("--days", type=int, default=1, help="instrument-days per GPU in one batch (BASELINE.json configs[2]: 32 days over the ranks, round-robin)")
    ap.add_argument("--vol-of-vol", action="store_true",
                    help="also run the vol-of-vol extension over 3 horizons (600 / 3000 / 30000 rows) inside every step (BASELINE.json configs[3]; no reference counterpart)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-timeshard", action="store_true", help="N > 1: skip the time-shard sub-record (one stream of rows*N snapshots with the halo exchange)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import mrp_b200
    from mrp_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = _bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic instrument-day, generated on the device (never crosses PCIe)
    rows = args.rows
    step_pre = None   # per-step work ahead of the kernels (time-sharded mode: the halo exchange)
    skip = 0
    if args.timeshard and world > 1:
        # ONE stream of rows*world snapshots split by time: handled by timeshard_leg() below (the instrument-day leg is skipped)
        line = timeshard_leg(args, rank, world, dev, rows, max(args.steps, 1), max(args.warmup, 3))
        if rank == 0:
            peak, peak_src = measured_peaks()
            bytes_row = algorithmic_bytes_per_row(LEVELS, TICKS_PER_ROW, 2.0)
            gbs = line["rows_total"] / world * bytes_row / (line["ms_per_step"] * 1e-3) / 1e9
            _emit({"metric": METRIC, "value": line["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                   "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                   "config": {"workload": "ONE stream of %d snapshots x %d levels split by time over %d GPUs (BASELINE.json configs[4])" % (line["rows_total"], LEVELS, world),
                              "rows_per_gpu": rows, "levels": LEVELS, "sharding": line["sharding"], "l2": "inputs far exceed the 126 MB L2; no explicit flush"},
                   "roofline": {"bound": "hbm", "kernel": "whole pass incl. halo exchange", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "traffic": None, "peak_source": peak_src},
                   "cpu_baseline": None, "e2e": None, "gpu_launches": line["gpu_launches"], "timeshard": line})
        dist.destroy_process_group()
        return
    if True:
        # `--days D`: D independent instrument-days of `rows` snapshots each in one batch on this rank (day d of the job's
        # D*world days lives on rank d mod world, scripts/multidate_feature_to_csv.sh:24 -> one GPU per instrument-day)
        specs = [synth.SynthSpec(n_rows=rows, levels=args.levels, ticks_per_row=TICKS_PER_ROW, seed=0x5EED + rank + world * d)
                 for d in range(args.days)]
        spec = specs[0]
        batch = synth.make_batch(specs, device=dev)
        rows = rows * args.days
    n_ticks, n_trades = batch["tick_offsets"][-1], batch["trade_offsets"][-1]
    q = n_trades / (batch["row_offsets"][-1])
    m_eff = n_ticks / (batch["row_offsets"][-1])
    torch.cuda.synchronize()

    eng = mrp_b200.Engine(mrp_b200.default_params(depth_levels=args.levels), device=local_rank)
    eng.bind_device(batch)
    stream = torch.cuda.current_stream().cuda_stream
    _run = eng.run

    VOV_HORIZONS = (600, 3000, 30000)

    def run_step(st):
        if step_pre is not None:
            step_pre()
        _run(st)
        if args.vol_of_vol:
            eng.vol_of_vol(VOV_HORIZONS, stream=st, download=False)

    eng.run = run_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        eng.run(stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = eng.launch_count()
    eng.enable_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        eng.run(stream)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    # per-kernel CUDA-event durations (same stream): mean over a second pass of the same steps, read back step by step
    kt = {}
    for _ in range(args.steps):
        eng.run(stream)
        torch.cuda.synchronize()
        for k, v in eng.kernel_times_ms().items():
            kt[k] = kt.get(k, 0.0) + v / args.steps
    eng.enable_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = rows * world / (ms_per_step * 1e-3)

    # ---- the one collective on this path: global column moments (19 x {n, sum, sum^2})
    mom = torch.from_numpy(eng.moments(stream)).to(dev)   # (time-sharded mode: halo rows are counted twice; diagnostic only)
    if world > 1:
        dist.all_reduce(mom, op=dist.ReduceOp.SUM)
    mom = mom.cpu().numpy()

    # ---- what the Python stage does next (StandardScaler over the z-scored columns of all dates): moments of the z-scores,
    # ONE all-reduce of 19 x 3 doubles, then every rank scales its own rows.  Outside the timed region; reported on its own.
    nm = torch.from_numpy(eng.norm_moments(stream)).to(dev)
    if world > 1:
        dist.all_reduce(nm, op=dist.ReduceOp.SUM)
    nm = nm.cpu().numpy()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.standardize(nm, stream, download=False)
    sv0.record()
    eng.standardize(nm, stream, download=False)
    sv1.record()
    torch.cuda.synchronize()
    scaler = {"standardize_ms": sv0.elapsed_time(sv1), "moments_combined_over_ranks": world,
              "mean_scale_log_return": [float(x) for x in mrp_b200.mean_scale(nm)[mrp_b200.FEATURES.index("log_return")]]}

    # ---- roofline of the dominant kernel (live CUDA events on the launching stream, last timed step)
    peak, peak_src = measured_peaks()
    bytes_row = algorithmic_bytes_per_row(args.levels, m_eff, q) + (48 if args.vol_of_vol else 0)   # + 3 x (raw + z) doubles
    # algorithmic bytes per kernel: what each kernel must read + write once
    kb = {
        # mids + spreads + trade dirs + both cursors, 8 feature columns
        "windows": rows * (16 * m_eff + q + 8 + 8 * 8),
        # ts, prices, sizes, buy/sell, adds, 11 feature columns
        "book": rows * (8 + 24 * args.levels + 16 + 4 + 11 * 8),
        # raw read once, z written once
        "normalize": rows * (2 * 19 * 8),
    }
    dom = max(kt, key=kt.get)
    whole_gbs = rows * bytes_row / (ms_per_step * 1e-3) / 1e9
    tr = ncu_traffic_per_row()
    traffic = tr["dram_bytes_per_row"][dom] * rows if tr and dom in tr.get("dram_bytes_per_row", {}) else None
    roof = {
        "bound": "hbm", "kernel": dom, "achieved": kb[dom] / (kt[dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": kb[dom] / (kt[dom] * 1e-3) / 1e9 / peak, "traffic": traffic,
        "traffic_source": (tr["source"] + " (per-row DRAM bytes of that capture x rows of this run)") if traffic else None,
        "algorithmic_bytes": kb[dom], "peak_source": peak_src,
        "whole_pass": {"bytes_per_row": bytes_row, "achieved": whole_gbs, "frac": whole_gbs / peak},
        "kernels_ms": kt,
        "kernels_gbs": {k: kb[k] / (kt[k] * 1e-3) / 1e9 for k in kt},
    }

    # ---- end to end through the C ABI with pinned host buffers
    e2e = None
    if not args.no_e2e and not (args.timeshard and world > 1) and args.days == 1 and args.levels == LEVELS:
        e2e_rows = args.e2e_rows or rows
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        need = e2e_rows * (algorithmic_bytes_per_row(LEVELS, m_eff, q) + 64)
        while need > 0.6 * avail / max(1, min(world, 8)) and e2e_rows > 1_000_000:
            e2e_rows //= 2
            need = e2e_rows * (algorithmic_bytes_per_row(LEVELS, m_eff, q) + 64)
        del eng
        if e2e_rows == rows:
            hb = batch
        else:
            del batch
            torch.cuda.empty_cache()
            hb = synth.make_batch([dataclass_replace(spec, n_rows=e2e_rows)], device=dev)
        host = {}
        for k, v in hb.items():
            if isinstance(v, torch.Tensor):
                if k == "trade_size":
                    continue
                host[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                host[k].copy_(v)
            else:
                host[k] = v
        del hb
        if e2e_rows == rows:
            del batch
        torch.cuda.empty_cache()
        hv = mrp_b200.HostView({k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in host.items()})
        raw_h = torch.empty((19, e2e_rows), dtype=torch.float64, pin_memory=True)
        norm_h = torch.empty((19, e2e_rows), dtype=torch.float64, pin_memory=True)
        eng2 = mrp_b200.Engine(mrp_b200.default_params(), device=local_rank)
        import ctypes as C

        def e2e_step():
            # the call a host pipeline makes: pinned host arrays in, pinned host features out; chunks of the stream are
            # uploaded / computed / downloaded on two alternating lanes so PCIe runs both ways while the SMs compute
            eng2.process_host(hv, raw_ptr=raw_h.data_ptr(), norm_ptr=norm_h.data_ptr(), chunk_rows=args.e2e_chunk)

        e2e_step()
        barrier()
        nst = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(nst):
            e2e_step()
        ev1.record()
        barrier()
        # mrp_fg_process_host is synchronous (it returns when the last chunk has landed in host memory): the wall clock and
        # the CUDA events around it agree; take the larger
        e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3 * 0.999) / nst
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        h2d = sum(v.numel() * v.element_size() for k, v in host.items() if isinstance(v, torch.Tensor))
        d2h = raw_h.numel() * 8 * 2
        e2e = {"value": e2e_rows * world / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "rows_per_gpu": e2e_rows, "ms_per_step": e_ms,
               "h2d_gbs": h2d / e_ms * 1e-6, "d2h_gbs": d2h / e_ms * 1e-6,
               "checksum": float(norm_h[0, :1000].sum().item())}
        eng2.close()
        # the wire's own ceiling: the SAME pinned buffers and byte counts moved by plain copies, no kernels —
        # both directions at once (what the pipeline does), then host-to-device alone
        def copy_loop(both: bool) -> float:
            piece = 256 << 20
            d_in = torch.empty(piece, dtype=torch.uint8, device=dev)
            d_out = torch.empty(piece, dtype=torch.uint8, device=dev)
            s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            srcs = [v.view(torch.uint8).reshape(-1) for k, v in host.items() if isinstance(v, torch.Tensor) and v.numel()]
            dsts = [raw_h.view(torch.uint8).reshape(-1), norm_h.view(torch.uint8).reshape(-1)]
            torch.cuda.synchronize()
            barrier()
            c0 = time.perf_counter()
            with torch.cuda.stream(s_in):
                for v in srcs:
                    for off in range(0, v.numel(), piece):
                        k = min(piece, v.numel() - off)
                        d_in[:k].copy_(v[off:off + k], non_blocking=True)
            if both:
                with torch.cuda.stream(s_out):
                    for v in dsts:
                        for off in range(0, v.numel(), piece):
                            k = min(piece, v.numel() - off)
                            v[off:off + k].copy_(d_out[:k], non_blocking=True)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - c0) * 1e3
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        # a consumer of the z-scored half only (the reference's Python stage reads `_norm.csv`, classifier.py:15-18): the raw matrix
        # stays on the device, half the device-to-host bytes
        eng3 = mrp_b200.Engine(mrp_b200.default_params(), device=local_rank)
        eng3.process_host(hv, norm_ptr=norm_h.data_ptr(), chunk_rows=args.e2e_chunk)
        barrier()
        n0 = time.perf_counter()
        eng3.process_host(hv, norm_ptr=norm_h.data_ptr(), chunk_rows=args.e2e_chunk)
        n_ms = (time.perf_counter() - n0) * 1e3
        if world > 1:
            t = torch.tensor([n_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_ms = float(t.item())
        eng3.close()
        e2e["norm_only"] = {"value": e2e_rows * world / (n_ms * 1e-3), "ms_per_step": n_ms, "d2h_bytes_per_step": d2h // 2,
                            "what": "mrp_fg_process_host with raw = NULL (not the headline: the reference emits both matrices)"}
        copy_loop(True)
        both_ms, h2d_ms = copy_loop(True), copy_loop(False)
        e2e["ceiling"] = {"both_directions_ms": both_ms, "h2d_only_ms": h2d_ms, "h2d_only_gbs": h2d / h2d_ms * 1e-6,
                          "both_h2d_gbs": h2d / both_ms * 1e-6, "both_d2h_gbs": d2h / both_ms * 1e-6,
                          "what": "the same pinned buffers and byte counts moved by plain cudaMemcpyAsync in 256 MB pieces on two streams, no kernels (max over ranks)"}
        e2e["frac_of_copy_ceiling"] = both_ms / e_ms
        del host, hv, raw_h, norm_h   # 37 GB of pinned memory per rank: released before anything else is staged

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if _ALL_CPUS:   # the CPU baseline gets every host core the launcher gave this process, not just the GPU's neighbours
            os.sched_setaffinity(0, _ALL_CPUS)
        cpu = cpu_baseline(args.levels, args.ref_rows)
    config0 = config0_leg() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    tshard = None
    if world > 1 and not args.no_timeshard and args.days == 1 and args.levels == LEVELS:
        # the interconnect-bearing curve (BASELINE.json configs[4]) beside the instrument-day value: free the instrument-day first
        try:
            del eng
        except NameError:
            pass
        try:
            del batch
        except NameError:
            pass
        torch.cuda.empty_cache()
        tshard = timeshard_leg(args, rank, world, dev, args.rows, max(2, min(args.steps, 3)), 3)
        tshard["efficiency_vs_instrument_day_rate"] = tshard["value"] / value   # value = N x this box's per-GPU instrument-day rate

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.days} instrument-day(s) per GPU, {rows // args.days} snapshots x {args.levels} levels each, {m_eff:.2f} ticks and {q:.2f} trades per snapshot"
                                   + (" (BASELINE.json configs[1])" if (rows == FULL_ROWS and args.days == 1 and args.levels == LEVELS) else " (not the headline config)")
                                   + (" + vol-of-vol over %s rows (extension, no reference counterpart)" % (VOV_HORIZONS,) if args.vol_of_vol else ""),
                       "rows_per_gpu": rows, "levels": args.levels, "days_per_gpu": args.days,
                       "sharding": "instrument-day per GPU, no data-path collective (the `timeshard` sub-record is the one stream split by time)",
                       "l2": "inputs (>= 20 GB at 50 M rows) far exceed the 126 MB L2; no explicit flush"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "host_affinity": numa, "global_standardisation": scaler,
            "global_moments_midprice": [float(x) for x in mom[0]], "timeshard": tshard, "config0": config0,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def dataclass_replace(spec, **kw):
    import dataclasses

    return dataclasses.replace(spec, **kw)


_ALL_CPUS = None


def _bind_to_gpu_numa_node(gpu_index: int):
    """One process per GPU: run on the CPUs next to that GPU, so that the pinned staging buffers of the end-to-end leg are
    first-touched in the memory its PCIe root complex hangs off (host plumbing; returns what was done, for the record)."""
    try:
        import pynvml

        global _ALL_CPUS
        _ALL_CPUS = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return {"cpus": f"{cpus[0]}-{cpus[-1]}", "n": len(cpus)}
    except Exception as e:  # not fatal: the benchmark just runs wherever the launcher put it
        return {"error": str(e)[:80]}


def _emit(line: dict):
    """The one JSON line, on the process's REAL stdout (see _quiet_stdout)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout():
    """Libraries print to fd 1 on their own (NCCL's version banner, torchrun notices): everything but the JSON line goes to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    main()

"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import numpy as np

import mrp_b200
from mrp_b200 import synth

# Tolerances (BASELINE.json north_star / BASELINE.md "Parity tolerances"):
RAW_RTOL = 1e-9        # raw floating-point features: 1e-9 relative ...
RAW_ATOL_SCALE = 1e-13  # ... plus a tiny absolute floor (x column max |value|) for cells that cross zero (ofi, stress)
NORM_TOL = 1e-9        # z-scores: |dz| <= 1e-9 * max(1,|z|) * conditioning allowance (see norm_allowance)


def make_host_batch(specs):
    return synth.to_numpy(synth.make_batch(specs))


def nan_pattern_equal(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))


def raw_errors(got, ref):
    """Per-column max of |got-ref| / (RAW_RTOL*|ref| + atol_c); <= 1 passes."""
    out = []
    for c in range(ref.shape[0]):
        g, r = got[c], ref[c]
        fin = np.isfinite(r)
        if not fin.any():
            out.append(0.0)
            continue
        scale = np.abs(r[fin]).max()
        tol = RAW_RTOL * np.abs(r[fin]) + RAW_ATOL_SCALE * max(scale, 1e-300)
        out.append(float((np.abs(g[fin] - r[fin]) / tol).max()))
    return np.array(out)


def norm_allowance(raw_ref, window=30000):
    """Conditioning of each z-score column: kappa = E[x^2]/Var over the rows.
    A relative perturbation eps of the window sums moves z by about eps*kappa*|z|/2 + eps*sqrt(kappa)."""
    k = []
    for c in range(raw_ref.shape[0]):
        x = raw_ref[c][np.isfinite(raw_ref[c])]
        v = x.var() if x.size else 0.0
        k.append(float((x * x).mean() / v) if v > 0 else np.inf)
    return np.array(k)


def norm_errors(got, ref):
    """Per-column max |dz| / max(1,|z|)."""
    out = []
    for c in range(ref.shape[0]):
        fin = np.isfinite(ref[c])
        if not fin.any():
            out.append(0.0)
            continue
        out.append(float((np.abs(got[c][fin] - ref[c][fin]) / np.maximum(1.0, np.abs(ref[c][fin]))).max()))
    return np.array(out)


def run_gpu(batch, params=None, mode="upload"):
    hv = mrp_b200.HostView(batch)
    p = params or mrp_b200.default_params(emit_dirs=1, emit_counts=1)
    eng = mrp_b200.Engine(p, device=0)
    if mode == "upload":
        eng.upload(hv)
    else:
        import torch

        dev = {}
        for k, v in hv.arrays.items():
            if k.endswith("_offsets"):
                dev[k] = v
            else:
                t = torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else (v.view(np.int32) if v.dtype == np.uint32 else v))
                if mode == "device_offset":
                    # every array starts one element past an allocation boundary: bases are only 8 / 4 / 1-byte aligned
                    flat = torch.empty(t.numel() + 1, dtype=t.dtype, device="cuda")
                    flat[1:].copy_(t.reshape(-1))
                    dev[k] = flat[1:].view(t.shape)
                else:
                    dev[k] = t.cuda()
        eng.bind_device(dev)
    eng.run()
    out = eng.download()
    eng.close()
    return hv, out


def exact_norm(raw, row_offsets, window=30000):
    """The z-scores the reference WOULD produce if its running sums did not drift: window sums from
    x87 extended-precision prefix sums rounded once to binary64, then the reference's own binary64
    formula (feature_normalizer.cpp:61-69).  Also returns the per-cell conditioning
    kappa = E[x^2]/Var of the window: a 1-ulp change of the sums moves z by about eps*kappa*max(1,|z|)."""
    ld = np.longdouble
    out = np.empty_like(raw)
    kappa = np.empty_like(raw)
    for s in range(len(row_offsets) - 1):
        r0, r1 = int(row_offsets[s]), int(row_offsets[s + 1])
        n = r1 - r0
        if n == 0:
            continue
        t = np.arange(n)
        lo = np.maximum(t + 1 - window, 0)          # window = rows [lo, t]
        cnt = (t + 1 - lo).astype(ld)
        for c in range(raw.shape[0]):
            x = raw[c, r0:r1].astype(ld)
            p1 = np.concatenate([[ld(0)], np.cumsum(x)])
            p2 = np.concatenate([[ld(0)], np.cumsum(x * x)])
            s1 = p1[t + 1] - p1[lo]
            s2 = p2[t + 1] - p2[lo]
            s1d, s2d, nd = s1.astype(np.float64), s2.astype(np.float64), cnt.astype(np.float64)
            xd = raw[c, r0:r1]
            with np.errstate(all="ignore"):
                mean = s1d / nd
                var = (s2d / nd) - (mean * mean)
                std = np.where(var > 0.0, np.sqrt(np.where(var > 0.0, var, 1.0)), 1.0)
                out[c, r0:r1] = (xd - mean) / std
                vx = (s2 / cnt - (s1 / cnt) ** 2).astype(np.float64)
                kappa[c, r0:r1] = np.where(vx > 0, (s2d / nd) / np.where(vx > 0, vx, 1.0), np.inf)
    return out, kappa


def norm_check(got_norm, ref_raw, ref_norm, row_offsets, window, z_tol=NORM_TOL, first_row=0):
    """Conditioning-aware z-score comparison.  Returns (worst_a, worst_b, drift) per column:
       a: |gpu - exact| / tol_cell                      tol_cell = max(1,|z|) * (z_tol + 16*eps*kappa)
       b: |gpu - ref|   / (tol_cell + 1.5*|ref - exact|)  — the reference's own drift, cell by cell
    Cells whose window is constant (kappa = inf: the variance is pure rounding noise in every
    implementation, feature_normalizer.cpp:63-67) are only required to be finite."""
    exact, kappa = exact_norm(ref_raw, row_offsets, window)
    if first_row:   # rows before first_row are history only (a range of a long stream with its warm-up rows in front)
        exact, kappa, ref_norm, got_norm = exact[:, first_row:], kappa[:, first_row:], ref_norm[:, first_row:], got_norm[:, first_row:]
    eps = np.finfo(np.float64).eps
    ok = np.isfinite(exact) & np.isfinite(ref_norm) & np.isfinite(kappa) & (kappa < 1e9)
    mag = np.maximum(1.0, np.abs(exact))
    tol = mag * (z_tol + 16 * eps * np.where(ok, kappa, 0.0))
    da = np.where(ok, np.abs(got_norm - exact) / tol, 0.0)
    dref = np.where(ok, np.abs(ref_norm - exact), 0.0)
    db = np.where(ok, np.abs(got_norm - ref_norm) / (tol + 1.5 * dref), 0.0)
    drift = np.where(ok, dref / mag, 0.0)
    return da.max(axis=1), db.max(axis=1), drift.max(axis=1), ok.mean()


# ---- golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the compiled reference)
def golden_cases():
    import glob
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(here, "*.npz")))


def load_golden(name):
    """Returns (spec, host batch, fixture dict).  Regenerates the inputs from the stored spec and checks
    their digest, so a drifting generator cannot silently invalidate the fixture."""
    import hashlib
    import json
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(here, name + ".npz"))
    spec = synth.SynthSpec(**json.loads(str(z["spec"])))
    batch = synth.to_numpy(synth.make_batch([spec]))
    h = hashlib.sha256()
    for k in sorted(batch):
        if k == "trade_size":
            continue
        h.update(k.encode())
        h.update(np.ascontiguousarray(batch[k]).tobytes())
    assert h.hexdigest() == str(z["digest"]), "synthetic generator no longer reproduces the fixture inputs"
    return spec, batch, {k: z[k] for k in ("rows", "raw", "norm", "bid_dir", "ask_dir")}


# ---- adversarial edits of a synthetic host batch (single stream)
def drop_rows(batch, lo, hi):
    """Remove snapshot rows [lo, hi): the next remaining row sees a long jump of its tick / trade cursors."""
    out = dict(batch)
    n = len(batch["ts_ns"])
    keep = np.r_[0:lo, hi:n]
    for k in ("ts_ns", "buy_vol", "sell_vol", "adds", "tick_cursor", "trade_cursor"):
        out[k] = np.ascontiguousarray(batch[k][keep])
    for k in ("bid_px", "ask_px", "bid_sz", "ask_sz"):
        out[k] = np.ascontiguousarray(batch[k][:, keep])
    out["row_offsets"] = np.array([0, len(keep)], dtype=np.int64)
    return out


def poke(batch, key, index, value):
    out = dict(batch)
    a = np.array(batch[key], copy=True)
    a[index] = value
    out[key] = a
    return out


# ---- oracle over selected row ranges of ONE long stream (benchmark-sized parity)
ORACLE_WARM_ROWS = 30000 + 2048   # one normaliser window + the EMA / lag warm-up: every state of the path has forgotten its start


def substream(batch, e0, b, mid_window=1800, dir_window=1500):
    """Rows [e0, b) of a single-stream host batch as a stream of their own: the ticks / trades those rows can see, cursors
    rebased to the slice (finite memory: SURVEY.md §8e)."""
    kc, jc = batch["tick_cursor"].astype(np.int64), batch["trade_cursor"].astype(np.int64)
    wmax = max(mid_window, dir_window + 1)
    tb = max(0, int(kc[e0]) - wmax - 2) if e0 > 0 else 0
    jb = max(0, int(jc[e0]) - dir_window - 2) if e0 > 0 else 0
    t1, j1 = int(kc[b - 1]), int(jc[b - 1])
    out = {}
    for k in ("ts_ns", "buy_vol", "sell_vol", "adds"):
        out[k] = batch[k][e0:b]
    for k in ("bid_px", "ask_px", "bid_sz", "ask_sz"):
        out[k] = batch[k][:, e0:b]
    out["tick_cursor"] = (kc[e0:b] - tb).astype(np.uint32)
    out["trade_cursor"] = (jc[e0:b] - jb).astype(np.uint32)
    out["mid"], out["spread"] = batch["mid"][tb:t1], batch["spread"][tb:t1]
    out["trade_dir"] = batch["trade_dir"][jb:j1]
    return out


def oracle_on_windows(batch, windows, params, warm=ORACLE_WARM_ROWS, runner=None):
    """The checker on the row ranges `windows` = [(a, b), ...] of a single long stream, each restarted `warm` rows earlier
    (or at row 0), all ranges side by side on the host cores.  Returns one dict per range holding only its rows [a, b)."""
    from oracle import binding as ob

    subs, skips = [], []
    for a, b in windows:
        e0 = max(0, a - warm)
        subs.append(substream(batch, e0, b, params.mid_window, params.dir_window))
        skips.append(a - e0)
    cat = {}
    for k in ("ts_ns", "buy_vol", "sell_vol", "adds", "tick_cursor", "trade_cursor", "mid", "spread", "trade_dir"):
        cat[k] = np.concatenate([s[k] for s in subs])
    for k in ("bid_px", "ask_px", "bid_sz", "ask_sz"):
        cat[k] = np.ascontiguousarray(np.concatenate([s[k] for s in subs], axis=1))
    cat["row_offsets"] = np.cumsum([0] + [len(s["ts_ns"]) for s in subs])
    cat["tick_offsets"] = np.cumsum([0] + [len(s["mid"]) for s in subs])
    cat["trade_offsets"] = np.cumsum([0] + [len(s["trade_dir"]) for s in subs])
    hv = mrp_b200.HostView(cat)
    ref = (runner or ob.run_oracle)(hv, params)
    outs = []
    for i, (a, b) in enumerate(windows):
        r0 = int(cat["row_offsets"][i]) + skips[i]
        r1 = int(cat["row_offsets"][i + 1])
        o = {k: (None if v is None else (v[:, r0:r1] if v.ndim == 2 else v[r0:r1])) for k, v in ref.items()}
        # the conditioning-aware z check needs the raw rows of the window in front of every row
        o["raw_with_warm"] = ref["raw"][:, int(cat["row_offsets"][i]):r1]
        o["norm_with_warm"] = ref["norm"][:, int(cat["row_offsets"][i]):r1]
        o["skip"] = skips[i]
        outs.append(o)
    return outs


def parity_record(name, **fields):
    """Append one JSON line to the parity report (MRP_PARITY_REPORT, default gpurun_out/parity_report.jsonl when that
    directory exists): per-column plain max|dz|/max(1,|z|) vs the checker, kappa, the checker-vs-exact drift."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.environ.get("MRP_PARITY_REPORT")
    if not path:
        d = os.path.join(root, "gpurun_out")
        if not os.path.isdir(d):
            return
        path = os.path.join(d, "parity_report.jsonl")

    def clean(v):
        if isinstance(v, np.ndarray):
            return [clean(x) for x in v.tolist()]
        if isinstance(v, (np.floating, float)):
            return float(v) if np.isfinite(v) else str(v)
        if isinstance(v, (np.integer,)):
            return int(v)
        if isinstance(v, dict):
            return {k: clean(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [clean(x) for x in v]
        return v

    with open(path, "a") as f:
        f.write(json.dumps({"test": name, **clean(fields)}) + "\n")


def column_report(got_raw, got_norm, ref_raw, ref_norm, ref_raw_hist, ref_norm_hist, skip, window):
    """Per-column numbers a reader can check against north_star's '1e-9': raw max relative error, PLAIN max|dz|/max(1,|z|)
    against the checker, the checker's own distance from the exactly summed definition (its drift), and kappa."""
    exact, kappa = exact_norm(ref_raw_hist, [0, ref_raw_hist.shape[1]], window)
    exact, kappa = exact[:, skip:], kappa[:, skip:]
    rep = {}
    for c, name in enumerate(mrp_b200.FEATURES):
        r, g = ref_raw[c], got_raw[c]
        fin = np.isfinite(r) & (r != 0)
        rel = float((np.abs(g[fin] - r[fin]) / np.abs(r[fin])).max()) if fin.any() else 0.0
        zr, zg, ze = ref_norm[c], got_norm[c], exact[c]
        ok = np.isfinite(zr) & np.isfinite(ze) & np.isfinite(kappa[c]) & (kappa[c] < 1e9)
        mag = np.maximum(1.0, np.abs(zr))
        rep[name] = {
            "raw_bit_exact": bool(np.array_equal(g, r, equal_nan=True)),
            "raw_max_rel": rel,
            "z_bit_exact": bool(np.array_equal(zg, zr, equal_nan=True)),
            "z_max_abs_over_max1": float((np.abs(zg - zr)[ok] / mag[ok]).max()) if ok.any() else 0.0,
            "z_gpu_vs_exact": float((np.abs(zg - ze)[ok] / mag[ok]).max()) if ok.any() else 0.0,
            "z_ref_vs_exact": float((np.abs(zr - ze)[ok] / mag[ok]).max()) if ok.any() else 0.0,
            "kappa_median": float(np.median(kappa[c][ok])) if ok.any() else float("inf"),
            "kappa_max": float(kappa[c][ok].max()) if ok.any() else float("inf"),
            "cells_compared": float(ok.mean()),
        }
    return rep


def z_consistency(a_norm, b_norm, raw, row_offsets, window, raw_rel=0.0, z_tol=NORM_TOL, allow_drift=False):
    """Two evaluations of the SAME z-scores (two GPU runs that cut the stream differently, or the checker started at two
    different rows): worst |a - b| / tol_cell with the conditioning-aware cell tolerance of norm_check,
        tol_cell = max(1,|z|) * (z_tol + 16 eps kappa + raw_rel sqrt(kappa))
    raw_rel = relative difference of the raw values the two runs normalise (0 when they are bit-identical); with
    allow_drift the distance of each side from the exactly summed definition is added (running-sum implementations)."""
    exact, kappa = exact_norm(raw, row_offsets, window)
    eps = np.finfo(np.float64).eps
    ok = np.isfinite(exact) & np.isfinite(a_norm) & np.isfinite(b_norm) & np.isfinite(kappa) & (kappa < 1e9)
    k = np.where(ok, kappa, 0.0)
    tol = np.maximum(1.0, np.abs(exact)) * (z_tol + 16 * eps * k + raw_rel * np.sqrt(k))
    if allow_drift:
        tol = tol + 1.5 * (np.where(ok, np.abs(a_norm - exact), 0.0) + np.where(ok, np.abs(b_norm - exact), 0.0))
    return float(np.where(ok, np.abs(a_norm - b_norm) / tol, 0.0).max()), float(ok.mean())

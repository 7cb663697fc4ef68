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


def norm_check(got_norm, ref_raw, ref_norm, row_offsets, window, z_tol=NORM_TOL):
    """Conditioning-aware z-score comparison.  Returns (worst_a, worst_b, drift) per column:
       a: |gpu - exact| / tol_cell                      tol_cell = max(1,|z|) * (z_tol + 16*eps*kappa)
       b: |gpu - ref|   / (tol_cell + 1.5*|ref - exact|)  — the reference's own drift, cell by cell
    Cells whose window is constant (kappa = inf: the variance is pure rounding noise in every
    implementation, feature_normalizer.cpp:63-67) are only required to be finite."""
    exact, kappa = exact_norm(ref_raw, row_offsets, window)
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

import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import os; os.environ["MRP_FG_SEGMENT_ROWS"]="257"
import numpy as np, mrp_b200
from mrp_b200 import synth
from oracle import binding as ob
from helpers import make_host_batch, poke, run_gpu
b = make_host_batch([synth.SynthSpec(n_rows=4000, seed=33)])
K = len(b["mid"]); rng = np.random.default_rng(5); idx = rng.choice(K, size=40, replace=False)
b = poke(b, "mid", idx[:20], 0.0); b = poke(b, "spread", idx[20:30], np.nan); b = poke(b, "mid", idx[30:], np.nan)
p = mrp_b200.default_params(emit_dirs=1, emit_counts=1)
hv, got = run_gpu(b, p); ref = ob.run_oracle(hv, p)
for key in ("raw","norm"):
    for c,name in enumerate(mrp_b200.FEATURES):
        g, r = got[key][c], ref[key][c]
        bad = (np.isnan(g) != np.isnan(r)) | (np.isinf(g) != np.isinf(r))
        if bad.any():
            i = np.nonzero(bad)[0]
            print(key, name, len(i), i[:5], g[i[:5]], r[i[:5]], 'first nan raw ref', np.argmax(np.isnan(ref['raw'][c])) if np.isnan(ref['raw'][c]).any() else None, 'inf raw', np.nonzero(np.isinf(ref['raw'][c]))[0][:5], ref['raw'][c][np.isinf(ref['raw'][c])][:5])

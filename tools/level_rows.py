"""Row-length statistics of the AMG levels at a saved SsN state (development aid): python tools/level_rows.py states.npz tag"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
from amg_prof import load_state  # noqa: E402

pd, m, n = load_state(sys.argv[1], sys.argv[2])
Ae, f = ssnamg.rescaled_system(pd)
ssnamg.rng_reset(); ssnamg.rand(m + n)
levels = ssnamg.amg_setup(Ae, dict(ssnamg.driver.CLASS1_AMG_OPTIONS, fnode=n, isnsp=1))
for k, (A, P) in enumerate(levels):
    S = A.to_scipy().tocsr(); rl = np.diff(S.indptr)
    msg = f"level {k}: N={S.shape[0]} nnz={S.nnz} avg {rl.mean():.1f} max {rl.max()} p90 {int(np.percentile(rl, 90))} p99 {int(np.percentile(rl, 99))} " \
          f"rows>8: {(rl > 8).sum()} >16: {(rl > 16).sum()} >32: {(rl > 32).sum()} >64: {(rl > 64).sum()}"
    if k == 0:
        nf = n
        Sf = S[:nf]; offd = np.diff(Sf.indptr) - 1; Sc = S[nf:]; offc = np.diff(Sc.indptr) - 1
        msg += f" | off-diagonal entries per F row: avg {offd.mean():.1f} max {offd.max()} >8: {(offd > 8).sum()}; per C row: avg {offc.mean():.1f} max {offc.max()} >8: {(offc > 8).sum()}"
    print(msg)

#!/usr/bin/env python3
"""One Baum-Welch iteration beyond 32 states (general statistics kernels): left-to-right 200-state x 4-mixture x 39-dim model,
1000 utterances of 200-400 frames (the config-3 batch with a large model).  One GPU; CUDA events."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "ln-asr_b200"), ROOT):
    sys.path.insert(0, p)


def main():
    import torch
    import bench
    from lnasr.em import EmTrainer
    from oracle import lnasr_oracle as O
    n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 4)
    mdl = O.random_gmmhmm(n, m, 39, np.random.default_rng(7), left_to_right=True)
    x, off = bench.synth_feats(2000, 1000, 200, 400)
    tr = EmTrainer(mdl).load_batch(torch.from_numpy(x).cuda(), off)
    ll0 = float(tr.iterate())
    for _ in range(2):
        tr.iterate()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ll = tr.iterate()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    T = int(off[-1])
    print(f"EM iteration, {n} states x {m} mixtures, {T} frames: {ms:.3f} ms = {T / ms / 1e3:.1f} M frames/s; "
          f"log-likelihood of the first iteration {ll0:.3f}")


if __name__ == "__main__":
    main()

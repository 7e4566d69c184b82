#!/usr/bin/env python3
"""Signed error of the tensor-core emission kernels at the config-4 model scale (one GPU; prints a table -> profiles/).

The per-utterance log-probability is a sum of ~T emission values, so what decides the 1e-3 bar is the MEAN signed error per
frame (a bias grows like T, zero-mean rounding like sqrt(T)).  Features are real MFCCs of the bench's synthetic audio; the
reference values come from the float64 emission kernel (itself 1e-13 from the oracle, tests/test_gmm_gpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "ln-asr_b200"), ROOT):
    sys.path.insert(0, p)


def main():
    import torch
    import bench
    from lnasr.gmm import PackedGmm
    from lnasr.gmm_tc import PackedGmmTC
    from lnasr.mfcc import MFCC
    mdl = bench.make_models("c4")[0]
    pcm, off = bench.synth_batch(1000, 16, 48000, 48000, n_draw=1250)
    mb = MFCC().calc_mfcc_batch(pcm, sample_offsets=off)
    feats = mb.feats
    T = feats.shape[0]
    b64 = PackedGmm(mdl["w"], mdl["mu"], mdl["var"]).loglik(feats)[0]                 # [T, N] float64
    top = b64 >= (b64.max(dim=1, keepdim=True).values - 30.0)                           # the states that carry the probability mass
    print(f"# c4 model, {T} frames x {b64.shape[1]} states; |b| mean {b64.abs().mean():.1f}; {top.double().mean() * 100:.1f} % of the "
          "entries are within 30 nats of their frame's maximum ('top')")
    ideal = (b64.float().double() - b64)
    print(f"float32 rounding of the exact value alone: mean {ideal.mean():+.3e}  rms {ideal.pow(2).mean().sqrt():.3e}")
    for fmt in ("f16", "tf32"):
        g = PackedGmmTC(mdl["w"], mdl["mu"], mdl["var"], fmt=fmt)
        b = g.loglik(feats).double()
        e = b - b64
        print(f"{fmt:5s}: mean signed err {e.mean():+.3e}  rms {e.pow(2).mean().sqrt():.3e}  max |err| {e.abs().max():.3e}   "
              f"top states: mean {e[top].mean():+.3e} rms {e[top].pow(2).mean().sqrt():.3e}   -> bias x 300 frames = {300 * e[top].mean():+.2e}")
        # by magnitude of (b - state constant): is the bias proportional to the accumulated magnitude?
        mag = (b64 - b64.max(dim=1, keepdim=True).values).abs()
        for lo, hi in ((0, 10), (10, 30), (30, 100), (100, 1e9)):
            sel = (mag >= lo) & (mag < hi)
            if sel.any():
                print(f"        {lo:>4.0f} <= max - b < {hi:<6.0f}: mean {e[sel].mean():+.3e}  rms {e[sel].pow(2).mean().sqrt():.3e}  ({int(sel.sum())} entries)")


if __name__ == "__main__":
    main()

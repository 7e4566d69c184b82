#!/usr/bin/env python3
"""BASELINE configs[4]: MFCC front-end sweep -- 16 kHz synthetic audio, utterances of 1 s to 60 s, batch 1 to 4096;
frames/s, achieved GB/s of the algorithmic traffic (476 B/frame) against the measured HBM peak, and the FP64 ALU
fraction.  Writes a markdown table (stdout) -- committed under profiles/.

    python tools/mfcc_sweep.py > gpurun_out/mfcc_sweep.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ln-asr_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from lnasr.mfcc import MFCC

torch.cuda.set_device(0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
fp64 = bench.fp64_peaks()
mf = MFCC()
print("| seconds/utt | batch | frames | ms (CUDA events, best of 5) | frames/s | GB/s (476 B/frame) | % of HBM peak | FP64 TFLOP/s | % of DFMA peak |")
print("|---|---|---|---|---|---|---|---|---|")
rng = np.random.default_rng(1234)
max_samples = 400_000_000
for secs in (1, 5, 10, 30, 60):
    for batch in (1, 8, 64, 512, 4096):
        S = 16000 * secs
        if S * batch > max_samples:
            continue
        pcm = torch.from_numpy(np.clip(np.round(3000.0 * rng.standard_normal(S * batch, dtype=np.float32)), -32767, 32767).astype(np.int16)).cuda()
        off = np.arange(batch + 1, dtype=np.int64) * S
        best = 1e30
        for rep in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mb = mf.calc_mfcc_batch(pcm, sample_offsets=off)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                best = min(best, e0.elapsed_time(e1))
        frames = int(mb.frame_offsets_host[-1])
        fps = frames / (best * 1e-3)
        gbs = fps * 476 / 1e9
        tf = fps * bench.MFCC_FP64_FLOPS_PER_FRAME / 1e12
        print(f"| {secs} | {batch} | {frames} | {best:.3f} | {fps:.3e} | {gbs:.1f} | {100 * gbs / hbm:.2f} | {tf:.2f} | {100 * tf / fp64['fp64_fma_tflops']:.1f} |")
        del pcm, mb
print(f"\nHBM peak {hbm} GB/s (MEASURED_PEAKS.json); FP64 DFMA peak measured in this run: {fp64['fp64_fma_tflops']:.1f} TFLOP/s; "
      "includes the host-side offset arithmetic of calc_mfcc_batch (small batches are launch/host bound).")

#!/usr/bin/env python3
"""BASELINE configs[4]: MFCC front-end sweep -- 16 kHz synthetic audio, utterances of 1 s to 60 s, batch 1 to 4096;
frames/s, achieved GB/s of the algorithmic traffic (476 B/frame) against the measured HBM peak, and the FP64 ALU
fraction.  Writes a markdown table (stdout) -- committed under profiles/.

    python tools/mfcc_sweep.py > gpurun_out/mfcc_sweep.md
    python tools/mfcc_sweep.py --only 1x1        (one case, e.g. under ncu: the batch-1 latency corner)

The audio is generated on the device (noise, int16) so that the 60 s x 4096 corner (3.9 G samples, 7.9 GB) needs no host
memory; every case of BASELINE configs[4] from 1 s x 1 to 60 s x 4096 runs.
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
only = None
if "--only" in sys.argv:
    a, b_ = sys.argv[sys.argv.index("--only") + 1].split("x")
    only = (int(a), int(b_))
gen = torch.Generator(device="cuda").manual_seed(1234)


def device_noise(n):
    out = torch.empty((n,), dtype=torch.int16, device="cuda")
    step = 1 << 27
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m] = torch.randn((m,), device="cuda", generator=gen).mul_(3000.0).round_().clamp_(-32767, 32767).to(torch.int16)
    return out


for secs in (1, 5, 10, 30, 60):
    for batch in (1, 8, 64, 512, 4096):
        if only is not None and (secs, batch) != only:
            continue
        S = 16000 * secs
        pcm = device_noise(S * batch)
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
        torch.cuda.empty_cache()
print(f"\nHBM peak {hbm} GB/s (MEASURED_PEAKS.json); FP64 DFMA peak measured in this run: {fp64['fp64_fma_tflops']:.1f} TFLOP/s; "
      "includes the host-side offset arithmetic of calc_mfcc_batch (small batches are launch/host bound).")

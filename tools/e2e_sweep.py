#!/usr/bin/env python3
"""End-to-end (pinned host PCM -> state paths on the host) timing of config 4 over the grouping parameters of
AcousticPipeline.submit_host: number of utterance groups, compute streams, tapered or uniform groups; one batch at a time
and two batches in flight.  One GPU; CUDA events around 10 steps."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "ln-asr_b200"), ROOT):
    sys.path.insert(0, p)


def main():
    import torch
    import bench
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    cfg = bench.WORKLOADS[wl]
    models = bench.make_models(wl)
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="tc16" if wl == "c4" else "exact"))
    pcm_np, off = bench.synth_batch(1000, cfg["n_utt"], cfg["min_s"], cfg["max_s"])
    pcm_host = torch.from_numpy(pcm_np).pin_memory()
    pcm_dev = pcm_host.cuda()
    rp = pipe.bank.K == 1
    res, mb = pipe.recognize_batch(pcm_dev, off, want_paths=True)
    frames = int(mb.frame_offsets_host[-1])
    steps = 10

    def timed(fn, fin=None):
        for _ in range(3):
            fn()
        if fin:
            fin()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if fin:
            fin()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps

    ms, _ = timed(lambda: pipe.recognize_batch(pcm_dev, off, want_paths=True))
    print(f"{wl}: device-resident {ms:.3f} ms per {frames} frames")
    for chunks in (1, 2, 3, 4, 6, 8):
        for streams in (1, 2):
            for taper in (False, True):
                if taper and chunks < 3:
                    continue
                kw = dict(n_chunks=chunks, want_paths=True, n_compute_streams=streams, taper=taper, return_paths=rp)
                ms_sync, _ = timed(lambda: pipe.recognize_batch_host(pcm_host, off, **kw))
                pend = [None]

                def step():
                    nxt = pipe.submit_host(pcm_host, off, **kw)
                    if pend[0] is not None:
                        pend[0].result()
                        if rp:
                            pend[0].paths()
                    pend[0] = nxt

                def fin():
                    if pend[0] is not None:
                        pend[0].result()
                        pend[0] = None
                ms_pipe, wall = timed(step, fin)
                print(f"  chunks {chunks} streams {streams} taper {int(taper)}: one batch at a time {ms_sync:7.3f} ms   two in flight {ms_pipe:7.3f} ms "
                      f"(wall {wall:7.3f})   -> {frames / min(ms_sync, ms_pipe) / 1e3:.1f} M frames/s")


if __name__ == "__main__":
    main()

"""Probe: H2D bandwidth from pinned memory and host-side enqueue cost of the e2e path (not part of the product)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ln-asr_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from lnasr.bank import AcousticPipeline, ModelBank
from lnasr.mfcc import MFCC

torch.cuda.set_device(0)
x = torch.empty(130_000_000 // 2, dtype=torch.int16).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (1, 4, 8):
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step = x.numel() // n
        for i in range(n):
            d[i * step:(i + 1) * step].copy_(x[i * step:(i + 1) * step], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"H2D {x.numel()*2/1e6:.0f} MB in {n} copies: {ms:.3f} ms = {x.numel()*2/ms/1e6:.1f} GB/s")

models = bench.make_models("c2")
pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="exact"))
pcm_np, off = bench.synth_batch(1000, 4096, 8000, 24000)
pcm_host = torch.from_numpy(pcm_np).pin_memory()
for chunks in (1, 2, 4, 8):
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w, b = pipe.recognize_batch_host(pcm_host, off, n_chunks=chunks, want_paths=True)
        t1 = time.perf_counter()
    print(f"chunks {chunks}: wall {1e3*(t1-t0):.3f} ms")
# host enqueue cost of one device-resident step (no sync inside)
pcm_dev = pcm_host.cuda()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, mb = pipe.recognize_batch(pcm_dev, off, want_paths=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(f"device step: enqueue {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms")
import cProfile, pstats
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(5):
    pipe.recognize_batch_host(pcm_host, off, n_chunks=8, want_paths=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

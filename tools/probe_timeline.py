"""Diagnostic (not part of the product): per-group timeline of AcousticPipeline.recognize_batch_host at config 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ln-asr_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from lnasr.bank import AcousticPipeline, ModelBank
from lnasr.mfcc import MFCC

torch.cuda.set_device(0)
pipe = AcousticPipeline(MFCC(), ModelBank(bench.make_models("c2"), emission="exact"))
pcm_np, off = bench.synth_batch(1000, 4096, 8000, 24000)
pcm_host = torch.from_numpy(pcm_np).pin_memory()
for chunks, streams in ((6, 2), (6, 1), (12, 2), (3, 2)):
    for rep in range(4):
        tl = []
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        pipe.recognize_batch_host(pcm_host, off, n_chunks=chunks, want_paths=True, n_compute_streams=streams, _timeline=tl)
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); torch.cuda.synchronize()
    print(f"chunks={chunks} streams={streams}: total {e0.elapsed_time(e1):.3f} ms")
    prev = 0.0
    for i, (c, d) in enumerate(tl):
        tc, td = e0.elapsed_time(c), e0.elapsed_time(d)
        print(f"   group {i}: copy done {tc:.3f}  compute done {td:.3f}  (compute span <= {td - max(tc, prev):.3f})")
        prev = td if streams == 1 else prev

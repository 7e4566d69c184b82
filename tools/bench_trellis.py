#!/usr/bin/env python3
"""Time the banded Viterbi decode at the config-4 shape (1250 utterances x 299 frames x 2000 states, float32 emissions):
trellis_band2_kernel (fused backtrace) against trellis_band_kernel + backtrace_kernel.  One GPU; CUDA events."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "ln-asr_b200"), ROOT):
    sys.path.insert(0, p)


def main():
    import torch
    import bench
    from lnasr import _capi
    from lnasr._trellis import TransitionBank
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    T = 299
    mdl = bench.make_models("c4")[0]
    bank = TransitionBank(mdl["A"], mdl["pi"])
    N = bank.N
    total = n_utt * T
    g = torch.Generator(device="cuda").manual_seed(1)
    b = (-160.0 - 20.0 * torch.rand((total, N), device="cuda", generator=g)).float()
    off = torch.arange(0, total + 1, T, dtype=torch.int64, device="cuda")
    lib = _capi.lib()
    ws = torch.empty((int(lib.lnasr_viterbi_workspace_bytes(total, n_utt, 1, N)),), dtype=torch.uint8, device="cuda")
    res = {}
    for ver in (2, 1, 2, 1):
        lib.lnasr_debug_set_band_kernel(ver)
        for _ in range(3):
            r = bank.viterbi(b, off, n_utt, total, want_path=True, want_score=True, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            r = bank.viterbi(b, off, n_utt, total, want_path=True, want_score=True, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.setdefault(ver, []).append((ms, r["path"].clone(), r["score"].clone()))
        alg = total * N * (4 + 4) / (ms * 1e-3) / 1e9
        print(f"band kernel v{ver}: {ms:.3f} ms per {total} frames  ({alg:.0f} GB/s of the algorithmic 8N B/frame, "
              f"{total * N * 4 / (ms * 1e-3) / 1e9:.0f} GB/s of emission reads)")
    lib.lnasr_debug_set_band_kernel(2)
    same = torch.equal(res[1][0][1], res[2][0][1]) and torch.equal(res[1][0][2], res[2][0][2])
    print("paths and scores identical between the two forms:", same)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""How does tcgen05.mma round?  (one GPU; prints a table -> profiles/)

The FP16-slice / 3xTF32 emission kernels accumulate in fp32 in TMEM.  The total log-probability of an utterance is a SUM of
emission values, so a rounding BIAS of the accumulator matters far more than its size: this probe measures the rounding
mode of (1) the sum inside one MMA and (2) the accumulation of a second MMA into TMEM, with lnasr_debug_umma_probe
(kind::tf32, M = N = 128, K = 8 per MMA): every row r of A holds one test case, B column 0 is all ones, so D[r, 0] is the
sum of row r's addends.  All addends are exactly representable in TF32 and every exact sum in float64.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ln-asr_b200"))


def kmajor(mat):
    rows, K = mat.shape
    return np.ascontiguousarray(mat.reshape(rows // 8, 8, K // 4, 4).transpose(0, 2, 1, 3)).ravel()


def run(a, second):
    """a [128, 16]: K chunk 0-7 = first MMA, 8-15 = second (accumulating) MMA when `second`."""
    import torch
    from lnasr import _capi
    lib = _capi.lib()
    b = np.zeros((128, 16), dtype=np.float32)
    b[0, :] = 1.0                                              # D[r, 0] = sum_k a[r, k]
    K = 16
    a_img = torch.from_numpy(kmajor(a.astype(np.float32))).cuda()
    b_img = torch.from_numpy(kmajor(b)).cuda()
    d = torch.zeros((128, 128), dtype=torch.float32, device="cuda")
    idesc = (1 << 4) | (2 << 7) | (2 << 10) | ((128 >> 3) << 17) | ((128 >> 4) << 24)
    sbo = (K // 4) * 128
    lib.lnasr_debug_set_probe_second_mma(256 if second else 0)          # K chunks 2, 3 = k 8..15
    try:
        _capi.check(lib.lnasr_debug_umma_probe(_capi.ptr(a_img), a_img.numel() * 4, _capi.ptr(b_img), b_img.numel() * 4,
                                               128, sbo, 128, sbo, 0, idesc, _capi.ptr(d), _capi.stream_ptr()), "probe")
        torch.cuda.synchronize()
    finally:
        lib.lnasr_debug_set_probe_second_mma(0)
    return d.cpu().numpy()[:, 0].astype(np.float64)


def classify(exact, got, ulp):
    """which rounding of `exact` to a multiple of `ulp` is `got`?"""
    lo = np.floor(exact / ulp) * ulp
    hi = np.ceil(exact / ulp) * ulp
    rn = np.round(exact / ulp) * ulp                          # ties to even
    rz = np.trunc(exact / ulp) * ulp
    tags = []
    if got == exact:
        tags.append("exact")
    if got == rn:
        tags.append("RN")
    if got == rz:
        tags.append("RZ")
    if got == lo:
        tags.append("RM(floor)")
    if got == hi:
        tags.append("RP(ceil)")
    return "/".join(tags) or "other"


def main():
    L = 2.0 ** 24                                             # ulp(L) = 2 in fp32
    fr = [0.5, 1.0, 1.5, 0.25, 1.75, 3.0, 2.5]
    cases = []
    for sign in (+1, -1):
        for f in fr:
            for fs in (+1, -1):
                cases.append((sign * L, fs * f))
    print("# tcgen05.mma kind::tf32 rounding probe; big = +-2^24 (fp32 ulp 2), small addends exact in TF32")
    print("## (1) inside ONE MMA: D = big + small (two of the K = 8 products, the rest zero)")
    a = np.zeros((128, 16))
    for r, (big, small) in enumerate(cases):
        a[r, 0], a[r, 1] = big, small
    got = run(a, second=False)
    for r, (big, small) in enumerate(cases):
        print(f"  {big:+.0f} {small:+5.2f} -> {got[r] - big:+5.1f}   [{classify(big + small, got[r], 2.0)}]")
    print("## (2) four small addends in ONE MMA: big + 4 x 0.5 (exact sum = big + 2 = one ulp), and big + 4 x 0.75 (= big + 3)")
    a = np.zeros((128, 16))
    rows = []
    for sign in (+1, -1):
        for f in (0.5, 0.75, -0.5, -0.75):
            r = len(rows)
            a[r, 0] = sign * L
            a[r, 1:5] = f
            rows.append((sign * L, f))
    got = run(a, second=False)
    for r, (big, f) in enumerate(rows):
        print(f"  {big:+.0f} + 4 x {f:+5.2f} -> {got[r] - big:+5.1f}   (exact {4 * f:+.1f})  [{classify(big + 4 * f, got[r], 2.0)}]")
    print("## (3) ACCUMULATION: first MMA writes big, a second MMA adds small (accumulate = 1)")
    a = np.zeros((128, 16))
    for r, (big, small) in enumerate(cases):
        a[r, 0], a[r, 8] = big, small
    got = run(a, second=True)
    for r, (big, small) in enumerate(cases):
        print(f"  {big:+.0f} then {small:+5.2f} -> {got[r] - big:+5.1f}   [{classify(big + small, got[r], 2.0)}]")
    print("## (4) ACCUMULATION of four small products at once: big, then a second MMA adding 4 x f")
    a = np.zeros((128, 16))
    for r, (big, f) in enumerate(rows):
        a[r, 0] = big
        a[r, 8:12] = f
    got = run(a, second=True)
    for r, (big, f) in enumerate(rows):
        print(f"  {big:+.0f} then 4 x {f:+5.2f} -> {got[r] - big:+5.1f}   (exact {4 * f:+.1f})  [{classify(big + 4 * f, got[r], 2.0)}]")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""bench.py -- throughput of the acoustic hot path (MFCC -> GMM emissions -> Viterbi word decision).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]

Metric (BASELINE.json): frames/sec through MFCC + GMM log-likelihood + Viterbi.  One "step" = one pass of the hot
path over one batch of synthetic 16 kHz utterances with random-init GMM-HMMs:
  c4 (default; BASELINE configs[3], the configuration the north star's 1e8 frames/s target is quoted on): one
      2000-state x 16-mixture x 39-dim left-to-right model, tensor-core emissions, 1250 utterances of 3 s per GPU (the
      per-GPU shard of 10 000 utterances on 8 GPUs); the decode result is the Viterbi state path.
  c2 (BASELINE configs[1]): isolated-word models, 10 words x 5 states x 4 mixtures x 39 dims, 4096 utterances of
      0.5-1.5 s per GPU; the word is the argmax over models of the Viterbi score.
  c3 (BASELINE configs[2]): one Baum-Welch EM iteration, 3 states x 8 mixtures x 39 dims, 1000 utterances per GPU,
      accumulators all-reduced over NCCL (lnasr.em.EmTrainer).
The default run measures c4 as the line's headline and adds short c2 and c3 measurements under "secondary" (same process
group, same N; --no-secondary skips them).
`value` is timed with the PCM already in HBM; `e2e` goes through the public API from pinned host buffers with the
host->device copy and the device->host read of the decode result (c4: state paths + scores; c2: words + scores) inside
the timed region.  Multi-GPU: one process per GPU (torchrun), utterances sharded by rank, no data-path collective
("weak" scaling: per-GPU batch fixed).

--impl reference times the UNMODIFIED reference (MFCC.calc_mfcc -> GMMHMM._viterbi(obs, True), staged by oracle/build_ref.py
into oracle/_ref/lnasr_reference.zip, which travels to the GPU box) on all host cores, on a bounded sample of the same
utterances; the numpy port (oracle/lnasr_oracle.py) is the fallback when the archive is absent (`kind` says which).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "ln-asr_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "frames/sec (MFCC+GMM loglik+Viterbi)"
UNIT = "frames/s"


# ------------------------------------------------------------------------------------------------ workloads
def synth_batch(seed: int, n_utt: int, min_s: int, max_s: int, n_draw: int = None):
    """SURVEY.md 8(d) synthetic audio: int16 = clip(round(3000 N(0,1)) + 2000 sin(2 pi f0 t)), f0 ~ U(80,400).
    n_draw: size of the batch the lengths are drawn for (the first n_utt utterances of it are generated: the noise is drawn
    per utterance from its own stream, so a prefix of a batch is identical to the start of the full batch)."""
    rng = np.random.default_rng(seed)
    n_draw = n_draw or n_utt
    lens = rng.integers(min_s, max_s + 1, size=n_draw)[:n_utt]
    f0 = rng.uniform(80.0, 400.0, size=n_draw)[:n_utt]
    off = np.zeros(n_utt + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    pcm = np.empty(total, dtype=np.float32)
    for u in range(n_utt):
        pcm[off[u]:off[u + 1]] = np.round(3000.0 * np.random.default_rng((seed, u)).standard_normal(int(lens[u]), dtype=np.float32))
    for u in range(n_utt):
        t = np.arange(lens[u], dtype=np.float32) / 16000.0
        pcm[off[u]:off[u + 1]] += 2000.0 * np.sin(2 * np.pi * f0[u] * t)
    return np.clip(pcm, -32767, 32767).astype(np.int16), off


def make_models(workload: str, seed: int = 1234):
    """Random-init GMM-HMMs with GMMHMM.reset('random') semantics (oracle.random_gmmhmm), scaled into the range of
    MFCC features (|x| up to ~200) so that every state has a sensible likelihood."""
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(seed)
    if workload == "c2":
        models = [O.random_gmmhmm(5, 4, 39, rng) for _ in range(10)]
    else:
        m = O.random_gmmhmm(2000, 16, 39, rng, left_to_right=True)
        m["pi"] = np.full(2000, -np.inf)
        m["pi"][0] = 0.0
        models = [m]
    for m in models:
        m["mu"] = m["mu"] * 100.0
        m["var"] = m["var"] * 400.0
    return models


WORKLOADS = {
    "c2": dict(desc="isolated-word GMM-HMM: 10 words x 5 states x 4 mixtures x 39 dims, batched Viterbi",
               n_utt=4096, min_s=8000, max_s=24000, emission="exact"),
    "c3": dict(desc="Baum-Welch EM: 3-state x 8-mixture x 39-dim monophone, 1000 utterances of 200-400 frames per GPU, "
                    "one EM iteration (emissions + forward + fused backward/accumulate + all-reduce + M-step)",
               n_utt=1000, min_t=200, max_t=400),
    "c4": dict(desc="large-vocab acoustic scoring: 2000 states x 16 mixtures x 39 dims (left-to-right), TC emissions",
               n_utt=1250, min_s=48000, max_s=48000, emission="tc"),
}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
_CPU_MODELS = None          # built in the parent before the fork: the c4 model's full si is 390 MB, shared copy-on-write
_CPU_REF = None
_CPU_KIND = "port"


def _full_si(var):
    """[n,m,d] variances -> the reference's [n,m,d,d] covariance array (diagonal)."""
    n, m, d = var.shape
    si = np.zeros((n, m, d, d))
    idx = np.arange(d)
    si[:, :, idx, idx] = var
    return si


def _cpu_prepare(workload):
    """Decide which CPU implementation runs (the staged unmodified reference if its archive is there, else the numpy port)
    and build the models once in the parent."""
    global _CPU_MODELS, _CPU_KIND
    from oracle import ref_loader
    _CPU_KIND = "reference" if ref_loader.available() else "port"
    _CPU_MODELS = make_models(workload)
    if _CPU_KIND == "reference":
        for m in _CPU_MODELS:
            m["si"] = _full_si(m["var"])


def _cpu_init():
    """Per-worker set-up (forked): single-threaded BLAS; the reference is imported as `lnasr` in the worker only."""
    global _CPU_REF
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    np.seterr(all="ignore")
    if _CPU_KIND == "reference":
        import warnings
        warnings.filterwarnings("ignore")
        from oracle import ref_loader
        R = ref_loader.load()
        mfcc = R["mfcc"].MFCC()
        hmms = [R["gmmhmm"].GMMHMM(m["w"].shape[0], m["w"].shape[1], 39, A=m["A"], pi=m["pi"], w=m["w"], mu=m["mu"], si=m["si"])
                for m in _CPU_MODELS]
        _CPU_REF = (mfcc, hmms)


def _cpu_worker(pcm):
    """One utterance through the CPU path: MFCC -> emissions of every model -> Viterbi -> best model / state path."""
    if _CPU_KIND == "reference":
        mfcc, hmms = _CPU_REF
        F = mfcc.calc_mfcc(pcm)[2]                                  # mfcc.py:108-175
        best = None
        for g in hmms:
            v, _, path = g._viterbi(F, True)                        # hmm.py:133-172 (what decode() runs, hmm.py:276-283)
            sc = float(np.max(v[-1]))
            if best is None or sc > best[0]:
                best = (sc, path)
        return F.shape[0]
    from oracle import lnasr_oracle as O
    _, _, F = O.mfcc(pcm)
    O.score_words_viterbi(F, _CPU_MODELS)
    return F.shape[0]


def cpu_sample(workload: str, n_utt_sample: int):
    """The first n utterances of rank 0's GPU batch (same seed, same lengths)."""
    cfg = WORKLOADS[workload]
    pcm, off = synth_batch(1000, n_utt_sample, cfg["min_s"], cfg["max_s"], n_draw=cfg["n_utt"])
    return [pcm[off[u]:off[u + 1]] for u in range(n_utt_sample)]


class CpuArm:
    """`cores` single-threaded worker processes (faster than threaded BLAS on these shapes)."""

    def __init__(self, workload, cores):
        import multiprocessing as mp
        _cpu_prepare(workload)
        self.kind, self.cores, self.workload = _CPU_KIND, cores, workload
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)

    def run(self, jobs):
        t0 = time.perf_counter()
        frames = sum(self.pool.map(_cpu_worker, jobs, chunksize=max(1, len(jobs) // (8 * self.cores))))
        return frames, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self):
        return ("unmodified reference (oracle/_ref/lnasr_reference.zip): MFCC.calc_mfcc + GMMHMM._viterbi(obs, True) per model"
                if self.kind == "reference" else "numpy port of the reference (oracle/lnasr_oracle.py)")


def cpu_sample_size(workload, cores):
    # c4: ~11 s per 3 s utterance in the reference (32 000 x det/inv of a 39 x 39 covariance + T x N Python-level Viterbi
    # steps) -> one utterance per core and step; c2: ~0.3 s per utterance
    return cores if workload == "c4" else max(32 * cores, 64)


def _em_cpu_worker(job):
    from oracle import lnasr_oracle as O
    x, mdl = job
    acc = O.em_accumulate_diag(x, mdl["pi"], mdl["A"], mdl["w"], mdl["mu"], mdl["var"])
    return x.shape[0], acc


def run_reference_em(args):
    """CPU restatement of one EM iteration of c3 (per-utterance statistics summed + M-step), one process per core."""
    import multiprocessing as mp
    from oracle import lnasr_oracle as O
    cores = os.cpu_count() or 1
    cfg = WORKLOADS["c3"]
    mdl = O.random_gmmhmm(3, 8, 39, np.random.default_rng(7))
    n_s = 2 * cores
    x, off = synth_feats(99, n_s, cfg["min_t"], cfg["max_t"])
    jobs = [(x[off[u]:off[u + 1]].astype(np.float64), mdl) for u in range(n_s)]
    vals = []
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_em_cpu_worker, jobs, chunksize=1)
            tot = None
            for _, acc in res:
                tot = acc if tot is None else {k: tot[k] + acc[k] for k in tot}
            O.em_finalize_diag(tot)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                vals.append((sum(r[0] for r in res), dt))
    frames, dt = sum(v[0] for v in vals), sum(v[1] for v in vals)
    value = frames / dt
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec (Baum-Welch EM iteration)", "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c3: " + cfg["desc"], "sample_utts_per_step": n_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_s} utterances/step x {len(vals)} steps, one single-threaded process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    return 0


def run_reference(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload == "c3":
        return run_reference_em(args)
    cores = os.cpu_count() or 1
    per_step = cpu_sample_size(args.workload, cores)
    jobs = cpu_sample(args.workload, per_step)
    arm = CpuArm(args.workload, cores)
    budget = float(args.cpu_budget_s)                       # the whole run must end within a few minutes
    vals = []
    t_all = time.perf_counter()
    n_warm = 0
    warm = args.warmup
    for i in range(args.warmup + args.steps):
        frames, dt = arm.run(jobs)
        if dt > 5.0:
            warm = min(warm, 1)                                 # a pass this long is warm after one repetition
        if i >= warm or (time.perf_counter() - t_all) + 2 * dt > budget:
            vals.append((frames, dt))                           # out of budget: the remaining passes are all timed ones
        else:
            n_warm += 1
        if len(vals) >= args.steps or (time.perf_counter() - t_all) + dt > budget:
            break
    arm.close()
    frames = sum(x[0] for x in vals)
    dt = sum(x[1] for x in vals)
    value = frames / dt
    sample = (f"{per_step} utterances/step (the first {per_step} of rank 0's GPU batch, {frames // len(vals)} frames) x {len(vals)} timed steps "
              f"after {n_warm} warm-up; {arm.describe()}; one single-threaded process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": n_warm, "ms_per_step": 1e3 * dt / max(1, len(vals)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + WORKLOADS[args.workload]["desc"], "sample_utts_per_step": per_step,
                   "steps_requested": args.steps, "warmup_requested": args.warmup, "wall_budget_s": budget},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """Process-group context shared by the workloads of one run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.numa = None
        if self.world > 1 and not args.no_numa_bind:
            from lnasr.dist import bind_to_gpu_numa_node
            self.numa = bind_to_gpu_numa_node(self.local)           # before the pinned host buffers are allocated
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm_peak = float(self.peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in self.peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch
        import torch.distributed as dist
        v = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    def sum_over_ranks(self, x):
        import torch
        import torch.distributed as dist
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
            dist.destroy_process_group()


def run_ours(args, ctx, workload, steps, warmup, with_cpu=True, load_s=1.5):
    """One decode workload (c2 / c4) on this rank's shard; every rank calls it, rank 0 gets the JSON dict back."""
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC

    rank, world, local = ctx.rank, ctx.world, ctx.local
    cfg = dict(WORKLOADS[workload])
    if cfg["emission"] == "tc" and args.tc_format == "f16":
        cfg["emission"] = "tc16"
        cfg["desc"] = cfg["desc"].replace("TC emissions", "tensor-core emissions (2 FP16 slices)")
    models = make_models(workload)
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission=cfg["emission"]))
    # every rank owns its own shard of utterances (weak scaling: per-GPU batch fixed)
    pcm_np, off = synth_batch(1000 + rank, cfg["n_utt"], cfg["min_s"], cfg["max_s"])
    n_utt = cfg["n_utt"]
    pcm_host = torch.from_numpy(pcm_np).pin_memory()
    pcm_dev = pcm_host.cuda()
    want_paths = True
    # what the caller gets back: one large model -> the state path IS the decode result; a bank of word models -> the word
    return_paths = pipe.bank.K == 1

    def step_device():
        res, mb = pipe.recognize_batch(pcm_dev, off, want_paths=want_paths)
        return res, mb

    # utterance groups of the host path: a copy-bound workload (c2: 2.4 ms of copies against 1.7 ms of compute) wants several
    # groups so that compute hides behind the copies of one batch; a compute-bound one (c4: 2.3 ms against 7 ms) wants whole
    # batches -- full waves for the per-utterance trellis CTAs -- with the NEXT batch's copy in flight (tools/e2e_sweep.py)
    n_chunks = args.e2e_chunks or (1 if workload == "c4" else 6)
    e2e_kw = dict(n_chunks=n_chunks, want_paths=want_paths, n_compute_streams=args.e2e_streams, taper=not args.e2e_uniform,
                  return_paths=return_paths)

    def step_e2e():
        # public host-facing API: pinned PCM in, decode result out; the H2D copy is chunked and overlapped
        return pipe.recognize_batch_host(pcm_host, off, **e2e_kw)

    res, mb = step_device()
    torch.cuda.synchronize()
    frames = int(mb.frame_offsets_host[-1])
    words0 = res.words.cpu().numpy().copy()
    paths0 = res.paths.cpu().numpy().copy() if return_paths else None

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    # clocks are sampled from before the warm-up until the sampler has seen >= 1.5 s of load (nvidia-smi needs a few
    # hundred ms to produce its first line; the timed region of a short run is over before that)
    sampler = ClockSampler(local) if rank == 0 else None
    t_load0 = time.perf_counter()
    ms_dev = timed(step_device, steps, warmup)
    ms_e2e_sync = timed(step_e2e, steps, max(3, warmup // 2))
    # streaming form of the same public API: two batches in flight (submit batch i+1, then collect batch i); every step still
    # copies its PCM from pinned host memory and reads its result back
    pend = [None]
    last = [None]

    def collect(h):
        out = h.result()
        last[0] = (out[0], h.paths() if return_paths else None)

    def step_e2e_pipelined():
        nxt = pipe.submit_host(pcm_host, off, **e2e_kw)
        if pend[0] is not None:
            collect(pend[0])
        pend[0] = nxt

    def timed_pipelined(steps, warmup):
        for _ in range(warmup):
            step_e2e_pipelined()
        collect(pend[0])
        pend[0] = None
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_e2e_pipelined()
        collect(pend[0])                                   # the last batch's result is on the host
        pend[0] = None
        e1.record()
        ctx.barrier()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1))
        assert np.array_equal(last[0][0].numpy(), words0), "pipelined e2e path decoded different words"
        if return_paths:
            assert np.array_equal(last[0][1].numpy(), paths0), "pipelined e2e path decoded different state paths"
        return ms

    ms_e2e = timed_pipelined(steps, 3)
    while time.perf_counter() - t_load0 < load_s:
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    chk = step_e2e()
    assert np.array_equal(chk[0].numpy(), words0), "e2e path decoded different words"
    if return_paths:
        assert np.array_equal(chk[2].numpy(), paths0), "e2e path decoded different state paths"

    total_frames = ctx.sum_over_ranks(frames)
    value = total_frames / (ms_dev / 1e3 / steps)
    e2e_value = total_frames / (ms_e2e / 1e3 / steps)
    e2e_sync_value = total_frames / (ms_e2e_sync / 1e3 / steps)
    d2h = n_utt * 12 + (4 * pipe.bank.K * frames if return_paths else 0)

    # ---- per-kernel shares + roofline of the dominant kernel (rank 0, CUDA events around each C-ABI call)
    line = None
    if rank == 0:
        stages = stage_times(pipe, pcm_dev, off, want_paths, reps=max(3, steps // 2))
        fp64 = fp64_peaks()
        if cfg["emission"] == "tc":
            fp64["tf32_tflops"] = tf32_peak()
        top = max(stages, key=lambda k: stages[k]["ms"])
        roof = roofline_for(top, stages[top], frames, cfg, models, ctx.hbm_peak, ctx.peaks, fp64)
        roof["kernel"] = top
        tr = ncu_traffic(workload, frames, top)
        if tr is not None:
            roof["traffic"] = tr["bytes"]
            roof["traffic_note"] = "bytes per launch of %s (ncu dram__bytes_read.sum + dram__bytes_write.sum, %s)" % (tr["kernel"], tr["source"])
            if roof.get("algorithmic_bytes_per_frame"):
                roof["traffic_note"] += "; algorithmic = %d" % (roof["algorithmic_bytes_per_frame"] * frames)
        roof["peak_source"] = ctx.peak_src
        if roof["bound"] == "tensor":
            # the denominator of a tensor-bound stage: the sustained (kernel timed inside a long step) 16-bit dense GEMM rate of
            # MEASURED_PEAKS.json for the tcgen05 kernels, the FP64 mma rate measured in this run for the exact kernel
            roof["peak_source"] = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if cfg["emission"] == "tc16" and "bf16_tflops_sustained" in ctx.peaks
                                   else "measured in this run (fp64_peaks_measured / cuBLAS TF32 probe)")
        roof["stage_ms"] = {k: round(v["ms"], 4) for k, v in stages.items()}
        roof["all_stages"] = {}
        for k in stages:
            r = roofline_for(k, stages[k], frames, cfg, models, ctx.hbm_peak, ctx.peaks, fp64)
            roof["all_stages"][k] = {"bound": r["bound"], "achieved": round(r["achieved"], 3), "peak": round(r["peak"], 1),
                                     "unit": r["unit"], "frac": round(r["frac"], 4)}
            if "fp64_alu" in r:
                roof["all_stages"][k]["fp64_alu_frac"] = round(r["fp64_alu"]["frac"], 4)
            if "executed_frac" in r:
                roof["all_stages"][k]["executed_frac"] = round(r["executed_frac"], 4)
        roof["fp64_peaks_measured"] = {k: round(v, 2) for k, v in fp64.items()}
        cores = os.cpu_count() or 1
        cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": "skipped"}
        if world == 1 and with_cpu and not args.no_cpu:
            os.environ["OMP_NUM_THREADS"] = "1"
            n_s = cpu_sample_size(workload, cores)
            jobs = cpu_sample(workload, n_s)
            arm = CpuArm(workload, cores)
            if workload != "c4":
                arm.run(jobs[:2 * cores])                   # warm the workers (imports, first-call costs)
            f_s, dt = arm.run(jobs)
            arm.close()
            cpu = {"value": f_s / dt, "unit": UNIT, "cores": cores, "kind": arm.kind,
                   "sample": f"the first {n_s} utterances of this batch ({f_s} frames) in {dt:.1f} s; {arm.describe()}; "
                             "one single-threaded process per core"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if cfg["emission"] == "exact" else ("f64 MFCC + trellis / 3xTF32 emissions" if cfg["emission"] == "tc"
                                                                else "f64 MFCC + trellis / 2xFP16-slice emissions (fp32 accumulate)"),
            "data": "synthetic",
            "config": {"workload": workload + ": " + cfg["desc"], "utts_per_gpu": n_utt, "frames_per_gpu_step": frames,
                       "pcm_bytes_per_gpu_step": int(pcm_np.nbytes), "l2": "inputs larger than L2 (PCM + intermediates > 126 MB per step)",
                       "parallelism": f"utterance-sharded x{world}, no data-path collective", "viterbi_paths": want_paths,
                       "host_numa_binding": ctx.numa},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(pcm_np.nbytes) + int(off.nbytes) * 2, "d2h_bytes_per_step": d2h,
                    "result": ("Viterbi state paths int32 [1, frames] + score per utterance" if return_paths
                               else "word index + score per utterance"),
                    "api": "AcousticPipeline.submit_host / PendingRecognition.result (+ .paths), two batches in flight",
                    "utterance_groups_per_batch": n_chunks,
                    "one_batch_at_a_time": {"value": e2e_sync_value, "ms_per_step": ms_e2e_sync / steps,
                                            "api": "AcousticPipeline.recognize_batch_host"}},
            "gpu_launches": launches_in(step_device) * steps,
            "roofline": roof, "cpu_baseline": cpu,
        }
    del pipe, pcm_dev, pcm_host, res, mb
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------ EM arm (c3)
def synth_feats(seed, n_utt, min_t, max_t, d=39):
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_t, max_t + 1, size=n_utt)
    off = np.zeros(n_utt + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    x = (0.5 * rng.standard_normal((int(off[-1]), d))).astype(np.float32)
    return x, off


def run_em(args, ctx, steps, warmup, load_steps=400):
    """BASELINE configs[2]: one Baum-Welch iteration over the rank's shard of utterances; the packed float64 accumulator is
    all-reduced over NCCL; every rank finalises the same M-step.  value = frames of ALL ranks / time of one iteration."""
    import torch
    from lnasr.dist import make_reduce_fn
    from lnasr.em import EmTrainer
    from oracle import lnasr_oracle as O

    rank, world, local = ctx.rank, ctx.world, ctx.local
    cfg = WORKLOADS["c3"]
    mdl = O.random_gmmhmm(3, 8, 39, np.random.default_rng(7))
    x_np, off = synth_feats(2000 + rank, cfg["n_utt"], cfg["min_t"], cfg["max_t"])
    x_host = torch.from_numpy(x_np).pin_memory()
    tr = EmTrainer(mdl).load_batch(x_host.cuda(), off)
    reduce_fn = make_reduce_fn() if world > 1 else None
    frames = int(off[-1])

    def step():
        return tr.iterate(reduce_fn)

    def step_e2e():
        tr._x.copy_(x_host, non_blocking=True)            # features from pinned host memory every iteration
        return float(tr.iterate(reduce_fn).item())       # log-likelihood read back

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev = timed(step, steps, warmup)
    ms_e2e = timed(step_e2e, steps, 3)
    for _ in range(load_steps):                           # ~1 s more load for the clock sampler; a FIXED count on every
        step()                                            # rank (each step holds a collective -- never loop on wall time)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    total_frames = ctx.sum_over_ranks(frames)
    n_launch = launches_in(step)                          # on every rank: the step holds a collective
    torch.cuda.synchronize()
    line = None
    if rank == 0:
        n, m, d = 3, 8, 39
        per = 16 * n + 4 * d + 8 * n * m          # b fwd + alpha w/r + b bwd + x + bm
        sec = ms_dev / 1e3 / steps
        ach = per * frames / sec / 1e9
        line = {
            "metric": "frames/sec (Baum-Welch EM iteration)", "value": total_frames / sec, "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c3: " + cfg["desc"], "utts_per_gpu": cfg["n_utt"], "frames_per_gpu_step": frames,
                       "l2": "x + b + bm + alpha = %.0f MB per step" % (frames * (4 * d + 8 * (2 * n + n * m)) / 1e6),
                       "parallelism": f"utterance-sharded x{world}, one all-reduce of {tr.acc.numel()} float64 per iteration"},
            "clocks": clocks,
            "e2e": {"value": total_frames / (ms_e2e / 1e3 / steps), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(x_np.nbytes), "d2h_bytes_per_step": 8},
            "gpu_launches": n_launch * steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": ach / ctx.hbm_peak, "traffic": None,
                         "algorithmic_bytes_per_frame": per, "kernel": "whole iteration", "peak_source": ctx.peak_src},
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "not run for c3"},
        }
    del tr
    torch.cuda.empty_cache()
    return line


def compact(line):
    """What a secondary workload contributes to the headline's JSON line."""
    if line is None:
        return None
    keep = {k: line[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "dtype")}
    keep["workload"] = line["config"]["workload"]
    keep["frames_per_gpu_step"] = line["config"]["frames_per_gpu_step"]
    keep["parallelism"] = line["config"]["parallelism"]
    keep["e2e"] = {k: line["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    keep["gpu_launches"] = line["gpu_launches"]
    r = line["roofline"]
    keep["roofline"] = {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "stage_ms", "all_stages",
                                          "fp64_alu", "fp64_peaks_measured") if k in r}
    return keep


def ncu_traffic(workload, frames, stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's dominant kernel, from the committed summary of an
    `ncu --set full` capture of THIS workload at THIS size (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the
    capture's raw page).  None when there is no capture of this shape -- never a constant carried in the source."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    ent = tab.get(workload)
    if not ent or int(ent.get("frames", -1)) != int(frames) or stage not in ent.get("stages", {}):
        return None
    st = ent["stages"][stage]
    return {"bytes": float(st["dram_read_bytes"]) + float(st["dram_write_bytes"]), "kernel": st["kernel"], "source": ent.get("source", "profiles/")}


def launches_in(fn):
    """Kernels of liblnasr_b200.so launched while fn() runs: the library counts its own launches (lnasr_launch_count)."""
    from lnasr import _capi
    lib = _capi.lib()
    n0 = int(lib.lnasr_launch_count())
    fn()
    return int(lib.lnasr_launch_count()) - n0


def stage_times(pipe, pcm_dev, off, want_paths, reps=5):
    """CUDA-event time of each C-ABI call of one step (current stream), averaged over `reps` full passes."""
    import torch
    from lnasr import _capi, _dev
    t = torch
    import ctypes as C
    import numpy as np
    acc = {"mfcc": 0.0, "gmm": 0.0, "viterbi": 0.0}
    warm = 3                       # the caching allocator settles (no cudaMalloc inside a timed stage)
    # the offsets and the workspace of the MFCC call are prepared once, outside the timed stage (as the host path of
    # AcousticPipeline does): the stage is the library's launches, not numpy offset arithmetic and a small upload
    s_off = np.asarray(off, dtype=np.int64)
    f_off = _dev.offsets_from_lengths(pipe.mfcc.frame_counts(s_off))
    meta = _dev.to_dev(np.concatenate([s_off, f_off]), t.int64)
    need = C.c_size_t(0)
    _capi.check(_capi.lib().lnasr_mfcc_workspace_bytes(int(f_off[-1]), len(s_off) - 1, C.byref(need)), "mfcc_workspace_bytes")
    mfcc_plan = (meta[:len(s_off)], meta[len(s_off):], f_off, _dev.empty((need.value,), t.uint8))
    for rep in range(reps + warm):
        ev = [t.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        mb = pipe.mfcc.calc_mfcc_batch(pcm_dev, sample_offsets=s_off, _plan=mfcc_plan)
        ev[1].record()
        b = pipe.bank.emissions(mb.feats)
        ev[2].record()
        n_utt, total = len(off) - 1, int(mb.frame_offsets_host[-1])
        r = pipe.bank.trans.viterbi(b, mb.frame_offsets, n_utt, total, want_path=want_paths, want_score=True)
        ev[3].record()
        t.cuda.synchronize()
        del mb, b, r
        if rep < warm:
            continue
        acc["mfcc"] += ev[0].elapsed_time(ev[1])
        acc["gmm"] += ev[1].elapsed_time(ev[2])
        acc["viterbi"] += ev[2].elapsed_time(ev[3])
    return {k: {"ms": v / reps} for k, v in acc.items()}


def fp64_peaks():
    """FP64 peaks of this GPU measured now (TFLOP/s): DFMA and mma.sync m8n8k4 f64 (csrc/capi.cu probes)."""
    import ctypes as C
    import torch
    from lnasr import _capi
    lib = _capi.lib()
    scratch = torch.empty(int(lib.lnasr_debug_fp64_peak_scratch_doubles()), dtype=torch.float64, device="cuda")
    out = {}
    for kind, name in ((0, "fp64_fma_tflops"), (1, "fp64_mma_tflops")):
        v = C.c_double(0.0)
        _capi.check(lib.lnasr_debug_fp64_peak(kind, _capi.ptr(scratch), C.byref(v), _capi.stream_ptr()), "fp64_peak")
        out[name] = v.value
    return out


def tf32_peak():
    """cuBLAS TF32 GEMM throughput measured now (8192^3, back to back for ~0.5 s): the denominator of the tcgen05 kind::tf32
    kernel's roofline (MEASURED_PEAKS.json has bf16 only)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * 8192 ** 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


MFCC_FP64_FLOPS_PER_FRAME = 18300     # 256-pt complex FFT 10.2k + split/power 4.1k + pre-emphasis/window 1.2k + mel 0.9k + log 0.8k + DCT 1.0k


def roofline_for(name, st, frames, cfg, models, hbm_peak, peaks, fp64):
    """Algorithmic bytes / flops per frame x frames per launch / measured duration."""
    K = len(models)
    n, m = models[0]["w"].shape
    N = K * n
    sec = st["ms"] / 1e3
    if name == "mfcc":
        per = 160 * 2 + 39 * 4                    # new int16 samples in + 39 fp32 features out (476 B)
        ach = per * frames / sec / 1e9
        alu = MFCC_FP64_FLOPS_PER_FRAME * frames / sec / 1e12
        return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_frame": per,
                "fp64_alu": {"achieved_tflops": alu, "peak_tflops": fp64["fp64_fma_tflops"], "frac": alu / fp64["fp64_fma_tflops"],
                             "algorithmic_flops_per_frame": MFCC_FP64_FLOPS_PER_FRAME},
                "note": "nominally HBM-bound (476 B/frame) but the float64 chain (parity, DESIGN.md 5) makes it FP64-ALU / "
                        "shared-memory bound: the fp64_alu entry is the roofline that binds"}
    if name == "viterbi":
        per = 8 * N                               # read b (4N as fp32-equivalent) + write backpointers (4N): SURVEY.md 8(d)
        ach = per * frames / sec / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_frame": per}
    # gmm
    flops = 2 * 79 * N * m                        # 2 (2D+1) N M: the dense contraction of SURVEY.md 8(d)
    if cfg["emission"] == "tc":
        ach = flops * frames / sec / 1e12         # executed = 3x (3xTF32), K padded to 80
        peak = fp64.get("tf32_tflops") or float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "algorithmic_flops_per_frame": flops, "executed_over_algorithmic": 3.04,
                "executed_frac": 3.04 * ach / peak,
                "note": "peak = cuBLAS TF32 GEMM (8192^3, sustained) measured in this run; MEASURED_PEAKS.json has no TF32 figure. "
                        "achieved counts ALGORITHMIC flops; the kernel executes 3.04x that (3xTF32 error compensation)"}
    if cfg["emission"] == "tc16":
        ach = flops * frames / sec / 1e12         # executed = 3 products of FP16 slices, K padded to 80
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "algorithmic_flops_per_frame": flops, "executed_over_algorithmic": 3.04,
                "executed_frac": 3.04 * ach / peak,
                "note": "tcgen05 kind::f16 on two FP16 slices per operand; peak = sustained 16-bit dense GEMM of MEASURED_PEAKS.json "
                        "(bf16; fp16 runs at the same rate). achieved counts ALGORITHMIC flops; the kernel executes 3.04x that"}
    ach = flops * frames / sec / 1e12
    peak = fp64["fp64_mma_tflops"]
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "algorithmic_flops_per_frame": flops, "executed_over_algorithmic": 80.0 / 79.0,
            "note": "FP64 tensor cores (mma.sync m8n8k4 f64); peak = the same instruction measured in this run "
                    "(lnasr_debug_fp64_peak), not in MEASURED_PEAKS.json which has no FP64 figure"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="headline workload only (no short c2 / c3 measurements)")
    ap.add_argument("--cpu-budget-s", type=float, default=200.0, help="--impl reference: wall-clock bound of the whole run")
    ap.add_argument("--tc-format", default="f16", choices=["f16", "tf32"],
                    help="c4: operand slices of the tensor-core emission kernel (two FP16 slices, or 3xTF32)")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not pin each rank to its GPU's NUMA node")
    ap.add_argument("--e2e-uniform", action="store_true", help="equal utterance groups instead of tapered ones")
    ap.add_argument("--e2e-streams", type=int, default=2, help="compute streams the utterance groups alternate between")
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="utterance groups the host->device copy is pipelined over (0 = by workload: c4 1, c2 6)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    ctx = Ctx(args)
    if args.workload == "c3":
        line = run_em(args, ctx, args.steps, args.warmup)
    else:
        line = run_ours(args, ctx, args.workload, args.steps, args.warmup)
    if not args.no_secondary and args.workload == "c4":
        # every rank takes part (c2 shards by rank, c3 all-reduces); short runs, no CPU leg
        st = min(args.steps, 10)
        sec = {"c2": compact(run_ours(args, ctx, "c2", st, 3, with_cpu=False, load_s=0.0)),
               "c3": compact(run_em(args, ctx, st, 3, load_steps=0))}
        if line is not None:
            line["secondary"] = sec
            line["gpu_launches_note"] = "gpu_launches counts the headline workload's timed steps; each secondary entry carries its own"
    if line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Banks of GMM-HMMs scored together, and the PCM -> word pipeline (additive API, SURVEY.md 8(b) "new-build
additions").

The reference scores one utterance against one model per call.  Isolated-word recognition (BASELINE config 2)
and large-vocabulary acoustic scoring (config 4) are batches of utterances against K models that share
(n, m, d): here all K*n*m Gaussians are one packed image, emissions are one [frames, K*n] matrix, and the
trellis kernels run every (utterance, model) pair in one launch.  Word decision = argmax over models of the
Viterbi score max_j v[T-1, j], lowest index on ties.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _capi, _dev
from ._trellis import TransitionBank
from .gmm import PackedGmm

__all__ = ["ModelBank", "AcousticPipeline", "BankResult", "PendingRecognition"]


class BankResult(NamedTuple):
    words: "object"                 # [n_utt] int32 CUDA tensor: index of the best model
    best: "object"                  # [n_utt] float64: its Viterbi score
    scores: "object"                # [n_utt, K] float64 Viterbi scores max_j v[T-1,j]
    paths: Optional["object"]       # [K, total_frames] int32 state paths per model, or None
    logprob: Optional["object"]     # [n_utt, K] float64 forward log-probabilities, or None


class PendingRecognition:
    """Handle of an asynchronous `AcousticPipeline.submit_host`: `.result()` -> (words int32 [n], best float64 [n]) on the host;
    `.paths()` -> the Viterbi state paths [K, total_frames] int32 on the host (submitted with return_paths=True)."""

    def __init__(self, words, best, event, keep, paths=None, path_blocks=None, n_models=1):
        self._words, self._best, self._event, self._keep = words, best, event, keep
        self._paths, self._path_blocks, self._K = paths, path_blocks, n_models

    def done(self) -> bool:
        return self._event is None or self._event.query()

    def _wait(self):
        """Wait for the batch and take the results out of the pipeline's pinned staging buffers (which the pipeline reuses for
        a later batch) into tensors this handle owns."""
        if self._event is not None:
            self._event.synchronize()
            self._words, self._best = self._words.clone(), self._best.clone()
            if self._paths is not None:
                self._paths = self._paths.clone()
            self._event, self._keep = None, ()

    def result(self):
        self._wait()
        return self._words, self._best

    def paths(self):
        """[K, total_frames] int32 CPU tensor (hmm.py:168-171 per utterance and model), frames in batch order.  The device
        writes one [K, frames of the group] block per utterance group; with one model (large-vocabulary scoring) that IS the
        batch-order path, with several the blocks are re-joined here."""
        if self._paths is None:
            raise _capi.LnasrError("submit_host(..., return_paths=True) was not requested for this batch")
        self._wait()
        if self._K == 1:
            return self._paths.view(1, -1)
        t = _dev.torch()
        out = t.empty((self._K, self._paths.numel() // self._K), dtype=t.int32)
        for f0, f1 in self._path_blocks:
            out[:, f0:f1] = self._paths[self._K * f0:self._K * f1].view(self._K, f1 - f0)
        return out


class _HostSlot:
    """Device and pinned-host buffers of one host batch in flight.  AcousticPipeline keeps a few of these and reuses them
    round-robin: a steady stream of batches then runs without a single allocation (a fresh multi-GB workspace from the caching
    allocator, or a fresh pinned buffer, can cost a cudaMalloc / cudaHostAlloc that synchronises the device)."""

    def __init__(self):
        self.bufs = {}
        self.done = None          # event: every device-side use of the slot's buffers is over
        self.handle = None        # weakref to the PendingRecognition whose results sit in the pinned buffers

    def get(self, key, numel, dtype, device=None, pinned=False):
        t = _dev.torch()
        b = self.bufs.get(key)
        if b is None or b.numel() < numel or b.dtype != dtype:
            if b is not None:
                t.cuda.synchronize()                                    # growing: nothing may still be using the old buffer
            b = t.empty((max(int(numel), 1),), dtype=dtype, pin_memory=True) if pinned else \
                t.empty((max(int(numel), 1),), dtype=dtype, device=device)
            self.bufs[key] = b
        return b[:numel]


class ModelBank:
    """K GMM-HMMs with common (n, m, d)."""

    def __init__(self, models: Sequence, emission: str = "auto"):
        """models: GMMHMM objects or dicts with keys A, pi, w, mu and si ([n,m,d,d]) or var ([n,m,d]).
        emission: 'exact' (float64, FP64 tensor cores), 'tc' (3xTF32 tcgen05), 'tc16' (two FP16 slices per operand, tcgen05
        kind::f16: half the MMAs of 'tc') or 'auto' (tc16 when K*n*m >= 2048)."""
        def get(mdl, k):
            return mdl[k] if isinstance(mdl, dict) else getattr(mdl, k)

        def cov(mdl):
            if isinstance(mdl, dict) and "var" in mdl:
                return np.asarray(mdl["var"], dtype=np.float64)
            si = np.asarray(get(mdl, "si"), dtype=np.float64)
            return si
        A = np.stack([np.asarray(get(m, "A"), dtype=np.float64) for m in models])
        pi = np.stack([np.asarray(get(m, "pi"), dtype=np.float64).reshape(-1) for m in models])
        w = np.concatenate([np.asarray(get(m, "w"), dtype=np.float64) for m in models], axis=0)
        mu = np.concatenate([np.asarray(get(m, "mu"), dtype=np.float64) for m in models], axis=0)
        cv = np.concatenate([cov(m) for m in models], axis=0)
        self.K, self.n = A.shape[0], A.shape[1]
        self.m, self.d = w.shape[1], mu.shape[2]
        self.trans = TransitionBank(A, pi)
        self.gmm = PackedGmm(w, mu, cv)
        if emission == "auto":
            emission = "tc16" if (self.K * self.n * self.m >= 2048 and not self.gmm.full) else "exact"
        if emission not in ("exact", "tc", "tc16"):
            raise ValueError("emission must be 'exact', 'tc', 'tc16' or 'auto'")
        self.emission = emission
        self._tc = None
        if emission in ("tc", "tc16") and self.gmm.full:
            raise _capi.LnasrError("the tensor-core emission kernels score DIAGONAL covariances; this bank has full `si` matrices "
                                   "(a model trained by the reference's train() is full-covariance): use emission='exact'")
        if emission in ("tc", "tc16"):
            from .gmm_tc import PackedGmmTC
            self._tc = PackedGmmTC(w, mu, cv if cv.ndim == 3 else np.diagonal(cv, axis1=2, axis2=3),
                                   fmt="f16" if emission == "tc16" else "tf32")

    def emissions(self, feats, b_out=None):
        """feats [T, d] float32/float64 device -> b [T, K*n] (float64 exact / float32 tc)."""
        if self._tc is not None:
            return self._tc.loglik(feats, b_out=b_out)
        return self.gmm.loglik(feats, b_out=b_out)[0]

    def score(self, feats, f_off_d, n_utt: int, total: int, want_paths=False, want_logprob=False,
              workspace=None, words_out=None, best_out=None) -> BankResult:
        t = _dev.torch()
        b = self.emissions(feats)
        r = self.trans.viterbi(b, f_off_d, n_utt, total, want_path=want_paths, want_score=True, workspace=workspace)
        words = words_out if words_out is not None else _dev.empty((n_utt,), t.int32)
        best = best_out if best_out is not None else _dev.empty((n_utt,), t.float64)
        _capi.check(_capi.lib().lnasr_argmax_rows(_capi.ptr(r["score"]), n_utt, self.K, _capi.ptr(words), _capi.ptr(best),
                                                 _capi.stream_ptr()), "argmax_rows")
        lp = self.trans.forward(b, f_off_d, n_utt, total)[1] if want_logprob else None
        return BankResult(words, best, r["score"], r["path"], lp)


class AcousticPipeline:
    """PCM -> MFCC -> GMM emissions -> Viterbi -> word, for a ragged batch (BASELINE north star (1)-(3))."""

    def __init__(self, mfcc, bank: ModelBank):
        self.mfcc = mfcc
        self.bank = bank

    def recognize_batch(self, pcm, sample_offsets, want_paths=False, want_logprob=False):
        """pcm: concatenated int16/float32 samples (numpy, pinned CPU tensor or CUDA tensor); sample_offsets [n+1] host.
        Returns (BankResult, MfccBatch); everything stays on the device."""
        mb = self.mfcc.calc_mfcc_batch(pcm, sample_offsets=sample_offsets)
        n_utt = len(mb.frame_offsets_host) - 1
        total = int(mb.frame_offsets_host[-1])
        res = self.bank.score(mb.feats, mb.frame_offsets, n_utt, total, want_paths=want_paths, want_logprob=want_logprob)
        return res, mb

    def recognize_batch_host(self, pcm_host, sample_offsets, n_chunks: int = 6, want_paths: bool = False,
                             n_compute_streams: int = 2, taper: bool = False, _timeline=None, return_paths: bool = False):
        """Synchronous form of `submit_host`: (words int32 [n], best float64 [n]) as CPU tensors, plus the state paths
        [K, total_frames] int32 when return_paths is set."""
        h = self.submit_host(pcm_host, sample_offsets, n_chunks=n_chunks, want_paths=want_paths, n_compute_streams=n_compute_streams,
                             taper=taper, _timeline=_timeline, return_paths=return_paths)
        return h.result() + ((h.paths(),) if return_paths else ())

    def submit_host(self, pcm_host, sample_offsets, n_chunks: int = 6, want_paths: bool = False,
                    n_compute_streams: int = 2, taper: bool = False, _timeline=None,
                    return_paths: bool = False) -> "PendingRecognition":
        """Asynchronous end-to-end entry point for HOST audio; returns a handle whose `.result()` waits for the words.
        Consecutive submissions pipeline: the copies of batch i+1 start while the last groups of batch i are still being
        decoded (a streaming server keeps two batches in flight).

        End-to-end entry point for HOST audio: `pcm_host` is a pinned CPU tensor (int16 or float32) of concatenated
        utterances, `sample_offsets` [n+1] host.  The batch is cut into `n_chunks` groups of whole utterances.  All
        host->device copies are queued up front on a copy stream (one event per group), so the copy engine streams
        the PCM without waiting for the host; MFCC / GMM / Viterbi of group i (compute stream) start when its copy
        has landed and overlap the copies of the later groups.  Offsets of every group travel in ONE pinned upload.
        Returns (words int32 [n], best float64 [n]) as CPU tensors.  return_paths=True also brings the Viterbi state paths
        of every model back to the host (`PendingRecognition.paths()`): for one large model the path IS the decode result."""
        t = _dev.torch()
        want_paths = want_paths or return_paths
        off = np.asarray(sample_offsets, dtype=np.int64)
        n_utt = len(off) - 1
        if n_utt <= 0:
            return PendingRecognition(t.empty((0,), dtype=t.int32), t.empty((0,), dtype=t.float64), None, (),
                                      paths=t.empty((0,), dtype=t.int32) if return_paths else None, path_blocks=[], n_models=self.bank.K)
        n_chunks = max(1, min(n_chunks, n_utt))
        # group boundaries at utterance borders.  Tapered sizes (1, 3, 4, .., 4, 3, 1): copy and compute are nearly balanced,
        # so the time to the first result (first copy) and the tail after the last copy (last group's compute) are what is
        # left to shave -- both want small groups at the ends, while few large groups in the middle keep the per-group
        # launch and tail costs low.
        if taper and n_chunks >= 3:
            wts = np.minimum(np.minimum(1 + 2 * np.arange(n_chunks), 1 + 2 * np.arange(n_chunks)[::-1]), 4).astype(np.float64)
        else:
            wts = np.ones(n_chunks)
        fr = np.cumsum(wts)[:-1] / wts.sum()
        targets = off[0] + (off[-1] - off[0]) * fr
        cuts = np.unique(np.concatenate([[0], np.searchsorted(off, targets), [n_utt]]))
        bounds = [(int(cuts[i]), int(cuts[i + 1])) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]
        # The path is bound by the one host->device copy engine, so (1) the first group's PCM goes out before any other host
        # work, (2) the per-group offsets are computed while it flies and queued right behind it (they are tiny, but a copy
        # queued after ALL the PCM would arrive last), (3) then the remaining groups.
        dev = _dev.device()
        cur = t.cuda.current_stream()
        if not hasattr(self, "_streams"):
            self._streams = (t.cuda.Stream(), t.cuda.Stream(), t.cuda.Stream())
            self._slots = [_HostSlot() for _ in range(3)]               # batches in flight / not yet collected
            self._stream_bufs = [_HostSlot(), _HostSlot()]              # per compute stream: workspace, scratch paths
            self._slot_i = 0
        copy_s, comp_a, comp_b = self._streams
        comps = (comp_a, comp_b) if n_compute_streams >= 2 else (comp_a,)
        slot = self._slots[self._slot_i % len(self._slots)]
        self._slot_i += 1
        prev = slot.handle() if slot.handle is not None else None
        if prev is not None:
            prev._wait()                                                 # its results leave the pinned buffers before they are reused
        copy_s.wait_stream(cur)
        if slot.done is not None:
            # the batch that used this slot three submissions ago has left the device (its pinned offset buffer is about to be
            # rewritten by the host, its device buffers by the copy stream); normally long complete
            slot.done.synchronize()
        total_samples = int(off[-1] - off[0])
        events = []
        pcm_dev = slot.get("pcm", total_samples, pcm_host.dtype, dev)

        def queue_pcm(lo, hi):
            s0, s1 = int(off[lo] - off[0]), int(off[hi] - off[0])
            pcm_dev[s0:s1].copy_(pcm_host[int(off[lo]):int(off[hi])], non_blocking=True)

        with t.cuda.stream(copy_s):
            queue_pcm(*bounds[0])
        # frame counts and offsets relative to each group's first sample / frame: [s_off | f_off] per group, one array
        n_fr = self.mfcc.frame_counts(off)
        meta, plan = [], []
        pos = 0
        for lo, hi in bounds:
            s_rel = off[lo:hi + 1] - off[lo]
            f_rel = _dev.offsets_from_lengths(n_fr[lo:hi])
            meta += [s_rel, f_rel]
            plan.append((pos, pos + len(s_rel), f_rel))
            pos += 2 * len(s_rel)
        meta_np = np.concatenate(meta)
        meta_host = slot.get("meta_host", meta_np.size, t.int64, pinned=True)
        meta_host.numpy()[:] = meta_np
        meta_dev = slot.get("meta", meta_np.size, t.int64, dev)
        max_frames = max(int(p[2][-1]) for p in plan)
        with t.cuda.stream(copy_s):
            meta_dev.copy_(meta_host, non_blocking=True)
            for i, (lo, hi) in enumerate(bounds):
                if i > 0:
                    queue_pcm(lo, hi)
                ev = t.cuda.Event(enable_timing=_timeline is not None)
                ev.record(copy_s)
                events.append(ev)
        for cs in comps:
            cs.wait_stream(cur)
        K = self.bank.K
        total_frames = int(n_fr.sum())
        words = slot.get("words", n_utt, t.int32, dev)
        best = slot.get("best", n_utt, t.float64, dev)
        paths_all = slot.get("paths", K * total_frames, t.int32, dev) if return_paths else None
        # groups alternate between the compute streams: the latency-bound tail of one group (Viterbi chains, kernel tails)
        # overlaps the head of the next.  One C call per group (lnasr_recognize_batch: MFCC -> emissions -> Viterbi -> word).
        # Workspaces belong to a compute STREAM (stream order makes their reuse by the next group or batch safe).
        lib = _capi.lib()
        rec = self._recognizer_struct()
        max_utt = max(h - l for l, h in bounds)
        need = C.c_size_t(0)
        _capi.check(lib.lnasr_recognize_workspace_bytes(max_frames, max_utt, C.byref(rec), int(want_paths), C.byref(need)),
                    "recognize_workspace_bytes")
        ws = [self._stream_bufs[i].get("ws", need.value, t.uint8, dev) for i in range(len(comps))]
        paths = [self._stream_bufs[i].get("paths", K * max_frames, t.int32, dev) if (want_paths and not return_paths) else None
                 for i in range(len(comps))]
        f_start = np.concatenate([[0], np.cumsum([int(p[2][-1]) for p in plan])])
        isz = pcm_dev.element_size()
        code = _dev.np_dtype_code(pcm_dev.dtype)
        for i, ((lo, hi), (p0, p1, f_rel), ev) in enumerate(zip(bounds, plan, events)):
            cs = comps[i % len(comps)]
            cs.wait_event(ev)
            s0 = int(off[lo] - off[0])
            nu = hi - lo
            _capi.check(lib.lnasr_recognize_batch(
                C.c_void_p(pcm_dev.data_ptr() + s0 * isz), code, C.c_void_p(meta_dev.data_ptr() + 8 * p0),
                C.c_void_p(meta_dev.data_ptr() + 8 * p1), nu, int(f_rel[-1]), C.byref(rec),
                C.c_void_p(words.data_ptr() + 4 * lo), C.c_void_p(best.data_ptr() + 8 * lo),
                C.c_void_p(paths_all.data_ptr() + 4 * K * int(f_start[i])) if return_paths else _capi.ptr(paths[i % len(comps)]),
                _capi.ptr(ws[i % len(comps)]), need.value, C.c_void_p(cs.cuda_stream)), "recognize_batch")
            if _timeline is not None:                      # diagnostic: (copy-done event, compute-done event) per group
                done = t.cuda.Event(enable_timing=True)
                done.record(cs)
                _timeline.append((ev, done))
        for cs in comps[1:]:
            comps[0].wait_stream(cs)
        out_w = slot.get("out_w", n_utt, t.int32, pinned=True)
        out_b = slot.get("out_b", n_utt, t.float64, pinned=True)
        out_p = slot.get("out_p", K * total_frames, t.int32, pinned=True) if return_paths else None
        with t.cuda.stream(comps[0]):
            out_w.copy_(words, non_blocking=True)
            out_b.copy_(best, non_blocking=True)
            if return_paths:
                out_p.copy_(paths_all, non_blocking=True)
            done = t.cuda.Event()
            done.record(comps[0])
        slot.done = done
        blocks = [(int(f_start[i]), int(f_start[i + 1])) for i in range(len(plan))]
        h = PendingRecognition(out_w, out_b, done, (pcm_host,), paths=out_p, path_blocks=blocks, n_models=K)
        slot.handle = weakref.ref(h)
        return h

    def _recognizer_struct(self):
        """lnasr_recognizer for the current device (cached; keeps the referenced structs alive)."""
        t = _dev.torch()
        dev = t.cuda.current_device()
        cache = self.__dict__.setdefault("_rec_cache", {})
        if dev not in cache:
            tb = self.mfcc._device_tables()
            bank = self.bank
            tc = bank._tc is not None
            if not tc and bank.gmm.full:
                raise _capi.LnasrError("lnasr_recognize_batch chains the diagonal-covariance emission kernels; a full-covariance bank "
                                       "is scored by the staged path (AcousticPipeline.recognize_batch)")
            packed = bank._tc.packed if tc else bank.gmm.packed
            rec = _capi.Recognizer(mfcc=C.pointer(tb), bank=C.pointer(bank.trans.struct), gmm_packed=packed.data_ptr(),
                                   gmm_kind=(2 if bank._tc.fmt == "f16" else 1) if tc else 0, n_mix=bank.m, dim=bank.d)
            cache[dev] = (rec, tb, packed)
        return cache[dev][0]

    def recognize_batch_fused(self, pcm, sample_offsets, want_paths=False):
        """Device-resident decode through the single C entry point lnasr_recognize_batch.  pcm: CUDA tensor (int16/float32);
        returns (words int32 [n], best float64 [n], paths int32 [K, total_frames] or None) on the device."""
        t = _dev.torch()
        lib = _capi.lib()
        off = np.asarray(sample_offsets, dtype=np.int64)
        n_utt = len(off) - 1
        if n_utt <= 0:
            return _dev.empty((0,), t.int32), _dev.empty((0,), t.float64), (_dev.empty((self.bank.K, 0), t.int32) if want_paths else None)
        f_off = _dev.offsets_from_lengths(self.mfcc.frame_counts(off))
        total = int(f_off[-1])
        meta = _dev.to_dev(np.concatenate([off - off[0], f_off]), t.int64)
        rec = self._recognizer_struct()
        need = C.c_size_t(0)
        _capi.check(lib.lnasr_recognize_workspace_bytes(total, n_utt, C.byref(rec), int(want_paths), C.byref(need)),
                    "recognize_workspace_bytes")
        ws = _dev.empty((need.value,), t.uint8)
        words, best = _dev.empty((n_utt,), t.int32), _dev.empty((n_utt,), t.float64)
        paths = _dev.empty((self.bank.K, total), t.int32) if want_paths else None
        pcm = _dev.to_dev(pcm)
        _capi.check(lib.lnasr_recognize_batch(
            _capi.ptr(pcm), _dev.np_dtype_code(pcm.dtype), _capi.ptr(meta), C.c_void_p(meta.data_ptr() + 8 * (n_utt + 1)), n_utt,
            total, C.byref(rec), _capi.ptr(words), _capi.ptr(best), _capi.ptr(paths), _capi.ptr(ws), need.value,
            _capi.stream_ptr()), "recognize_batch")
        return words, best, paths

    def recognize(self, signals: Sequence[np.ndarray]):
        """Convenience: list of utterances -> numpy array of word indices."""
        arrs = [np.asarray(s) for s in signals]
        i16 = all(a.dtype == np.int16 for a in arrs)
        pcm, off = _dev.pack_ragged(arrs, np.int16 if i16 else np.float32)
        res, _ = self.recognize_batch(pcm, off)
        return res.words.cpu().numpy()

"""MFCC front end -- API mirror of the reference's lnasr/mfcc.py, computed by csrc/mfcc.cu on a B200.

`MFCC(...)` takes the reference constructor arguments (mfcc.py:23-48), exposes the same attributes and
methods, and `calc_mfcc(signal)` returns the same `(frames_power, frames_cepstrum, features)` tuple
(mfcc.py:108-175) as float64 numpy arrays.  Additions: `calc_mfcc_batch` for ragged batches of
utterances that stay on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _capi, _dev

__all__ = ["MFCC", "MfccBatch"]


class MfccBatch(NamedTuple):
    """Device-resident result of `MFCC.calc_mfcc_batch`."""
    feats: "object"            # [total_frames, 39] float32 CUDA tensor
    frame_offsets: "object"    # [n_utt+1] int64 CUDA tensor
    frame_offsets_host: np.ndarray
    power: Optional["object"]  # [total_frames, 257] float32 or None
    ceps: Optional["object"]   # [total_frames, 40] float32 or None


class MFCC:
    """39-dimensional MFCC features: 12 cepstra + log energy, with deltas and delta-deltas."""

    D = 39   # mfcc.py:21

    def __init__(self, fs=16000, frame_T=25e-3, frame_stride=10e-3, hpf_alpha=0.97, fft_N=512,
                 mfbank_num=40, cepstrum_num=12):
        self.fs = fs
        self.frame_T = frame_T
        self.frame_L = int(self.fs * self.frame_T)
        self.frame_stride = frame_stride
        self.frame_step = int(self.fs * self.frame_stride)
        self.hpf_alpha = hpf_alpha
        self.fft_N = fft_N
        self.fft_size = self.fft_N // 2 + 1
        self.mfbank_num = mfbank_num
        self.cepstrum_num = cepstrum_num
        self._tables = {}          # device index -> (struct, keep-alive tensors)

    # ------------------------------------------------------------------ small host formulas
    def calc_mel(self, hz):
        """Hertz -> mel, 2595 ln(1 + hz/700): natural log, as the reference (mfcc.py:66-72)."""
        return 2595 * np.log(1 + hz / 700)

    def calc_mel_inv(self, mel):
        """mel -> Hertz (mfcc.py:74-76)."""
        return 700 * (np.exp(mel / 2595) - 1)

    def create_filter_bank(self) -> np.ndarray:
        """[mfbank_num, fft_N/2+1] triangular filters on FFT bins (mfcc.py:78-106).  A constant table:
        built on the host once, uploaded in banded form for the kernel."""
        M, N = self.mfbank_num, self.fft_N
        edges = np.floor((N / self.fs) * self.calc_mel_inv(
            np.linspace(self.calc_mel(0), self.calc_mel(self.fs / 2), M + 2))).astype(np.int64)
        bank = np.zeros((M, N // 2 + 1))
        for m in range(M):
            lo, mi, hi = edges[m:m + 3]
            k_up, k_dn = np.arange(lo, mi), np.arange(mi, hi)
            bank[m, k_up] = (k_up - lo) / (mi - lo) if mi > lo else 0.0
            bank[m, k_dn] = (hi - k_dn) / (hi - mi) if hi > mi else 0.0
        return bank

    # ------------------------------------------------------------------ device tables
    def _device_tables(self):
        t = _dev.torch()
        dev = t.cuda.current_device()
        if dev in self._tables:
            return self._tables[dev][0]
        if self.fft_N != 512:
            raise _capi.LnasrError("this build supports fft_N=512 only")
        L = self.frame_L
        window = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(L) / (L - 1))            # utils.py:138
        ang = 2 * np.pi * np.arange(self.fft_N) / self.fft_N
        twiddle = np.stack([np.cos(ang), -np.sin(ang)], axis=1)
        bank = self.create_filter_bank()
        nz = bank != 0
        if not nz.any(axis=1).all():
            raise _capi.LnasrError("a mel filter has no non-zero bin; unsupported filter-bank geometry")
        start = nz.argmax(axis=1)
        stop = bank.shape[1] - nz[:, ::-1].argmax(axis=1)
        length = stop - start
        stride = int(max(1, length.max()))
        if stride > 64:
            raise _capi.LnasrError("mel filter wider than 64 bins is not supported by this build")
        weights = np.zeros((self.mfbank_num, stride))
        for m in range(self.mfbank_num):
            weights[m, :length[m]] = bank[m, start[m]:stop[m]]
        # DCT-II, norm='ortho' (scipy.fftpack.dct as used at mfcc.py:152), stored transposed [n, k]
        Nm = self.mfbank_num
        n = np.arange(Nm)
        k = np.arange(Nm)[:, None]
        dct = 2.0 * np.cos(np.pi * k * (2 * n + 1) / (2 * Nm)) * np.sqrt(1.0 / (2 * Nm))
        dct[0] *= np.sqrt(0.5)
        keep = {
            "window": _dev.to_dev(window, t.float64),
            "twiddle": _dev.to_dev(twiddle, t.float64),
            "fb_start": _dev.to_dev(start.astype(np.int32), t.int32),
            "fb_len": _dev.to_dev(length.astype(np.int32), t.int32),
            "fb_weight": _dev.to_dev(weights, t.float32),
            "dct_t": _dev.to_dev(dct.T, t.float64),
        }
        tb = _capi.MfccTables(
            frame_len=self.frame_L, frame_step=self.frame_step, fft_n=self.fft_N, n_mel=self.mfbank_num,
            n_ceps=self.cepstrum_num, hpf_alpha=float(self.hpf_alpha), fb_stride=stride,
            window=keep["window"].data_ptr(), twiddle=keep["twiddle"].data_ptr(),
            fb_start=keep["fb_start"].data_ptr(), fb_len=keep["fb_len"].data_ptr(),
            fb_weight=keep["fb_weight"].data_ptr(), dct_t=keep["dct_t"].data_ptr(), mel_plan=None)
        # the mel-product plan of this filter bank, derived once on the device (lnasr_mfcc_mel_plan)
        lib = _capi.lib()
        keep["mel_plan"] = t.empty((int(lib.lnasr_mfcc_mel_plan_bytes()),), dtype=t.uint8, device="cuda")
        _capi.check(lib.lnasr_mfcc_mel_plan(C.byref(tb), _capi.ptr(keep["mel_plan"]), _capi.stream_ptr()), "mfcc_mel_plan")
        # the tables are used from any stream afterwards (AcousticPipeline's compute streams): the plan must be complete first
        t.cuda.current_stream().synchronize()
        tb.mel_plan = keep["mel_plan"].data_ptr()
        self._tables[dev] = (tb, keep)
        return tb

    def num_frames(self, n_samples: int) -> int:
        """ceil(|S - (L - step)| / step), the reference's frame count (utils.py:118)."""
        return int(_capi.lib().lnasr_mfcc_num_frames(int(n_samples), self.frame_L, self.frame_step))

    # ------------------------------------------------------------------ reference methods
    def calc_high_pass_filter(self, signal: np.ndarray) -> np.ndarray:
        """y[0]=x[0]; y[t]=x[t]-alpha*x[t-1] (mfcc.py:50-64); float64 result like the reference."""
        t = _dev.torch()
        sig = np.asarray(signal)
        if sig.dtype == np.int16:
            x, code = _dev.to_dev(sig, t.int16), _capi.DT_I16
        else:
            x, code = _dev.to_dev(sig.astype(np.float64), t.float64), _capi.DT_F64
        out = _dev.empty((x.numel(),), t.float64)
        _capi.check(_capi.lib().lnasr_preemphasis(_capi.ptr(x), code, x.numel(), float(self.hpf_alpha),
                                                  _capi.ptr(out), _capi.stream_ptr()), "preemphasis")
        return out.cpu().numpy()

    def calc_mfcc(self, signal: np.ndarray):
        """(frames_power [F,257], frames_cepstrum [F,40], features [F,39]) of one utterance (mfcc.py:108-175)."""
        res = self.calc_mfcc_batch([signal], want_power=True, want_ceps=True)
        return (res.power.double().cpu().numpy(), res.ceps.double().cpu().numpy(), res.feats.double().cpu().numpy())

    # ------------------------------------------------------------------ batched, device-resident
    def frame_counts(self, sample_offsets) -> np.ndarray:
        """Frames per utterance from host sample offsets [n+1]: ceil(|S-(L-step)|/step) (utils.py:118), vectorised."""
        lens = np.diff(np.asarray(sample_offsets, dtype=np.int64))
        step = self.frame_step
        n_fr = (np.abs(lens - (self.frame_L - step)) + step - 1) // step
        if len(n_fr) and n_fr.min() < 2:
            # the reference indexes row 1 of the feature matrix (mfcc.py:164): fewer than 2 frames is an IndexError there
            raise IndexError("calc_mfcc needs at least 2 frames per utterance (index 1 is out of bounds)")
        return n_fr

    def calc_mfcc_batch(self, signals, sample_offsets=None, want_power=False, want_ceps=False, feats_out=None,
                        _plan=None) -> MfccBatch:
        """MFCC of a ragged batch.

        signals: a list of 1-D arrays (numpy, int16 or float), or one concatenated 1-D CUDA/CPU tensor / array
        together with `sample_offsets` ([n_utt+1], host).  Returns device tensors (see MfccBatch).
        `_plan` = (sample_offsets_dev, frame_offsets_dev, frame_offsets_host, workspace) skips the host-side
        offset arithmetic and uploads (used by AcousticPipeline.recognize_batch_host)."""
        t = _dev.torch()
        lib = _capi.lib()
        if sample_offsets is None:
            arrs = [np.asarray(s) for s in signals]
            # int16 and float32 audio stay as they are; anything else is scored from float64 samples like the reference does
            if all(a.dtype == np.int16 for a in arrs):
                nd, td = np.int16, t.int16
            elif all(a.dtype == np.float32 for a in arrs):
                nd, td = np.float32, t.float32
            else:
                nd, td = np.float64, t.float64
            pcm_host, s_off = _dev.pack_ragged(arrs, nd)
            pcm = _dev.to_dev(pcm_host, td)
        else:
            s_off = sample_offsets if _plan is not None else np.asarray(sample_offsets, dtype=np.int64)
            if isinstance(signals, t.Tensor):
                pcm = signals if signals.dtype in (t.int16, t.float32, t.float64) else signals.to(t.float64)
                pcm = _dev.to_dev(pcm)
            else:
                a = np.asarray(signals)
                pcm = _dev.to_dev(a if a.dtype in (np.int16, np.float32, np.float64) else a.astype(np.float64))
        ws = None
        if _plan is not None:
            s_off_d, f_off_d, f_off, ws = _plan
            n_utt = len(f_off) - 1
        else:
            n_utt = len(s_off) - 1
            f_off = _dev.offsets_from_lengths(self.frame_counts(s_off))
            meta = _dev.to_dev(np.concatenate([s_off, f_off]), t.int64)      # one upload for both offset arrays
            s_off_d, f_off_d = meta[:n_utt + 1], meta[n_utt + 1:]
        total = int(f_off[-1])
        ncol = 3 * (self.cepstrum_num + 1)
        feats = feats_out if feats_out is not None else _dev.empty((total, ncol), t.float32)
        power = _dev.empty((total, self.fft_size), t.float32) if want_power else None
        ceps = _dev.empty((total, self.mfbank_num), t.float32) if want_ceps else None
        if total:
            need = C.c_size_t(0)
            _capi.check(lib.lnasr_mfcc_workspace_bytes(total, n_utt, C.byref(need)), "mfcc_workspace_bytes")
            if ws is None or ws.numel() < need.value:
                ws = _dev.empty((need.value,), t.uint8)
            tb = self._device_tables()
            _capi.check(lib.lnasr_mfcc_batch(
                _capi.ptr(pcm), _dev.np_dtype_code(pcm.dtype), _capi.ptr(s_off_d), _capi.ptr(f_off_d), n_utt, total,
                C.byref(tb), _capi.ptr(feats), feats.stride(0), _capi.ptr(power), _capi.ptr(ceps),
                _capi.ptr(ws), ws.numel(), _capi.stream_ptr()), "mfcc_batch")
        return MfccBatch(feats, f_off_d, f_off, power, ceps)

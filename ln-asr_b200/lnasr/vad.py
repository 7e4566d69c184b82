"""Voice activity detection by long-term spectral divergence -- API mirror of the reference's third/vadlstd/VadLstd.py
(class VadLtsd, VadLstd.py:21-74), computed by csrc/vad.cu on a B200.  The step before the acoustic path:
it shares the framing rule and the Hamming window of the MFCC front end (lnasr.utils.split_frames / create_hamming).
"""
from __future__ import annotations

import numpy as np

from . import _capi, _dev

__all__ = ["VadLtsd"]


class VadLtsd:
    """LTSD in dB per frame of `winsize` samples every `stepsize` (VadLstd.py:24-42: same constructor arguments)."""

    def __init__(self, freq=16000, winsize=2048, stepsize=1024, order=6, threshold=-6, alpha=None):
        self.freq = freq
        self.winsize = winsize
        self.stepsize = stepsize
        self.order = order
        self.threshold = threshold
        self.alpha = alpha
        self.fft_size = self.winsize // 2 + 1
        self._tables = {}

    def _device_tables(self):
        t = _dev.torch()
        dev = t.cuda.current_device()
        if dev not in self._tables:
            N = self.winsize
            window = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(N) / (N - 1))          # utils.py:138
            ang = 2 * np.pi * np.arange(N // 2) / N
            tw = np.stack([np.cos(ang), -np.sin(ang)], axis=1)
            self._tables[dev] = (_dev.to_dev(window, t.float64), _dev.to_dev(tw, t.float64))
        return self._tables[dev]

    def detect(self, data: np.ndarray) -> np.ndarray:
        """LTSD of every frame, float64 [frames] (VadLstd.py:44-74); frames closer than `order` to either end stay 0."""
        t = _dev.torch()
        lib = _capi.lib()
        x = np.asarray(data)
        if x.dtype == np.int16:
            xd = _dev.to_dev(x, t.int16)
        elif x.dtype == np.float32:
            xd = _dev.to_dev(x, t.float32)
        else:
            xd = _dev.to_dev(x.astype(np.float64), t.float64)
        n = int(xd.numel())
        F = int(lib.lnasr_vad_num_frames(n, self.winsize, self.stepsize))
        out = _dev.zeros((F,), t.float64)
        if F == 0:
            return out.cpu().numpy()
        nbytes = int(lib.lnasr_vad_workspace_bytes(n, self.winsize, self.stepsize))
        ws = _dev.empty((nbytes,), t.uint8)
        window, tw = self._device_tables()
        _capi.check(lib.lnasr_vad_ltsd(_capi.ptr(xd), _dev.np_dtype_code(xd.dtype), n, self.winsize, self.stepsize, self.order,
                                       float(self.threshold), int(self.alpha is not None), float(self.alpha or 0.0),
                                       _capi.ptr(window), _capi.ptr(tw), _capi.ptr(out), _capi.ptr(ws), nbytes,
                                       _capi.stream_ptr()), "vad_ltsd")
        return out.cpu().numpy()

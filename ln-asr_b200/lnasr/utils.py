"""Host-side helpers with the names of the reference's lnasr/utils.py (star-imported by recognizer.py:11).

On the acoustic hot path the work of `split_frames`, `create_hamming` and `lse` is folded into the CUDA
kernels (csrc/mfcc.cu, csrc/trellis.cu, csrc/em.cu).  The functions below exist for API parity: they are
small numpy utilities for callers that used them directly (table building, scripts), and audio file I/O.
Nothing in lnasr.mfcc / lnasr.hmm / lnasr.gmmhmm routes data through them.
"""
from __future__ import annotations

import functools

import numpy as np

__all__ = [
    "Punctuation_Ascii_Set", "Punctuation_Unicode_Set", "recording", "read_wave", "read_pcm",
    "split_frames", "create_hamming", "lse", "lse2", "np",
]

# sentence punctuation stripped by the n-gram tokenizer (reference lnasr/utils.py:13-49)
Punctuation_Ascii_Set = set(".,?!\"':;()[]{}")
Punctuation_Unicode_Set = set("。，？！：；、‘’“”—《》（）【】")


def recording(time, pcmfile=None, wavefile=None) -> np.ndarray:
    """Record `time` seconds of 16 kHz / 16 bit / mono PCM from the default input (utils.py:51-88).

    Needs `pyaudio`, imported lazily so that the package works where it is absent."""
    import wave
    import pyaudio
    pa = pyaudio.PyAudio()
    stream = pa.open(rate=16000, channels=1, format=pyaudio.paInt16, input=True, frames_per_buffer=160)
    print("start recording...")
    chunks = [stream.read(160) for _ in range(int(16000 / 160 * time))]
    print("stop recording")
    stream.stop_stream()
    stream.close()
    width = pa.get_sample_size(pyaudio.paInt16)
    pa.terminate()
    data = np.frombuffer(b"".join(chunks), dtype=np.int16)
    if pcmfile is not None:
        data.tofile(pcmfile)
    if wavefile is not None:
        with wave.open(wavefile, "wb") as fp:
            fp.setnchannels(1)
            fp.setsampwidth(width)
            fp.setframerate(16000)
            fp.writeframes(data)
    return data


def read_wave(filename: str):
    """(int16 samples at 16 kHz mono, 16000) (utils.py:90-98).  Uses librosa when present, else the
    standard `wave` module for files that already are 16 kHz / 16 bit / mono."""
    try:
        import librosa
    except ImportError:
        import wave
        with wave.open(filename, "rb") as fp:
            if fp.getframerate() != 16000 or fp.getnchannels() != 1 or fp.getsampwidth() != 2:
                raise RuntimeError("read_wave without librosa handles only 16 kHz / 16 bit / mono files")
            return np.frombuffer(fp.readframes(fp.getnframes()), dtype=np.int16), 16000
    wdata, sr = librosa.core.load(filename, sr=16000, mono=True)
    return (wdata * pow(2, 15)).astype(np.int16), sr


def read_pcm(filename: str) -> np.ndarray:
    """Raw PCM file: mono, 16 bit, little endian (utils.py:100-104)."""
    return np.fromfile(filename, dtype="<i2")


def num_frames(n_samples: int, L: int, S: int) -> int:
    """Frame-count rule of split_frames, including its abs() (utils.py:118)."""
    return int(np.ceil(np.abs(n_samples - (L - S)) / S))


def split_frames(signal: np.ndarray, L, S):
    """[n_frames, L] view-copy of `signal` with hop S, the tail zero-padded (utils.py:106-130)."""
    signal = np.asarray(signal)
    n = num_frames(len(signal), L, S)
    need = n * S + (L - S)
    if need > len(signal):
        signal = np.append(signal, np.zeros(need - len(signal)))
    return signal[np.arange(L)[None, :] + (S * np.arange(n))[:, None]]


def create_hamming(N: int) -> np.ndarray:
    """0.54 - 0.46 cos(2 pi n / (N-1)) (utils.py:132-138)."""
    return 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(N) / (N - 1))


def _fold(seq, axis, op):
    seq = np.asarray(seq)
    if axis is None:
        return _fold(seq.ravel(), 0, op)
    if isinstance(axis, tuple):
        for ax in reversed(axis):          # ascending tuple, reduced from the last axis (utils.py:166-167)
            seq = _fold(seq, ax, op)
        return seq
    if seq.shape[axis] == 0:
        return np.full(tuple(d for d in seq.shape if d != 0), -np.inf)
    # left-to-right fold, the reference's summation order (utils.py:176-179)
    return functools.reduce(op, np.moveaxis(seq, axis, 0))


def lse(seq: np.ndarray, axis=None) -> np.ndarray:
    """log(sum(exp(seq))) over `axis` (None, int or ascending tuple) (utils.py:140-179)."""
    return _fold(seq, axis, np.logaddexp)


def lse2(seq: np.ndarray, axis=None) -> np.ndarray:
    """log2(sum(2**seq)) over `axis` (utils.py:181-204)."""
    return _fold(seq, axis, np.logaddexp2)

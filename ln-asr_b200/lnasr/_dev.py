"""Device plumbing shared by the API mirror: numpy <-> CUDA tensors, ragged batches.  PyTorch is used for
memory, streams and torch.distributed only; every arithmetic step of the hot path is a call into
liblnasr_b200.so (lnasr/_capi.py)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import _capi


def torch():
    return _capi.require_cuda()


def device():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def to_dev(x, dtype=None, dev=None):
    """numpy / tensor -> contiguous CUDA tensor of `dtype` (torch dtype)."""
    t = torch()
    dev = dev or device()
    if isinstance(x, t.Tensor):
        y = x.to(device=dev, dtype=dtype or x.dtype)
    else:
        a = np.ascontiguousarray(np.asarray(x))
        y = t.from_numpy(a).to(device=dev, dtype=dtype)
    return y.contiguous()


def empty(shape, dtype, dev=None):
    t = torch()
    return t.empty(shape, dtype=dtype, device=dev or device())


def zeros(shape, dtype, dev=None):
    t = torch()
    return t.zeros(shape, dtype=dtype, device=dev or device())


def offsets_from_lengths(lengths: Sequence[int]) -> np.ndarray:
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=off[1:])
    return off


def pack_ragged(seqs: Sequence, dtype) -> Tuple[np.ndarray, np.ndarray]:
    """list of [T_u, ...] arrays -> (concatenated [sum T_u, ...], offsets [n+1] int64)."""
    arrs = [np.asarray(s) for s in seqs]
    off = offsets_from_lengths([len(a) for a in arrs])
    if len(arrs) == 0:
        return np.zeros((0,), dtype=dtype), off
    return np.ascontiguousarray(np.concatenate(arrs, axis=0).astype(dtype, copy=False)), off


def split_ragged(x: np.ndarray, off: np.ndarray) -> List[np.ndarray]:
    return [x[off[u]:off[u + 1]] for u in range(len(off) - 1)]


def np_dtype_code(t_dtype) -> int:
    t = torch()
    if t_dtype == t.float32:
        return _capi.DT_F32
    if t_dtype == t.float64:
        return _capi.DT_F64
    if t_dtype == t.int16:
        return _capi.DT_I16
    raise _capi.LnasrError(f"unsupported dtype {t_dtype}")

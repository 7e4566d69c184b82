"""A stand-in for the part of h5py the reference's model files use (hmm.py:252-266, gmmhmm.py:190-208, recognizer.py:17):

    with h5py.File(name, 'w') as f: f.create_dataset('A', data=arr)
    f = h5py.File(name, 'r'); f['A'] -> dataset (array-like: np.array(ds), ds[...], ds.shape, ds.dtype)

h5py itself is absent from this image, so the package's save() / load() code could not be executed by any test.  This
module implements that API over a small self-describing container (NOT the HDF5 byte format) so that the code paths run:
tests install it as sys.modules['h5py'].  Interoperability with real .h5 files needs the real h5py."""
import json
import struct

import numpy as np

MAGIC = b"LNASRH5SHIM\0"


class Dataset:
    def __init__(self, arr):
        self._a = arr

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __getitem__(self, idx):
        return self._a[idx]

    def __len__(self):
        return len(self._a)


class File:
    def __init__(self, name, mode="r"):
        self.filename, self.mode = str(name), mode
        self._d = {}
        self._open = True
        if mode == "r":
            with open(self.filename, "rb") as f:
                if f.read(len(MAGIC)) != MAGIC:
                    raise OSError("not a file written by the h5py test shim")
                (n,) = struct.unpack("<Q", f.read(8))
                index = json.loads(f.read(n).decode())
                blob = f.read()
            for key, (dt, shape, off, size) in index.items():
                self._d[key] = Dataset(np.frombuffer(blob[off:off + size], dtype=np.dtype(dt)).reshape(shape).copy())
        elif mode != "w":
            raise ValueError("the shim supports modes 'r' and 'w'")

    def create_dataset(self, name, data=None, **kw):
        if self.mode != "w":
            raise OSError("file is not open for writing")
        if name in self._d:
            raise ValueError(f"dataset {name!r} already exists")
        self._d[name] = Dataset(np.ascontiguousarray(np.asarray(data)))
        return self._d[name]

    def __getitem__(self, name):
        return self._d[name]

    def __contains__(self, name):
        return name in self._d

    def keys(self):
        return self._d.keys()

    def close(self):
        if self._open and self.mode == "w":
            index, blobs, off = {}, [], 0
            for key, ds in self._d.items():
                raw = np.asarray(ds).tobytes()
                index[key] = (np.asarray(ds).dtype.str, list(np.asarray(ds).shape), off, len(raw))
                blobs.append(raw)
                off += len(raw)
            head = json.dumps(index).encode()
            with open(self.filename, "wb") as f:
                f.write(MAGIC + struct.pack("<Q", len(head)) + head + b"".join(blobs))
        self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

#!/usr/bin/env python3
"""Stage the UNMODIFIED reference hot-path modules for the GPU box  --  TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference (yehuohan/ln-asr) is pure Python: there is nothing to compile.  What `bench.py --impl reference` and the
`cpu_baseline` leg need on the GPU box -- where /root/reference does not exist -- is the reference's own code for the path:
lnasr/{__init__,utils,mfcc,gmm,hmm,gmmhmm}.py.  This script reads those six files from where they lie under /root/reference
and writes them, byte for byte, into ONE archive  oracle/_ref/lnasr_reference.zip  (Python imports packages straight from a
zip).  oracle/_ref/ is git-ignored (no reference source enters the repository or its history) but not gpurun-ignored, so the
archive travels with the snapshot like the built .so does.  `oracle/ref_loader.py` imports it under the import shims of
SURVEY.md 8(c).
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "lnasr_reference.zip")
MODULES = ("__init__", "utils", "mfcc", "gmm", "hmm", "gmmhmm")


def build(reference_root: str = None) -> str:
    """Returns the archive path, or '' when the reference tree is not present (the GPU box: the prebuilt archive is used)."""
    root = reference_root or os.environ.get("LNASR_REFERENCE", "/root/reference")
    src = os.path.join(root, "lnasr")
    if not os.path.isdir(src):
        return OUT if os.path.exists(OUT) else ""
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = {}
    tmp = OUT + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for m in MODULES:
            path = os.path.join(src, m + ".py")
            data = open(path, "rb").read()
            manifest[m + ".py"] = hashlib.sha256(data).hexdigest()
            info = zipfile.ZipInfo("lnasr/" + m + ".py", date_time=(2020, 1, 1, 0, 0, 0))     # fixed date: reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(info, data)
        z.writestr(zipfile.ZipInfo("MANIFEST.json", date_time=(2020, 1, 1, 0, 0, 0)),
                   json.dumps({"source": "yehuohan/ln-asr, unmodified", "sha256": manifest}, indent=1))
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    out = build(sys.argv[1] if len(sys.argv) > 1 else None)
    print(out or "reference tree not found and no prebuilt archive")

#!/usr/bin/env python3
"""Populate oracle/_ref/ with the reference's own two hot-path modules so that the CPU baseline of bench.py runs the UNMODIFIED reference code
(``kind: "reference"``) on the GPU box, where /root/reference does not exist.

TEST / BENCH INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (never part of the repository history) but is not gpurun-ignored, so it travels
with the snapshot like the built .so files.  Only ``model.py`` (ModelB_2 and its blocks) and ``utils.py`` (generate_psf_kernel,
downscale_LST_SR_to_LR, get_output_ftm, upsampling) are taken, byte for byte.  Run by ``__graft_entry__.build()`` whenever /root/reference
is present; without it bench.py falls back to the oracle port and says ``kind: "port"``.

    python oracle/make_ref.py [reference_root]
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("model.py", "utils.py")


def populate(ref_root: str = "/root/reference") -> bool:
    if not all(os.path.isfile(os.path.join(ref_root, f)) for f in FILES):
        return False
    os.makedirs(DEST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(ref_root, f), os.path.join(DEST, f))
        with open(os.path.join(DEST, f), "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}\n")
    with open(os.path.join(DEST, "SHA256SUMS"), "w") as fh:
        fh.writelines(lines)
    return True


if __name__ == "__main__":
    ok = populate(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("oracle/_ref populated" if ok else "reference not found: oracle/_ref left as is")

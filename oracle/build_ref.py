"""Recipe for oracle/_ref/: verbatim copies of the reference files the hot path consists of, so that the UNMODIFIED
reference PyTorch-CPU path can be timed on the GPU box's host cores (bench.py --impl reference, cpu_baseline) and used as
the checker there.  /root/reference exists only in the build container; oracle/_ref/ is git-ignored (the reference
sources never enter this repo's history) but NOT gpurun-ignored, so it travels with the snapshot like a built .so.

  python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Nothing is modified: the files are byte-identical (a sha256 manifest is written next to them and checked by
tests/test_ref_loader.py when both trees are present).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import REF_FILES  # noqa: E402

SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
# the two JSON configs of the vocoder are test fixtures already (tests/golden/nsf_configs); nothing else is needed
EXTRA = ("LICENSE",)


def build_ref(force=False):
    """-> True if oracle/_ref is in place (copied now or already there), False if /root/reference is absent."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    manifest = {}
    for rel in REF_FILES + EXTRA:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if force or not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
            shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "fishaudio/fish-diffusion (/root/reference), verbatim", "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build_ref(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference)")

"""Recipe for oracle/_ref: a verbatim copy of the reference's own implementation of the hot path (TEST INFRASTRUCTURE).

    python oracle/make_ref.py            (also run by __graft_entry__.build() whenever /root/reference is mounted)

The reference is pure Python (no build system, nothing to compile): the "build" is a byte-for-byte copy of the modules that
`import vit` pulls in — vit.py, layers.py, criterions.py and what layers.py imports at module level (autoencoders.py,
hamburger/, nnmf/) — from where they lie under /root/reference into oracle/_ref/.  That directory is git-ignored (no reference
source enters the history) but travels to the GPU box with the snapshot, so `bench.py --impl reference` / `--impl eager`
and the cpu_baseline leg time the reference ITSELF there (kind "reference") instead of the oracle port, and the live-reference
tests of tests/test_oracle.py can run on the box too.  A manifest with the sha256 of every copied file is written next to them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

SRC = os.environ.get("VITB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["vit.py", "layers.py", "criterions.py", "autoencoders.py"]
PACKAGES = ["hamburger", "nnmf"]


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "vit.py")):
        if verbose:
            print(f"make_ref: {SRC} not mounted; keeping whatever is in {DST}")
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    todo = [(f, f) for f in FILES]
    for pkg in PACKAGES:
        for root, _, names in os.walk(os.path.join(SRC, pkg)):
            for n in names:
                if n.endswith(".py"):
                    rel = os.path.relpath(os.path.join(root, n), SRC)
                    todo.append((rel, rel))
    for src_rel, dst_rel in todo:
        s, d = os.path.join(SRC, src_rel), os.path.join(DST, dst_rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[dst_rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"make_ref: copied {len(manifest)} files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() or os.path.isdir(DST) else 1)

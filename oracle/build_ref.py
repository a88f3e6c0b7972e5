"""Recipe for oracle/_ref: the UNMODIFIED reference objective.py placed where it travels to the GPU box.

    python oracle/build_ref.py            # in the build container (the only place /root/reference exists)

The reference's hot path is one pure-Python/PyTorch file without a build step, so "building" it is a byte-for-byte
copy of /root/reference/objective.py into oracle/_ref/objective.py plus a sha256 manifest.  oracle/_ref/ is listed in
.gitignore (reference sources never enter the history) but NOT in .gpurunignore, so it ships with the snapshot like the
built .so files.  `bench.py --impl reference` and the cpu_baseline leg import it (kind "reference"); when it is absent
(a checkout without the reference) they fall back to the oracle's dense port (kind "port").  TEST / BENCH INFRASTRUCTURE.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF_PATH", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(quiet: bool = False) -> bool:
    src = os.path.join(REF, "objective.py")
    if not os.path.isfile(src):
        if not quiet:
            print(f"{src} not present: oracle/_ref not (re)built", file=sys.stderr)
        return os.path.isfile(os.path.join(OUT, "objective.py"))
    os.makedirs(OUT, exist_ok=True)
    dst = os.path.join(OUT, "objective.py")
    shutil.copyfile(src, dst)
    digest = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": digest, "note": "verbatim copy, never committed (see .gitignore)"}, f)
    if not quiet:
        print(f"oracle/_ref/objective.py <- {src} (sha256 {digest[:16]}...)")
    return True


def load():
    """The unmodified reference module from oracle/_ref, or None when it has not been built."""
    path = os.path.join(OUT, "objective.py")
    if not os.path.isfile(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_objective_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if build() else 1)

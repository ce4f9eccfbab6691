#!/usr/bin/env python3
"""Install the UNMODIFIED reference model files into the git-ignored baseline/_ref/ so that the GPU box (where
/root/reference does not exist) can run the reference nn.Module under stock PyTorch -- the "existing Blackwell
path" of SURVEY.md section 2.4 / BASELINE.md section 3 (`bench.py --impl torch-eager`).

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` cannot work;
this copies the handful of files the module needs, byte for byte, into the same package layout.  baseline/_ref is in
.gitignore (never committed) and NOT in .gpurunignore (travels with the snapshot).  Run in the build container only.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "team_code/mmfn_utils/models/model_rad.py",
    "team_code/mmfn_utils/models/model_vec.py",
    "team_code/mmfn_utils/models/model_img.py",
    "team_code/mmfn_utils/datasets/config.py",
    "team_code/benchmarks/transfuser/model.py",
    "team_code/benchmarks/transfuser/config.py",
]


def main():
    if not os.path.isdir(REF):
        print(f"{REF} is not present: nothing installed")
        return 1
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    print(f"installed {len(FILES)} reference files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Stage the UNMODIFIED reference envs under oracle/_ref/ so they can travel to the GPU box (TEST INFRASTRUCTURE ONLY).

The reference is pure Python, so there is nothing to compile: the four files its env package needs
(`swarm_marl/__init__.py`, `envs/{__init__,common,single_drone_env,drone_swarm_env}.py`) are copied byte for byte from
/root/reference into oracle/_ref/src/, which is git-ignored (no reference source enters the repo's history) but not
gpurun-ignored.  `bench.py --impl reference` and its `cpu_baseline` leg time these files on the GPU box's host cores
through oracle/ref_runner.py.  Run by `__graft_entry__.build()` wherever /root/reference exists.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SWARM_REFERENCE_ROOT", "/root/reference")
FILES = ("swarm_marl/__init__.py", "swarm_marl/envs/__init__.py", "swarm_marl/envs/common.py",
         "swarm_marl/envs/single_drone_env.py", "swarm_marl/envs/drone_swarm_env.py")


def main() -> bool:
    src_root = os.path.join(REFERENCE_ROOT, "src")
    if not os.path.isdir(os.path.join(src_root, "swarm_marl", "envs")):
        return False
    dst_root = os.path.join(HERE, "_ref", "src")
    manifest = {}
    for rel in FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": src_root, "sha256": manifest}, open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref staged" if main() else f"no reference tree at {REFERENCE_ROOT}: nothing staged")

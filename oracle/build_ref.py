#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference (SelennLamson/AntsRL) compiled from its sources where they lie.

The reference is pure Python; its "binary" is CPython bytecode.  This script byte-compiles the modules of the step-loop
path (environment/, environment/rewards/, generator/, utils.py) and the agents that consume it (agents/) from
``/root/reference`` into ONE archive of sourceless ``.pyc`` files with the reference's package layout,
``oracle/_ref/reference.zip`` (an archive because loose ``*.pyc`` files do not survive the snapshot to the GPU box), so

    sys.path.insert(0, "oracle/_ref/reference.zip"); import environment.RL_api

imports the reference's own code (zipimport) without its checkout.  No reference source is copied: ``oracle/_ref/``
holds build outputs only, is git-ignored and travels to the GPU box with the snapshot like the built ``.so``.  Test infrastructure: used by ``oracle/ref_harness.py`` (tests, ``bench.py --impl reference`` and bench's
``cpu_baseline`` leg) when ``/root/reference`` is absent.  The bytecode is tied to the interpreter that wrote it
(``MANIFEST.json`` records the magic number); the GPU box runs the same image.

    python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import hashlib
import importlib.util
import json
import os
import py_compile
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
# packages / modules of the path and of its callers; everything else of the reference (GUI, driver, keras agents)
# stays behind
MODULES = [
    "utils.py",
    "environment/__init__.py", "environment/RL_api.py", "environment/environment.py", "environment/ants.py",
    "environment/pheromone.py", "environment/walls.py", "environment/food.py", "environment/anthill.py",
    "environment/circle_obstacles.py",
    "environment/rewards/__init__.py", "environment/rewards/reward.py", "environment/rewards/reward_custom.py",
    "generator/environment_generator.py", "generator/map_generators.py",
    "agents/agent.py", "agents/replay_memory.py", "agents/explore_agent_pytorch.py", "agents/collect_agent.py",
    "agents/collect_agent_rework.py", "agents/collect_agent_memory.py",
]


ARCHIVE = os.path.join(OUT, "reference.zip")


def build(reference="/root/reference", force=False):
    """-> True if oracle/_ref is usable afterwards.  Without the reference checkout an existing build is kept."""
    manifest_path = os.path.join(OUT, "MANIFEST.json")
    magic = importlib.util.MAGIC_NUMBER.hex()
    if not os.path.isdir(reference):
        return os.path.exists(manifest_path) and os.path.exists(ARCHIVE)
    digests = {}
    for rel in MODULES:
        with open(os.path.join(reference, rel), "rb") as f:
            digests[rel] = hashlib.sha256(f.read()).hexdigest()
    if not force and os.path.exists(manifest_path) and os.path.exists(ARCHIVE):
        try:
            old = json.load(open(manifest_path))
            if old.get("magic") == magic and old.get("sha256") == digests and old.get("archive") == "reference.zip":
                return True
        except Exception:
            pass
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as zf:
        dirs = set()
        for rel in MODULES:
            d = os.path.dirname(rel)
            while d and d not in dirs:                       # explicit directory entries: zipimport needs them for the
                dirs.add(d)                                  # namespace packages (generator/, agents/ have no __init__.py)
                d = os.path.dirname(d)
        for d in sorted(dirs):
            zf.writestr(d + "/", b"")
        for k, rel in enumerate(MODULES):
            pyc = os.path.join(tmp, "%d.pyc" % k)
            py_compile.compile(os.path.join(reference, rel), cfile=pyc, dfile="<reference>/" + rel, doraise=True,
                               optimize=0, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            zf.write(pyc, rel + "c")                         # module.pyc where module.py would be
    with open(manifest_path, "w") as f:
        json.dump({"what": "SelennLamson/AntsRL byte-compiled by oracle/build_ref.py (unmodified reference, sourceless)",
                   "archive": "reference.zip", "python": sys.version.split()[0], "magic": magic, "modules": MODULES,
                   "sha256": digests}, f, indent=1)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force)
    print("oracle/_ref:", "ready" if ok else "NOT built (no reference checkout at %s)" % a.reference)
    sys.exit(0 if ok else 1)

#!/usr/bin/env python
"""Stage the UNMODIFIED reference tree under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that it
travels to the GPU box, where /root/reference does not exist.

    python baseline/stage_reference.py            # copies /root/reference -> baseline/_ref

The reference (dglai/dgl-0.5-benchmark) is 66 Python/markdown/docker files with no build system and no
package metadata, so "installing" it is a plain copy; its one dependency on this path, DGL v0.6.1
(docker/build.dockerfile:14), is not installable here and is what dgl-0.5-benchmark_b200/dgl replaces.
The `-m gpu` tests (tests/test_gpu_reference_scripts.py) execute these unchanged files on cuda:0 through
dgl-0.5-benchmark_b200/run_reference.py; a sha256 manifest is written next to them so the tests can
assert that what ran is byte-identical to what was staged.  Nothing under baseline/_ref is committed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("DGLB200_REFERENCE", "/root/reference")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(src=SRC, dest=DEST):
    """Copy the tree; returns the manifest {relative path: sha256}.  No-op (returns the existing manifest,
    or None) when the reference tree is not present, e.g. on the GPU box."""
    man_path = os.path.join(dest, "MANIFEST.json")
    if not os.path.isdir(src):
        return json.load(open(man_path)) if os.path.exists(man_path) else None
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    manifest = {}
    for root, _dirs, files in os.walk(src):
        for name in files:
            a = os.path.join(root, name)
            rel = os.path.relpath(a, src)
            b = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(b), exist_ok=True)
            shutil.copyfile(a, b)
            manifest[rel] = _sha(b)
    with open(man_path, "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    return manifest


def verify(dest=DEST):
    """True iff every staged file still matches the manifest written at staging time."""
    man_path = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(man_path):
        return False
    manifest = json.load(open(man_path))
    return all(os.path.exists(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == h
               for rel, h in manifest.items())


if __name__ == "__main__":
    m = stage()
    if m is None:
        sys.exit("reference tree %s not found and nothing staged" % SRC)
    print("staged %d files under %s" % (len(m), DEST))

"""Populate baseline/_ref/ with an UNMODIFIED copy of the reference's Python tree (the "install" of the reference arm).

The reference (ymxlzgy/echoscene) has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
has nothing to build: the install is a verbatim copy of its importable packages (model/, helpers/, dataset/, scripts/) and
config/.  baseline/_ref/ is git-ignored (the reference's sources never enter this repo's history) but NOT gpurun-ignored, so
it travels to the GPU box, where /root/reference does not exist.  Run here (build container) by __graft_entry__.build();
`python baseline/install_ref.py` does the same by hand.  Records what it copied in baseline/_ref/INSTALL.json.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("ECHOSCENE_REF_SRC", "/root/reference")
SUBTREES = ("model", "helpers", "dataset", "scripts", "config")
KEEP_EXT = (".py", ".yaml", ".yml", ".json", ".txt")


def install(force: bool = False) -> bool:
    """-> True when baseline/_ref is populated (now or earlier); False when there is no reference tree to copy from."""
    stamp = os.path.join(DEST, "INSTALL.json")
    if os.path.exists(stamp) and not force:
        return True
    if not os.path.isdir(os.path.join(SRC, "model")):
        return False
    os.makedirs(DEST, exist_ok=True)
    files, digest = [], hashlib.sha256()
    for sub in SUBTREES:
        for root, _dirs, names in os.walk(os.path.join(SRC, sub)):
            for n in sorted(names):
                if not n.endswith(KEEP_EXT):
                    continue
                src = os.path.join(root, n)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                with open(src, "rb") as f:
                    digest.update(rel.encode() + b"\0" + f.read())
                files.append(rel)
    with open(stamp, "w") as f:
        json.dump({"source": SRC, "files": len(files), "sha256": digest.hexdigest(), "modified": False,
                   "note": "verbatim copy of the reference's Python packages and configs; nothing patched"}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "installed" if ok else f"no reference tree at {SRC}")

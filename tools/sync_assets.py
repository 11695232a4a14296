#!/usr/bin/env python
"""Make the reference's assets available to the scene programs.

* copies ``/root/reference/assets`` (OBJ meshes, JPEG/PNG textures) into the
  git-ignored ``assets/`` directory — data files, not sources; they travel to
  the GPU box with the snapshot but stay out of history;
* bakes the meshes the core parity tests need into ``tests/golden/meshes/*.ptmesh``
  (committed): the exact vertex/index arrays the tobj-0.1.7-style OBJ loader of
  ``portrayer_b200/host/scene.cpp`` produces, so a checkout without ``assets/``
  can still build those scenes.

Run here (the container that has /root/reference); ``__graft_entry__.build()`` calls it.
"""
from __future__ import annotations

import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ASSETS = os.environ.get("PORTRAYER_REFERENCE_ASSETS", "/root/reference/assets")
BAKED = ["monkey", "prim_castle_door", "cow", "castle", "castle_door", "castle_door_arch", "castle_glass_ceilings",
         "castle_hill", "castle_stairs_side", "castle_tapestry", "castle_water_dirt", "castle_window_frames",
         "puppet_castle_left_tower", "puppet_castle_right_tower"]


def main() -> int:
    if not os.path.isdir(REFERENCE_ASSETS):
        print(f"{REFERENCE_ASSETS} not present: nothing to sync (baked meshes and stand-in textures will be used)")
        return 0
    dst = os.path.join(REPO, "assets")
    os.makedirs(dst, exist_ok=True)
    n = 0
    for root, _dirs, files in os.walk(REFERENCE_ASSETS):
        rel = os.path.relpath(root, REFERENCE_ASSETS)
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for f in files:
            if f == "README.md":
                continue
            src, out = os.path.join(root, f), os.path.join(dst, rel, f)
            if not os.path.exists(out) or os.path.getsize(out) != os.path.getsize(src):
                shutil.copyfile(src, out)
                n += 1
    print(f"assets/: {n} file(s) copied from {REFERENCE_ASSETS}")

    sys.path.insert(0, REPO)
    from portrayer_b200 import _ffi, assets  # noqa: F401

    golden = os.path.join(REPO, "tests", "golden", "meshes")
    os.makedirs(golden, exist_ok=True)
    for name in BAKED:
        obj = os.path.join(dst, name + ".obj")
        if not os.path.exists(obj):
            continue
        rc = _ffi.host.pth_bake_obj(obj.encode(), os.path.join(golden, name + ".ptmesh").encode())
        if rc != 0:
            print("bake failed:", name, _ffi.host.pth_last_error().decode())
            return 1
    print(f"tests/golden/meshes: {len(BAKED)} baked mesh(es)")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Write the inputs of tools/leafstats (blob, camera, params, background) for one scene program.
    python tools/leafstats/dump.py graphics-castle 960 540 1 /tmp/ls_castle
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np

import portrayer_b200 as pt
from portrayer_b200.render import _background_arg, make_params

name, w, h, samples, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
os.makedirs(out, exist_ok=True)
sc = pt.Scene.example(name)
sc.width, sc.height = w, h
bg, bg_mode = _background_arg(sc, w, h)
p = make_params(w, h, samples, "hash", 1, bg_mode=bg_mode, flags=32)
sc.blob.tofile(os.path.join(out, "blob.bin"))
open(os.path.join(out, "camera.bin"), "wb").write(bytes(sc.camera()))
open(os.path.join(out, "params.bin"), "wb").write(bytes(p))
np.ascontiguousarray(bg, dtype=np.float64).tofile(os.path.join(out, "background.bin"))
print("wrote", out)

import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import parity, portrayer_b200 as pt
from portrayer_b200 import _ffi
for name in ("big-scene", "graphics-castle"):
    scene = pt.Scene.example(name)
    ds = pt.DeviceScene(scene.blob)
    for flags, label in ((0, "prune"), (_ffi.PT_RENDER_EXACT_WALK, "exact"), (_ffi.PT_RENDER_TOLERATE_KD_PLANE, "prune+tol")):
        for rep in range(3):
            img = pt.Image(scene.width, scene.height)
            st = img.render(scene, samples=2, rng="hash", flags=flags | _ffi.PT_RENDER_TOLERATE_KD_PLANE, dscene=ds)
        print(name, label, "device_ms", round(st.device_ms, 3), "launches", st.kernel_launches)

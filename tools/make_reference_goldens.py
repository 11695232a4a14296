#!/usr/bin/env python
"""Freeze small fixtures of the reference's OWN published renders (/root/reference/render/*.png, produced upstream
at SAMPLES=100 with OS-seeded jitter) under tests/golden/reference_renders/, so that the oracle can be checked
against reference OUTPUT on machines where /root/reference does not exist.

    python tools/make_reference_goldens.py

Each fixture is the render box-filtered down to 1/7 (910x512 -> 130x73) and stored as an RGB8 PNG: small,
and insensitive to the per-pixel jitter noise the upstream image contains.  The comparison (tests/
test_reference_renders.py) is statistical (mean absolute error / PSNR), not bit-exact — that is all an image
rendered with OS randomness can pin.  Run here (this container has the reference); the fixtures are committed."""
import os

import numpy as np
from PIL import Image

REF = "/root/reference/render"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_renders")
FACTOR = 7
# reference render -> example scene program of the host mirror
RENDERS = {
    "01a_primitives-simple.png": "primitives-simple",
    "01b_primitives.png": "primitives",
    "02_smooth-shading.png": "smooth-shading",
    "04a_normal-mapping.png": "normal-mapping",
    "04b_normal-mapping-left.png": "normal-mapping-left",
    "04c_normal-mapping-right.png": "normal-mapping-right",
    "06a_water-glass.png": "water-glass",
    "06b_transmission-refraction.png": "transmission-refraction",
    "entering-the-mirror-dimension.png": "entering-the-mirror-dimension",
    # upstream published four colour variants; the example source as shipped is the green one
    "10_robot-alarm-clock_green.png": "robot-alarm-clock",
    # ... the other three are the same program with the commented-out diffuse lines (examples/robot-alarm-clock.rs:98-100)
    "10_robot-alarm-clock.png": "robot-alarm-clock-cyan",
    "10_robot-alarm-clock_dark_blue.png": "robot-alarm-clock-dark-blue",
    "10_robot-alarm-clock_red.png": "robot-alarm-clock-red",
    # (03_antialiasing.png is a hand-made 4x zoom montage of crops, not the output of examples/antialiasing.rs: mean
    # absolute error 64 LSB against the program's 300x250 render at either of its sample counts — nothing to pin with it;
    # 05a / 05b need the three assets missing upstream)
    "07_glossy-reflection.png": "glossy-reflection",
    "08_soft-shadows.png": "soft-shadows",
    "09a_kdtree.png": "big-scene",
}


def box_down(img: np.ndarray, f: int) -> np.ndarray:
    h, w = (img.shape[0] // f) * f, (img.shape[1] // f) * f
    x = img[:h, :w].astype(np.float64).reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3))
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    for fname, example in RENDERS.items():
        src = np.asarray(Image.open(os.path.join(REF, fname)).convert("RGB"))
        small = box_down(src, FACTOR)
        Image.fromarray(small).save(os.path.join(OUT, f"{example}.png"), optimize=True)
        print(f"{fname}: {src.shape[1]}x{src.shape[0]} -> {small.shape[1]}x{small.shape[0]}  ({example})")


if __name__ == "__main__":
    main()

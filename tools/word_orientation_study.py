"""Evidence for DESIGN.md section 11 item 1: how many mask-word flushes (REDs) one scan of each BASELINE config costs for the
three possible word orientations of the update grid (CPU model, exact DDA).   python tools/word_orientation_study.py"""
import os, struct, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdb_mapping_b200 import scans

out = os.path.join(ROOT, "tools", "build")
os.makedirs(out, exist_ok=True)
exe = os.path.join(out, "word_orientation_study")
subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tools", "word_orientation_study.cpp"), "-o", exe], check=True)
for cfg in (1, 2, 3):
    c = scans.CONFIGS[cfg]
    pts, origin = scans.make_scan(cfg, 3)
    p16 = np.ones((pts.shape[0], 4), dtype=np.float32); p16[:, :3] = pts[:, :3]
    path = os.path.join(out, f"scan_cfg{cfg}.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("7d", c.resolution, c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max, 1))
        f.write(struct.pack("3d", *origin)); f.write(struct.pack("I", p16.shape[0])); f.write(p16.tobytes())
    print(c.name, subprocess.run([exe, path], capture_output=True, text=True).stdout.strip())

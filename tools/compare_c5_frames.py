"""Parity of the full-size config-5 runs at different world sizes: the final frames written by
tools/scale_c5_full.py (gpurun_out/c5_frame_world{N}_n{frames}.npy) against the single-GPU frame.
    python tools/compare_c5_frames.py [frames] [c5|c4]   ->  lines for profiles/"""
import glob
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "4000"
tag = sys.argv[2] if len(sys.argv) > 2 else "c5"          # c5 | c4
files = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_frame_world*_n{n}.npy")))
frames = {int(re.search(r"world(\d+)_", f).group(1)): np.load(f).astype(np.float64) for f in files}
if 1 not in frames:
    sys.exit("no single-GPU frame (world 1) to compare with")
ref = frames[1]
for w in sorted(frames):
    if w == 1:
        continue
    m = ~np.isnan(ref)
    err = np.max(np.abs(frames[w][m] - ref[m])) / np.max(np.abs(ref[m]))
    print(f"world {w} vs world 1, {n} frames: max|diff| / max|frame| = {err:.2e}")

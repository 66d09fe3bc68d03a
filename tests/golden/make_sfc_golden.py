"""Generate tests/golden/sfc_coords.npz: the id -> (cz, cy, cx) tables of the reference's OWN ChunkMap
(chunkmap.cpp:176-191 over sfc.cpp:97-169), printed by host/_build/demo (which compiles the reference's
chunkmap.cpp / sfc.cpp where they lie under /root/reference).  Run here, where the reference exists:

    make -C host && python tests/golden/make_sfc_golden.py
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DIMS = [(2, 2, 2), (4, 4, 4), (8, 8, 8), (16, 16, 16), (16, 8, 8), (16, 16, 8), (4, 4, 2), (2, 4, 4), (4, 2, 4),
        (3, 5, 7), (7, 5, 3), (5, 7, 3), (6, 6, 6), (1, 4, 4), (4, 1, 4), (4, 4, 1), (1, 1, 8), (1, 8, 1), (8, 1, 1),
        (1, 1, 1), (2, 3, 1), (1, 5, 3), (9, 2, 1), (10, 12, 14), (2, 2, 16), (12, 4, 4), (1, 2, 3), (2, 3, 4)]

out = {}
for cd in DIMS:
    txt = subprocess.run([os.path.join(ROOT, "host", "_build", "demo"), "coord", *map(str, cd)],
                         capture_output=True, text=True, check=True).stdout
    out["c_%d_%d_%d" % cd] = np.array([[int(v) for v in ln.split()] for ln in txt.strip().splitlines()], dtype=np.int16)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sfc_coords.npz"), **out)
print("wrote", len(out), "tables")

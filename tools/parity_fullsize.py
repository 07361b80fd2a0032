#!/usr/bin/env python
"""Run the whole-path parity report (oracle/fullsize_check.py) for the BASELINE configs and print one JSON line
per workload.  GPU only.   python tools/parity_fullsize.py [cfg3 cfg2 posed cfg5 bg]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gaussian_splatting_3d_b200 import synthetic as S  # noqa: E402
from oracle import fullsize_check as F  # noqa: E402
from oracle import ref_gpu  # noqa: E402

ext = ref_gpu.load_reference_extension()
assert ext is not None, "oracle/_ref/_gs_ref*.so missing"
which = sys.argv[1:] or ["cfg3", "cfg2", "posed", "cfg5", "bg"]
for w in which:
    if w == "cfg3":
        r = F.compare_whole_path(ext, "cfg3")
    elif w == "cfg2":
        r = F.compare_whole_path(ext, "cfg2")
    elif w == "posed":
        r = F.compare_whole_path(ext, "cfg2", N=1_000_000, seed=1, c2w=S.ring_cameras(8)[1])
        r["workload"] = "cfg2-posed-1M"
    elif w == "cfg5":
        r = F.compare_whole_path(ext, "cfg5", backward=False)
    elif w == "bg":
        r = F.compare_whole_path(ext, "cfg3", N=20_000, seed=3, bg_rgb=(1.0, 0.5, 0.25))
        r["workload"] = "cfg3-20k-bg"
    else:
        raise SystemExit(f"unknown workload {w}")
    r["pass"] = F.passes(r)
    print(json.dumps(r), flush=True)

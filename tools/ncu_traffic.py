#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of ONE training step (tools/one_step.py): DRAM
bytes per stage (summed over the stage's kernels, per launch of the stage) and the utilisation figures of each
stage's dominant kernel, stamped with the hash of the library that was profiled.

    python tools/ncu_traffic.py gpurun_out/r2_step_metrics.csv profiles/ncu_traffic.json "<source note>"
(input: an .ncu-rep, or the CSV of `ncu --metrics ... --csv --page raw --log-file x.csv`, tools/gpu_r2_ncu_step.sh)
"""
import csv
import hashlib
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
STAGE = [  # first match wins
    ("project_cull_fused", "K1_project_cull"), ("frustum_kernel", "K1_project_cull"),
    ("radix_", "K2_binning"), ("scatter_kernel", "K2_binning"), ("count_rects", "K2_binning"),
    ("scan_block_sums", "K2_binning"), ("emit_", "K2_binning"), ("ranges_kernel", "K2_binning"),
    ("keys64", "K2_binning"), ("init_depth_keys", "K2_binning"), ("count_sorted", "K2_binning"),
    ("composite_fwd", "K3_composite_fwd"), ("composite_bwd", "K4a_composite_bwd"),
    ("project_backward", "K4b_project_bwd"), ("rows_zero_marked", "X_rows_zero"),
]
UTIL = {"issue_slot_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smem_wavefront_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"}


def source_hash():
    """Hash of the kernel sources the library is built from (csrc/*.cu, *.cuh, the C header): stable across
    rebuilds, unlike the .so (nvcc output is not byte-reproducible)."""
    h = hashlib.sha256()
    files = sorted((ROOT / "gaussian_splatting_3d_b200" / "csrc").glob("*.cu*")) + [ROOT / "include" / "gs3d_b200.h"]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    if rep.endswith(".csv"):  # `ncu --csv --page raw --log-file x.csv` written on the GPU box
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(txt)) if r and not r[0].startswith("==")]
    while rows and "Kernel Name" not in rows[0]:  # (the profiled program's own stdout may precede the table)
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6,
             "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

    def val(r, name):
        i = col.get(name)
        if i is None or i >= len(r):
            return None
        v = num(r[i])
        return None if v is None else v * scale.get(units[i], 1.0)

    stages = {}
    for r in rows[2:]:
        kn = r[col["Kernel Name"]]
        st = next((s for pat, s in STAGE if pat in kn), None)
        if st is None:
            continue
        d = stages.setdefault(st, {"dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "time_ms": 0.0, "launches": 0,
                                   "_top": (0.0, None)})
        rd, wr, t = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0, \
            val(r, "gpu__time_duration.sum") or 0.0
        d["dram_bytes_read"] += rd
        d["dram_bytes_write"] += wr
        d["time_ms"] += t
        d["launches"] += 1
        if t > d["_top"][0]:
            d["_top"] = (t, r)
    sha = source_hash()
    res = {"_meta": {"capture": rep, "note": note, "src_sha256_16": sha,
                     "method": "ncu --set full --clock-control none, one eager cfg-2 step (tools/one_step.py); "
                               "cold-cache serialised replays: bytes are per stage per step, times are NOT bench times"}}
    for st, d in stages.items():
        top = d.pop("_top")[1]
        d["traffic"] = d["dram_bytes_read"] + d["dram_bytes_write"]
        d["kernel"] = top[col["Kernel Name"]] if top else None
        d["source"] = f"{out} <- {rep} ({note})" if note else f"{out} <- {rep}"
        for k, m in UTIL.items():
            v = val(top, m) if top else None
            if v is not None:
                d[k] = v
        res[st] = d
    Path(out).write_text(json.dumps(res, indent=1))
    for st, d in res.items():
        if st != "_meta":
            print(f"{st:20s} {d['launches']:3d} launches  {d['time_ms']:.3f} ms (ncu)  DRAM {d['traffic'] / 1e6:8.1f} MB  {d['kernel'][:60]}")


if __name__ == "__main__":
    main()

"""Regenerates profiles/ncu_traffic.json from an `ncu --page raw --csv` export: DRAM bytes read + written of the dominant
SLOS launch (the longest launch whose kernel name matches), tagged with a hash of the kernel sources so that bench.py can
tell a capture of an older kernel from a current one (roofline.traffic is null when the hash does not match).

    python tools/ncu_traffic.py profiles/r2_slos_thin6_ncu_raw.csv [kernel-name-regex]
"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["perceval_b200/csrc/slos.cu", "perceval_b200/csrc/slos_thin.cu", "perceval_b200/csrc/slos_tile.cuh", "perceval_b200/csrc/slos_mu.cu"]


def source_hash() -> str:
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"slos_thin6_kernel|slos_tile_kernel")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    best = None
    for r in rows[2:]:
        if not pat.search(r[col["Kernel Name"]]):
            continue
        dur = num(r[col["gpu__time_duration.sum"]])
        if dur is not None and (best is None or dur > best[0]):
            best = (dur, r)
    assert best, "no matching launch"
    dur, r = best

    def bytes_of(name):
        v, u = num(r[col[name]]), units[col[name]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
        return v * scale

    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    out = {"slos_last_layer_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel": r[col["Kernel Name"]],
           "duration_under_ncu": f"{dur} {units[col['gpu__time_duration.sum']]}", "source_csv": os.path.relpath(path, ROOT),
           "source_sha": source_hash(),
           "how": "ncu --set full --clock-control none --import-source on, exported with --page raw --csv; this file is written by tools/ncu_traffic.py, never by hand"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

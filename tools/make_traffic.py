#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the step kernel: dram__bytes_read.sum + dram__bytes_write.sum per
launch (mean over the captured launches), stamped with the SHA-1 of the kernel sources so that bench.py can tell a stale number.

    tools/make_traffic.py gpurun_out/prof_r2base.ncu-rep b [--kernel k_fast_step]
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha1():
    h = hashlib.sha1()
    d = os.path.join(ROOT, "starfish_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith(".cuh"):  # the kernels (sf_gpu.cu is host code)
            h.update(open(os.path.join(d, fn), "rb").read())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("workload")
    ap.add_argument("--kernel", default="k_fast_step")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in rows[2:] if a.kernel in r[kn]]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tj = json.load(open(path))
    except Exception:
        tj = {}
    sha = csrc_sha1()
    if tj.get("csrc_sha1") != sha:
        tj = {"csrc_sha1": sha}
    tj[a.workload] = sum(vals) / len(vals)
    tj["source"] = os.path.basename(a.report) + ": " + a.kernel + f", {len(vals)} launches, dram__bytes_read.sum + dram__bytes_write.sum per launch"
    json.dump(tj, open(path, "w"), indent=1)
    print(tj)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-source-line view of an ncu capture: joins `ncu --page source --csv` (SASS rows with executed-instruction counts and
stall samples) with the line table of the cubin (`nvdisasm -g`), because the CSV export of the CUDA-source view carries no
metrics.

usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [--so starfish_b200/libstarfish_gpu.so] [--top 40]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def line_table(so, kernel):
    """[(offset, file, line, sass)] of the kernel's instructions, in address order."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
        txt = subprocess.run(["nvdisasm", "-g", "-c"] + cubins, check=True, capture_output=True, text=True).stdout
    out, cur, inside = [], ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur[0], cur[1], m.group(2).strip()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--so", default="starfish_b200/libstarfish_gpu.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sass", action="store_true", help="also list the hottest SASS instructions")
    ap.add_argument("--sym", default=None, help="substring of the mangled symbol in the cubin when KERNEL (a regex on the demangled name) is ambiguous there")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "-k", "regex:" + a.kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    sass = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            break
        sass.append(r)
    lt = line_table(a.so, a.sym or a.kernel)
    if len(lt) != len(sass):
        print("warning: %d SASS rows in the report, %d in the cubin (rebuilt since the capture?)" % (len(sass), len(lt)))
    n = min(len(lt), len(sass))
    agg = collections.defaultdict(lambda: collections.Counter())
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_inst = tot_samp = 0
    hot = []
    for k in range(n):
        r, (_, f, line, text) = sass[k], lt[k]
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["# Samples"]] or 0))
        c = agg[(f, line)]
        c["inst"] += inst
        c["samples"] += samp
        c["shared_wf"] += int(float(r[col["L1 Wavefronts Shared"]] or 0))
        c["local_sectors"] += int(float(r[col["L2 Theoretical Sectors Local"]] or 0)) if "L2 Theoretical Sectors Local" in col else 0
        for h in stall_cols:
            v = int(float(r[col[h]] or 0))
            if v:
                c[h] += v
        tot_inst += inst
        tot_samp += samp
        hot.append((samp, inst, f, line, text))
    print("total: %d warp instructions, %d samples" % (tot_inst, tot_samp))
    print("%-22s %6s %12s %6s %8s %10s %10s  top stalls" % ("file:line", "inst%", "inst", "samp%", "samples", "shared_wf", "local_sec"))
    for (f, line), c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[: a.top]:
        st = sorted(((v, h) for h, v in c.items() if h.startswith("stall_")), reverse=True)[:3]
        print("%-22s %6.2f %12d %6.2f %8d %10d %10d  %s" % ("%s:%d" % (f, line), 100.0 * c["inst"] / max(tot_inst, 1), c["inst"], 100.0 * c["samples"] / max(tot_samp, 1),
                                                              c["samples"], c["shared_wf"], c["local_sectors"], ", ".join("%s %d" % (h[6:], v) for v, h in st)))
    if a.sass:
        print("\nhottest SASS:")
        for samp, inst, f, line, text in sorted(hot, reverse=True)[: a.top]:
            print("%7d %10d  %s:%d  %s" % (samp, inst, f, line, text))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Instruction mix and stall hot spots from an `ncu --page source --csv` export (SASS view).
    python tools/ncu_sass_mix.py file_source.csv.gz [top_n]"""
import collections, csv, gzip, re, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    rows = list(csv.reader(f))
# possibly several kernels concatenated: split at "Kernel Name" rows
k = 0
while k < len(rows):
    assert rows[k][0] == "Kernel Name"
    name = rows[k][1]; head = rows[k + 1]; k += 2
    body = []
    while k < len(rows) and rows[k][0] != "Kernel Name":
        body.append(rows[k]); k += 1
    ix = {n: i for i, n in enumerate(head)}
    mix = collections.Counter(); samp = collections.Counter()
    tot_i = tot_s = 0
    for r in body:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_\.]+)", r[ix["Source"]])
        opc = m.group(2).split(".")[0] if m else "?"
        n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
        mix[opc] += n; samp[opc] += s; tot_i += n; tot_s += s
    print("== %s\n   %d SASS lines, %d warp instructions, %d samples" % (name[:100], len(body), tot_i, tot_s))
    print("   opcode mix (warp insts):  " + "  ".join("%s %.1f%%" % (o, 100.0 * n / tot_i) for o, n in mix.most_common(18)))
    print("   stall samples by opcode:  " + "  ".join("%s %.1f%%" % (o, 100.0 * n / max(tot_s, 1)) for o, n in samp.most_common(12)))
    hot = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:topn]
    stall_cols = [n for n in head if n.startswith("stall_") and "Not Issued" not in n]
    for r in hot:
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("   %6s samp %9s exec  %-70s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                                   " ".join("%s:%d" % (c, v) for v, c in st if v)))

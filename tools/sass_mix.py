#!/usr/bin/env python
"""Dynamic SASS opcode mix of one kernel from `ncu -i rep --page source --csv --print-source sass`.

    python tools/sass_mix.py src.csv <sites per launch>
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    sites = float(sys.argv[2])
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    cnt, smp, tot = collections.Counter(), collections.Counter(), 0
    first = True
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) < len(hdr) or not r[iE].isdigit():
            if "Kernel Name" in r[:1] and not first:
                break
            continue
        first = False
        s = r[iS].strip().split()
        op = (s[1] if s[0].startswith("@") else s[0]).rstrip(";")
        base = op.split(".")[0]
        if base == "IMAD":
            base = ".".join(op.split(".")[:2])
        e = int(r[iE])
        cnt[base] += e
        smp[base] += int(r[iSm])
        tot += e
    print(f"total warp instr {tot}  thread-instr per site {tot * 32 / sites:.2f}")
    for k, v in cnt.most_common(32):
        print(f"{k:14s} {v:10d} {v / tot * 100:5.1f}%  per-site {v * 32 / sites:5.2f}  samples {smp[k]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Static instruction count of the hot loop(s) of a kernel: for every backward branch, the
instructions between its target and itself minus the blocks a forward branch skips that
contain a CALL (the rare tie path).

    python tools/loop_count.py <object file> <mangled kernel name>
"""
import re
import subprocess
import sys


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:P\d, )?0x([0-9a-f]+)", t)
        if not m or "BRA.DIV" in t:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_idx:
            continue
        j = addr_idx[tgt]
        n = i - j + 1
        if n < 100:
            continue
        # forward branches inside whose skipped block holds a CALL
        rare = 0
        k = j
        while k < i:
            mm = re.search(r"BRA\s+0x([0-9a-f]+)", ins[k][1])
            if mm and "BRA.DIV" not in ins[k][1]:
                t2 = int(mm.group(1), 16)
                if t2 > ins[k][0] and t2 in addr_idx and addr_idx[t2] <= i:
                    blk = ins[k + 1:addr_idx[t2]]
                    if any("CALL" in x for _, x in blk):
                        rare += len(blk)
                        k = addr_idx[t2]
                        continue
            k += 1
        mix = {}
        for _, x in ins[j:i + 1]:
            op = x.split()[1] if x.startswith("@") else x.split()[0]
            op = op.split(".")[0]
            mix[op] = mix.get(op, 0) + 1
        top = sorted(mix.items(), key=lambda kv: -kv[1])[:12]
        print(f"loop 0x{tgt:x}..0x{a:x}: {n} instr, rare {rare}, executed ~{n - rare}  ({(n - rare) / 16:.2f} per site)  {top}")


if __name__ == "__main__":
    main()

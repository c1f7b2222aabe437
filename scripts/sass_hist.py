#!/usr/bin/env python
"""Instruction histogram of a kernel's hot loop from `cuobjdump -sass` (dev tool, no GPU needed).

    python scripts/sass_hist.py file.cubin|lib.so [kernel-name-regex] [--rows 4]

The hot tile of `march_step` is one long basic block (R latitudes unrolled): the longest basic blocks are printed with their
opcode counts, per block and (--rows R) per latitude."""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


FP_OPS = ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "DADD", "DMUL", "DFMA")


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else re.compile("march_kernel")
    rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 4
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 2
    for name, body in kernels(path):
        if not pat.search(name):
            continue
        print(f"== {name}: {len(body)} instructions")
        # basic blocks: split at branches / exits and at branch targets; the hot tile is the longest one
        targets = set()
        for addr, text in body:
            m2 = re.search(r"\b(?:BRA|BSSY|CALL|JMP)\b.*0x([0-9a-f]+)", text)
            if m2:
                targets.add(int(m2.group(1), 16))
        blocks, cur = [], []
        for addr, text in body:
            if addr in targets and cur:
                blocks.append(cur)
                cur = []
            cur.append((addr, text))
            if re.search(r"\b(BRA|EXIT|RET|BRX|JMP)\b", text):
                blocks.append(cur)
                cur = []
        if cur:
            blocks.append(cur)
        blocks.sort(key=len, reverse=True)
        for blk in blocks[:top]:
            hist = collections.Counter()
            for _, text in blk:
                t = re.sub(r"^@!?U?P\d+\s+", "", text)
                op = t.split()[0].split(".")[0]
                hist[op] += 1
            fp = sum(v for o, v in hist.items() if o in FP_OPS)
            n = len(blk)
            print(f"-- block {blk[0][0]:#x}..{blk[-1][0]:#x}: {n} instructions ({n / rows:.1f} per latitude at {rows} rows), FP {fp} ({fp / rows:.1f})")
            print("   " + ", ".join(f"{o} {v}" for o, v in hist.most_common(30)))


if __name__ == "__main__":
    main()

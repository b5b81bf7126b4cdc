"""Static SASS accounting for one kernel of libcolvo_b200.so (no GPU needed).

The three main kernels are issue-bound, so the number of SASS instructions on the hot
path is the quantity to drive down; this prints, for a kernel whose mangled name contains the given substrings,
the instruction count per opcode over (a) the whole function and (b) every loop (backward branch target ..
branch), so a change can be judged on the CPU container before GPU time is spent.

usage: python scripts/sass_count.py k_photo_bwdILi2ELb0ELb0 [--lib path] [--loops] [--dump out.sass]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_sass(lib, pat):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = out.split("\t\tFunction : ")
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if pat in name:
            return name, b
    raise SystemExit(f"no kernel matching {pat}")


INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")


def parse(body):
    ins = []
    for line in body.split("\n"):
        m = INS.match(line)
        if m:
            addr = int(m.group(1), 16)
            text = m.group(2).strip()
            ins.append((addr, text))
    return ins


def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split()[0].split(".")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern")
    ap.add_argument("--lib", default=os.path.join(ROOT, "coivo_b200", "libcolvo_b200.so"))
    ap.add_argument("--loops", action="store_true")
    ap.add_argument("--dump")
    ap.add_argument("--top", type=int, default=18)
    a = ap.parse_args()
    name, body = kernel_sass(a.lib, a.pattern)
    ins = parse(body)
    if a.dump:
        with open(a.dump, "w") as f:
            for addr, t in ins:
                f.write(f"{addr:06x}  {t}\n")
    print(name[:100])
    c = collections.Counter(opcode(t) for _, t in ins)
    print(f"static instructions: {len(ins)}")
    print("  " + ", ".join(f"{k} {v}" for k, v in c.most_common(a.top)))
    if a.loops:
        addr_idx = {ad: i for i, (ad, _) in enumerate(ins)}
        loops = []
        for i, (ad, t) in enumerate(ins):
            if opcode(t) == "BRA":
                m = re.search(r"0x([0-9a-f]+)", t)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt <= ad and tgt in addr_idx:
                        loops.append((addr_idx[tgt], i))
        for s, e in sorted(loops, key=lambda x: x[0] - x[1]):
            sub = ins[s:e + 1]
            cc = collections.Counter(opcode(t) for _, t in sub)
            print(f"loop {ins[s][0]:#x}..{ins[e][0]:#x}: {len(sub)} instr")
            print("    " + ", ".join(f"{k} {v}" for k, v in cc.most_common(a.top)))


if __name__ == "__main__":
    main()

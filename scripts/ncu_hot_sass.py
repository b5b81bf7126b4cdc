"""Top stalled SASS instructions and per-opcode sample/exec shares from an ncu report's source page."""
import collections
import csv
import subprocess
import sys


def main(rep, kernel_regex, topn=18):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel_regex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # one block per kernel instance: "Kernel Name" row, header row, data rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for blk in blocks[:1]:
        idx = {h: i for i, h in enumerate(blk["hdr"])}
        data = blk["data"]
        tot = sum(int(r[idx["# Samples"]]) for r in data)
        byop, execs = collections.Counter(), collections.Counter()
        for r in data:
            t = r[idx["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            byop[op] += int(r[idx["# Samples"]])
            execs[op] += int(r[idx["Instructions Executed"]])
        te = sum(execs.values())
        print(blk["name"][:80], "| samples", tot, "| sass lines", len(data), "| warp-inst", te)
        print("  samples by opcode:", ", ".join(f"{o} {c / tot * 100:.1f}%" for o, c in byop.most_common(14)))
        print("  exec by opcode:   ", ", ".join(f"{o} {c / te * 100:.1f}%" for o, c in execs.most_common(16)))
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:topn]:
            print(f"  {r[idx['# Samples']]:>6s} {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 18)

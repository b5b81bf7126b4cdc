"""profiles/traffic.json from an ncu capture of one step: per kernel and per step the ncu-measured DRAM bytes, the
executed warp instructions and the COUNTED fp32 flops (predicated-on thread instructions of every SASS line x flops of
its opcode), stamped with the hash of the kernel sources the capture was taken from (bench.py reports a record whose
hash differs from the current sources as stale).

  ncu --set full --import-source on --clock-control none -k regex:k_ -s <launches of one step> -c <the same> \
      -o gpurun_out/step python bench.py --config 2 --profile --steps 2 --warmup 1
  python scripts/make_traffic.py --config 2 gpurun_out/step.ncu-rep          # merges into profiles/traffic.json
"""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coivo_b200 import _lib  # noqa: E402

FLOPS = {"FFMA": 2, "FMUL": 1, "FADD": 1, "FFMA2": 4, "FMUL2": 2, "FADD2": 2, "MUFU": 1, "FMNMX": 1, "FSEL": 0,
         "DFMA": 2, "DADD": 1, "DMUL": 1}       # fp64 counted separately


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("colvo::", "").strip()
    return n.split("<")[0]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]

    def val(r, key, scale_unit=True):
        v = float(r[idx[key]].replace(",", ""))
        u = units[idx[key]]
        if scale_unit:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1)
        return v
    res = []
    for r in rows[2:]:
        res.append({"name": short(r[idx["Kernel Name"]]), "us": val(r, "gpu__time_duration.sum"),
                    "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                    "warp_inst": val(r, "smsp__inst_executed.sum", False)})
    return res


def source_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": short(r[1]), "fp32": 0.0, "fp64": 0.0, "ops": collections.Counter(), "rows": []}
            res.append(cur)
            hdr = None
        elif cur is not None and hdr is None:
            hdr = {h: i for i, h in enumerate(r)}
        elif cur is not None and len(r) == len(hdr):
            cur["rows"].append(r)
            t = r[hdr["Source"]].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            thr = float(r[hdr["Predicated-On Thread Instructions Executed"]])
            cur["ops"][op] += float(r[hdr["Instructions Executed"]])
            if op in FLOPS:
                cur["fp64" if op[0] == "D" else "fp32"] += thr * FLOPS[op]
    # ncu lists a launch once per function it contains (the kernel and e.g. a called division helper share one
    # listing): drop a block that repeats its predecessor row for row
    out = []
    for b in res:
        if out and out[-1]["name"] == b["name"] and out[-1]["rows"] == b["rows"]:
            continue
        out.append(b)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    a = ap.parse_args()
    raw, src = raw_page(a.report), source_page(a.report)
    assert len(raw) == len(src), (len(raw), len(src))
    kernels = collections.OrderedDict()
    for r, s in zip(raw, src):
        assert r["name"] == s["name"], (r["name"], s["name"])
        k = kernels.setdefault(r["name"], {"launches_per_step": 0, "us": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0, "fp32_flop": 0.0,
                                           "fp64_flop": 0.0, "packed_fp32x2_warp_inst": 0.0})
        k["launches_per_step"] += 1
        k["us"] += r["us"]
        k["dram_bytes"] += r["dram_read"] + r["dram_write"]
        k["warp_inst"] += r["warp_inst"]
        k["fp32_flop"] += s["fp32"]
        k["fp64_flop"] += s["fp64"]
        k["packed_fp32x2_warp_inst"] += s["ops"]["FFMA2"] + s["ops"]["FMUL2"] + s["ops"]["FADD2"]
    for k in kernels.values():
        k["dram_bytes_per_launch"] = k["dram_bytes"] / k["launches_per_step"]
    step = {key: sum(k[key] for k in kernels.values()) for key in ("us", "dram_bytes", "warp_inst", "fp32_flop", "fp64_flop")}
    rec = {}
    if os.path.exists(a.out):
        with open(a.out) as f:
            rec = json.load(f)
        if "configs" not in rec:
            rec = {}
    rec.setdefault("configs", {})[str(a.config)] = {
        "src_sha16": _lib.source_hash(), "report": os.path.basename(a.report),
        "how": "ncu --set full --import-source on --clock-control none, one step (serialised, cold caches); scripts/make_traffic.py",
        "kernels": kernels, "step": step}
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps({"config": a.config, "step": step, "kernels": {n: {"us": round(k["us"], 1), "dram_MB": round(k["dram_bytes"] / 1e6, 1),
                                                                      "Minst": round(k["warp_inst"] / 1e6, 2), "GF": round(k["fp32_flop"] / 1e9, 3)}
                                                                  for n, k in kernels.items()}}, indent=1))


if __name__ == "__main__":
    main()

"""Condense `ncu --set full` reports into the JSON / markdown that profiles/ keeps (runs anywhere `ncu -i` works, no GPU needed).

    python tools/ncu_summary.py --out profiles/r02_ncu_summary.json --layer-json profiles/r02_ncu_layer.json \
        gpurun_out/r02_frame_full.ncu-rep [more.ncu-rep ...]

Per kernel name (template arguments kept, anonymous namespace stripped): launches, median duration, median DRAM read / write
bytes, registers, grid x block, SM / DRAM throughput %, issue-active %, tensor-pipe %.  --layer-json writes the per-SRA-layer
DRAM traffic (chain kernel + attention kernel, median launch each) that bench.py reports as roofline.traffic."""
import argparse
import csv
import io
import json
import re
import statistics
import subprocess

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0,
        "msecond": 1e3, "second": 1e6}
COLS = {
    "us": "gpu__time_duration.sum",
    "dram_read_B": "dram__bytes_read.sum",
    "dram_write_B": "dram__bytes_write.sum",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "inst": "smsp__inst_executed.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
}


def short(name):
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    return re.sub(r"^void ", "", re.sub(r"\(.*\)$", "", name))


def rows(rep):
    if rep.endswith(".csv"):   # already exported with `ncu -i X.ncu-rep --page raw --csv`
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    while r and "Kernel Name" not in r[0]:   # ==PROF== banner lines
        r.pop(0)
    head, unit = r[0], r[1]
    for line in r[2:]:
        d = {"name": short(line[head.index("Kernel Name")])}
        for k, col in COLS.items():
            if col not in head:
                continue
            i = head.index(col)
            try:
                v = float(line[i].replace(",", ""))
            except ValueError:
                continue
            d[k] = v * UNIT.get(unit[i], 1.0) if k in ("us", "dram_read_B", "dram_write_B") else v
        yield d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--layer-json")
    ap.add_argument("--md")
    a = ap.parse_args()
    per = {}
    for rep in a.reports:
        for d in rows(rep):
            per.setdefault(d["name"], []).append(dict(d, report=rep.split("/")[-1]))
    out = {}
    for name, ls in sorted(per.items(), key=lambda kv: -sum(x.get("us", 0) for x in kv[1])):
        e = {"launches": len(ls), "reports": sorted({x["report"] for x in ls})}
        for k in COLS:
            vs = [x[k] for x in ls if k in x]
            if vs:
                e[k] = round(statistics.median(vs), 3)
        e["us_total"] = round(sum(x.get("us", 0) for x in ls), 2)
        if e.get("us"):
            e["dram_GBps"] = round((e.get("dram_read_B", 0) + e.get("dram_write_B", 0)) / e["us"] * 1e-3, 1)
        out[name] = e
    json.dump({"note": "ncu --set full --clock-control none; per-launch medians; durations are serialised + cold-cache, use shares not absolutes",
               "kernels": out}, open(a.out, "w"), indent=1)
    if a.layer_json:
        chain = next((v for k, v in out.items() if k.startswith("sra_chain2_kernel<true>") or k.startswith("sra_chain2_kernel<(bool)1>")), None)
        if chain is None:
            chain = next((v for k, v in out.items() if "sra_chain2_kernel" in k), None)
        attn = next((v for k, v in out.items() if "win_attn_batch_kernel" in k), None)
        if chain and attn:
            tr = chain["dram_read_B"] + chain["dram_write_B"] + attn["dram_read_B"] + attn["dram_write_B"]
            json.dump({"dram_bytes_per_layer": tr, "chain": chain, "attention": attn,
                       "source": [r.split("/")[-1] for r in a.reports], "how": "median launch of the chain kernel + median launch of the attention kernel"},
                      open(a.layer_json, "w"), indent=1)
    if a.md:
        with open(a.md, "w") as f:
            f.write("| kernel | launches | µs (median) | DRAM rd MB | DRAM wr MB | DRAM GB/s | regs | grid×block | SM % | DRAM % | issue % | tensor % |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for k, v in out.items():
                f.write("| `%s` | %d | %.2f | %.3f | %.3f | %.0f | %d | %d×%d | %.1f | %.1f | %.1f | %.1f |\n" % (
                    k[:90], v["launches"], v.get("us", 0), v.get("dram_read_B", 0) / 1e6, v.get("dram_write_B", 0) / 1e6, v.get("dram_GBps", 0), v.get("regs", 0),
                    v.get("grid", 0), v.get("block", 0), v.get("sm_pct", 0), v.get("dram_pct", 0), v.get("issue_active_pct", 0), v.get("tensor_pipe_pct", 0)))
    print("kernels:", len(out))


if __name__ == "__main__":
    main()

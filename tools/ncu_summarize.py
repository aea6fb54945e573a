"""Summarise an .ncu-rep (read here, on the CPU box): launch metrics of the captured kernel + where its instructions go.

  python tools/ncu_summarize.py gpurun_out/x.ncu-rep profiles/x.json [--samples N] [--note "..."]

Writes: the raw metrics the roofline needs (duration, instructions, issue-active, DRAM bytes, registers, occupancy limits,
stall ratios), the per-source-line top list and the SASS execution-count plateaus (contiguous instruction ranges executed the
same number of times = loop bodies / branches) with their share of all issued warp-instructions and their lane utilisation."""
import csv
import io
import json
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    samples = None
    note = ""
    if "--samples" in sys.argv:
        samples = float(sys.argv[sys.argv.index("--samples") + 1])
    if "--note" in sys.argv:
        note = sys.argv[sys.argv.index("--note") + 1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    raw = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
    want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    metrics = {}
    for k in want:
        if k in raw:
            v, u = raw[k]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            metrics[k] = {"value": v, "unit": u}

    def num(k):
        return metrics[k]["value"] if k in metrics and isinstance(metrics[k]["value"], float) else None

    def to_bytes(k):
        if k not in metrics:
            return None
        v, u = metrics[k]["value"], metrics[k]["unit"].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv"))))
    h = src[1]
    iS, iI, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    ins = [(r[iS].strip(), int(r[iI]), int(r[iT])) for r in src[2:] if len(r) > iT and r[iI].isdigit()]
    tot = sum(i[1] for i in ins) or 1
    seg = []
    for k, (s, c, t) in enumerate(ins):
        if seg and abs(c - seg[-1]["exec_per_inst"]) <= 0.02 * max(c, seg[-1]["exec_per_inst"]) + 5:
            seg[-1]["last"] = k; seg[-1]["warp_inst"] += c; seg[-1]["thread_inst"] += t
        else:
            seg.append({"first": k, "last": k, "exec_per_inst": c, "warp_inst": c, "thread_inst": t, "starts_with": s[:60]})
    plateaus = [{"sass_range": [s["first"], s["last"]], "n_sass": s["last"] - s["first"] + 1, "executions": s["exec_per_inst"],
                 "share_pct": round(100.0 * s["warp_inst"] / tot, 2), "lanes_active": round(s["thread_inst"] / max(1, s["warp_inst"]), 1),
                 "starts_with": s["starts_with"]} for s in seg if s["warp_inst"] > 0.01 * tot]
    cs = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    lines, cur = [], None
    if len(cs) > 3:
        hh = cs[2]
        jI, jT = hh.index("Instructions Executed"), hh.index("Thread Instructions Executed")
        for r in cs:
            if len(r) >= 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
                continue
            if len(r) > jT and r[0] not in ("", "Line No") and r[2] == "-":
                try:
                    lines.append((cur, int(r[0]), r[1].strip()[:100], int(r[jI]), int(r[jT])))
                except ValueError:
                    pass
    ltot = sum(x[3] for x in lines) or 1
    lines.sort(key=lambda x: -x[3])
    top_lines = [{"file": f, "line": ln, "share_pct": round(100.0 * wi / ltot, 2), "lanes_active": round(ti / max(1, wi), 1), "source": s}
                 for f, ln, s, wi, ti in lines[:40]]
    inst = num("smsp__inst_executed.sum")
    summary = {
        "report": rep, "note": note, "kernel": metrics.get("Kernel Name", {}).get("value"),
        "duration_ms": None if num("gpu__time_duration.sum") is None else num("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(metrics["gpu__time_duration.sum"]["unit"], 1),
        "inst_executed_per_launch": inst, "thread_inst_per_launch": num("smsp__thread_inst_executed.sum") or (sum(i[2] for i in ins) or None),
        "smsp_issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "registers_per_thread": num("launch__registers_per_thread"),
        "dram_bytes_per_launch": None if to_bytes("dram__bytes_read.sum") is None else to_bytes("dram__bytes_read.sum") + (to_bytes("dram__bytes_write.sum") or 0),
        "samples_per_launch": samples,
        "thread_inst_per_sample": None if not samples else (sum(i[2] for i in ins) / samples),
        "warp_inst_per_sample": None if not samples or not inst else inst / samples,
        "metrics": metrics, "sass_plateaus": plateaus, "top_source_lines": top_lines,
    }
    json.dump(summary, open(out, "w"), indent=1)
    print(out, "kernel", summary["kernel"], "duration_ms", summary["duration_ms"], "warp-inst", inst, "issue %", summary["smsp_issue_active_pct"],
          "lanes/inst", summary["threads_per_inst"], "thread-inst/sample", summary["thread_inst_per_sample"])


if __name__ == "__main__":
    main()

"""Summarise `ncu --set full` reports into the small CSVs kept under profiles/ (and profiles/traffic.json).

    python scripts/ncu_summary.py OUT.csv REPORT.ncu-rep [REPORT2.ncu-rep ...] [--traffic WORKLOAD:KERNELKEY ...]

One row per profiled kernel launch: duration, DRAM bytes, L2 / shared-memory / tensor-pipe utilisation, issue rate,
registers, shared memory.  `--traffic cfg4_b64:rvq_search` records the first matching row's DRAM bytes in
profiles/traffic.json under that workload / kernel key (bench.py's roofline.traffic reads it).
Runs in the dev container: `ncu -i` only reads the report.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar2sm_read_bytes"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_pipe_pct"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc_smem_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct_elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct_active"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_dim_x", "cluster"),
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def rows_of(report):
    if report.endswith(".csv"):           # a raw page exported on the GPU box (scripts/ncu_round.sh)
        out = open(report).read()
    else:
        out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = [i for i, r in enumerate(rd) if r and r[0] == "ID"]
    if not hdr:
        return []
    h = rd[hdr[0]]
    units = rd[hdr[0] + 1]
    res = []
    for r in rd[hdr[0] + 2:]:
        if len(r) != len(h):
            continue
        d = {}
        for k, u, v in zip(h, units, r):
            d[k] = (v, u)
        res.append(d)
    return res


def num(cell):
    v, u = cell
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    return x * UNIT_SCALE.get(u, 1.0)


def main():
    args = sys.argv[1:]
    traffic = []
    while "--traffic" in args:
        i = args.index("--traffic")
        traffic.append(args[i + 1])
        del args[i:i + 2]
    out_csv, reports = args[0], args[1:]
    lines = []
    for rep in reports:
        for d in rows_of(rep):
            name = d.get("Kernel Name", ("?", ""))[0]
            row = {"report": os.path.basename(rep), "kernel": name[:90]}
            for k, short in KEYS:
                if k in d:
                    row[short] = num(d[k])
            lines.append(row)
    cols = ["report", "kernel"] + [s for _, s in KEYS]
    with open(out_csv, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=cols)
        w.writeheader()
        for r in lines:
            w.writerow({c: r.get(c, "") for c in cols})
    for r in lines:
        print({k: (f"{v:.4g}" if isinstance(v, float) else v) for k, v in r.items()})
    if traffic:
        tpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
        for spec in traffic:
            wl, key, *pat = spec.split(":")
            pat = pat[0] if pat else key
            for r in lines:
                if pat in r["kernel"] and "dram_read" in r:
                    tj.setdefault(wl, {})[key] = {"dram_read": int(r["dram_read"]), "dram_write": int(r["dram_write"]),
                                                  "duration_us_under_ncu": round(r.get("duration", 0) * 1e6, 1),
                                                  "source": f"profiles/{os.path.basename(out_csv)} ({r['kernel'][:60]})"}
                    break
        json.dump(tj, open(tpath, "w"), indent=2)


if __name__ == "__main__":
    main()

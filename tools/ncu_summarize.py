#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the markdown summary kept under profiles/ and, optionally, the measured
DRAM traffic per audio-second for bench.py's roofline.traffic (profiles/traffic.json).

  python tools/ncu_summarize.py gpurun_out/x.ncu-rep "title" [--audio-s 191.98 --kinds front_fused,seanet_conv3,...]
"""
import argparse, csv, io, json, os, subprocess, sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("title")
    ap.add_argument("--audio-s", type=float, default=None)
    ap.add_argument("--kinds", default="")
    ap.add_argument("--traffic-out", default=None)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kinds = [k for k in a.kinds.split(",") if k]
    traffic = {}
    print(f"# {a.title}\n\nSource report: {a.rep} (scratch).\n")
    for n, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"## launch {n}: `{name}`" + (f" ({kinds[n]})" if n < len(kinds) else "") + "\n")
        for m in METRICS:
            if m in idx:
                print(f"- {m}: {r[idx[m]]} {units[idx[m]]}")
        if a.audio_s and n < len(kinds):
            b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            traffic[kinds[n]] = {"dram_bytes_per_audio_s": b / a.audio_s, "source": f"{os.path.basename(a.rep)} launch {n} ({a.audio_s:.2f} audio-s)"}
            print(f"- DRAM bytes per audio-second: {b / a.audio_s:,.0f}")
        print()
    if a.traffic_out and traffic:
        old = json.load(open(a.traffic_out)) if os.path.exists(a.traffic_out) else {}
        old.update(traffic)
        json.dump(old, open(a.traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main()

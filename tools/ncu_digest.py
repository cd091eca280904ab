#!/usr/bin/env python
"""Digest .ncu-rep captures (ncu --set full) into the per-kernel summary tables kept under profiles/ (no GPU needed).

usage: python tools/ncu_digest.py <rep> [label1,label2,...]      -> markdown table on stdout + one JSON line per launch
Labels (optional) name the launches in capture order and select the algorithmic work for the roofline columns:
  gemm:M:N:K   attn:B:S:NH:HD (causal prefill)   bytes:<algorithmic bytes>   (anything else: label only)
"""
import csv
import json
import subprocess
import sys

METRICS = {
    "dur_us": ("gpu__time_duration.sum", 1e-3),
    "dram_rd": ("dram__bytes_read.sum", 1),
    "dram_wr": ("dram__bytes_write.sum", 1),
    "dram_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "tensor_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    "xu_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", 1),
    "l2_hit": ("lts__t_sector_hit_rate.pct", 1),
    "regs": ("launch__registers_per_thread", 1),
    "smem_dyn": ("launch__shared_mem_per_block_dynamic", 1),
    "grid": ("launch__grid_size", 1),
    "block": ("launch__block_size", 1),
    "cluster": ("launch__cluster_size", 1),
    "sm_mhz": ("sm__cycles_elapsed.avg.per_second", 1e-6),
    "warps_active": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "smem_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
              "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9, "hz": 1, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9,
              "Kbyte/block": 1e3, "byte/block": 1, "Mbyte/block": 1e6}


def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {"kernel": r[ix["Kernel Name"]]}
        for k, (m, sc) in METRICS.items():
            if m in ix and r[ix[m]] not in ("", "n/a"):
                v = float(r[ix[m]].replace(",", ""))
                u = units[ix[m]]
                d[k] = v * UNIT_SCALE.get(u, 1) * sc
        out.append(d)
    return out


def work(label):
    p = label.split(":")
    if p[0] == "gemm":
        M, N, K = map(float, p[1:4])
        return {"flops": 2 * M * N * K, "bytes": 2 * (M * K + N * K) + (4 if "f32" in label else 2) * M * N}
    if p[0] == "attn":
        B, S, NH, HD = map(float, p[1:5])
        return {"flops": 4 * B * NH * HD * S * (S + 1) / 2}
    if p[0] == "bytes":
        return {"bytes": float(p[1])}
    return {}


def main():
    rep = sys.argv[1]
    labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
    ks = load(rep)
    print("| # | kernel | label | µs | grid×block | regs | DRAM rd+wr MB | DRAM % | tensor % | XU % | L2 hit % | achieved |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for i, d in enumerate(ks):
        lab = labels[i] if i < len(labels) else ""
        w = work(lab)
        dur = d.get("dur_us", 0.0) * 1e-6
        ach = ""
        if "flops" in w and dur:
            d["tflops"] = w["flops"] / dur / 1e12
            ach = f"{d['tflops']:.0f} TFLOP/s"
        traffic = d.get("dram_rd", 0) + d.get("dram_wr", 0)
        d["dram_bytes"] = traffic
        if dur:
            d["dram_gbs"] = traffic / dur / 1e9
        if "bytes" in w and dur:
            d["alg_bytes"] = w["bytes"]
            d["alg_gbs"] = w["bytes"] / dur / 1e9
            ach += f" {d['alg_gbs']:.0f} GB/s alg ({traffic / w['bytes']:.2f}x traffic)"
        name = d["kernel"].replace("void ", "").replace("b2::", "")
        name = name[:name.index("(")] if "(" in name else name
        d["label"] = lab
        print(f"| {i} | `{name}` | {lab} | {d.get('dur_us', 0):.1f} | {int(d.get('grid', 0))}×{int(d.get('block', 0))} | {int(d.get('regs', 0))} | "
              f"{traffic / 1e6:.1f} | {d.get('dram_pct', 0):.1f} | {d.get('tensor_pct', 0):.1f} | {d.get('xu_pct', 0):.1f} | {d.get('l2_hit', 0):.1f} | {ach} |")
    print()
    for d in ks:
        print("JSON " + json.dumps(d))


if __name__ == "__main__":
    main()

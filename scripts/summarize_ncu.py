"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list:
per kernel the number of launches, the summed duration, its share, and DRAM bytes per launch.

    python scripts/summarize_ncu.py gpurun_out/launches.csv [--traffic-json profiles/r1_traffic.json] [--compact out.csv]
"""
import argparse
import collections
import csv
import json

CLASS_OF = (("syrk_nhwc_kernel<1>", "syrk_nhwc_bf16"), ("syrk_nhwc_kernel<0>", "syrk_nhwc_tf32"), ("syrk_tc_kernel", "syrk_staged_nchw"),
            ("syrk_tc_tma_kernel", "syrk_staged_nchw"), ("syrk_tc_reduce_kernel", "syrk_split_reduce"), ("syrk_sk_reduce_kernel", "syrk_split_reduce"),
            ("syrk_sk_reduce_taps_kernel", "syrk_split_reduce"), ("cast_bf16_kernel", "cast_prepass"),
            ("round_tf32_kernel", "cast_prepass"), ("split_bf16_kernel", "cast_prepass"), ("nchw_to_nhwc_bf16_kernel", "cast_prepass"),
            ("pack_smallc_kernel", "cast_prepass"), ("syrk_simt_kernel", "syrk_simt_fp32"), ("gemm_chain_kernel", "gemm_chain"),
            ("gemm_tc_kernel", "gemm_tc"))


def kclass(name):
    for pat, c in CLASS_OF:
        if pat in name:
            return c
    return name.split("(")[0][-40:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic-json")
    ap.add_argument("--compact")
    ap.add_argument("--skip", type=int, default=0, help="ignore the first N launches (warm-up)")
    args = ap.parse_args()
    lines = [l for l in open(args.csv) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per_id = collections.OrderedDict()
    for r in rows:
        d = per_id.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        elif m.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d[m] = v * mult
    launches = list(per_id.values())[args.skip:]
    agg = collections.OrderedDict()
    for d in launches:
        a = agg.setdefault(kclass(d["name"]), {"launches": 0, "us": 0.0, "dram": 0.0})
        a["launches"] += 1
        a["us"] += d.get("us", 0.0)
        a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values()) or 1.0
    print(f"{'kernel class':24s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s} {'DRAM MB/launch':>15s}")
    out = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k:24s} {a['launches']:8d} {a['us'] / 1e3:10.3f} {a['us'] / total:7.3f} {a['us'] / a['launches']:9.1f} "
              f"{a['dram'] / a['launches'] / 1e6:15.1f}")
        out[k] = {"launches": a["launches"], "total_ms": a["us"] / 1e3, "share": a["us"] / total,
                  "dram_bytes_per_launch": a["dram"] / a["launches"] if a["dram"] else None}
    if args.traffic_json:
        json.dump(out, open(args.traffic_json, "w"), indent=1)
    if args.compact:
        with open(args.compact, "w") as f:
            f.write("id,kernel,grid,block,duration_us,dram_read_bytes,dram_write_bytes\n")
            for i, d in enumerate(launches):
                f.write(f"{i},{kclass(d['name'])},\"{d['grid']}\",\"{d['block']}\",{d.get('us', 0):.3f},"
                        f"{d.get('dram__bytes_read.sum', 0):.0f},{d.get('dram__bytes_write.sum', 0):.0f}\n")


if __name__ == "__main__":
    main()

"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv``) of
bench.py: share of the step per kernel family and the launch-weighted DRAM bytes of the convolution kernel, which bench.py reports
as ``roofline.traffic`` (profiles/conv_traffic.json).

usage: python tools/launch_summary.py profiles/<list>.csv.gz [--write-traffic profiles/conv_traffic.json]"""
import collections
import csv
import gzip
import io
import json
import re
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = [ln for ln in op(path, "rt") if ln.startswith('"')]
    launches = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(rows))):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(launches.values())


def family(name):
    m = re.search(r"(conv_f16x3\w*kernel<\d+>)", name)
    if m:
        return m.group(1)
    m = re.search(r"rpe::(\w+)", name)
    if m:
        return m.group(1)
    return "torch: " + re.sub(r"<.*", "", name.replace("void ", ""))[:50]


def main():
    path = sys.argv[1]
    L = load(path)
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in L:
        f = fam[family(d["name"])]
        f[0] += 1
        f[1] += d.get("gpu__time_duration.sum", 0.0)
        f[2] += d.get("dram__bytes_read.sum", 0.0)
        f[3] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(f[1] for f in fam.values())
    print(f"# {path}: {len(L)} launches, {total / 1e6:.1f} ms of kernel time under ncu (serialised, cold caches: compare shares)")
    print(f"{'kernel':52s} {'launches':>8s} {'share':>7s} {'avg us':>9s} {'rd MB':>9s} {'wr MB':>9s}   (per launch)")
    for name, (n, t, rd, wr) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        if t / total < 0.002:
            continue
        print(f"{name:52s} {n:8d} {100 * t / total:6.1f}% {t / n / 1e3:9.1f} {rd / n / 1e6:9.1f} {wr / n / 1e6:9.1f}")
    conv = [d for d in L if "conv_f16x3" in d["name"] and "kernel<7>" not in d["name"]]
    traffic = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in conv) / max(len(conv), 1)
    print(f"convolution kernel (all kinds but the correlation volume): {len(conv)} launches, {traffic / 1e6:.1f} MB of DRAM traffic per launch, "
          f"{sum(d['gpu__time_duration.sum'] for d in conv) / max(len(conv), 1) / 1e3:.1f} us per launch under ncu")
    if "--write-traffic" in sys.argv:
        out = sys.argv[sys.argv.index("--write-traffic") + 1]
        with open(out, "w") as f:
            json.dump({"config": {"chunk": 32, "precision": "fp16x3", "pairs": 64}, "dram_bytes_per_launch": traffic, "launches": len(conv),
                       "avg_launch_us_under_ncu": sum(d["gpu__time_duration.sum"] for d in conv) / max(len(conv), 1) / 1e3,
                       "source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over `bench.py --steps 2 "
                                 "--warmup 1` (every conv_f16x3_* launch except the correlation-volume kind <7>), launch-weighted mean of "
                                 "dram__bytes_read.sum + dram__bytes_write.sum"}, f, indent=1)
        print("wrote", out)


if __name__ == "__main__":
    main()

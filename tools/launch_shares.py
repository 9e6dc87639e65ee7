"""Per-kernel share of a bench step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/launch_shares.py gpurun_out/TAG_launches.csv > profiles/rNN_x_launch_shares.txt
The times are cold-cache and serialised (ncu replays every launch alone): it is the SHARES that are compared with the live
CUDA-event figures of bench.py (`roofline.step_shares`), not the absolute durations."""
import collections
import csv
import sys


def main(path):
    tot, cnt = collections.Counter(), collections.Counter()
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[k] == "Kernel Name":
            continue
        try:
            t = float(r[v].replace(",", ""))
        except ValueError:
            continue
        t *= {"ns": 1e-3, "us": 1., "ms": 1e3}.get(r[u], 1e-3)
        name = r[k].split("(")[0]
        tot[name] += t
        cnt[name] += 1
    whole = sum(tot.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'mean us':>9s}")
    for name, t in tot.most_common(12):
        print(f"{name[:60]:60s} {cnt[name]:8d} {t:12.1f} {100 * t / whole:6.1f}% {t / cnt[name]:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])

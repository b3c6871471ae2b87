"""Per-kernel totals of an ncu launch list (CPU only):
    ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <command>
    python tools/summarize_launches.py launches.csv "<command>" > summary.txt
Launches are serialised and cache-cold under ncu: compare SHARES of the step, not absolute times."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
        name = r[ik].split("(")[0][:48]
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    if note:
        print(note)
    print("%-48s %9s %12s %7s" % ("kernel", "launches", "total us", "share"))
    for name, t in tot.most_common(40):
        print("%-48s %9d %12.1f %6.1f%%" % (name, cnt[name], t, 100.0 * t / total))
    print("%-48s %9d %12.1f" % ("all kernels", sum(cnt.values()), total))


if __name__ == "__main__":
    main()

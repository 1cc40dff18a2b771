"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total and share."""
import collections
import csv
import sys


def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)   # -> microseconds
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    lines_out = [f"# {path}: {len(rows)} launches, {T / 1e3:.2f} ms total device time (cold-cache, serialised: compare shares)",
                 f"{'share':>7} {'total_ms':>10} {'launches':>9} {'avg_us':>9}  kernel"]
    for k, v in tot.most_common():
        lines_out.append(f"{100 * v / T:6.2f}% {v / 1e3:10.3f} {cnt[k]:9d} {v / cnt[k]:9.1f}  {k}")
    text = "\n".join(lines_out) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

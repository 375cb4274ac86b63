"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and mean
device time and share of the listed launches.  Usage: python tools/summarize_launches.py in.csv [skip_prefix_launches]
Writes markdown to stdout (committed under profiles/)."""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\((const |float|int|long|unsigned|CUtensorMap|T\d|mhe_|__nv|tc::GemmShape).*$', '', name)
    return name[:120]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((int(r['ID']), r['Kernel Name'], r['Grid Size'], r['Block Size'], float(r['Metric Value'].replace(',', '')) / 1e3))
    rows = [r for r in rows if r[0] >= skip]
    agg = collections.OrderedDict()
    for _, name, grid, block, us in rows:
        k = short(name)
        a = agg.setdefault(k, [0, 0.0, set()])
        a[0] += 1
        a[1] += us
        a[2].add(f'{grid}x{block}')
    total = sum(a[1] for a in agg.values())
    print(f'launches listed: {len(rows)}, summed device time {total:.1f} us (cold-cache, serialised under ncu: compare shares)\n')
    print('| kernel | launches | total us | mean us | share | grid x block |')
    print('|---|---:|---:|---:|---:|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        shapes = sorted(a[2])
        print(f'| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.2f} | {100 * a[1] / total:.1f}% | {"; ".join(shapes[:3])}{" ..." if len(shapes) > 3 else ""} |')


if __name__ == '__main__':
    main()

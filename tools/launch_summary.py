"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py:
cuts the list into train steps at the `sgd_momentum` launch that ends each one and
prints the per-kernel table of one step (default: the last graph-replayed step, i.e. the
one before the eager roofline pass).
Usage: python tools/launch_summary.py profiles/r1_launches_v18_step.csv [--step -2]"""
import argparse
import collections
import csv
import re


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hdr]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    return [(r[ki], float(r[vi].replace(',', '')) / 1e3) for r in rows[hdr + 1:] if len(r) > vi]


def short(name):
    m = re.search(r'cmr::(?:<unnamed>::)?(\w+)(<[^(]*>)?\(', name)
    if m:
        return m.group(1) + (m.group(2) or '').replace('cmr::<unnamed>::', '')
    return re.sub(r'^void ', '', re.sub(r'[<(].*', '', name))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--step', type=int, default=-2)
    args = ap.parse_args()
    launches = load(args.csv)
    steps, cur = [], []
    for name, us in launches:
        cur.append((name, us))
        if 'sgd_momentum' in name:
            steps.append(cur)
            cur = []
    print('%d launches, %d steps (%s launches each), %d trailing' % (
        len(launches), len(steps), ','.join(str(len(s)) for s in steps), len(cur)))
    step = steps[args.step]
    agg, cnt = collections.Counter(), collections.Counter()
    for name, us in step:
        agg[short(name)] += us
        cnt[short(name)] += 1
    tot = sum(agg.values())
    print('step %d: %d launches, %.2f ms of kernel time' % (args.step, len(step), tot / 1e3))
    for n, v in agg.most_common():
        print('| `%s` | %d | %.0f | %.1f %% |' % (n, cnt[n], v, 100 * v / tot))


if __name__ == '__main__':
    main()

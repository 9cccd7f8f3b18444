"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step
(the launches between two consecutive update kernels) and per-kernel totals."""
import collections
import csv
import sys


def main(path, step_index=-2):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
    data = [(int(r[ii]), r[ki], float(r[vi].replace(',', '')) / 1000.0) for r in rows[1:]]
    ends = [i for i, d in enumerate(data) if 'dp_update_kernel' in d[1]]       # fused exchange+update ends a step
    if not ends:
        ends = [i for i, d in enumerate(data) if 'clip_sgd_kernel' in d[1]]
        ends = [e for j, e in enumerate(ends) if j % 2 == 1]      # second clip_sgd of each step
    lo, hi = ends[step_index - 1] + 1, ends[step_index] + 1
    step = data[lo:hi]
    total = sum(d[2] for d in step)
    print(f"# one step: launches {step[0][0]}..{step[-1][0]}  ({len(step)} launches, {total:.1f} us summed, cold-cache serialised)")
    for d in step:
        print(f"{d[0]:5d} {d[2]:8.2f} us  {100 * d[2] / total:5.1f}%  {d[1][:90]}")
    agg = collections.defaultdict(float)
    for d in step:
        agg[d[1].split('(')[0]] += d[2]
    print("# per kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"{v:8.2f} us {100 * v / total:5.1f}%  {k}")


if __name__ == '__main__':
    main(sys.argv[1])

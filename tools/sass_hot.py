#!/usr/bin/env python
"""Summarise an `ncu --page source --print-source=sass --csv` export: executed instructions and stall samples by code
segment and the hottest instructions.   python tools/sass_hot.py file.csv [segment_size] [top_n]"""
import csv
import sys


def main():
    path = sys.argv[1]
    seg = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rows = list(csv.reader(open(path)))
    h = rows[1]
    body = [r for r in rows[2:] if len(r) == len(h)]
    ci = {c: i for i, c in enumerate(h)}
    ex = [int(r[ci['Instructions Executed']]) for r in body]
    sm = [int(r[ci['# Samples']]) for r in body]
    print("kernel:", rows[0][1][:90])
    print("instructions %d, executed %.1f M, samples %d" % (len(body), sum(ex) / 1e6, sum(sm)))
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    for s in range(0, len(body), seg):
        e, m = sum(ex[s:s + seg]), sum(sm[s:s + seg])
        if m < 0.002 * sum(sm) and e < 0.002 * sum(ex):
            continue
        tot = {c: sum(int(r[ci[c]]) for r in body[s:s + seg]) for c in stall_cols}
        top = sorted(tot.items(), key=lambda kv: -kv[1])[:3]
        print("%5d  exec %7.1f M  samples %6d  %s   | %s" % (s, e / 1e6, m, ", ".join("%s %d" % (k[6:], v) for k, v in top),
                                                           body[s][ci['Source']].strip()[:40]))
    print("hottest instructions:")
    for r in sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:topn]:
        i = body.index(r)
        tot = sorted(((c, int(r[ci[c]])) for c in stall_cols), key=lambda kv: -kv[1])[:2]
        print("%5d %-60s samples %6s exec %9s  %s" % (i, r[ci['Source']].strip()[:60], r[ci['# Samples']], r[ci['Instructions Executed']],
                                                     ", ".join("%s %d" % (k[6:], v) for k, v in tot)))


if __name__ == "__main__":
    main()

"""Summarise `ncu --page source --csv` output: top stall locations per kernel (used to write profiles/*.md)."""
import csv
import sys


def main(path, kernel_idx=1, n=30):
    rows = list(csv.reader(open(path)))
    hdr, out, k = None, [], 0
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            k += 1
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and k == kernel_idx and len(r) == len(hdr):
            out.append(r)
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot = sum(int(r[si] or 0) for r in out)
    print("total samples", tot, "instructions", len(out))
    agg = {}
    for r in out:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(out, key=lambda r: -int(r[si] or 0))[:n]:
        st = sorted([(int(r[i] or 0), hdr[i]) for i in stall_cols], reverse=True)[:2]
        print(r[si].rjust(7), r[ie].rjust(9), r[src][:100], st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1, int(sys.argv[3]) if len(sys.argv) > 3 else 30)

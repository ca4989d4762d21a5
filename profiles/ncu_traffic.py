"""Turn an `ncu --page raw --csv` export into profiles/kernel_traffic.json (DRAM bytes per launch, per kernel class).

usage: python profiles/ncu_traffic.py <raw.csv> [<raw2.csv> ...] > profiles/kernel_traffic.json
Kernel classes are the names bench.py's kernel-span profiler uses (krasis_b200/csrc/prof.cuh); when a class was captured
more than once the launches are averaged.  `dram__bytes_read.sum + dram__bytes_write.sum` is the `roofline.traffic` figure.
"""
import csv
import json
import os
import sys

CLASSES = [("dense_gemm_kernel<0", "dense_gemm<bf16>"), ("dense_gemm_kernel<1", "dense_gemm<bf16>"), ("dense_gemm_kernel<2", "dense_gemm<int8>"),
           ("grouped_gemm_kernel<0, 1>", "grouped_gemm<gate_up+silu_mul>"), ("grouped_gemm_kernel<0, 0>", "grouped_gemm<down>"),
           ("grouped_gemm_kernel<1, 1>", "grouped_gemm<gate_up+silu_mul>"), ("grouped_gemm_kernel<1, 0>", "grouped_gemm<down>"),
           ("gdn_scan_tc", "gdn_chunk_scan"), ("gdn_chunk_scan", "gdn_chunk_scan"), ("gdn_prepare_tc", "gdn_chunk_prepare"),
           ("gdn_chunk_prepare", "gdn_chunk_prepare"), ("gqa_fmha_kernel", "fmha"), ("combine_kernel", "combine"), ("gdn_prep_vec", "gdn_prep(conv+l2norm+gates)")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main(paths):
    acc = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        kn, rd, wr, tm = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        for r in rows[2:]:
            cls = next((c for pat, c in CLASSES if pat in r[kn]), None)
            if cls is None:
                continue
            a = acc.setdefault(cls, {"n": 0, "bytes": 0.0, "us": 0.0, "src": set()})
            a["n"] += 1
            a["bytes"] += to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
            a["us"] += float(r[tm].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[tm], 1)  # ncu prints usecond as "us"
            a["src"].add("profiles/" + os.path.basename(path))
    out = {c: {"dram_bytes_per_launch": a["bytes"] / a["n"], "launches_captured": a["n"], "ncu_time_us_per_launch": a["us"] / a["n"],
               "source": ", ".join(sorted(a["src"])) + " (ncu --set full, --clock-control none, one kernel per replay: cold-cache, serialised)"}
           for c, a in sorted(acc.items())}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1:])

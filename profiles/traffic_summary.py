"""DRAM bytes of ONE whole render from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--cache-control none --csv` launch list (caches kept between kernels): per kernel group and in total.
usage: python profiles/traffic_summary.py capture.csv [--json]"""
import collections
import csv
import json
import re
import sys

GROUPS = [
    (r"pass_strided_kernel<\d+, \d+, 0, \d+, 15, 0>", "fft:olsb first pass (strided forward, signal windows in)"),
    (r"pass_mid", "fft:olsb middle pass (contiguous forward x IR spectrum x contiguous inverse)"),
    (r"pass_last_pipe_kernel|pass_strided_kernel<\d+, \d+, 1, \d+, 0, 7>", "fft:olsb last pass (strided inverse, stereo frames + maxima out)"),
    (r"final_kernel|pan_max_kernel", "final_kernel (guards, pan, map, clip, PCM16, sums)"),
    (r"loudness_kernel|hop_combine_kernel|gate_kernel", "loudness_kernel (K-weighting stages + hop energies, one pass)"),
    (r"^air_", "air fold chain (air kernel table, far taps, fold)"),
    (r"^ir_|pass_strided_kernel<\d+, \d+, 0, \d+, 16, 0>|pass_contig_kernel", "ir synthesis chain (taps, smoothed tail, envelope, normalisation)"),
]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = rows[0]
    iid, ik, im, iu, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", re.sub(r"^void ", "", r[ik])).replace("fft::", "")
        d = launches.setdefault(int(r[iid]), {"name": name})
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1}.get(u, 1)
        d[r[im]] = v
    ids = list(launches)
    starts = [i for i in ids if launches[i]["name"].startswith("ir_scatter_kernel")]
    if len(starts) < 2:
        sys.exit("need two render starts (ir_scatter_kernel) in the capture")
    render = [launches[i] for i in ids if starts[0] <= i < starts[1]]
    per = collections.OrderedDict()
    tot_r = tot_w = tot_ns = 0.0
    for l in render:
        g = next((g for pat, g in GROUPS if re.search(pat, l["name"])), "other: " + l["name"])
        p = per.setdefault(g, [0.0, 0.0, 0.0, 0])
        rd, wr, ns = l.get("dram__bytes_read.sum", 0.0), l.get("dram__bytes_write.sum", 0.0), l.get("gpu__time_duration.sum", 0.0)
        p[0] += rd; p[1] += wr; p[2] += ns; p[3] += 1
        tot_r += rd; tot_w += wr; tot_ns += ns
    out = {"dram_bytes_per_render": round(tot_r + tot_w, -5), "dram_read_bytes": round(tot_r, -5), "dram_write_bytes": round(tot_w, -5),
           "launches": len(render), "serialised_us": round(tot_ns / 1e3, 1),
           "per_kernel": {g: round(p[0] + p[1], -5) for g, p in per.items()}}
    if "--json" in sys.argv:
        print(json.dumps(out, indent=1))
        return
    for g, p in per.items():
        print(f"{(p[0] + p[1]) / 1e6:9.1f} MB  (read {p[0] / 1e6:8.1f}, write {p[1] / 1e6:8.1f})  {p[2] / 1e3:8.1f} us  n={p[3]:2d}  {g}")
    print(f"{(tot_r + tot_w) / 1e6:9.1f} MB  (read {tot_r / 1e6:8.1f}, write {tot_w / 1e6:8.1f})  {tot_ns / 1e3:8.1f} us  n={len(render)}  one whole render")


if __name__ == "__main__":
    main()

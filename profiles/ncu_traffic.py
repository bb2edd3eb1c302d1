"""Write profiles/traffic.json from an `ncu --set full ... --page raw --csv` dump.

usage: python profiles/ncu_traffic.py <raw.csv> <kernel-name substring> [label]

`dram_bytes_per_launch` = dram__bytes_read.sum + dram__bytes_write.sum of the named kernel, averaged over the
launches of it in the capture.  The file carries the source stamp of the tree it is written from (bench.py's
source_stamp(): a hash of csrc/ and include/), and bench.py reports `roofline.traffic` only when that stamp
equals the stamp of the build it is benchmarking -- so run this right after the capture, before editing kernels.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, needle = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(raw)
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    k = hdr.index("Kernel Name")
    cols = [hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")]
    dur = hdr.index("gpu__time_duration.sum")
    per_launch, times, name = [], [], None
    for r in rows[2:]:
        if needle not in r[k]:
            continue
        name = r[k]
        per_launch.append([float(r[c].replace(",", "")) * UNIT[units[c]] for c in cols])
        times.append(float(r[dur].replace(",", "")))
    if not per_launch:
        raise SystemExit("no launch of a kernel matching %r in %s" % (needle, raw))
    rd = sum(p[0] for p in per_launch) / len(per_launch)
    wr = sum(p[1] for p in per_launch) / len(per_launch)
    import bench
    out = {
        "kernel": name, "launches_in_capture": len(per_launch),
        "dram_bytes_per_launch": int(round(rd + wr)), "dram_read_bytes": int(round(rd)), "dram_write_bytes": int(round(wr)),
        "ncu_duration_%s" % units[dur]: sum(times) / len(times),
        "source": "ncu --set full --clock-control none, %s: dram__bytes_read.sum + dram__bytes_write.sum per launch" % label,
        "source_stamp": bench.source_stamp(),
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

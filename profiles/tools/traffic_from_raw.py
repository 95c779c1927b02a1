"""DRAM bytes per launch from the committed `ncu --set full` raw pages (profiles/r01_prof_*_raw.csv,
made with `ncu -i X.ncu-rep --page raw --csv`) -> profiles/r01_traffic.json, which bench.py reads for
roofline.traffic."""
import csv
import json
import os

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(name):
    rows = list(csv.reader(open(os.path.join(HERE, name))))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[k]].replace(",", "")) * SCALE[units[ix[k]]]
        out.append((r[ix["Kernel Name"]], b, float(r[ix["gpu__time_duration.sum"]].replace(",", "")), units[ix["gpu__time_duration.sum"]]))
    return out


def main():
    res = {}
    f = launches("r01_prof_factor_c2_raw.csv")
    res["c2_factor_kernels"] = {"dram_bytes_per_launch": sum(b for _, b, _, _ in f),
                                "what": "sum over the %d factor2_kernel launches of one GN iteration" % len(f),
                                "source": "ncu --set full --clock-control none (round 1, refreshed); profiles/r01_prof_factor_c2_raw.csv"}
    a = launches("r01_prof_assoc_c4_raw.csv")
    res["c4_assoc_bulk_grid_kernel"] = {"dram_bytes_per_launch": sum(b for _, b, _, _ in a) / len(a),
                                        "what": "assoc_bulk_grid_kernel, one frame (mean of %d launches)" % len(a),
                                        "source": "profiles/r01_prof_assoc_c4_raw.csv"}
    c3 = launches("r01_prof_asm_c3_raw.csv")
    res["c3_assemble_kernels"] = {"dram_bytes_per_launch": 2.0 * sum(b for _, b, _, _ in c3),
                                  "what": "pose + landmark assembly kernel, 2,048 replicas captured, scaled x2 to 4,096",
                                  "source": "profiles/r01_prof_asm_c3_raw.csv"}
    c5 = launches("r01_prof_asm_c5_raw.csv")
    res["c5_assemble_kernels"] = {"dram_bytes_per_launch": sum(b for _, b, _, _ in c5),
                                  "what": "pose + landmark assembly kernel, one assembly",
                                  "source": "profiles/r01_prof_asm_c5_raw.csv"}
    json.dump(res, open(os.path.join(HERE, "r01_traffic.json"), "w"), indent=1)
    for k, v in res.items():
        print(k, "%.1f MB" % (v["dram_bytes_per_launch"] / 1e6))
    for nm, rows in (("factor c2", f), ("assoc c4", a), ("asm c3", c3), ("asm c5", c5)):
        for n, b, t, u in rows:
            print("  %-10s %-60s %10.1f MB %10.2f %s" % (nm, n[:60], b / 1e6, t, u))


if __name__ == "__main__":
    main()

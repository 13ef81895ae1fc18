"""Turns `ncu -i step.ncu-rep --page raw --csv` (one eager step captured with `bench.py --ncu-step`) into
(1) the launch list profiles/<name>_launches.csv (kernel, duration, DRAM bytes) and
(2) profiles/<name>.json: per kernel kind the launch count, summed duration and DRAM traffic - the source of
bench.py's `roofline.traffic`.   usage: ncu_step_summary.py raw.csv out_prefix batch"""
import csv
import re
import json
import sys

KIND = [("decode", "decode"), ("pw_tc_kernel", "pw"), ("pw_ts_kernel", "pw"), ("pw_simt_kernel", "pw"), ("dw_tma_kernel", "dw"), ("dw_kernel", "dw"),
        ("se_fc_kernel", "se"), ("se_kernel", "se"), ("stem_kernel", "stem"), ("rfcr_kernel", "rfcr"),
        ("resample_kernel", "resample"), ("nms_kernel", "nms"), ("pack_kernel", "pack")]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    raw, prefix, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = [r for r in csv.reader(open(raw)) if r and not r[0].startswith("==")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    kinds, launches = {}, []
    per_id = {}
    hdr = rows[0]
    if "Metric Name" in hdr:  # long format (ncu --csv --log-file): one row per (launch, metric)
        ci = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
        for r in rows[1:]:
            if len(r) != len(hdr):
                continue
            e = per_id.setdefault(r[ci["ID"]], {"name": r[ci["Kernel Name"]]})
            e[r[ci["Metric Name"]]] = num(r[ci["Metric Value"]]) * scale.get(r[ci["Metric Unit"]], 1.0)
        items = [(e["name"], e.get("gpu__time_duration.sum", 0.0), e.get("dram__bytes_read.sum", 0.0),
                  e.get("dram__bytes_write.sum", 0.0)) for _, e in sorted(per_id.items(), key=lambda kv: int(kv[0]))]
    else:  # wide format (ncu -i rep --page raw --csv)
        units, data = rows[1], rows[2:]
        col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
        items = []
        for r in data:
            if len(r) != len(hdr):
                continue
            items.append((r[col["Kernel Name"]],
                          num(r[col["gpu__time_duration.sum"]]) * scale.get(units[col["gpu__time_duration.sum"]], 1.0),
                          num(r[col["dram__bytes_read.sum"]]) * scale.get(units[col["dram__bytes_read.sum"]], 1.0),
                          num(r[col["dram__bytes_write.sum"]]) * scale.get(units[col["dram__bytes_write.sum"]], 1.0)))
    for name, t, rd, wr in items:
        kind = next((k for pat, k in KIND if pat in name), "other")
        m = re.search(r"pw_ts_kernel<([^>]*)>", name)
        if m:
            args = [a.replace("(bool)", "").replace("(int)", "").strip() for a in m.group(1).split(",")]
            if len(args) >= 2 and args[1] not in ("0", "false"):  # <DBG, FRONT, CG>: FRONT = depthwise stride
                kind = "dwpw"  # the depthwise-front instantiations of the tcgen05 pointwise kernel (YR_OP_DWPW)
        launches.append((name.split("(")[0], kind, t, rd, wr))
        k = kinds.setdefault(kind, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
        k["launches"] += 1
        k["time_us"] += t
        k["dram_bytes"] += rd + wr
    tot = sum(k["time_us"] for k in kinds.values())
    for k in kinds.values():
        k["share"] = round(k["time_us"] / tot, 4)
        k["time_us"] = round(k["time_us"], 1)
        k["dram_bytes"] = int(k["dram_bytes"])
    with open(prefix + "_launches.csv", "w") as f:
        f.write("kernel,kind,gpu__time_duration_us,dram_read_bytes,dram_write_bytes\n")
        for n, kind, t, rd, wr in launches:
            f.write('"%s",%s,%.2f,%d,%d\n' % (n, kind, t, rd, wr))
    json.dump({"batch": batch, "total_time_us": round(tot, 1), "kinds": kinds,
               "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                      "--profile-from-start off python bench.py --ncu-step (cold-cache, serialised launches: compare shares)"},
              open(prefix + ".json", "w"), indent=1)
    print(json.dumps(kinds, indent=1))


if __name__ == "__main__":
    main()

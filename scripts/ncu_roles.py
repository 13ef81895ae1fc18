"""Per-role warp-state sampling of one warp-specialised kernel from an `ncu --set full` report (source page, SASS):
the kernel's instructions are cut into address segments, each segment is labelled by the marker instructions it
contains (UTMALDG = TMA producer, FFMA2/STTM = converters, LDTM/STG = epilogue, UTCHMMA/UTCBAR = MMA issuer) and its
samples are broken down by stall reason.  usage: ncu_roles.py report.ncu-rep [segment length]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h2 = {h: i for i, h in enumerate(raw[0])}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError):
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("report %s: %s, %s us, %s instructions, issue slots busy %s %%, %d samples" % (
    rep.split("/")[-1], rows[0][1][:60], raw[2][h2["gpu__time_duration.sum"]], raw[2][h2["smsp__inst_executed.sum"]],
    raw[2][h2["smsp__issue_active.avg.pct_of_peak_sustained_active"]][:5], tot))
stalls = [k for k in hdr if k.startswith("stall_") and "Not" not in k]
marks = ("UTMALDG", "UBLKCP", "FFMA2", "STTM", "LDTM", "STG", "UTCHMMA", "UTCBAR", "BAR")
print("| first address | samples | share | max executions | markers | stall reasons (>= 4 % of the segment) |")
print("|---|---|---|---|---|---|")
for a in range(0, len(data), step):
    seg = data[a:a + step]
    s = sum(f(r, "# Samples") for r in seg)
    if s < 0.004 * tot:
        continue
    ex = max(f(r, "Instructions Executed") for r in seg)
    agg = {k: sum(f(r, k) for r in seg) for k in stalls}
    top = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.04 * s)
    ops = set()
    for r in seg:
        w = r[1].split()
        if w:
            ops.add((w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0])
    print("| ...%s | %d | %.1f %% | %d | %s | %s |" % (seg[0][0][-5:], s, 100 * s / tot, ex, " ".join(m for m in marks if m in ops), top))

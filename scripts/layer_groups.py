"""Per-group roofline fractions from a per-launch timing file of `Engine.profile_layers` (profiles/r2_layers_final.json):
bytes / CUDA-event time against the measured HBM peak, in the executed accounting (what the launch must move) and in
SURVEY.md 8d's per-layer accounting (as if fused intermediates existed).  usage: layer_groups.py layers.json [peak GB/s]"""
import json
import sys

L = json.load(open(sys.argv[1]))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6547.8
S2 = ("block_1", "block_3", "block_6", "block_13")  # the stride-2 blocks of MobileNetV2


def grp(l):
    n, k = l["name"], l["kind"]
    if k == "dwpw":
        return "fused depthwise->pointwise, stride 2" if n.split("_depthwise")[0] in S2 else "fused depthwise->pointwise, stride 1"
    if k == "pw":
        return "backbone expand 1x1" if n.startswith("block_") and "expand" in n else "head / FPN / RFCR 1x1"
    return k


G = {}
for l in L:
    a = G.setdefault(grp(l), [0, 0.0, 0, 0])
    a[0] += l.get("launches", 1)
    a[1] += l["ms"]
    a[2] += l.get("bytes", 0)
    a[3] += l.get("ref_bytes", l.get("bytes", 0))
print("| group | launches | ms | executed GB | fraction of HBM peak | (per-layer accounting) |")
print("|---|---|---|---|---|---|")
for g, a in sorted(G.items(), key=lambda x: -x[1][1]):
    f = lambda b: b / 1e9 / a[1] * 1e3 / peak if a[1] else 0.0
    print("| %s | %d | %.3f | %.2f | %.2f | %.2f |" % (g, a[0], a[1], a[2] / 1e9, f(a[2]), f(a[3])))
print("\nsum of the per-launch times: %.3f ms (event-timed one by one; the graph replay of the same launches is shorter)" % sum(a[1] for a in G.values()))

import json, sys
for path in sys.argv[1:]:
    line = [l for l in open(path) if l.startswith("RESULT ")]
    if not line:
        print(path, "no RESULT"); print(open(path).read()[-2000:]); continue
    d = json.loads(line[-1][7:])
    print(path, {k: round(v, 2) for k, v in d["traced_ms"].items() if k.startswith("rec")}, "untraced", {k: round(v, 2) for k, v in d["untraced_ms"].items() if k.startswith("rec")})
    for l in ("layer0", "layer1"):
        print(" ", l, "step cycles", d[l]["step_cycles_median"])
        for k, v in d[l]["median_offset_from_h_ready_seen"].items():
            print("     %-45s %7.0f" % (k, v))
        print("     arrivals", [int(x) for x in d[l]["gate_warp_arrivals_median"]])
        for k in d[l]:
            if k.startswith("aux_warp"): print("     " + k, [int(x) for x in d[l][k]])

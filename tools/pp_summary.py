import json, sys
for block in open(sys.argv[1]).read().split("== ")[1:]:
    head, _, rest = block.partition("\n")
    line = [l for l in rest.splitlines() if l.startswith("RESULT ")]
    if not line:
        print(head, "FAILED", rest[:300]); continue
    r = json.loads(line[0][7:])
    print(head, "rec0 %.1f ms rec1 %.1f ms" % (r["traced_ms"]["rec0_ms"], r["traced_ms"]["rec1_ms"]))
    for layer in ("layer0", "layer1"):
        for t in ("tile0", "tile1"):
            o = r[layer][t]["offsets"]
            keys = ["issuer: r guard passed", "issuer: r queued", "issuer: n queued", "issuer: r commit seen by issuer (L0 only)", "issuer: logits queued", "relay: r arrived", "relay: z arrived", "relay: n arrived", "gate: r ld done", "gate: r math done", "gate: z ld done", "gate: z math done", "gate: n ld done", "gate: h written", "gate: arrived H"]
            print("   %s %s period %5.0f | " % (layer, t, r[layer][t]["period"]) + " ".join("%5.0f" % o[k] for k in keys))

import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    k = d.get("kernel_ms_per_step", {})
    print(path, "| GCells/s %.0f | ms/step %.2f | samples/s %.0f | e2e ms %.2f | launches %s | roof %.3f (%.2f TCells/s) | lik GB/s %.0f | truth %s/%s" % (
        d["value"], d["ms_per_step"], d["samples_per_s"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"],
        d["roofline"]["cells_per_s"]/1e12, d["roofline_likelihood"]["achieved"], d["parity"]["genes_matching_generator_truth"], d["parity"]["genes"]))
    print("   kernels ms/step:", {a: round(b, 2) for a, b in k.items()}, "clocks", d.get("clocks"))

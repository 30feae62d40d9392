"""Per-kernel table of a bench line: python tools/kernel_table.py profiles/<bench>.json > profiles/<name>_kernel_table.md"""
import json
import sys

d = json.load(open(sys.argv[1]))
peak = d["roofline"]["peak"]
steps = d["steps"]
print("# per-kernel CUDA-event times of `%s` (%s, %d GPU)\n" % (sys.argv[1].split("/")[-1], d["config"]["workload"], d["n_gpus"]))
print("step %.3f ms resident (%.2f G PCM frames/s), %.2f ms end to end; HBM peak %.1f GB/s (%s)\n" % (
    d["ms_per_step"], d["value"] / 1e9, d["e2e"]["ms_per_step"], peak, d["roofline"]["peak_source"]))
print("| kernel | launches per step | ms per launch | ms per step | share of step | SURVEY 8(d) GB/s | frac of HBM peak | moved GB/s (incl. intermediates) |")
print("|---|---|---|---|---|---|---|---|")
tot = 0.0
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_total"]):
    a = v["algo_GBps"]
    print("| `%s` | %.2f | %.4f | %.3f | %.1f %% | %s | %s | %s |" % (
        k, v["launches"] / steps, v["ms_per_launch"], v["ms_total"] / steps, 100 * v["share_of_step"],
        "%.0f" % a if a else "—", "%.3f" % (a / peak) if a else "—", "%.0f" % v["moved_GBps"] if v["moved_GBps"] else "—"))
    tot += v["ms_total"] / steps
print("| all kernels | | | %.3f | %.1f %% | | | |" % (tot, 100 * tot / d["ms_per_step"]))
print("| host between launches | | | %.3f | %.1f %% | | | |" % (d["ms_per_step"] - tot, 100 * (1 - tot / d["ms_per_step"])))
r = d["roofline"]
print("\nroofline: dominant kernel `%s` %.0f GB/s of %.1f = **%.3f** (SURVEY 8(d) bytes); whole path %.0f GB/s = %.3f; embed kernel %.3f; `k_sync_gather` moves %.0f GB/s = %.2f of peak." % (
    r["kernel"], r["achieved"], peak, r["frac"], r["path"]["achieved"], r["path"]["frac"], r["embed_kernel"]["frac"],
    r["hbm_bound_kernel"]["moved_GBps"], r["hbm_bound_kernel"]["frac_of_peak_moved"]))

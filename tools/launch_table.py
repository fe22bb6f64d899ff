"""Per-kernel table of ONE bench step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/launch_table.py gpurun_out/launches.csv [out.csv.gz]

The step starts at the last `pack_keys_kernel<float>` launch of the list (the timed step's encode); everything before it is
warm-up.  Prints a markdown table (launches, ms, share) and optionally writes the step's launches as a gzipped csv."""
import collections
import csv
import gzip
import re
import sys


def short(k: str) -> str:
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"\(.*$", "", k)
    return k.replace("(int)", "").replace("<unnamed>::", "")


def main() -> int:
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = list(csv.reader(lines))
    ix = {n: i for i, n in enumerate(r[0])}
    L = [(x[ix["Kernel Name"]], float(x[ix["Metric Value"]].replace(",", "")), x[ix["Grid Size"]], x[ix["Block Size"]]) for x in r[1:]
         if len(x) > ix["Metric Value"]]
    starts = [i for i, x in enumerate(L) if "pack_keys_kernel<float>" in x[0]]
    S = L[starts[-1]:]
    tot = sum(x[1] for x in S) / 1e6
    print(f"launches in the step: {len(S)}, sum of kernel times {tot:.1f} ms\n")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, t, _, _ in S:
        a = agg[short(k)]
        a[0] += 1
        a[1] += t / 1e6
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {t:.3f} | {100 * t / tot:.1f} % |")
    conv = sum(t for k, (c, t) in agg.items() if k.startswith(("spconv_", "sp_straggler", "sp_centre")))
    print(f"\nAll sparse-conv launches together (spconv_fwd_* + sp_straggler + sp_centre): {conv:.1f} ms = {100 * conv / tot:.1f} % of the kernel time.")
    if len(sys.argv) > 2:
        with gzip.open(sys.argv[2], "wt", newline="") as g:
            w = csv.writer(g)
            w.writerow(["kernel", "grid", "block", "gpu__time_duration.sum [ns]"])
            for k, t, gr, bl in S:
                w.writerow([short(k), gr, bl, int(t)])
    return 0


if __name__ == "__main__":
    sys.exit(main())

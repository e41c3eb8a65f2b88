"""profiles/<run>_bench_launch_summary.md from the ncu launch list of a bench run (run here, no GPU).

    python tools/launch_summary.py profiles/r05r_bench_launches.csv profiles/r05r_bench_1gpu.json profiles/r05r_bench_launch_summary.md

The list is the `--metrics gpu__time_duration.sum --clock-control none` pass of B200_PROFILING.md over
`python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e`.  One step = the forward + backward of objective 0 and
of objective 1, each from its xprep_kernel to its finalize_kernel; medians over the last three steps of the list."""
import collections, csv, json, statistics, sys


def main(csv_path, bench_path, out_path):
    rows = list(csv.reader(open(csv_path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    L = [(r[kn].split("(")[0].replace("void ", "").replace("dkg::", "").replace("<unnamed>::", ""), float(r[mv].replace(",", "")) / 1e3)
         for r in rows[hi + 1:] if len(r) > mv]
    fin = [i for i, (n, _) in enumerate(L) if n.startswith("finalize_kernel")]
    seqs = []
    for f in fin:
        s = f
        while s > 0 and L[s][0] != "xprep_kernel":
            s -= 1
        seqs.append(L[s:f + 1])

    def agg(seq):
        d, c = collections.defaultdict(float), collections.Counter()
        for n, t in seq:
            key = n
            if n == "ozaki_kernel":  # the same kernel serves the covariance contraction (~0.6 ms) and the T solve products (~35 us)
                key = "ozaki_kernel (covariance contraction)" if t > 400 else "ozaki_kernel (T solve products)"
            d[key] += t
            c[key] += 1
        return d, c

    per_obj = [[], []]
    for k, s in enumerate(seqs[-6:]):
        per_obj[k % 2].append(agg(s))
    keys = []
    for o in (0, 1):
        for d, c in per_obj[o]:
            for k in d:
                if k not in keys:
                    keys.append(k)
    tab = [(k, [statistics.median([c.get(k, 0) for d, c in per_obj[o]]) for o in (0, 1)],
            [statistics.median([d.get(k, 0.0) for d, c in per_obj[o]]) for o in (0, 1)]) for k in keys]
    tot = sum(u[0] + u[1] for _, _, u in tab)
    b = json.loads(open(bench_path).read().strip().splitlines()[-1])
    name = csv_path.split("/")[-1]
    out = [f"# {name.split('_')[0]} — launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e` under",
           "# `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes).",
           f"# Raw list: {name} ({len(L)} launches: plan builds, warm-up, timed and profiled steps, the int8 MMA-peak launches of",
           "# dkg_int8_peak and the cuBLAS DGEMM peak measurement).  Below: the kernels of ONE step = the forward + backward of objective 0 and of",
           "# objective 1 (each from its xprep_kernel to its finalize_kernel), median over the last three steps of the list.", "",
           "| kernel | launches (obj 0 / 1) | us per step, objective 0 | objective 1 | share of the step |", "|---|---|---|---|---|"]
    for k, n, u in sorted(tab, key=lambda x: -(x[2][0] + x[2][1])):
        out.append(f"| {k} | {n[0]:.0f} / {n[1]:.0f} | {u[0]:.1f} | {u[1]:.1f} | {(u[0] + u[1]) / tot:.3f} |")
    out += ["", f"Serialised step (sum): {tot / 1e3:.2f} ms under ncu; bench (CUDA events, two streams): {b['ms_per_step']:.2f} ms; "
                f"per-category CUDA-event sums {sum(b['kernel_ms_per_step'].values()):.2f} ms.",
            f"Covariance contraction: share of the step in this list vs `roofline.share_of_step` of the bench line "
            f"({b['roofline']['share_of_step']:.3f}): see the first row."]
    open(out_path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])

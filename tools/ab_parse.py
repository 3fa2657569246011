import json, sys, glob, statistics as st
# usage: ab_parse.py NVARIANTS   -- per variant: median over repetitions of the per-run MEDIAN kernel times
n = int(sys.argv[1])
for k in range(n):
    rows = []
    for f in sorted(glob.glob(f"gpurun_out/ab_{k}_r*.log")):
        try:
            d = json.loads(open(f).read().strip().splitlines()[-1])
            ks = d["roofline"]["kernels"]
            rows.append((ks["stft_fwd_kernel"]["ms_min_median_max"][1], ks["istft_inv_kernel"]["ms_min_median_max"][1]))
        except Exception as e:
            rows.append((float("nan"), float("nan")))
    f1 = [r[0] for r in rows]; f2 = [r[1] for r in rows]
    print(f"variant {k}: K1 med {st.median(f1):.3f} | K2 med {st.median(f2):.3f} | runs {[(round(a,3), round(b,3)) for a,b in rows]}")

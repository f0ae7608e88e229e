#!/bin/bash
# second pass over the V-cycle knobs at 256^3 + per-kernel launch list of the multigrid iteration (ncu, cold-cache serialised times: shares only)
O=gpurun_out
run() {
    tag=$1; shift
    env "$@" timeout 100 python tools/run_poisson3d.py --nx ${NX:-256} --precond mg --repeat 2 > $O/r2_mg_$tag.json 2> $O/r2_mg_$tag.err
    python - "$tag" $O/r2_mg_$tag.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:28s} it {d['iterations']:4d}  loop {d['krylov_loop_ms']:8.1f} ms  {d['ms_per_iteration']:6.3f} ms/it  launches {d['launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run lev4_alpha4 PB200_MG_LEVELS=4 PB200_MG_ALPHA=4
run lev4_alpha3 PB200_MG_LEVELS=4 PB200_MG_ALPHA=3
run lev4_alpha2 PB200_MG_LEVELS=4 PB200_MG_ALPHA=2
run lev5_alpha3 PB200_MG_LEVELS=5 PB200_MG_ALPHA=3
run lev3_alpha3 PB200_MG_LEVELS=3 PB200_MG_ALPHA=3
run deg1_alpha2 PB200_MG_DEG=1 PB200_MG_ALPHA=2 PB200_MG_LEVELS=4
run deg1_alpha3 PB200_MG_DEG=1 PB200_MG_ALPHA=3 PB200_MG_LEVELS=4
PB200_NO_GRAPH=1 PB200_MG_LEVELS=4 PB200_MG_ALPHA=4 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 260 --csv --log-file $O/r2_mg_launches_256.csv \
    python tools/run_poisson3d.py --nx 256 --precond mg --repeat 1 > $O/r2_mg_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2_mg_launches_256.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    k = r[ki].split("(")[0][:60]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} {n:4d} launches {t:10.1f} us  {100 * t / tot:5.1f} %")
PY

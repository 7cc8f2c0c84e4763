#!/bin/bash
# Turn the files a `round_evidence.sh <tag>` run left in gpurun_out/ into the tracked summaries
# under profiles/ (run here, no GPU).  Usage: bash scripts/collect_evidence.sh <tag> [<old tag to replace>]
tag=$1; old=$2
set -e
for w in C4 C2 C3 C3tri C5dis C5nn; do
  python scripts/ncu_summary.py gpurun_out/${tag}_$w.ncu-rep "round 2 ($tag): ncu --set full --clock-control none of \`python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0\`, first launch(es) after 3 skipped" > profiles/${tag}_ncu_$w.txt
done
python scripts/make_traffic.py $tag C4:1095:f32:f64:gpurun_out/${tag}_C4.ncu-rep C2:3441:f32:f64:gpurun_out/${tag}_C2.ncu-rep C3:96:f32:f64:gpurun_out/${tag}_C3.ncu-rep C5dis:64:f32:f64:gpurun_out/${tag}_C5dis.ncu-rep C5nn:64:f32:f64:gpurun_out/${tag}_C5nn.ncu-rep > /dev/null
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
p = 'profiles/traffic.json'
d = json.load(open(p))
for c in d["captures"]:
    c["source"] = "profiles/%s_ncu_%s.txt" % (tag, c["workload"])
json.dump(d, open(p, 'w'), indent=1)
rows = list(csv.reader(open('gpurun_out/%s_launches_c4.csv' % tag)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
col = {h: i for i, h in enumerate(rows[hdr])}
out, tot = [], {}
for r in rows[hdr + 1:]:
    if len(r) < len(col):
        continue
    name = r[col["Kernel Name"]].split("(")[0][:70]
    v = float(r[col["Metric Value"]].replace(",", "")); u = r[col["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
    out.append((r[col["ID"]], name, r[col["Grid Size"]], ms)); tot[name] = tot.get(name, 0) + ms
with open('profiles/%s_launches_c4_default.txt' % tag, 'w') as f:
    f.write("# round 2 (%s): every launch of `python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0` with its device time\n# (ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised -- compare shares)\n" % tag)
    for o in out:
        f.write(f"{o[0]:>4} {o[3]:10.4f} ms  grid {o[2]:<16} {o[1]}\n")
    f.write("# totals per kernel\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"# {v:10.4f} ms  {k}\n")
print("traffic hash", d["kernel_source_hash"])
PY
cp gpurun_out/${tag}_bench_c4_n1.json profiles/${tag}_bench_c4_n1.json
cp gpurun_out/${tag}_bench_reference.json profiles/${tag}_bench_c4_reference_arm.json
python scripts/sass_summary.py > profiles/r02_sass_summary.txt
if [ -n "$old" ]; then git rm -q profiles/${old}_* ; sed -i "s/${old}_/${tag}_/g" profiles/README.md; fi
tail -2 gpurun_out/${tag}_pytest.log

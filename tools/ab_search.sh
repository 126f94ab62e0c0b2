#!/bin/bash
# A/B of the search-kernel variants on the bench workload: tools/ab_search.sh 0 2 4 6
for v in "$@"; do
  PPCR_SEARCH_VARIANT=$v python bench.py --no-cpu --steps 5 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_v$v.json"))
k = d["roofline"]["kernels"]
print("variant $v: %.2f ms/step, e2e %.2f ms, outer %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["outer_iterations"]),
      {n: (round(x["avg_ms"], 4), round(x.get("isolated_avg_ms", 0), 4)) for n, x in k.items()})
PY
done

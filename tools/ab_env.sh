#!/bin/bash
# A/B of environment settings on the bench workload: tools/ab_env.sh "PPCR_SEARCH_QUEUED=0" "PPCR_SEARCH_QUEUED=1" ...
n=0
for e in "$@"; do
  n=$((n+1))
  env $e python bench.py --no-cpu --steps 5 > gpurun_out/bench_e$n.json 2> gpurun_out/bench_e$n.err || tail -3 gpurun_out/bench_e$n.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_e$n.json"))
k = d["roofline"]["kernels"]
print("$e: %.2f ms/step, e2e %.2f ms, outer %d, K %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["outer_iterations"], d["config"]["correspondences_per_pair"]),
      {n: (round(x["avg_ms"], 4), round(x.get("isolated_avg_ms", 0), 4)) for n, x in k.items()})
PY
done

"""One c5 pair: bench kernel breakdown + set-up trace (debug aid)."""
import json, sys
d = json.load(open(sys.argv[1]))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["outer_iterations"], d["config"]["correspondences_per_pair"], d["gpu_launches"])
for k, v in d["roofline"]["kernels"].items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
print(d["roofline"]["share_of_step_ms"])

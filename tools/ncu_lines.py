"""Per-source-line instruction counts of one kernel from an .ncu-rep (run here, no GPU needed).

    python tools/ncu_lines.py rep.ncu-rep kernel_mangled_substring [top] [inst|samples]

Joins ncu's SASS page (address -> executed instructions, stall samples) with nvdisasm's line table of the built
library (address -> file:line, inlining included), because `ncu --page source --csv` carries no metrics for the CUDA-C view.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "probabilistic_point_clouds_registration_b200", "csrc", "libppcr_cuda.so")

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
order = 2 if (len(sys.argv) > 4 and sys.argv[4].startswith("s")) else 0

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()

# address -> (file, line) for the wanted function
line_of = {}
in_fn = False
cur = ("?", 0)
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        in_fn = kern in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

# a capture may hold several kernels (the search launches k_search and k_search_q back to back): NCU_KERNEL=<regex> picks one
sel = ["--kernel-name", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
raw = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
base = None
per_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
total = [0.0, 0.0, 0.0]
for r in rows[hdr_i + 1:]:
    if len(r) <= ci["Thread Instructions Executed"]:
        continue
    try:
        addr = int(r[ci["Address"]], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    off = addr - base
    vals = [float(r[ci["Instructions Executed"]] or 0), float(r[ci["Thread Instructions Executed"]] or 0), float(r[ci["# Samples"]] or 0)]
    key = line_of.get(off, ("?", 0))
    for k in range(3):
        per_line[key][k] += vals[k]
        total[k] += vals[k]

src_cache = {}


def src_text(f, l):
    if f not in src_cache:
        for d in ("probabilistic_point_clouds_registration_b200/csrc", "include"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                src_cache[f] = open(p).read().splitlines()
                break
        else:
            src_cache[f] = []
    t = src_cache[f]
    return t[l - 1].strip()[:90] if 0 < l <= len(t) else ""


print(f"total warp instructions {total[0]:.0f}, thread instructions {total[1]:.0f} (avg {total[1] / max(total[0], 1):.1f} lanes), samples {total[2]:.0f}")
for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][order])[:top]:
    print(f"{100 * v[0] / total[0]:5.1f}% inst {100 * v[2] / max(total[2], 1):5.1f}% smp  lanes {v[1] / max(v[0], 1):4.1f}  {key[0]}:{key[1]:<4d} {src_text(*key)}")

"""The C-ABI library loads without a GPU and exports every symbol include/ppcr.h declares; the product never
touches oracle/."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "probabilistic_point_clouds_registration_b200")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ppcr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppcr_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(capi):
    declared = _declared_symbols()
    assert len(declared) >= 20
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTED_SYMBOLS) == declared


def test_version_and_defaults_need_no_gpu(capi):
    assert b"sm_100a" in capi.lib().ppcr_version()
    p = capi.make_params()
    # struct defaults, params.hpp:6-17
    assert (p.max_neighbours, p.dof, p.radius, p.n_iter, p.cost_drop_thresh, p.n_cost_drop_it) == (20, 5.0, 1.0, 1000, 0.01, 5.0)
    assert list(p.initial_rotation) == [1.0, 0.0, 0.0, 0.0] and p.source_filter_size == 0 and p.target_filter_size == 0


def test_product_never_references_the_oracle():
    offenders = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", ".cpp")):
                path = os.path.join(base, f)
                if re.search(r"\boracle\b", open(path, errors="ignore").read()):
                    offenders.append(os.path.relpath(path, ROOT))
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            if re.search(r"\boracle\b", open(os.path.join(base, f), errors="ignore").read()):
                offenders.append(f)
    assert not offenders, offenders


def test_cuda_library_holds_sm100a_code(capi):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out

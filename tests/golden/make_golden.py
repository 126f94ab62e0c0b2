"""Writes the fixtures under tests/golden/.

  reference_weights.json   G1/G2: the known-answer vectors of the reference's own unit test, transcribed from
                           /root/reference/test/ProbabilisticWeightsTest.cc (lines cited per entry).  The reference
                           cannot be compiled in this image (PCL / Ceres / Eigen / GTest absent), so these are
                           transcriptions of its constants, not outputs of a run.
  reference_fixture.npz    G3/G4: the analytic cloud + transform of test/PointCloudRegistrationTest.cc:12-37 as float32
                           arrays, and its acceptance bound.
  oracle_c1_small.npz      output of OUR oracle (oracle/ppcr_oracle.cpp) on a reduced BASELINE config 1: guards the
                           oracle and the CUDA path against drifting together.  Regenerate with this script only when
                           the restated algorithm is deliberately changed.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from probabilistic_point_clouds_registration_b200 import synth  # noqa: E402


def main():
    weights = {
        "source": "test/ProbabilisticWeightsTest.cc",
        "pattern_rows": [[0, 2, 3], [0, 1, 2, 3]],                # :19-33  2 x 4 sparsity pattern
        "squared_errors": [1, 1, 1, 1, 4, 9, 16],                # :16
        "dimension": 1, "max_neighbours": 4,                     # :39, :55
        "t_distribution": {"dof": 5, "row0": [1 / 3, 1 / 3, 1 / 3],
                           "row1": [0.7151351, 0.1412613, 0.0241258, 0.0047656], "lines": "35-49", "tol": 1e-6},
        "gaussian": {"dof": "inf", "row0": [1 / 3, 1 / 3, 1 / 3],
                     "row1": [0.805153702921689, 0.179654074677018, 0.0147469044726408, 0.000445317928652638],
                     "lines": "51-66", "tol": 1e-6},
    }
    json.dump(weights, open(os.path.join(HERE, "reference_weights.json"), "w"), indent=1)

    src = synth.reference_test_cloud()
    T = synth.reference_test_transform()
    np.savez_compressed(os.path.join(HERE, "reference_fixture.npz"), source=src, T=T, target=synth.apply_T_like_pcl(src, T),
                        mean_error_bound=1e-6, function_tolerance=10e-5, max_neighbours=3)

    s, t, _ = synth.config1_plane_sphere(seed=77, n_plane=600, n_sphere=400)
    out = {}
    for name, dof, radius in (("t5_r1", 5.0, 1.0), ("gauss_r1", np.inf, 1.0), ("t5_r3", 5.0, 3.0)):
        r = O.align(s, t, O.make_params(max_neighbours=20, dof=dof, radius=radius), O.make_options(inner_kind=1))
        out[name + "_history"] = r.history
        out[name + "_K"] = np.array([x["n_correspondences"] for x in r.stats])
        out[name + "_cost"] = np.array([[x["initial_cost"], x["final_cost"]] for x in r.stats])
        out[name + "_lm"] = np.array([x["lm_iterations"] for x in r.stats])
    np.savez_compressed(os.path.join(HERE, "oracle_c1_small.npz"), seed=77, n_plane=600, n_sphere=400, **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()

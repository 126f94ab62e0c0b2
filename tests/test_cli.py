"""The reference-facing C++ layer: PCD reader/writer, the tclap-compatible command line and (on a GPU) the whole
prob_point_cloud_registration binary against the oracle.  Reference behaviour: src/prob_point_cloud_registration_ex.cc."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import pose_delta
from pcd import read_pcd_xyz, write_pcd
from probabilistic_point_clouds_registration_b200 import build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def cli():
    build.build_all()
    assert os.path.exists(build.CLI_BIN)
    return build.CLI_BIN


@pytest.fixture(scope="session")
def pcd_tool():
    here = os.path.join(ROOT, "tests", "emu")
    exe = os.path.join(here, "pcd_tool")
    src = os.path.join(here, "pcd_tool.cpp")
    hdr = os.path.join(ROOT, "include", "ppcr_compat", "pcl", "io", "pcd_io.h")
    if not os.path.exists(exe) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(exe):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include", "ppcr_compat"), "-o", exe, src])
    return exe


def _run(cmd, cwd):
    return subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("kind", ["ascii", "binary", "binary_compressed"])
@pytest.mark.parametrize("extra", [False, True])
def test_pcd_reader_and_writer(pcd_tool, tmp_path, kind, extra):
    src, _, _ = synth.config1_plane_sphere(n_plane=150, n_sphere=117)
    write_pcd(tmp_path / "in.pcd", src, kind, extra_field=extra)
    for mode in ("ascii", "binary"):
        r = _run([pcd_tool, "in.pcd", "out.pcd", mode], tmp_path)
        assert r.returncode == 0 and r.stdout.strip() == str(len(src))
        back = read_pcd_xyz(tmp_path / "out.pcd")
        if mode == "binary" or kind != "ascii":
            assert np.array_equal(back.view(np.uint32), src[:, :3].view(np.uint32)) or mode == "ascii"
        np.testing.assert_allclose(back, src[:, :3], rtol=2e-7, atol=0)


def test_pcd_reader_rejects_garbage(pcd_tool, tmp_path):
    (tmp_path / "bad.pcd").write_text("VERSION 0.7\nFIELDS a b\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n")
    assert _run([pcd_tool, "bad.pcd", "o.pcd"], tmp_path).returncode == 1          # no x/y/z fields
    (tmp_path / "short.pcd").write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary\n1234")
    assert _run([pcd_tool, "short.pcd", "o.pcd"], tmp_path).returncode == 1        # truncated payload
    assert _run([pcd_tool, "missing.pcd", "o.pcd"], tmp_path).returncode == 1


def test_cli_argument_errors_exit_like_the_reference(cli, tmp_path):
    src, tgt, _ = synth.config1_plane_sphere(n_plane=40, n_sphere=40)
    write_pcd(tmp_path / "s.pcd", src)
    write_pcd(tmp_path / "t.pcd", tgt)
    r = _run([cli, "s.pcd"], tmp_path)                                   # tclap ArgException -> stderr, EXIT_FAILURE
    assert r.returncode == 1 and r.stderr.startswith("error: ") and "for arg" in r.stderr
    r = _run([cli, "s.pcd", "t.pcd", "-m", "abc"], tmp_path)
    assert r.returncode == 1 and "for arg" in r.stderr
    r = _run([cli, "s.pcd", "t.pcd", "--no_such_flag"], tmp_path)
    assert r.returncode == 1
    r = _run([cli, "nope.pcd", "t.pcd"], tmp_path)                       # message on stdout, EXIT_FAILURE (CLI:113-116)
    assert r.returncode == 1 and "Could not load source cloud, closing" in r.stdout
    r = _run([cli, "s.pcd", "nope.pcd"], tmp_path)
    assert r.returncode == 1 and "Could not load target cloud, closing" in r.stdout
    r = _run([cli, "--version"], tmp_path)
    assert r.returncode == 0 and "1.0" in r.stdout


@pytest.mark.gpu
def test_cli_end_to_end_matches_oracle(cli, oracle, tmp_path):
    """Flags of BASELINE config 1 at CLI defaults (r=3, m=20, dof=5) plus -v --dump -g: history, aligned cloud,
    summary file and ground-truth metric, against the oracle on the same clouds."""
    src, tgt, T_true = synth.config1_plane_sphere(n_plane=2500, n_sphere=2500)
    gt = synth.apply_T_like_pcl(src, T_true)
    write_pcd(tmp_path / "source_scan.pcd", src, "binary")
    write_pcd(tmp_path / "target_scan.pcd", tgt, "ascii")
    write_pcd(tmp_path / "gt.pcd", gt, "binary")
    r = _run([cli, "source_scan.pcd", "target_scan.pcd", "-v", "--dump", "-g", "gt.pcd"], tmp_path)
    assert r.returncode == 0, r.stderr
    hist = re.findall(r"^T: (.*?) \|\|\| R: (.*)$", r.stdout, flags=re.M)
    ref = oracle.align(src, tgt, oracle.make_params(max_neighbours=20, dof=5.0, radius=3.0), oracle.make_options(inner_kind=1))
    assert abs(len(hist) - ref.n_outer) <= 1
    t = np.array([float(v) for v in hist[-1][0].split(",")])
    q = np.array([float(v) for v in hist[-1][1].split(",")])           # x, y, z, w
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    rot, tr = pose_delta(T, ref.transformation)
    assert rot < 2e-4 and tr < 2e-4                                   # printed with 6 significant digits
    aligned = read_pcd_xyz(tmp_path / "aligned_source_scan.pcd")
    want = synth.apply_T_like_pcl(src, ref.transformation)[:, :3]
    assert np.max(np.abs(aligned - want)) < 1e-3
    lines = (tmp_path / "source_scan_target_scan_summary.txt").read_text().splitlines()
    assert lines[0].startswith("Source: source_scan.pcd with filter size: 0")
    assert lines[3] == "iter, n_success_steps, initial_cost, final_cost, tx, ty, tz, roll, pitch, yaw, mse_prev_iter, mse_gtruth"
    rows = [ln.split(", ") for ln in lines[4:]]
    assert len(rows) == len(hist) and all(len(rw) == 12 for rw in rows)
    np.testing.assert_allclose(float(rows[0][2]), ref.stats[0]["initial_cost"], rtol=1e-4)
    m = re.findall(r"MSE w.r.t. ground truth: ([0-9.eE+-]+)", r.stdout)
    want_mse = oracle.calculate_mse(synth.apply_T_like_pcl(src, ref.transformation), gt)   # calculateMSE, CLI:186
    assert len(m) >= 2 and abs(float(m[-1]) - want_mse) < 1e-3 and float(m[-1]) < float(m[0])
    assert abs(float(rows[-1][11]) - want_mse) < 1e-3


@pytest.mark.gpu
def test_cli_gaussian_with_filters(cli, oracle, tmp_path):
    src, tgt, _ = synth.lidar_pair(3, 24, 500, outlier_frac=0.2)
    write_pcd(tmp_path / "a.pcd", src, "binary")
    write_pcd(tmp_path / "b.pcd", tgt, "binary")
    r = _run([cli, "a.pcd", "b.pcd", "-u", "-s", "0.3", "-t", "0.3", "-v"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Using gaussian model" in r.stdout and "Filtering source point cloud with leaf of size 0.3" in r.stdout
    hist = re.findall(r"^T: (.*?) \|\|\| R: (.*)$", r.stdout, flags=re.M)
    leaf = float(np.float32(0.3))                                      # the CLI reads floats (CLI:39-42)
    ref = oracle.align(src, tgt, oracle.make_params(max_neighbours=20, dof=np.inf, radius=3.0, source_filter_size=leaf,
                                                    target_filter_size=leaf), oracle.make_options(inner_kind=1))
    assert abs(len(hist) - ref.n_outer) <= 1
    t = np.array([float(v) for v in hist[-1][0].split(",")])
    assert np.linalg.norm(t - ref.transformation[:3, 3]) < 2e-4


@pytest.mark.gpu
def test_cli_unlimited_neighbours_and_its_failure_exit(cli, oracle, tmp_path):
    """`-m 0` reaches pcl's "every target within the radius" (CLI:43-44 -> registration.cc:74-75): served while no row needs
    128 neighbours or more; a row that would stops the program with the failure exit code instead of a truncated answer."""
    src, tgt, _ = synth.config1_plane_sphere(seed=5, n_plane=1500, n_sphere=1000)
    write_pcd(tmp_path / "s.pcd", src, "binary")
    write_pcd(tmp_path / "t.pcd", tgt, "binary")
    r = _run([cli, "s.pcd", "t.pcd", "-m", "0", "-r", "0.7", "-v"], tmp_path)
    assert r.returncode == 0, r.stderr
    hist = re.findall(r"^T: (.*?) \|\|\| R: (.*)$", r.stdout, flags=re.M)
    ref = oracle.align(src, tgt, oracle.make_params(max_neighbours=0, dof=5.0, radius=0.7), oracle.make_options(inner_kind=1))
    assert abs(len(hist) - ref.n_outer) <= 1
    t = np.array([float(v) for v in hist[-1][0].split(",")])
    assert np.linalg.norm(t - ref.transformation[:3, 3]) < 2e-4
    r = _run([cli, "s.pcd", "t.pcd", "-m", "0", "-r", "3"], tmp_path)
    assert r.returncode == 1 and "128" in (r.stderr + r.stdout)

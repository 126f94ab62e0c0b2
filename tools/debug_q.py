"""Debug aid: association after n_iter outer iterations, saved per search mode (PPCR_SEARCH_QUEUED), then compared.
    PPCR_SEARCH_QUEUED=0 python tools/debug_q.py run a c5 6 ; PPCR_SEARCH_QUEUED=1 python tools/debug_q.py run b c5 6 ; python tools/debug_q.py cmp a b"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out")
if sys.argv[1] == "run":
    import bench
    from probabilistic_point_clouds_registration_b200 import capi
    tag, w, n_iter = sys.argv[2], sys.argv[3], int(sys.argv[4])
    src, tgt = bench.make_pair(w, 0)
    prm = dict(bench.WORKLOADS[w]["params"])
    with capi.Registration(src, tgt, capi.make_params(n_iter=n_iter, **prm)) as reg:
        reg.align()
        idx, cnt = reg.association()
        st = reg.iteration_stats()
        inc = reg.increment_history()
        if len(sys.argv) > 5:
            np.save(os.path.join(out, "moved_src.npy"), reg.filtered_source())
    np.savez(os.path.join(out, f"assoc_{tag}.npz"), idx=idx, cnt=cnt, K=np.array([s["n_correspondences"] for s in st]), inc=inc)
    print(tag, "K per iteration", [s["n_correspondences"] for s in st])
else:
    a = np.load(os.path.join(out, f"assoc_{sys.argv[2]}.npz")); b = np.load(os.path.join(out, f"assoc_{sys.argv[3]}.npz"))
    print("K equal:", np.array_equal(a["K"], b["K"]), (a["K"] - b["K"]).tolist())
    bad = np.nonzero(a["cnt"] != b["cnt"])[0]
    print("rows with different count:", bad[:20], len(bad))
    sa = np.sort(np.where(np.arange(a["idx"].shape[1])[None, :] < a["cnt"][:, None], a["idx"], -1), axis=1)
    sb = np.sort(np.where(np.arange(b["idx"].shape[1])[None, :] < b["cnt"][:, None], b["idx"], -1), axis=1)
    diff = np.nonzero((sa != sb).any(axis=1))[0]
    print("rows with different sets:", diff[:20], len(diff))
    for r in diff[:5]:
        print(r, a["cnt"][r], b["cnt"][r], sa[r].tolist(), sb[r].tolist())

    if len(diff) and len(sys.argv) > 4:
        import bench
        from probabilistic_point_clouds_registration_b200 import synth
        w = sys.argv[4]
        src, tgt = bench.make_pair(w, 0)
        prm = bench.WORKLOADS[w]["params"]
        cur = src.copy()
        for T in a["inc"][:-1]:  # the last search ran before the last increment was applied
            cur = synth.apply_T_like_pcl(cur, T)
        dev = np.load(os.path.join(out, "moved_src.npy"))  # the device's own moved cloud (run with one iteration less)
        print("replayed cloud == device cloud:", np.array_equal(cur[:, :3], dev[:, :3]), np.abs(cur[:, :3] - dev[:, :3]).max())
        cur = dev
        r2 = np.float32(prm["radius"] * prm["radius"])
        m = prm["max_neighbours"]
        t = tgt[:, :3]
        for r in diff[:8]:
            q = cur[r, :3]
            dx, dy, dz = q[0] - t[:, 0], q[1] - t[:, 1], q[2] - t[:, 2]
            d2 = (dx * dx + dy * dy) + dz * dz
            assert d2.dtype == np.float32
            inside = np.nonzero(d2 < r2)[0]
            order = inside[np.lexsort((inside, d2[inside]))][:m]
            truth = np.sort(order)
            extra = [j for j in sb[r] if j >= 0 and j not in truth.tolist()] + [j for j in sa[r] if j >= 0 and j not in truth.tolist()]
            print("   q", q.tolist(), "r2", float(r2), "extras", [(int(j), float(d2[j]), t[j].tolist()) for j in extra])
            print("row", r, "truth", truth.tolist(), "heap ok" if np.array_equal(truth, sa[r][sa[r] >= 0]) else "heap WRONG",
                  "queued ok" if np.array_equal(truth, sb[r][sb[r] >= 0]) else "queued WRONG")

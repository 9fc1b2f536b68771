"""Where a visit-kernel launch spends its time: per-CTA clock stamps (MGCFD_VISIT_DEBUG=1, k_visit<.., DBG>) of one launch on each level
shape of a workload, reduced to medians over the CTAs.
usage: visit_timeline.py [c2|c3s] [level ...]     (environment: MGCFD_VISIT_K / MGCFD_VISIT_R / MGCFD_VISIT_RESIDENT as for the solver)"""
import json
import os
import sys

os.environ["MGCFD_VISIT_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import mgcfd_b200 as M

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, dims, variant, _ = bench.WORKLOADS[wl]
levels = [int(x) for x in sys.argv[2:]] or list(range(len(dims)))
GHZ = 1.965
for l in levels:
    mesh = M.Mesh.generate(kind, [dims[l]], mesh_variant=variant)
    s = M.Solver.from_mesh(mesh, use_graph=False, visit=True)
    vi = s.visit_info(0)
    if not vi["visit"]:
        print(json.dumps({"level": l, "visit": 0}))
        continue
    s.run_cycles(3)
    d = s.visit_debug()[:vi["ctas"]].astype(np.float64)
    K = vi["supers_per_cta"]
    Q = 3 * K
    us = lambda a, b: float(np.median(d[:, b] - d[:, a])) / GHZ / 1e3
    out = {"level": l, "nodes": int(np.prod(dims[l])) * (6 if kind == 2 else 1), "cfg": {k: int(v) for k, v in vi.items()}, "premin": os.environ.get("MGCFD_PREMIN", "1"),
           "total_us": us(0, 59), "prologue_to_wait_us": us(0, 1), "min_dt_and_barrier0_arrive_us": us(1, 2),
           "sum_ring_wait_us(thread0)": float(np.median(d[:, 56])) / GHZ / 1e3, "sum_edge_rounds_us(thread0)": float(np.median(d[:, 57])) / GHZ / 1e3,
           "sum_boundary_update_us(thread0)": float(np.median(d[:, 58])) / GHZ / 1e3, "sum_produce_us(thread0)": float(np.median(d[:, 62])) / GHZ / 1e3,
           "sum_tile_prologue_us(thread0)": float(np.median(d[:, 63])) / GHZ / 1e3, "iterations": []}
    for q in range(min(Q, 13)):
        b = 3 + 4 * q
        nxt = (3 + 4 * (q + 1)) if q + 1 < min(Q, 13) else 59
        out["iterations"].append({"q": q, "stage": q // K, "wait_records_us": round(us(b, b + 1), 2), "tiles_us": round(us(b + 1, b + 2), 2),
                                  "cta_sync_us": round(us(b + 2, b + 3), 2), "barrier_and_issue_us": round(us(b + 3, nxt), 2)})
    # spread over CTAs of the total
    tot = (d[:, 59] - d[:, 0]) / GHZ / 1e3
    out["total_us_min_max"] = [float(tot.min()), float(tot.max())]
    print(json.dumps(out), flush=True)
    s.close()

"""GPU: how the keyframe front-end and a back-to-back pose-graph solve share one GPU (SM budget sweep)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from omniswarm_b200 import synth, host, lib
L = lib.load()
W, H = 640, 480
comp, mean = synth.pca_matrices(0)
fe = host.KeyframeFrontend(synth.flatten_sp_weights(synth.superpoint_weights(0)), comp, mean,
                           synth.flatten_nv_weights(synth.netvlad_weights(0)), width=W, height=H, db_capacity=8192)
up = torch.from_numpy(np.stack([synth.image(d, H, W) for d in range(4)])).cuda()
dn = torch.from_numpy(np.stack([synth.image(10 + d, H, W) for d in range(4)])).cuda()
rec = torch.zeros(lib.RECORD_BYTES, dtype=torch.uint8, device="cuda"); res = torch.zeros(lib.RESULT_BYTES, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def kf(i):
    fe.extract(up.data_ptr(), dn.data_ptr(), i, rec.data_ptr(), st, device_images=True)
    fe.ingest(rec.data_ptr(), 1, -1, st); fe.query(rec.data_ptr(), res.data_ptr(), st)
g = synth.pose_graph(5, 100, seed=0)
sv = host.PoseGraphSolver(1024, 8192)
sv.graph_clear(); sv.graph_add_nodes(g["init"], g["fixed"]); sv.graph_add_factors(g["ftype"], g["ia"], g["ib"], g["payload"], g["huber"])
sv.solve_resident()
for i in range(5): kf(i)
fe.finish(st)
def run(budget, with_solver, n=80):
    L.osb_set_sm_budget(budget)
    stop = [False]; cnt = [0]
    def loop():
        while not stop[0]:
            sv.graph_set_poses(0, g["init"]); sv.solve_resident(); cnt[0] += 1
    th = threading.Thread(target=loop) if with_solver else None
    if th: th.start()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        kf(100 + i)
        if (i & 7) == 7: fe.finish(st)
    fe.finish(st); dt = time.perf_counter() - t0
    stop[0] = True
    if th: th.join()
    L.osb_set_sm_budget(0)
    return n / dt, cnt[0] / dt
for budget, ws in ((0, False), (132, False), (0, True), (132, True), (128, True), (116, True)):
    k, s = run(budget, ws)
    print(f"budget {budget:3d} solver {'on ' if ws else 'off'}: {k:7.1f} keyframes/s, {s:6.1f} solves/s", flush=True)

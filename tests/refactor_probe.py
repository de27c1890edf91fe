"""GPU probe (not a test): the C2 bench batch on device-resident inputs under several refactor periods (tier 1).
Prints LP/s, inversions, pivots and the largest relative difference of optF / x against the default period. With the
robust option the warm variant of the kernel runs (simplex_wave_reg_warm). Measured: the period does not matter (one
inversion per LP = the initial basis; 8 LPs of 4096 reach pivot 100), the warm variant is 3 % slower (720 vs 744 k LP/s)."""
import json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench as B
import gomilp_b200 as gm
gm.init(0)
dev = torch.device("cuda", 0)
c, A, b = B.make_batch(0)
dc, dA, db = (torch.from_numpy(a).to(dev) for a in (c, A, b))
n, M, N = B.BATCH, B.M, B.N
st = torch.zeros(n, dtype=torch.int32, device=dev); F = torch.zeros(n, dtype=torch.float64, device=dev)
x = torch.zeros(n, N, dtype=torch.float64, device=dev); bs = torch.zeros(n, M, dtype=torch.int64, device=dev)
ss = torch.zeros(n, 8, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)
ref = None
for period, robust in ((0, False), (64, False), (128, False), (100000, False), (0, True), (0, False), (0, True)):
    gm.set_options(refactor_period=period, robust=robust)
    def step():
        gm.simplex_batch_device(n, dc.data_ptr(), dA.data_ptr(), db.data_ptr(), M, N, 0.0, st.data_ptr(), F.data_ptr(),
                                x.data_ptr(), bs.data_ptr(), ss.data_ptr(), stream.cuda_stream)
    for _ in range(3): step()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10): step()
        e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / 10
    s = ss.cpu().numpy(); Fh = F.cpu().numpy(); xh = x.cpu().numpy(); sth = st.cpu().numpy()
    if ref is None: ref = (Fh.copy(), xh.copy())
    rel = lambda a, r: float(np.max(np.abs(a - r) / np.maximum(1.0, np.abs(r))))
    print(json.dumps({"period": period, "robust": robust, "ms": ms, "lp_per_s": n / ms * 1e3, "ok": int((sth == 0).sum()), "pivots": int(s[:, :2].sum()),
                      "bland": int(s[:, 2].sum()), "inversions": int(s[:, 3].sum()), "dF": rel(Fh, ref[0]), "dx": rel(xh, ref[1])}), flush=True)
gm.set_options()

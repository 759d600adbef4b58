import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, time, numpy as np, torch
icp = importlib.import_module("3dtk_b200")
ctx = icp.Context(0)
n = 1_000_000
m = torch.from_numpy(icp.synth_scene(7, 42, n, 0.5)).pin_memory()
for rep in range(5):
    torch.cuda.synchronize(); t = time.perf_counter()
    s = icp.Scan.from_host_pointers(ctx, m.data_ptr(), None, n, 0.0, 25.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s.destroy(); torch.cuda.synchronize(); t2 = time.perf_counter()
    s2 = icp.Scan.from_host_pointers(ctx, m.data_ptr(), None, n, 4.778, 25.0)   # explicit cell edge: no trials
    torch.cuda.synchronize(); t3 = time.perf_counter(); s2.destroy()
    print("scan_create auto %.2f ms   destroy %.2f ms   explicit-cell %.2f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))

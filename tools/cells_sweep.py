"""cfg2 kNN (50k x 50k, k = 11) and 1M x 1M under the cell-grid knobs."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import sbp_env_b200 as sg
from sbp_env_b200 import _lib
from sbp_env_b200.runtime import tune  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for n in (50_000, 1_000_000):
    tree = sg.NodeStore(2, capacity=n)
    tree.extend(torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * torch.tensor([579., 581.], device=dev, dtype=torch.float64))
    X = tree.rows_device(0, n)
    ref = None
    for mode, dens, qpw in ((3, 150, 32), (3, 150, 28), (3, 150, 24), (3, 150, 20), (3, 150, 18), (3, 150, 16)):
        tune("knn_cells", mode); tune("knn_cell_density", dens); tune("knn_cells_qpw", qpw)
        gr = torch.cuda.CUDAGraph()
        tree.knn(X, 11, return_dist=False); torch.cuda.synchronize()
        with torch.cuda.graph(gr):
            out = tree.knn(X, 11, return_dist=False)
        ms = timeit(gr.replay)
        idx = out[0] if isinstance(out, tuple) else out
        if ref is None: ref = idx.clone()
        print(n, "mode", mode, "density", dens, "qpw", qpw, "ms %.4f" % ms, "same", bool(torch.equal(ref, idx)), flush=True)

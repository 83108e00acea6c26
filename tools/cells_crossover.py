"""Warp-per-query (mode 1) vs thread-per-query (mode 3) cell-grid kNN for small query batches."""
import sys, torch
sys.path.insert(0, "/root/repo")
import sbp_env_b200 as sg
from sbp_env_b200 import _lib
from sbp_env_b200.runtime import tune  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n = 50_000
scale = torch.tensor([579., 581.], device=dev, dtype=torch.float64)
tree = sg.NodeStore(2, capacity=n)
tree.extend(torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale)
for m in (256, 512, 1024, 2048, 4096, 8192, 16384):
    X = torch.rand((m, 2), generator=g, device=dev, dtype=torch.float64) * scale
    row = []
    for mode in (1, 3, 0):
        tune("knn_cells", mode)
        gr = torch.cuda.CUDAGraph()
        tree.knn(X, 11, return_dist=False); torch.cuda.synchronize()
        with torch.cuda.graph(gr):
            tree.knn(X, 11, return_dist=False)
        row.append(timeit(gr.replay))
    print(f"m={m:6d}  warp {row[0]*1e3:7.1f} us   thread {row[1]*1e3:7.1f} us   brute force {row[2]*1e3:7.1f} us", flush=True)
tune("knn_cells", 2)

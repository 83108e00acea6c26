import sys; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import torch, numpy as np, sbp_env_b200 as sg
from sbp_env_b200 import _lib
lib=_lib.load()
g=torch.Generator(device="cuda"); g.manual_seed(0)
n=50000
tree=sg.NodeStore(2,capacity=n); tree.extend(torch.rand((n,2),generator=g,device="cuda",dtype=torch.float64)*580)
X=tree.rows_device(0,n)
def t(reps=10):
    tree.knn(X,11,return_dist=False); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): tree.knn(X,11,return_dist=False)
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
for S in (0,1,2,3,4):
    lib.sbp_tune(b"knn_chunks",S); print("chunks",S,"ms",t())
lib.sbp_tune(b"knn_chunks",0)
for seed,pf in ((1,1),(0,1),(1,0),(0,0)):
    lib.sbp_tune(b"knn_seed",seed); lib.sbp_tune(b"knn_prefilter",pf); print("seed",seed,"prefilter",pf,"ms",t())

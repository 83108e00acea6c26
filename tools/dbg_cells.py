import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import sbp_env_b200 as sg
from sbp_env_b200 import _lib
import c_oracle as orc
n, layout = 2048, "duplicates"
rng = np.random.default_rng(n + len(layout))
m, k = 1300, 11
poses = np.floor(rng.random((n, 2)) * 9) * 64.0
X = np.concatenate([poses[rng.integers(0, n, m // 2)] + rng.standard_normal((m // 2, 2)) * 0.3,
                    rng.random((m - m // 2, 2)) * 900 - 150])
t = sg.NodeStore(2); t.extend(poses)
lib = _lib.load()
for kk in (32, 21, 11):
    ei, ed = orc.knn(poses, X, kk)
    for cells, limit in ((1, 1024), (1, 0), (1, 1 << 30), (0, 1024)):
        lib.sbp_tune(b"knn_cells", cells); lib.sbp_tune(b"knn_cells_limit", limit)
        idx, dist = t.knn(X, kk)
        bad = np.nonzero((idx != ei).any(1))[0]
        print(kk, cells, limit, "bad rows", len(bad), bad[:10])
        if len(bad) and cells == 1 and limit == 1024:
            q = bad[0]
            print(" q", X[q], "\n got", idx[q], "\n exp", ei[q], "\n gd", dist[q][-6:], "\n ed", ed[q][-6:])
            print(" pos got", poses[idx[q][-4:]], "pos exp", poses[ei[q][-4:]])

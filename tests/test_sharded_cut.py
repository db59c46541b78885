"""The rank-based global Stage-1 cut of the sharded engine (W - 1 binary searches) against the general two-sort form, on random tied lists."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infidex_b200.dist import ShardedSearchEngine as S
def test_global_cut_forms_agree():
    rng = np.random.default_rng(3)
    for W, nq, K in ((2, 50, 16), (4, 200, 40), (8, 100, 64), (3, 30, 1)):
        lists = []
        for r in range(W):
            n = rng.integers(0, K + 1, nq); key = np.zeros((nq, K), np.int64); sc = np.zeros((nq, K), np.float32)
            for q in range(nq):
                m = n[q]; s_ = np.sort(rng.choice(np.array([0.5, 1.0, 1.5, 2.25, 3.0, 7.5], np.float32), m))[::-1]; k_ = rng.choice(10**6, m, replace=False) * W + r
                # list order: score desc, key asc
                o = np.lexsort((k_, -s_)); key[q, :m] = k_[o]; sc[q, :m] = s_[o]
            lists.append((torch.from_numpy(key).view(-1), torch.from_numpy(sc).view(-1), torch.from_numpy(n.astype(np.int32))))
        for me in range(W):
            e = S.__new__(S); e.torch = torch; e.world = W; e.rank = me
            ks, ss, ns = [l[0] for l in lists], [l[1] for l in lists], [l[2] for l in lists]
            a = e._global_cut(ks, ss, ns, nq, K); b = e._global_cut_sort(ks, ss, ns, nq, K)
            # rows with n == 0 everywhere: gmax 0 in both
            assert torch.equal(a[0], b[0]), (W, me, (a[0] != b[0]).nonzero()[:5])
            assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), (W, me)

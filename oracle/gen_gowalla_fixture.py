"""TEST INFRASTRUCTURE — packs the reference's bundled gowalla data into tests/golden/gowalla.npz.

train.txt is missing from the reference mount, but data/gowalla/s_adj_mat.npz (the un-normalised
bipartite adjacency the reference itself saved) is present, so the train set is exactly
R = A[:29858, 29858:].  test.txt is bundled.  Item ids fit uint16.
Runs only in the build container.
"""
import os
import numpy as np
import scipy.sparse as sp

REF = '/root/reference/LightGCN_work/data/gowalla'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'gowalla.npz')

A = sp.load_npz(os.path.join(REF, 's_adj_mat.npz')).tocsr()
nu, ni = 29858, 40981
assert A.shape == (nu + ni, nu + ni) and A.nnz == 1620256
R = A[:nu, nu:].tocsr()
R.sort_indices()
assert R.nnz == 810128 and np.all(R.data == 1.0)
test_users, test_indptr, test_items = [], [0], []
with open(os.path.join(REF, 'test.txt')) as f:
    for line in f:
        cols = line.split()
        if len(cols) < 2:
            continue
        test_users.append(int(cols[0]))
        test_items.extend(int(x) for x in cols[1:])
        test_indptr.append(len(test_items))
np.savez_compressed(OUT, n_users=nu, m_items=ni,
                    train_indptr=R.indptr.astype(np.int32), train_items=R.indices.astype(np.uint16),
                    test_users=np.array(test_users, dtype=np.int32), test_indptr=np.array(test_indptr, dtype=np.int32),
                    test_items=np.array(test_items, dtype=np.uint16),
                    kat_precision=0.0001875544, kat_recall=0.0005374941, kat_ndcg=0.00040836)
print(OUT, os.path.getsize(OUT), 'bytes; test pairs', len(test_items), 'test users', len(test_users))

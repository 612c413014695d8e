"""TEST INFRASTRUCTURE — not part of the product.

CPU restatement of the reference's LightGCN hot path (saamiya225/Graph-and-sequential-
recommendation-systems, LightGCN_work/code).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
(graph-and-sequential-recommendation-systems_b200/) never does.

Pinning status: the reference has no tests, golden vectors or fixtures of its own (SURVEY.md §4,
§8c).  The oracle is pinned against (a) outputs of the reference itself, imported from
/root/reference in the build container by oracle/gen_golden.py and committed as tests/golden/*.npz,
and (b) the one known answer recoverable from the author's run artefacts: the step-0 gowalla
evaluation (Precision@20 0.0001875544, Recall@20 0.0005374941, NDCG@20 0.00040836), reproduced in
tests/test_oracle.py::test_gowalla_step0_kat.
"""

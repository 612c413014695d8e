"""GPU: every kernel of liblgcn_b200.so against the oracle, called through the C ABI wrappers.

Tolerances: integer/index outputs bit-exact; fp32 outputs within 1e-5 norm-wise relative
(rel_err = max |a-b| / (|b| + rms(b))) unless stated otherwise.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope='module')
def lg():
    import lgcn_b200
    assert lgcn_b200.ops.device_info()['cc'] >= 100, "these kernels are built for sm_100a"
    return lgcn_b200


@pytest.fixture(scope='module')
def orc():
    from oracle import lightgcn_oracle
    return lightgcn_oracle


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build(lg, tu, ti, nu, ni, seg_len=128):
    return lg.ops.csr_build(dev(np.asarray(tu, np.int64)), dev(np.asarray(ti, np.int64)), nu, ni, seg_len=seg_len)


def random_edges(rng, nu, ni, E, dup=0):
    u = rng.integers(0, nu, E); i = rng.integers(0, ni, E)
    if dup:
        u = np.concatenate([u, u[:dup]]); i = np.concatenate([i, i[:dup]])
    return u.astype(np.int64), i.astype(np.int64)


# ------------------------------------------------------------------------------------ K4
def test_csr_build_matches_reference_output(lg, orc, golden):
    nu, ni = int(golden['n_users']), int(golden['m_items'])
    g = build(lg, golden['train_user'], golden['train_item'], nu, ni)
    N = nu + ni
    ref_indptr = np.concatenate([[0], np.cumsum(np.bincount(golden['adj_row'], minlength=N))])
    assert np.array_equal(g.indptr.cpu().numpy(), ref_indptr)                 # bit-exact vs the reference
    assert np.array_equal(g.indices.cpu().numpy(), golden['adj_col'])
    assert np.max(np.abs(g.vals.cpu().numpy() - golden['adj_val']) / golden['adj_val']) < 3e-7
    o_indptr, o_indices, o_vals, o_deg, o_dinv = orc.build_norm_adj(golden['train_user'], golden['train_item'], nu, ni)
    assert np.array_equal(g.vals.cpu().numpy(), o_vals)                       # bit-exact vs the oracle's rounding rule
    assert np.array_equal(g.deg.cpu().numpy(), o_deg) and np.array_equal(g.dinv.cpu().numpy(), o_dinv)


@pytest.mark.parametrize("nu,ni,E,dup", [(1, 1, 1, 0), (5, 7, 0, 0), (50, 30, 400, 40), (3000, 2000, 60000, 500), (70000, 300, 5000, 0)])
def test_csr_build_random(lg, orc, nu, ni, E, dup):
    rng = np.random.default_rng(nu * 7 + E)
    tu, ti = random_edges(rng, nu, ni, E, dup)
    g = build(lg, tu, ti, nu, ni)
    o_indptr, o_indices, o_vals, o_deg, o_dinv = orc.build_norm_adj(tu, ti, nu, ni)
    assert g.nnz == o_indices.size
    assert np.array_equal(g.indptr.cpu().numpy(), o_indptr)
    assert np.array_equal(g.indices.cpu().numpy(), o_indices)
    assert np.array_equal(g.deg.cpu().numpy(), o_deg)
    assert np.array_equal(g.vals.cpu().numpy(), o_vals)


def test_csr_build_rejects_out_of_range_ids(lg):
    with pytest.raises(RuntimeError):
        build(lg, [0, 5], [0, 1], 3, 4)


def test_csr_build_full_size_properties(lg, orc):
    gr = lg.synth.make_graph('yelp2018')
    nu, ni = gr['n_users'], gr['m_items']
    g = build(lg, gr['train_user'], gr['train_item'], nu, ni)
    indptr, indices = g.indptr.cpu().numpy().astype(np.int64), g.indices.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(nu + ni), np.diff(indptr))
    key = rows * (nu + ni) + indices
    assert np.all(np.diff(key) > 0)                                           # sorted, duplicate-free
    assert np.array_equal(np.sort(indices * (nu + ni) + rows), key)           # structurally symmetric
    assert g.nnz == 2 * np.unique(gr['train_user'] * ni + gr['train_item']).size
    o = orc.build_norm_adj(gr['train_user'], gr['train_item'], nu, ni)
    assert np.array_equal(indptr, o[0]) and np.array_equal(indices, o[1]) and np.array_equal(g.vals.cpu().numpy(), o[2])
    # bipartite: user rows only reference item columns and vice versa
    assert indices[:indptr[nu]].min() >= nu and indices[indptr[nu]:].max() < nu


@pytest.mark.parametrize("parts,chunked", [(1, False), (2, False), (3, True), (8, True)])
def test_row_block_build_equals_rows_of_the_full_build(lg, orc, parts, chunked):
    """Memory-partitioned row partition: a rank assembles ONLY its rows (lgcn_degree_accumulate / lgcn_csr_rows_emit /
    lgcn_csr_rows_finish) — each block must be the corresponding rows of lgcn_csr_build's CSR and of the oracle's, bit for bit,
    also when the edge stream arrives in chunks, with duplicate pairs, zero-degree nodes and an empty block."""
    rng = np.random.default_rng(parts)
    nu, ni = 700, 900
    tu, ti = random_edges(rng, nu, ni - 40, 9000, dup=300)            # last 40 items: zero-degree rows
    full = build(lg, tu, ti, nu, ni)
    o_indptr, o_indices, o_vals, o_deg, o_dinv = orc.build_norm_adj(tu, ti, nu, ni)
    dtu, dti = dev(tu), dev(ti)

    def chunks():
        if not chunked:
            yield dtu, dti
        else:
            for lo in range(0, tu.size, 2500):
                yield dtu[lo:lo + 2500].contiguous(), dti[lo:lo + 2500].contiguous()
    b = lg.ops.RowBlockBuilder(nu, ni, chunks)
    assert np.array_equal(b.deg.cpu().numpy(), o_deg) and np.array_equal(b.dinv.cpu().numpy(), full.dinv.cpu().numpy())
    bounds = lg.engine.balanced_row_bounds(b.cost_prefix, parts)
    if parts == 3:
        bounds[1] = bounds[2]                                           # an empty block in the middle
    N = nu + ni
    assert bounds[0] == 0 and bounds[-1] == N
    tot = 0
    for r0, r1 in zip(bounds, bounds[1:]):
        g = b.build(r0, r1)
        ip = g.indptr.cpu().numpy()
        assert g.n_rows == r1 - r0 and g.n_cols == N
        assert np.array_equal(ip, o_indptr[r0:r1 + 1] - o_indptr[r0])
        assert np.array_equal(g.indices.cpu().numpy(), o_indices[o_indptr[r0]:o_indptr[r1]])
        assert np.array_equal(g.vals.cpu().numpy(), o_vals[o_indptr[r0]:o_indptr[r1]])
        tot += g.nnz
        if r1 > r0:                                                     # and the block drives K1 like the view of the full CSR
            X = torch.randn(N, 64, device='cuda')
            ya, yb = torch.empty(r1 - r0, 64, device='cuda'), torch.empty(r1 - r0, 64, device='cuda')
            lg.ops.spmm(g, X, ya); lg.ops.spmm(full.rows(r0, r1), X, yb)
            assert torch.equal(ya, yb)
    assert tot == full.nnz


@pytest.mark.parametrize("seg_len,slab_rows", [(4096, 97), (4096, 700), (16, 211), (128, 5000)])
def test_spmm_column_slab_blocking_equals_unblocked(lg, orc, seg_len, slab_rows):
    """K1 with column-slab blocking (graphs whose gathered table exceeds L2): one launch per slab, running sums carried
    between launches.  A row's summation order is unchanged, so with no row segmented inside a slab (seg_len 4096) the
    result is BIT-identical to the unblocked kernel; with segmentation it agrees to rounding.  Covers empty rows, slabs
    without any entry, hub rows, the mean epilogue with own-row addends, and the Adam epilogue."""
    rng = np.random.default_rng(seg_len + slab_rows)
    nu, ni, d = 900, 1300, 64
    tu, ti = random_edges(rng, nu, ni - 300, 20000, dup=50)            # 300 items never seen: empty rows and an empty slab
    tu = np.concatenate([tu, np.zeros(1500, np.int64)]); ti = np.concatenate([ti, rng.integers(0, ni - 300, 1500)])   # a hub user
    N = nu + ni
    g0 = build(lg, tu, ti, nu, ni, seg_len=seg_len)
    g1 = build(lg, tu, ti, nu, ni, seg_len=seg_len)
    plans = g1.block_plans(d, slab_bytes=slab_rows * d * 4)
    assert plans is not None and len(plans) >= 1 and len(plans) <= -(-N // slab_rows)
    X = torch.randn(N, d, device='cuda'); Z1 = torch.randn(N, d, device='cuda'); Z2 = torch.randn(N, d, device='cuda')
    Y0, Y1 = torch.full((N, d), 7.0, device='cuda'), torch.full((N, d), -3.0, device='cuda')
    lg.ops.spmm(g0, X, Y0, 0.25, 0.5, [Z1, Z2]); lg.ops.spmm(g1, X, Y1, 0.25, 0.5, [Z1, Z2])
    if seg_len == 4096:
        assert torch.equal(Y0, Y1)
    else:
        assert rel_err(Y1.cpu().numpy(), Y0.cpu().numpy()) < 2e-6
    ref = 0.25 * orc.spmm(*(a.cpu().numpy() for a in (g0.indptr, g0.indices, g0.vals)), X.cpu().numpy().astype(np.float64)) + 0.5 * (Z1 + Z2).cpu().numpy()
    assert rel_err(Y1.cpu().numpy(), ref) < TOL
    # row mask: masked-out rows are left untouched by every slab launch
    mask = torch.from_numpy(rng.integers(0, 2**31, (N + 31) // 32).astype(np.int32)).cuda()
    Ym0, Ym1 = torch.full((N, d), 7.0, device='cuda'), torch.full((N, d), 7.0, device='cuda')
    lg.ops.spmm(g0, X, Ym0, row_mask=mask); lg.ops.spmm(g1, X, Ym1, row_mask=mask)
    assert torch.equal(Ym0, Ym1) if seg_len == 4096 else rel_err(Ym1.cpu().numpy(), Ym0.cpu().numpy()) < 2e-6
    # Adam epilogue
    sc = lg.ops.adam_scalars(X.device, 1e-3); lg.ops.adam_tick(sc)
    P0 = torch.randn(N, d, device='cuda'); M0 = torch.zeros_like(P0); V0 = torch.zeros_like(P0)
    P1, M1, V1 = P0.clone(), M0.clone(), V0.clone()
    lg.ops.spmm_adam(g0, X, P0, M0, V0, sc, 1.0, 0.25, [Z1]); lg.ops.spmm_adam(g1, X, P1, M1, V1, sc, 1.0, 0.25, [Z1])
    if seg_len == 4096:
        assert torch.equal(P0, P1) and torch.equal(M0, M1) and torch.equal(V0, V1)
    else:
        assert rel_err(M1.cpu().numpy(), M0.cpu().numpy()) < 2e-6 and rel_err(P1.cpu().numpy(), P0.cpu().numpy()) < 1e-4


def test_spmm_l2_hinted_gathers_are_bit_identical(lg):
    """K1 with L2 eviction hints (hot columns evict-last, the rest evict-first: a cache policy, not arithmetic)."""
    rng = np.random.default_rng(3)
    nu, ni, d = 800, 1100, 64
    tu, ti = random_edges(rng, nu, ni, 15000, dup=20)
    g0, g1 = build(lg, tu, ti, nu, ni), build(lg, tu, ti, nu, ni)
    g1.col_weight = g1.deg.to(torch.int32)
    h = g1.hint_indices(d, hot_bytes=100 * d * 4)
    assert h is not None and int((h < 0).sum()) > 0 and torch.equal(h & 0x7fffffff, g1.indices)
    N = nu + ni
    X = torch.randn(N, d, device='cuda'); Z = torch.randn(N, d, device='cuda')
    Y0, Y1 = torch.empty(N, d, device='cuda'), torch.empty(N, d, device='cuda')
    lg.ops.spmm(g0, X, Y0, 0.5, 0.25, [Z]); lg.ops.spmm(g1, X, Y1, 0.5, 0.25, [Z])
    assert torch.equal(Y0, Y1)


def test_training_step_with_blocked_graph_matches_unblocked(lg):
    """The whole fused step (forward, K2, backward, Adam) with K1 column-blocked == unblocked, bit for bit (tiny graph, no
    row long enough to be segmented inside a slab), through the reference-facing API in deterministic mode."""
    g = load_golden('tiny')
    nu = int(g['n_users'])
    res = []
    for blocked in (False, True):
        cfg = dict(lg.world.config)
        cfg.update(latent_dim_rec=int(g['d']), lightGCN_n_layers=int(g['L']), bpr_batch_size=len(g['users']), decay=float(g['decay']),
                   lr=float(g['lr']), deterministic=True, spmm_seg_len=4096)
        ds = lg.InteractionDataset(nu, int(g['m_items']), g['train_user'], g['train_item'], g['test_user'], g['test_item'], config=cfg)
        m = lg.LightGCN(cfg, ds)
        with torch.no_grad():
            m.embedding_user.weight.copy_(torch.from_numpy(g['E0'][:nu])); m.embedding_item.weight.copy_(torch.from_numpy(g['E0'][nu:]))
        if blocked:
            assert ds.getCSRGraph().block_plans(int(g['d']), slab_bytes=150 * int(g['d']) * 4) is not None
        bpr = lg.utils.BPRLoss(m, cfg)
        losses = [bpr.stageOne(*(torch.from_numpy(np.roll(g[k], s * 17)).long() for k in ('users', 'pos', 'neg'))) for s in range(3)]
        res.append((losses, torch.cat([m.embedding_user.weight, m.embedding_item.weight]).detach().cpu().numpy()))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    assert rel_err(res[1][1], g['params_after'][2]) < 1e-4


def test_rank_metrics_is_bitwise_repeatable_and_split_invariant(lg):
    """lgcn_rank_metrics adds the per-row values in a fixed order: the same bits on every call — the property the multi-GPU
    Test relies on (ranked lists gathered from the ranks, then this kernel on every rank)."""
    rng = np.random.default_rng(0)
    Bt, k, ni = 5000, 20, 3000
    topk = np.stack([rng.choice(ni, k, replace=False) for _ in range(Bt)]).astype(np.int64)
    lens = rng.integers(0, 30, Bt)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    items = np.concatenate([np.sort(rng.choice(ni, l, replace=False)) for l in lens]).astype(np.int32)
    a = lg.ops.rank_metrics(dev(topk), dev(indptr), dev(items), [5, 20]).cpu().numpy()
    for _ in range(5):
        assert np.array_equal(lg.ops.rank_metrics(dev(topk), dev(indptr), dev(items), [5, 20]).cpu().numpy(), a)


def test_coo_to_csr(lg, orc):
    g = load_golden('tiny')
    N = int(g['n_users']) + int(g['m_items'])
    c = lg.ops.coo_to_csr(dev(g['adj_row'].astype(np.int64)), dev(g['adj_col'].astype(np.int64)), dev(g['adj_val']), N, N)
    ref_indptr = np.concatenate([[0], np.cumsum(np.bincount(g['adj_row'], minlength=N))])
    assert np.array_equal(c.indptr.cpu().numpy(), ref_indptr) and np.array_equal(c.indices.cpu().numpy(), g['adj_col'])


# ------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("d", [8, 16, 32, 64, 128, 256])
@pytest.mark.parametrize("seg_len", [128, 8, 2])
def test_spmm_vs_oracle(lg, orc, d, seg_len):
    rng = np.random.default_rng(d + seg_len)
    nu, ni = 700, 900
    tu, ti = random_edges(rng, nu, ni, 9000, 100)
    tu[:800] = 3; ti[:800] = rng.permutation(ni)[:800]          # a hub row with 800 non-zeros
    g = build(lg, tu, ti, nu, ni, seg_len=seg_len)
    if seg_len <= 8:
        assert g.n_segs > 0 and g.n_long > 0
    if seg_len == 2:
        assert g.max_item_len == 2                       # 400 parts for the hub row: below the 2048-part cap
    N = nu + ni
    X = rng.normal(0, 0.1, (N, d)).astype(np.float32)
    Z1 = rng.normal(0, 0.1, (N, d)).astype(np.float32); Z2 = rng.normal(0, 0.1, (N, d)).astype(np.float32)
    indptr, indices, vals = (t.cpu().numpy() for t in (g.indptr, g.indices, g.vals))
    AX = orc.spmm_scipy(indptr, indices, vals.astype(np.float64), X.astype(np.float64))
    Y = torch.empty((N, d), device='cuda')
    lg.ops.spmm(g, dev(X), Y)
    assert rel_err(Y.cpu().numpy(), AX) < TOL
    lg.ops.spmm(g, dev(X), Y, alpha=0.25, beta=0.5, zs=[dev(Z1), dev(Z2)])
    assert rel_err(Y.cpu().numpy(), 0.25 * AX + 0.5 * (Z1.astype(np.float64) + Z2)) < TOL
    # second launch on the same plan: arrival counters must have reset themselves
    lg.ops.spmm(g, dev(X), Y)
    assert rel_err(Y.cpu().numpy(), AX) < TOL


def test_spmm_hub_row_hits_the_part_cap(lg, orc):
    """A row is cut into at most 2048 parts (the last-arriving part adds the partials in part order, several loads in
    flight): with seg_len 2 a 5000-non-zero row gets 3-element parts, and the result still matches the oracle."""
    rng = np.random.default_rng(77)
    nu, ni, d = 40, 6000, 64
    tu = np.concatenate([np.full(5000, 7), rng.integers(0, nu, 3000)]).astype(np.int64)
    ti = np.concatenate([rng.permutation(ni)[:5000], rng.integers(0, ni, 3000)]).astype(np.int64)
    g = build(lg, tu, ti, nu, ni, seg_len=2)
    assert g.max_item_len == 3 and g.n_long > 0
    N = nu + ni
    X = rng.normal(0, 0.1, (N, d)).astype(np.float32)
    indptr, indices, vals = (t.cpu().numpy() for t in (g.indptr, g.indices, g.vals))
    ref = orc.spmm_scipy(indptr, indices, vals.astype(np.float64), X.astype(np.float64))
    Y = torch.empty((N, d), device='cuda')
    lg.ops.spmm(g, dev(X), Y)
    assert rel_err(Y.cpu().numpy(), ref) < TOL
    Y2 = torch.empty_like(Y); lg.ops.spmm(g, dev(X), Y2)
    assert torch.equal(Y, Y2)                                    # fixed summation order: bitwise repeatable


def test_spmm_empty_rows_and_zero_degree_nodes(lg, orc):
    nu, ni, d = 40, 30, 64
    tu = np.array([0, 0, 5], np.int64); ti = np.array([1, 2, 1], np.int64)
    g = build(lg, tu, ti, nu, ni)
    X = np.random.default_rng(0).normal(0, 1, (nu + ni, d)).astype(np.float32)
    Y = torch.full((nu + ni, d), 7.0, device='cuda')
    lg.ops.spmm(g, dev(X), Y)
    indptr, indices, vals = (t.cpu().numpy() for t in (g.indptr, g.indices, g.vals))
    exp = orc.spmm(indptr, indices, vals.astype(np.float64), X.astype(np.float64))
    assert rel_err(Y.cpu().numpy(), exp) < TOL
    assert float(Y[10].abs().max()) == 0.0                # empty row is written (zeros), not skipped


def test_spmm_full_size_properties(lg):
    """amazon-book shape (hub user ~10 k): linearity and symmetry <y, A x> = <A y, x>."""
    gr = lg.synth.make_graph('amazon-book')
    g = build(lg, gr['train_user'], gr['train_item'], gr['n_users'], gr['m_items'])
    assert g.n_long > 0
    N, d = g.n_rows, 64
    gen = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn((N, d), device='cuda', generator=gen); y = torch.randn((N, d), device='cuda', generator=gen)
    Ax, Ay, Axy = (torch.empty_like(x) for _ in range(3))
    lg.ops.spmm(g, x, Ax); lg.ops.spmm(g, y, Ay); lg.ops.spmm(g, (x + y).contiguous(), Axy)
    assert rel_err(Axy.cpu().numpy(), (Ax + Ay).cpu().numpy()) < TOL
    lhs = (y.double() * Ax.double()).sum().item(); rhs = (Ay.double() * x.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * (abs(lhs) + (y.double().norm() * Ax.double().norm()).item() * 1e-2)
    # against torch's own CSR SpMM (cuSPARSE) on the same matrix
    ref = torch.sparse.mm(g.to_torch_sparse_csr(), x)
    assert rel_err(Ax.cpu().numpy(), ref.cpu().numpy()) < TOL
    # the plan only changes scheduling: without it (one item per row, natural order) every row that was
    # not segmented is bit-identical, segmented rows agree to rounding
    g.use_plan = False
    Ax2 = torch.empty_like(x); lg.ops.spmm(g, x, Ax2)
    g.use_plan = True
    short = (torch.diff(g.indptr) <= g.seg_len)
    assert torch.equal(Ax[short], Ax2[short])
    assert rel_err(Ax.cpu().numpy(), Ax2.cpu().numpy()) < TOL
    # every tuning variant of the d=64 kernel computes the same thing
    lib = lg._lib.load()
    for v in range(1, 10):
        lib.lgcn_debug_spmm_variant(v)
        Av = torch.empty_like(x); lg.ops.spmm(g, x, Av)
        lib.lgcn_debug_spmm_variant(0)
        assert rel_err(Av.cpu().numpy(), Ax.cpu().numpy()) < 1e-6, v


def _bits(mask_np, n):
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    idx = np.flatnonzero(mask_np)
    np.bitwise_or.at(words, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return torch.from_numpy(words.view(np.int32)).cuda()


@pytest.mark.parametrize("d", [64, 128])
def test_spmm_row_and_column_masks(lg, orc, d):
    rng = np.random.default_rng(11 + d)
    nu, ni = 600, 500
    tu, ti = random_edges(rng, nu, ni, 9000)
    tu[:400] = 2; ti[:400] = rng.permutation(ni)[:400]
    g = build(lg, tu, ti, nu, ni, seg_len=32)
    N = nu + ni
    indptr, indices, vals = (t.cpu().numpy() for t in (g.indptr, g.indices, g.vals))
    # column mask: X is zero outside the marked rows -> identical to the unmasked product, bit for bit
    nz_rows = rng.random(N) < 0.1
    X = rng.normal(0, 0.1, (N, d)).astype(np.float32); X[~nz_rows] = 0
    Z = rng.normal(0, 0.1, (N, d)).astype(np.float32)
    Yf = torch.empty((N, d), device='cuda'); Ym = torch.empty((N, d), device='cuda')
    lg.ops.spmm(g, dev(X), Yf, 0.25, 0.5, [dev(Z)])
    lg.ops.spmm(g, dev(X), Ym, 0.25, 0.5, [dev(Z)], col_mask=_bits(nz_rows, N))
    assert torch.equal(Yf, Ym)
    assert rel_err(Ym.cpu().numpy(), 0.25 * orc.spmm_scipy(indptr, indices, vals.astype(np.float64), X.astype(np.float64)) + 0.5 * Z) < TOL
    # row mask: marked rows equal the full product, the others are left untouched
    rows = rng.random(N) < 0.3; rows[2] = True
    X2 = rng.normal(0, 0.1, (N, d)).astype(np.float32)
    lg.ops.spmm(g, dev(X2), Yf)
    Ym.fill_(7.0)
    lg.ops.spmm(g, dev(X2), Ym, row_mask=_bits(rows, N))
    r = torch.from_numpy(rows).cuda()
    assert torch.equal(Ym[r], Yf[r]) and bool((Ym[~r] == 7.0).all())
    # a second full launch must not be disturbed by the skipped segments (arrival counters untouched)
    lg.ops.spmm(g, dev(X2), Ym)
    assert torch.equal(Ym, Yf)


def test_batch_masks(lg):
    rng = np.random.default_rng(5)
    nu, ni, B = 400, 300, 100
    tu, ti = random_edges(rng, nu, ni, 5000)
    g = build(lg, tu, ti, nu, ni)
    N = nu + ni
    users = rng.integers(0, nu, B); pos = rng.integers(0, ni, B); neg = rng.integers(0, ni, B)
    ctl = torch.tensor([3, B, B + 3, 0], dtype=torch.int32, device='cuda')
    pad = lambda a: dev(np.concatenate([np.zeros(3, np.int64), a.astype(np.int64), np.zeros(50, np.int64)]))
    m0 = torch.full(((N + 31) // 32,), -1, dtype=torch.int32, device='cuda'); m1 = m0.clone()
    lg.ops.batch_masks(pad(users), pad(pos), pad(neg), 128, ctl, nu, g, m0, m1)
    rows = np.unique(np.concatenate([users, nu + pos, nu + neg]))
    indptr, indices = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    nb = np.unique(np.concatenate([indices[indptr[r]:indptr[r + 1]] for r in rows] + [rows]))
    unpack = lambda m: np.flatnonzero(np.unpackbits(m.cpu().numpy().view(np.uint8), bitorder='little')[:N])
    assert np.array_equal(unpack(m0), rows) and np.array_equal(unpack(m1), nb)


# ------------------------------------------------------------------------------------ Adam
def test_adam_matches_torch_and_oracle(lg, orc):
    rng = np.random.default_rng(0)
    n, d = 500, 64
    p0 = rng.normal(0, 0.1, (n, d)).astype(np.float32)
    P = dev(p0); M = torch.zeros_like(P); V = torch.zeros_like(P)
    tp = torch.nn.Parameter(dev(p0).clone()); opt = torch.optim.Adam([tp], lr=1e-3)
    sc = lg.ops.adam_scalars(P.device, 1e-3)
    p64, m64, v64 = p0.astype(np.float64), np.zeros((n, d)), np.zeros((n, d))
    for t in range(1, 6):
        gnp = rng.normal(0, 1e-3, (n, d)).astype(np.float32); gnp[::7] = 0
        G = dev(gnp)
        lg.ops.adam_tick(sc); lg.ops.adam(P, M, V, G, sc)
        tp.grad = G.clone(); opt.step()
        p64, m64, v64 = orc.adam_step(p64, m64, v64, gnp.astype(np.float64), t)
    assert lg.ops.adam_step_count(sc) == 5
    assert rel_err(P.cpu().numpy(), tp.detach().cpu().numpy()) < 2e-6
    assert rel_err(P.cpu().numpy(), p64) < 2e-6
    assert rel_err(M.cpu().numpy(), m64) < TOL and rel_err(V.cpu().numpy(), v64) < TOL


@pytest.mark.parametrize("d", [8, 16, 32, 64])
def test_spmm_adam_epilogue_equals_spmm_then_adam(lg, d):
    rng = np.random.default_rng(2)
    nu, ni = 300, 200
    tu, ti = random_edges(rng, nu, ni, 4000)
    g = build(lg, tu, ti, nu, ni, seg_len=16)
    N = nu + ni
    X = dev(rng.normal(0, 1e-3, (N, d)).astype(np.float32)); Z = dev(rng.normal(0, 1e-3, (N, d)).astype(np.float32))
    p0 = rng.normal(0, 0.1, (N, d)).astype(np.float32)
    m0 = rng.normal(0, 1e-3, (N, d)).astype(np.float32); v0 = np.abs(rng.normal(0, 1e-6, (N, d))).astype(np.float32)
    sc = lg.ops.adam_scalars(X.device, 1e-3, step=3)
    Pa, Ma, Va = dev(p0), dev(m0), dev(v0)
    Ygrad = torch.empty((N, d), device='cuda')
    lg.ops.spmm_adam(g, X, Pa, Ma, Va, sc, alpha=1.0, beta=0.25, zs=[Z], Y=Ygrad)
    Pb, Mb, Vb = dev(p0), dev(m0), dev(v0)
    Gd = torch.empty((N, d), device='cuda')
    lg.ops.spmm(g, X, Gd, 1.0, 0.25, [Z]); lg.ops.adam(Pb, Mb, Vb, Gd, sc)
    assert torch.equal(Ygrad, Gd)
    assert torch.equal(Pa, Pb) and torch.equal(Ma, Mb) and torch.equal(Va, Vb)


# ------------------------------------------------------------------------------------ K2
def _bpr_case(rng, nu, ni, d, B):
    out = rng.normal(0, 0.3, (nu + ni, d)).astype(np.float32)
    users = rng.integers(0, nu, B); pos = rng.integers(0, ni, B); neg = rng.integers(0, ni, B)
    users[: B // 8] = users[0]; pos[: B // 8] = pos[min(1, B - 1)]          # heavy row collisions (atomics / owner-computes)
    return out, users.astype(np.int64), pos.astype(np.int64), neg.astype(np.int64)


@pytest.mark.parametrize("d", [8, 16, 64, 256])
@pytest.mark.parametrize("B", [1, 37, 2048])
@pytest.mark.parametrize("deterministic", [False, True])
def test_bpr_vs_oracle(lg, orc, d, B, deterministic):
    rng = np.random.default_rng(d * B + deterministic)
    nu, ni, decay = 500, 800, 1e-4
    out, users, pos, neg = _bpr_case(rng, nu, ni, d, B)
    bpr, reg, Gb, Gr = orc.bpr_loss(out, users, pos, neg, nu)
    B_cap = 2048
    ws = lg.ops.bpr_workspace(B_cap, d, 'cuda')
    pad = lambda a: dev(np.concatenate([np.zeros(5, np.int64), a, np.zeros(B_cap, np.int64)]))   # window at offset 5
    ctl = torch.tensor([5, B, 5 + B, 0], dtype=torch.int32, device='cuda')
    loss = torch.zeros(4, device='cuda'); G = torch.zeros((nu + ni, d), device='cuda')
    o = dev(out)
    for rep in range(2):                                       # 2nd call: arrival counter reset, running sum
        G.zero_()
        lg.ops.bpr_fwd_bwd(o, pad(users), pad(pos), pad(neg), B_cap, ctl, nu, ni, 0.0, decay, 1.0, decay, loss, G, ws,
                           deterministic=deterministic)
    l = loss.cpu().numpy()
    assert abs(l[0] - bpr) < TOL * abs(bpr) and abs(l[1] - reg) < TOL * abs(reg)
    assert abs(l[2] - (bpr + decay * reg)) < TOL * abs(bpr) and abs(l[3] - 2 * l[2]) < 1e-6
    assert rel_err(G.cpu().numpy(), Gb + decay * Gr) < TOL
    # separate coefficients (generic autograd path) and forward-only call
    G.zero_()
    lg.ops.bpr_fwd_bwd(o, pad(users), pad(pos), pad(neg), B_cap, ctl, nu, ni, 1.0 / B, 0.0, 0.5, 2.0, loss, G, ws,
                       deterministic=deterministic)
    assert rel_err(G.cpu().numpy(), 0.5 * Gb + 2.0 * Gr) < TOL
    lg.ops.bpr_fwd_bwd(o, pad(users), pad(pos), pad(neg), B_cap, ctl, nu, ni, 1.0 / B, 0.0, 0.0, 0.0, loss, None, ws)
    assert abs(loss[0].item() - bpr) < TOL * abs(bpr)
    # clear_rows zeroes exactly the touched rows
    lg.ops.bpr_clear_rows(G, pad(users), pad(pos), pad(neg), B_cap, ctl, nu)
    assert float(G.abs().max()) == 0.0


def test_bpr_owner_range_and_global_norm(lg, orc):
    """Row-partition filter (own_begin/own_end) and data-parallel normalisation via ctl[3]."""
    rng = np.random.default_rng(9)
    nu, ni, d, B = 300, 400, 64, 512
    out, users, pos, neg = _bpr_case(rng, nu, ni, d, B)
    _, _, Gb, Gr = orc.bpr_loss(out, users, pos, neg, nu)
    ws = lg.ops.bpr_workspace(B, d, 'cuda')
    o = dev(out); loss = torch.zeros(4, device='cuda')
    G = torch.zeros((nu + ni, d), device='cuda')
    ctl = torch.tensor([0, B, B, 0], dtype=torch.int32, device='cuda')
    lg.ops.bpr_fwd_bwd(o, dev(users), dev(pos), dev(neg), B, ctl, nu, ni, 0.0, 0.0, 1.0, 0.0, loss, G, ws, own=(100, 450))
    exp = Gb.copy(); exp[:100] = 0; exp[450:] = 0
    assert rel_err(G.cpu().numpy(), exp) < TOL
    # two shards normalised by the global batch add up to the full-batch gradient
    G.zero_(); h = B // 2
    for lo in (0, h):
        ctl = torch.tensor([lo, h, B, B], dtype=torch.int32, device='cuda')
        lg.ops.bpr_fwd_bwd(o, dev(users), dev(pos), dev(neg), B, ctl, nu, ni, 0.0, 0.0, 1.0, 0.0, loss, G, ws)
    assert rel_err(G.cpu().numpy(), Gb) < TOL


def test_bpr_deterministic_mode_is_bitwise_repeatable(lg):
    rng = np.random.default_rng(4)
    nu, ni, d, B = 200, 100, 64, 2048                        # tiny item set -> many collisions
    out, users, pos, neg = _bpr_case(rng, nu, ni, d, B)
    ws = lg.ops.bpr_workspace(B, d, 'cuda'); o = dev(out); loss = torch.zeros(4, device='cuda')
    ctl = torch.tensor([0, B, B, 0], dtype=torch.int32, device='cuda')
    outs = []
    for _ in range(3):
        G = torch.zeros((nu + ni, d), device='cuda')
        lg.ops.bpr_fwd_bwd(o, dev(users), dev(pos), dev(neg), B, ctl, nu, ni, 0.0, 1e-4, 1.0, 1e-4, loss, G, ws, deterministic=True)
        outs.append((G.clone(), loss[:3].clone()))
    assert all(torch.equal(outs[0][0], x[0]) and torch.equal(outs[0][1], x[1]) for x in outs[1:])


def test_batch_advance(lg):
    ctl = torch.tensor([0, 0, 5000, 0], dtype=torch.int32, device='cuda')
    seen = []
    for _ in range(4):
        lg.ops.batch_advance(ctl, 2048); seen.append(ctl.cpu().tolist()[:2])
    assert seen == [[0, 2048], [2048, 2048], [4096, 904], [5000, 0]]


# ------------------------------------------------------------------------------------ K3
@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
def test_score_topk_bit_exact(lg, orc, d):
    rng = np.random.default_rng(d)
    nu, ni, k = 333, 1301, 20
    tu, ti = random_edges(rng, nu, ni, 12000)
    tu[:900] = 7; ti[:900] = rng.permutation(ni)[:900]         # user 7 has 900 train items
    g = build(lg, tu, ti, nu, ni)
    out = rng.normal(0, 0.1, (nu + ni, d)).astype(np.float32)
    out[nu + 5] = out[nu + 9]                                   # exact score ties between items 5 and 9
    indptr, indices = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    o = dev(out)
    for users in (None, rng.permutation(nu)[:150].astype(np.int64), np.array([7, 7, 3], np.int64)):
        exp_idx, exp_val = orc.score_topk_exact(out[:nu], out[nu:], users, k, indptr, indices, nu)
        idx, val = lg.ops.score_topk(o[:nu], o[nu:], None if users is None else dev(users), k, g.indptr, g.indices, nu)
        assert np.array_equal(idx.cpu().numpy(), exp_idx)
        assert np.array_equal(val.cpu().numpy(), exp_val)       # scores are bit-identical too
    exp_idx, exp_val = orc.score_topk_exact(out[:nu], out[nu:], None, k)
    idx, val = lg.ops.score_topk(o[:nu], o[nu:], None, k)       # no mask
    assert np.array_equal(idx.cpu().numpy(), exp_idx) and np.array_equal(val.cpu().numpy(), exp_val)


def test_score_topk_k_exceeds_unmasked_items(lg, orc):
    g0 = load_golden('edge')
    nu, ni = int(g0['n_users']), int(g0['m_items'])
    g = build(lg, g0['train_user'], g0['train_item'], nu, ni)
    out = g0['out_after']; o = dev(out)
    for k in (1, 5, 20, 60):
        exp_idx, exp_val = orc.score_topk_exact(out[:nu], out[nu:], None, k, g.indptr.cpu().numpy(), g.indices.cpu().numpy(), nu)
        idx, val = lg.ops.score_topk(o[:nu], o[nu:], None, k, g.indptr, g.indices, nu)
        assert np.array_equal(idx.cpu().numpy(), exp_idx) and np.array_equal(val.cpu().numpy(), exp_val)
    assert (val[0] == -1024.0).sum().item() == 45              # user 0: 45 masked train items among all 60
    with pytest.raises(RuntimeError):
        lg.ops.score_topk(o[:nu], o[nu:], None, 61)


@pytest.mark.parametrize("scale", [0.1, 1.0])
def test_score_topk_tensor_core_is_bit_identical_to_exact(lg, scale):
    """tcgen05 filter + exact rescoring + certificate == the exact kernel, indices AND scores."""
    rng = np.random.default_rng(int(scale * 10))
    nu, ni, d, k = 700, 20011, 64, 20
    tu, ti = random_edges(rng, nu, ni, 60000)
    tu[:ni - 11] = 5; ti[:ni - 11] = rng.permutation(ni)[:ni - 11]    # user 5: only 11 unmasked items -> must be flagged
    g = build(lg, tu, ti, nu, ni)
    out = (scale * rng.normal(0, 1, (nu + ni, d))).astype(np.float32)
    out[nu + 7] = out[nu + 3]                                   # exact ties
    o = dev(out)
    for users in (None, dev(rng.permutation(nu)[:130].astype(np.int64)), dev(np.array([5, 6, 5], np.int64))):
        ei, ev = lg.ops.score_topk(o[:nu], o[nu:], users, k, g.indptr, g.indices, nu)
        ti_, tv, redone = lg.ops.score_topk_tc(o[:nu], o[nu:], users, k, g.indptr, g.indices, nu)
        assert torch.equal(ti_, ei) and torch.equal(tv, ev)
        n_rows = nu if users is None else users.numel()
        assert redone < max(3, n_rows // 20), redone           # the certificate almost always holds
    ei, ev = lg.ops.score_topk(o[:nu], o[nu:], None, k)
    ti_, tv, _ = lg.ops.score_topk_tc(o[:nu], o[nu:], None, k)   # no mask
    assert torch.equal(ti_, ei) and torch.equal(tv, ev)


@pytest.mark.parametrize("shape", ["random", "trained-like"])
def test_score_topk_tensor_core_directly_against_the_c_oracle(lg, orc, shape):
    """The tcgen05 path against oracle/c/score_topk_ref.c DIRECTLY (not through the exact kernel): 200 sampled users over
    a 40,981-item table (gowalla's size), masks from a real-shaped train CSR, indices AND scores bit for bit."""
    rng = np.random.default_rng(5 if shape == "random" else 6)
    gr = lg.synth.make_graph('gowalla', seed=2020)
    nu, ni, d, k = gr['n_users'], gr['m_items'], 64, 20
    g = build(lg, gr['train_user'], gr['train_item'], nu, ni)
    if shape == "random":
        U = rng.normal(0, 0.1, (nu, d)).astype(np.float32); V = rng.normal(0, 0.1, (ni, d)).astype(np.float32)
    else:       # popular items (small ids) have large norms and score high for everybody; a few exact ties
        common = rng.normal(0, 1, d).astype(np.float32)
        pop = (1.0 / (1.0 + np.arange(ni) / 300.0)).astype(np.float32)[:, None]
        V = (0.05 * rng.normal(0, 1, (ni, d)) + pop * common).astype(np.float32)
        U = (0.1 * rng.normal(0, 1, (nu, d)) + 0.3 * common).astype(np.float32)
        V[11] = V[4]; V[2000] = V[1999]
    users = np.sort(rng.choice(nu, 200, replace=False)).astype(np.int64)
    indptr, indices = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    exp_idx, exp_val = orc.score_topk_exact(U, V, users, k, indptr, indices, nu)
    ti_, tv, redone = lg.ops.score_topk_tc(dev(U), dev(V), dev(users), k, g.indptr, g.indices, nu)
    assert np.array_equal(ti_.cpu().numpy(), exp_idx) and np.array_equal(tv.cpu().numpy(), exp_val)
    assert redone <= 10, redone                                  # and it is the tensor-core path that answered, not the fallback


def test_score_topk_tensor_core_popular_items_with_adjacent_ids(lg):
    """What a trained model looks like: every user's best items are the popular ones, popular items have small ADJACENT
    ids and large norms, and a user's train items are among its best.  In id order that put most candidates into a few
    64-item blocks of one split (70 % of the rows overflowed on real gowalla); the interleaved item layout must keep the
    tensor-core path certified — and bit-identical — here."""
    rng = np.random.default_rng(21)
    nu, ni, d, k = 900, 24001, 64, 20
    c = rng.normal(0, 1, d).astype(np.float32); c /= np.linalg.norm(c)
    pop = (4.0 / (1.0 + np.arange(ni) / 150.0)).astype(np.float32)                 # popularity decays with the id
    V = pop[:, None] * (c[None, :] + 0.35 * rng.normal(0, 1, (ni, d)).astype(np.float32) / np.sqrt(d)) \
        + 0.05 * rng.normal(0, 1, (ni, d)).astype(np.float32)
    U = (1.0 + rng.random(nu).astype(np.float32))[:, None] * (c[None, :] + 0.5 * rng.normal(0, 1, (nu, d)).astype(np.float32) / np.sqrt(d))
    deg = rng.integers(5, 120, nu)
    tu = np.repeat(np.arange(nu), deg)
    ti = np.concatenate([rng.choice(2000, n, replace=False, p=(pop[:2000] / pop[:2000].sum())) for n in deg])    # train items: popular ones
    g = build(lg, tu.astype(np.int64), ti.astype(np.int64), nu, ni)
    o = dev(np.concatenate([U, V]).astype(np.float32))
    ei, ev = lg.ops.score_topk(o[:nu], o[nu:], None, k, g.indptr, g.indices, nu)
    ti_, tv, redone = lg.ops.score_topk_tc(o[:nu], o[nu:], None, k, g.indptr, g.indices, nu)
    assert torch.equal(ti_, ei) and torch.equal(tv, ev)
    assert redone <= nu // 50, redone
    assert int(ei[:, 0].max()) < 2000                          # the best items really are the small ids


@pytest.mark.parametrize("ni,k", [(16384, 20), (16385, 1), (16511, 24), (16512, 7), (33000, 20), (40961, 20)])
def test_score_topk_tensor_core_item_count_edges(lg, ni, k):
    """The interleaved layout has 128*T - m_items holes in its last block and a residue order that depends on T: smallest
    table (T = 128, no holes), one item more (127 holes), last tile nearly full / exactly full, T even and odd; k = 1 and
    k = 24; user batches that are not a multiple of anything.  Always bit-identical to the exact kernel, holes never
    appear as items, nearly all rows certified."""
    rng = np.random.default_rng(ni + k)
    nu = 333
    tu, ti = random_edges(rng, nu, ni, 12000)
    ti[:400] = ni - 1 - rng.integers(0, 300, 400)              # train items among the last ids (the block next to the holes)
    g = build(lg, tu, ti, nu, ni)
    out = rng.normal(0, 0.2, (nu + ni, 64)).astype(np.float32)
    out[nu + ni - 200:] *= 3.0                                  # ... and the best-scoring items are the last ids
    o = dev(out)
    for users in (None, dev(rng.permutation(nu)[:77].astype(np.int64))):
        ei, ev = lg.ops.score_topk(o[:nu], o[nu:], users, k, g.indptr, g.indices, nu)
        ti_, tv, redone = lg.ops.score_topk_tc(o[:nu], o[nu:], users, k, g.indptr, g.indices, nu)
        assert torch.equal(ti_, ei) and torch.equal(tv, ev)
        assert int(ti_.max()) < ni and int(ti_.min()) >= 0
        assert redone <= max(2, (nu if users is None else 77) // 25), redone


def test_score_topk_tensor_core_too_few_tiles_still_exact(lg):
    """Item tables below 16 k items (fewer than 128 tiles) are not taken by the tensor-core path at all: the call is
    served by the exact kernel, whatever min_items says."""
    rng = np.random.default_rng(11)
    nu, ni, k = 300, 3001, 20
    tu, ti = random_edges(rng, nu, ni, 9000)
    g = build(lg, tu, ti, nu, ni)
    o = dev(rng.normal(0, 0.3, (nu + ni, 64)).astype(np.float32))
    ei, ev = lg.ops.score_topk(o[:nu], o[nu:], None, k, g.indptr, g.indices, nu)
    ti_, tv, redone = lg.ops.score_topk_tc(o[:nu], o[nu:], None, k, g.indptr, g.indices, nu, min_items=0)
    assert torch.equal(ti_, ei) and torch.equal(tv, ev)
    assert redone == nu


def test_score_topk_tensor_core_small_item_table(lg):
    g0 = load_golden('edge')                                    # 60 items (< one 256-item tile), d = 32 -> exact path
    nu, ni = int(g0['n_users']), int(g0['m_items'])
    rng = np.random.default_rng(1)
    out = rng.normal(0, 0.3, (nu + ni, 64)).astype(np.float32); o = dev(out)
    g = build(lg, g0['train_user'], g0['train_item'], nu, ni)
    ei, ev = lg.ops.score_topk(o[:nu], o[nu:], None, 20, g.indptr, g.indices, nu)
    ti_, tv, redone = lg.ops.score_topk_tc(o[:nu], o[nu:], None, 20, g.indptr, g.indices, nu)
    assert torch.equal(ti_, ei) and torch.equal(tv, ev)
    assert redone >= 1                                          # user 0 has 15 unmasked items < k


def test_score_dense_bit_exact(lg, orc):
    rng = np.random.default_rng(3)
    nu, ni, d = 130, 700, 64
    out = rng.normal(0, 0.1, (nu + ni, d)).astype(np.float32); o = dev(out)
    users = rng.integers(0, nu, 100).astype(np.int64)
    got = lg.ops.score_dense(o[:nu], o[nu:], dev(users))
    assert np.array_equal(got.cpu().numpy(), orc.score_dense_exact(out[:nu], out[nu:], users))


@pytest.mark.parametrize("Bt,ni", [(100, 40981), (333, 16384 + 77), (2048, 38048), (5, 130)])
def test_score_dense_tensor_core_3xtf32_accuracy(lg, orc, Bt, ni):
    """getUsersRating on tcgen05 (3xTF32: hi/lo operand split, fp32 TMEM accumulation) against the fp64 product:
    |err| <= 2e-6 |u||v| (fp32 class; plain TF32 would be ~1e-3), every cell of the Bt x M matrix written, ragged sizes."""
    rng = np.random.default_rng(Bt + ni)
    nu, d = max(Bt, 400), 64
    U = (rng.normal(0, 1, (nu, d)) * rng.uniform(0.01, 3.0, (nu, 1))).astype(np.float32)
    V = (rng.normal(0, 1, (ni, d)) * rng.uniform(0.01, 3.0, (ni, 1))).astype(np.float32)
    users = rng.permutation(nu)[:Bt].astype(np.int64)
    got = lg.ops.score_dense_tc(dev(U), dev(V), dev(users)).cpu().numpy()
    assert got.shape == (Bt, ni) and np.isfinite(got).all()
    ref = U[users].astype(np.float64) @ V.astype(np.float64).T
    bound = np.linalg.norm(U[users], axis=1)[:, None] * np.linalg.norm(V, axis=1)[None, :]
    assert np.max(np.abs(got - ref) / bound) < 2e-6
    exact = lg.ops.score_dense(dev(U), dev(V), dev(users)).cpu().numpy()          # the bit-exact FMA-chain kernel
    assert np.max(np.abs(got - exact) / bound) < 2e-6
    got2 = lg.ops.score_dense_tc(dev(U[:Bt]), dev(V), None).cpu().numpy()         # users = None: rows 0..Bt-1
    assert np.max(np.abs(got2 - U[:Bt].astype(np.float64) @ V.astype(np.float64).T) / (np.linalg.norm(U[:Bt], axis=1)[:, None] * np.linalg.norm(V, axis=1)[None, :])) < 2e-6


def test_score_topk_full_size_against_fp64(lg):
    """yelp2018 shape, all users: against fp64 scores the indices may differ only at near-ties."""
    gr = lg.synth.make_graph('yelp2018')
    nu, ni, d, k = gr['n_users'], gr['m_items'], 64, 20
    g = build(lg, gr['train_user'], gr['train_item'], nu, ni)
    gen = torch.Generator(device='cuda').manual_seed(0)
    out = (0.1 * torch.randn((nu + ni, d), device='cuda', generator=gen)).contiguous()
    idx, val = lg.ops.score_topk(out[:nu], out[nu:], None, k, g.indptr, g.indices, nu)
    assert bool((val[:, :-1] >= val[:, 1:]).all())             # sorted descending
    tidx, tval, redone = lg.ops.score_topk_tc(out[:nu], out[nu:], None, k, g.indptr, g.indices, nu)
    assert torch.equal(tidx, idx) and torch.equal(tval, val)    # tensor-core path: bit-identical at full size
    assert redone < nu // 100
    sub = torch.arange(0, nu, 37, device='cuda')
    S = out[:nu][sub].double() @ out[nu:].double().T
    indptr = g.indptr.cpu().numpy(); indices = g.indices.cpu().numpy()
    for b, u in enumerate(sub.cpu().tolist()):
        S[b, torch.from_numpy(indices[indptr[u]:indptr[u + 1]] - nu).cuda().long()] = -1024.0
    ref_val, ref_idx = torch.topk(S, k, dim=1)
    got = idx[sub]
    mism = got != ref_idx
    if mism.any():
        a = torch.gather(S, 1, got)[mism]; b = ref_val[mism]
        assert bool(((a - b).abs() <= 1e-6 * b.abs() + 1e-9).all())
    assert mism.float().mean().item() < 1e-3
    # masked items never appear
    rows = sub.cpu().numpy()
    for b in range(0, len(rows), 97):
        u = rows[b]
        assert np.intersect1d(got[b].cpu().numpy(), indices[indptr[u]:indptr[u + 1]] - nu).size == 0


def test_device_sampler_properties(lg):
    """K5 keeps the reference sampler's contract (per-user counts, positives in / negatives out of the row) and
    emits a permutation; the stream is counter-based, so parity is distributional, not bit-level."""
    gr = lg.synth.make_graph('tiny', seed=2)
    nu, ni = gr['n_users'], gr['m_items']
    g = build(lg, gr['train_user'], gr['train_item'], nu, ni)
    E = gr['train_user'].size
    S = lg.ops.sample_bpr(g, nu, ni, E, seed=2020, epoch=0).cpu().numpy()
    per = E // nu
    assert S.shape == (3, per * nu)
    assert np.array_equal(np.bincount(S[0], minlength=nu), np.full(nu, per))           # exactly per_user triples per user
    indptr, indices = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    rows = [set((indices[indptr[u]:indptr[u + 1]] - nu).tolist()) for u in range(nu)]
    assert all(p in rows[u] and n not in rows[u] and 0 <= n < ni for u, p, n in S.T[::7])
    assert not np.array_equal(S[0], np.sort(S[0]))                                      # shuffled, not in user order
    S2 = lg.ops.sample_bpr(g, nu, ni, E, seed=2020, epoch=0).cpu().numpy()
    S3 = lg.ops.sample_bpr(g, nu, ni, E, seed=2020, epoch=1).cpu().numpy()
    assert np.array_equal(S, S2) and not np.array_equal(S, S3)                          # reproducible, epoch-dependent
    # positives are uniform over the row: chi-square-ish check on the busiest user
    u = int(np.argmax(np.diff(indptr[:nu + 1])))
    big = np.concatenate([lg.ops.sample_bpr(g, nu, ni, E, seed=1, epoch=e).cpu().numpy()[:, :] for e in range(40)], axis=1)
    pu = big[1][big[0] == u]
    counts = np.array([np.sum(pu == it) for it in sorted(rows[u])])
    assert counts.min() > 0 and counts.max() < 4 * counts.mean() + 10
    # negatives are uniform over the complement
    nn_ = big[2]
    assert abs(nn_.mean() - (ni - 1) / 2) < 0.05 * ni


def test_device_sampler_refuses_users_without_train_items(lg):
    """A user with no train interaction would get a fabricated positive (item 0): the device sampler reports it and the
    wrapper raises, like the host sampler does for the same input."""
    tu = np.array([0, 0, 2, 2, 2], np.int64); ti = np.array([1, 3, 0, 2, 4], np.int64)       # user 1 has nothing
    g = build(lg, tu, ti, 3, 6)
    with pytest.raises(RuntimeError, match="no train item"):
        lg.ops.sample_bpr(g, 3, 6, 6, seed=1, epoch=0)
    S = lg.ops.sample_bpr(g, 3, 6, 6, seed=1, epoch=0, check=False)
    assert S.shape == (3, 6)


def test_rank_metrics_vs_oracle(lg, orc):
    rng = np.random.default_rng(5)
    n, m, kmax = 257, 400, 20
    gts = [rng.choice(m, size=int(rng.integers(1, 40)), replace=False) for _ in range(n)]
    topk = np.stack([rng.choice(m, size=kmax, replace=False) for _ in range(n)])
    for b in range(0, n, 3):
        topk[b, :3] = gts[b][:3] if len(gts[b]) >= 3 else topk[b, :3]
    indptr = np.concatenate([[0], np.cumsum([len(x) for x in gts])]).astype(np.int32)
    items = np.concatenate([np.sort(x) for x in gts]).astype(np.int32)
    ks = [5, 10, 20]
    sums = lg.ops.rank_metrics(dev(topk.astype(np.int64)), dev(indptr), dev(items), ks).cpu().numpy()
    exp = orc.metrics_at_k(topk, [x.tolist() for x in gts], ks)
    for j, name in enumerate(('precision', 'recall', 'ndcg')):
        assert np.allclose(sums[:, j] / n, exp[name], rtol=1e-12, atol=1e-14)

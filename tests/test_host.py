"""CPU: host-side logic and the C-ABI boundary (no device compute)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    import lgcn_b200 as lg
    header = open(os.path.join(ROOT, 'include', 'lgcn_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(lgcn_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 20
    lib = lg._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/lgcn_b200.h but not exported"
    assert declared == set(lg._lib.EXPORTED_SYMBOLS), declared ^ set(lg._lib.EXPORTED_SYMBOLS)
    assert lib.lgcn_abi_version() == 1


def _header_prototypes():
    """{name: [parameter declarations]} of every function prototype in include/lgcn_b200.h."""
    header = open(os.path.join(ROOT, 'include', 'lgcn_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    header = re.sub(r'typedef\s+struct\s*\{.*?\}\s*\w+\s*;', '', header, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(lgcn_[a-z0-9_]+)\s*\(([^()]*)\)\s*;', header):
        params = [p.strip() for p in m.group(2).split(',')]
        if params == ['void'] or params == ['']:
            params = []
        protos[m.group(1)] = params
    return protos


def test_ctypes_signatures_match_the_header_arity_and_kinds():
    """Every export's ctypes argtypes list has exactly the header's parameter count, and pointer / integer /
    floating kinds agree position by position (a missing argtype lets ctypes pass a Python int as a 32-bit C int)."""
    import ctypes
    import lgcn_b200 as lg
    protos = _header_prototypes()
    assert set(protos) == set(lg._lib.EXPORTED_SYMBOLS)
    lg._lib.load()
    for name, params in protos.items():
        _, argtypes = lg._lib._SIGNATURES[name]
        assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters, _lib.py declares {len(argtypes)}"
        for i, (decl, at) in enumerate(zip(params, argtypes)):
            is_ptr_decl = '*' in decl or 'lgcn_stream_t' in decl
            is_ptr_ct = at in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(at, 'contents')
            assert is_ptr_decl == is_ptr_ct, f"{name} arg {i}: '{decl}' vs {at}"
            if not is_ptr_decl:
                is_float_decl = bool(re.match(r'(float|double)\b', decl))
                assert is_float_decl == (at in (ctypes.c_float, ctypes.c_double)), f"{name} arg {i}: '{decl}' vs {at}"
                if not is_float_decl:
                    want64 = bool(re.match(r'(int64_t|uint64_t|size_t)\b', decl))
                    assert want64 == (ctypes.sizeof(at) == 8), f"{name} arg {i}: '{decl}' vs {at}"


def test_ctypes_structs_have_the_headers_layout(tmp_path):
    """The structs that cross the C ABI by pointer (plan, peers, Adam scalars, feature-partition exchange): sizeof and every
    offsetof of the ctypes mirrors in _lib.py equal what a C compiler makes of include/lgcn_b200.h."""
    import ctypes
    import shutil
    import subprocess
    import lgcn_b200  # noqa: F401  (registers the package under its importable name)
    from lgcn_b200 import _lib
    cc = shutil.which('gcc') or shutil.which('cc')
    if cc is None:
        pytest.skip("no C compiler")
    pairs = {'lgcn_spmm_plan_t': _lib.SpmmPlan, 'lgcn_spmm_peers_t': _lib.SpmmPeers, 'lgcn_adam_scalars_t': _lib.AdamScalars,
             'lgcn_bpr_feat_t': _lib.BprFeat}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lgcn_b200.h"', 'int main(void) {']
    for cname, st in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include')
    subprocess.run([cc, '-std=c99', '-I', inc, str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(pairs)
    for line in out:
        cname, *nums = line.split()
        st = pairs[cname]
        want = [ctypes.sizeof(st)] + [getattr(st, f).offset for f, _ in st._fields_]
        assert [int(x) for x in nums] == want, f"{cname}: header {nums} vs ctypes {want}"


def test_host_side_queries_without_gpu():
    import lgcn_b200 as lg
    lib = lg._lib.load()
    assert lib.lgcn_csr_build_workspace_bytes(1000, 10, 20) > 2 * 8 * 2000
    assert lib.lgcn_bpr_workspace_bytes(2048, 64) >= 16 + 2048 * 4
    assert lib.lgcn_score_topk_workspace_bytes(100, 1000, 20) >= 100 * 32 * 20 * 8
    # argument validation happens before any launch
    assert lib.lgcn_spmm_f32(None, None, None, 1, 64, None, None, 1.0, 0.0, None, 0, None, None, None, None, None) != 0
    assert b'null' in lib.lgcn_last_error()


def test_sampler_matches_reference_cpp_bit_exact():
    """Same glibc rand() stream and draw order as the reference's sources/sampling.cpp (golden made by
    compiling that file, oracle/gen_golden.py::sampler_golden)."""
    import ctypes
    import lgcn_b200 as lg
    g = load_golden('sampler')
    lib = lg._lib.load()
    lg.utils.sampler_seed(2020)
    nu, ni, tn = int(g['n_users']), int(g['m_items']), int(g['train_num'])
    out = np.empty(((tn // nu) * nu, 3), dtype=np.int32)
    indptr, items = np.ascontiguousarray(g['indptr']), np.ascontiguousarray(g['items'])
    rows = lib.lgcn_sample_negative(nu, ni, tn, indptr.ctypes.data_as(ctypes.c_void_p),
                                    items.ctypes.data_as(ctypes.c_void_p), 1, out.ctypes.data_as(ctypes.c_void_p))
    assert rows == out.shape[0]
    assert np.array_equal(out, g['S'])


def test_sampling_module_abi_matches_reference_cpp():
    """The whole ABI of the reference's pybind11 `sampling` module (code/sources/sampling.cpp:95-106): seed,
    sample_negative, randint, sample_negative_ByUser — one rand() stream, recorded from the reference's compiled file."""
    import lgcn_b200 as lg
    g = load_golden('sampler')
    nu, ni, tn = int(g['n_users']), int(g['m_items']), int(g['train_num'])
    all_pos = np.split(g['items'], g['indptr'][1:-1])
    lg.sampling.seed(2020)
    assert np.array_equal(lg.sampling.sample_negative(nu, ni, tn, all_pos, 1), g['S'])
    assert [lg.sampling.randint(1000) for _ in range(16)] == g['randints'].tolist()
    S2 = lg.sampling.sample_negative_ByUser(g['by_users'].tolist(), ni, all_pos, 2)
    assert S2.dtype == np.int32 and np.array_equal(S2, g['S_by_user'])
    with pytest.raises(RuntimeError):
        lg.sampling.sample_negative_ByUser([nu + 3], ni, all_pos, 1)


def test_sampler_generator_is_glibc_rand():
    """The library carries its own copy of glibc's rand() (TYPE_3 additive feedback generator): same stream as this machine's
    libc for several seeds, so the reference's sampler stream no longer depends on the platform's C library."""
    import ctypes
    import lgcn_b200 as lg
    lib = lg._lib.load()
    libc = ctypes.CDLL(None)
    libc.rand.restype = ctypes.c_int
    for seed in (0, 1, 2020, 123456789, 0xffffffff):
        libc.srand(ctypes.c_uint(seed)); lg.sampling.seed(seed)
        assert [libc.rand() % 1000003 for _ in range(3000)] == [lg.sampling.randint(1000003) for _ in range(3000)]


def test_epoch_prefetch_keeps_the_sampler_stream_exact():
    """Procedure.BPR_train_original draws the NEXT epoch's sample in the background.  Whatever is called next, the stream is
    the one the reference's module would have produced: the same dataset gets the prefetched sample (= what a direct call
    would have drawn), anything else rewinds the generator first."""
    import lgcn_b200 as lg
    ds = lg.synth.make_dataset('tiny')
    ds2 = lg.synth.make_dataset('tiny', seed=5)
    lg.utils.sampler_seed(11)
    ref = [lg.utils.UniformSample_original(ds) for _ in range(3)]                      # plain sequential draws
    ref_ri = lg.sampling.randint(1000)
    ref_other = lg.utils.UniformSample_original(ds2)
    lg.utils.sampler_seed(11)
    a = lg.utils.UniformSample_original(ds, prefetch_next=True)
    b = lg.utils.UniformSample_original(ds, prefetch_next=True)                        # served by the prefetch
    c = lg.utils.UniformSample_original(ds, prefetch_next=True)
    assert all(np.array_equal(x, y) for x, y in zip(ref, (a, b, c)))
    assert lg.sampling.randint(1000) == ref_ri                                         # pending prefetch rewound before the draw
    assert np.array_equal(lg.utils.UniformSample_original(ds2, prefetch_next=True), ref_other)
    lg.utils.sampler_seed(11)                                                          # reseeding drops a pending prefetch as well
    assert np.array_equal(lg.utils.UniformSample_original(ds), ref[0])


def test_oracle_c_sampler_matches_reference_cpp():
    """oracle/c/sampler_ref.c (what bench.py --impl reference draws its triples with) against the same recording."""
    from oracle import lightgcn_oracle as orc
    g = load_golden('sampler')
    nu, ni, tn = int(g['n_users']), int(g['m_items']), int(g['train_num'])
    assert np.array_equal(orc.sample_negative_ref(2020, nu, ni, tn, g['indptr'], g['items'], 1), g['S'])
    assert [orc.randint_ref(1000) for _ in range(16)] == g['randints'].tolist()
    assert np.array_equal(orc.sample_negative_by_user_ref(g['by_users'], ni, g['indptr'], g['items'], 2), g['S_by_user'])


def test_sampler_through_dataset_api():
    import lgcn_b200 as lg
    ds = lg.synth.make_dataset('tiny')
    lg.utils.sampler_seed(7)
    S = lg.utils.UniformSample_original(ds)
    per = ds.trainDataSize // ds.n_users
    assert S.shape == (ds.n_users * per, 3) and S.dtype == np.int32
    assert np.array_equal(S[:, 0], np.repeat(np.arange(ds.n_users), per))
    ap = ds.allPos
    for u, p, n in S[::37]:
        assert p in ap[u] and n not in ap[u] and 0 <= n < ds.m_items


def test_loader_parses_reference_format(tmp_path):
    import lgcn_b200 as lg
    (tmp_path / 'train.txt').write_text("0 1 2 3\n1 0\n\n2\n3 4 4\n")
    (tmp_path / 'test.txt').write_text("0 5\n1 2 3\n4\n")
    ds = lg.Loader(lg.world.config, path=str(tmp_path))
    assert (ds.n_users, ds.m_items) == (4, 6)            # max id over train+test, +1; item-less lines skipped
    assert ds.trainDataSize == 6 and ds.testDataSize == 3
    assert list(ds.testDict.keys()) == [0, 1] and ds.testDict[1] == [2, 3]
    assert [a.tolist() for a in ds.allPos] == [[1, 2, 3], [0], [], [4]]
    assert ds.users_D.tolist() == [3, 1, 1, 2] and ds.items_D[4] == 2     # duplicates counted, zero -> 1
    assert ds.getUserItemFeedback([0, 1], [1, 1]).tolist() == [1, 0]


def test_native_parser_matches_the_references_line_loop(tmp_path):
    """lgcn_parse_interactions against a restatement of the reference's parsing loop (code/dataloader.py:82-97):
    tabs, runs of blanks, CRLF, a last line without newline, item-less and empty lines; errors name the line."""
    import ctypes
    import lgcn_b200 as lg
    from lgcn_b200 import dataloader
    rng = np.random.default_rng(3)
    lines = []
    for u in rng.permutation(300):
        k = int(rng.integers(0, 6))
        sep = ['  ', ' ', '\t'][int(rng.integers(0, 3))]
        lines.append(sep.join([str(u)] + [str(int(x)) for x in rng.integers(0, 100000, k)]) + ['', ' ', '\r'][int(rng.integers(0, 3))])
    lines.insert(5, ''); lines.insert(9, '   ')
    text = '\n'.join(lines)                                    # no trailing newline
    f = tmp_path / 'train.txt'
    f.write_text(text)
    eu, ei = [], []
    for l in text.split('\n'):                                  # the reference's loop
        if not l.strip():
            continue
        cols = l.strip().split()
        items = [int(i) for i in cols[1:]]
        if not items:
            continue
        eu.extend([int(cols[0])] * len(items)); ei.extend(items)
    users, items = dataloader._parse_interactions(str(f))
    assert users.tolist() == eu and items.tolist() == ei
    lib = lg._lib.load()
    mu, mi = ctypes.c_int64(), ctypes.c_int64()
    n = lib.lgcn_parse_interactions(str(f).encode(), None, None, 0, ctypes.byref(mu), ctypes.byref(mi))
    assert n == len(eu) and mu.value == max(eu) and mi.value == max(ei)
    (tmp_path / 'bad.txt').write_text("0 1 2\n1 x7\n")
    with pytest.raises(RuntimeError, match="line 2"):
        dataloader._parse_interactions(str(tmp_path / 'bad.txt'))
    with pytest.raises(RuntimeError, match="cannot open"):
        dataloader._parse_interactions(str(tmp_path / 'missing.txt'))
    (tmp_path / 'empty.txt').write_text("")
    users, items = dataloader._parse_interactions(str(tmp_path / 'empty.txt'))
    assert users.size == 0 and items.size == 0


def test_test_csr_matches_testDict(monkeypatch):
    """The vectorised test CSR (what the device metric kernel reads) is testDict in key order with sorted items."""
    import torch
    import lgcn_b200 as lg
    monkeypatch.setattr(lg.world, 'device', torch.device('cpu'))
    g = lg.synth.make_graph('tiny', seed=9)
    perm = np.random.default_rng(0).permutation(g['test_user'].size)          # users interleaved, not grouped
    ds = lg.InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'][perm], g['test_item'][perm])
    users, indptr, items = (t.numpy() for t in ds.test_csr())
    td = ds.testDict
    assert users.tolist() == list(td.keys())
    for j, u in enumerate(users):
        assert items[indptr[j]:indptr[j + 1]].tolist() == sorted(td[int(u)])


def test_loader_round_trip_of_synthetic_graph(tmp_path):
    import lgcn_b200 as lg
    g = lg.synth.make_graph('tiny', seed=5)
    lg.synth.write_txt(g, str(tmp_path))
    ds = lg.Loader(lg.world.config, path=str(tmp_path))
    a = lg.InteractionDataset(g['n_users'], g['m_items'], g['train_user'], g['train_item'], g['test_user'], g['test_item'])
    assert ds.trainDataSize == a.trainDataSize
    assert all(np.array_equal(x, y) for x, y in zip(ds.allPos, a.allPos))
    assert {k: sorted(v) for k, v in ds.testDict.items()} == {k: sorted(v) for k, v in a.testDict.items()}


def test_synthetic_shapes_match_calibration():
    import lgcn_b200 as lg
    g = lg.synth.make_graph('yelp2018')
    assert (g['n_users'], g['m_items']) == (31668, 38048)
    E = g['train_user'].size
    assert abs(E - 1237259) / 1237259 < 0.03
    du = np.bincount(g['train_user'], minlength=g['n_users'])
    assert du.min() >= 1 and 24 <= np.median(du) <= 30
    key = g['train_user'] * g['m_items'] + g['train_item']
    assert np.unique(key).size == key.size                          # de-duplicated
    tkey = g['test_user'] * g['m_items'] + g['test_item']
    assert np.intersect1d(key, tkey).size == 0                      # train/test disjoint
    g2 = lg.synth.make_graph('yelp2018')
    assert np.array_equal(g['train_item'], g2['train_item'])        # deterministic in the seed


def test_minibatch_shuffle_timer():
    import lgcn_b200 as lg
    u = lg.utils
    assert [list(b) for b in u.minibatch(list(range(5)), batch_size=2)] == [[0, 1], [2, 3], [4]]
    a, b = np.arange(10), np.arange(10) * 2
    assert [tuple(x.tolist() for x in t) for t in u.minibatch(a, b, batch_size=6)][1] == ([6, 7, 8, 9], [12, 14, 16, 18])
    np.random.seed(3); sa, sb = u.shuffle(a, b)
    np.random.seed(3); perm = np.arange(10); np.random.shuffle(perm)
    assert np.array_equal(sa, a[perm]) and np.array_equal(sb, b[perm])
    with pytest.raises(ValueError):
        u.shuffle(a, b[:3])
    u.timer.zero()
    with u.timer(name="Sample"):
        pass
    assert u.timer.dict().startswith("|Sample:")
    u.timer.zero()
    with u.timer():
        pass
    assert u.timer.get() >= 0


def test_host_metrics_match_oracle():
    import lgcn_b200 as lg
    from oracle import lightgcn_oracle as orc
    rng = np.random.default_rng(0)
    lg.world.configure(topks=[5, 20])
    try:
        for _ in range(20):
            gt = rng.choice(100, size=int(rng.integers(1, 30)), replace=False).tolist()
            top = rng.choice(100, size=20, replace=False)
            got = lg.Procedure.test_one_batch((top, gt))
            exp = orc.metrics_at_k(top[None, :], [gt], [5, 20])
            for k in ('precision', 'recall', 'ndcg'):
                assert np.allclose(got[k], exp[k], atol=1e-7)
    finally:
        lg.world.configure(topks=[20])


def test_device_paths_fail_loudly_without_gpu():
    import torch
    import lgcn_b200 as lg
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ds = lg.synth.make_dataset('tiny')
    with pytest.raises(RuntimeError):
        ds.getSparseGraph()
    with pytest.raises(RuntimeError):
        lg.ops.spmm(None, torch.zeros(4, 64), torch.zeros(4, 64))


def test_world_has_reference_defaults():
    import lgcn_b200 as lg
    c = lg.world.config
    assert (c['latent_dim_rec'], c['lightGCN_n_layers'], c['bpr_batch_size'], c['test_u_batch_size']) == (64, 3, 2048, 100)
    assert (c['lr'], c['decay'], lg.world.topks, lg.world.seed) == (0.001, 1e-4, [20], 2020)


def test_rebalance_by_time_moves_boundaries_toward_equal_time():
    """Row partition: equal cost is not equal time (item rows gather from the larger table); given measured block times
    the boundaries move so that, with time proportional to cost inside a block, every block takes the same time."""
    import torch
    from lgcn_b200.engine import balanced_row_bounds, rebalance_by_time
    indptr = torch.arange(0, 101 * 10, 10, dtype=torch.int64)             # 100 rows, 10 non-zeros each
    b = balanced_row_bounds(indptr, 2)
    assert b == [0, 50, 100]
    nb = rebalance_by_time(indptr, b, [1.0, 3.0])                         # block 1 is 3x slower per unit of cost
    assert nb[0] == 0 and nb[-1] == 100 and 65 <= nb[1] <= 68             # 50 + (1/3) * 50
    # predicted times after the move are equal
    t0 = 1.0 + 3.0 * (nb[1] - 50) / 50; t1 = 3.0 * (100 - nb[1]) / 50
    assert abs(t0 - t1) < 0.15
    assert rebalance_by_time(indptr, b, [2.0, 2.0]) == b                  # already balanced: unchanged
    b4 = balanced_row_bounds(indptr, 4, row_cost=5)
    nb4 = rebalance_by_time(indptr, b4, [1.0, 1.0, 1.0, 5.0], row_cost=5)
    assert nb4[0] == 0 and nb4[-1] == 100 and all(x <= y for x, y in zip(nb4, nb4[1:])) and nb4[3] > b4[3]


@pytest.mark.parametrize("m", [16384, 16385, 16511, 16512, 20011, 38048, 40981, 91599])
def test_tensor_core_item_layout_is_a_bijection_with_holes_in_the_last_slot(m):
    """The interleaved item order of the tensor-core ranking (host view of the mapping the kernels use): every item has
    one position, positions map back, the 128*T - m unused positions are slot 127 / 63 of block 127 (the only places the
    kernel masks holes), whether an item is in the sampled half of its tile (slots 0-63) alternates from id to id, and
    neighbouring item ids never share a tile."""
    import lgcn_b200 as lg
    lib = lg._lib.load()
    space = lib.lgcn_score_topk_tc_position_space(m)
    T = (m + 127) // 128
    assert space == 128 * T
    pos = np.array([lib.lgcn_score_topk_tc_host_position(i, m) for i in range(m)], dtype=np.int64)
    assert pos.min() >= 0 and pos.max() < space and np.unique(pos).size == m
    back = np.array([lib.lgcn_score_topk_tc_host_item(int(p), m) for p in pos[:: max(1, m // 4000)]])
    assert np.array_equal(back, np.arange(m)[:: max(1, m // 4000)])
    used = np.zeros(space, dtype=bool); used[pos] = True
    holes = np.nonzero(~used)[0]
    assert holes.size == space - m and np.all(np.isin(holes % 128, (63, 127)))
    assert all(lib.lgcn_score_topk_tc_host_item(int(h), m) == -1 for h in holes[:50])
    blk = np.arange(m) // T
    assert np.array_equal((pos % 128) < 64, (blk + np.arange(m) % T) % 2 == 0)  # sampled half: alternates from id to id
    sampled = (pos % 128) < 64
    assert abs(sampled[: T].mean() - 0.5) < 0.02 and abs(sampled[T: 2 * T].mean() - 0.5) < 0.02    # every id range is half sampled
    tile = pos // 128
    assert np.all(tile[1:][blk[1:] == blk[:-1]] != tile[:-1][blk[1:] == blk[:-1]])   # neighbours in different tiles
    far = np.abs(tile[1:] - tile[:-1])[blk[1:] == blk[:-1]]
    assert np.median(far) > T // 4                                              # ... and far apart
    assert lib.lgcn_score_topk_tc_host_position(0, 1000) == -1                  # small tables are not laid out at all


def test_native_parser_fuzz_against_python_split(tmp_path):
    """Property test (hypothesis): any file made of non-negative integers separated by blanks/tabs and line ends parses
    to the same pairs as Python's line.split() loop, and the id maxima come from lines that have items."""
    import ctypes
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    import lgcn_b200 as lg
    from lgcn_b200 import dataloader
    lib = lg._lib.load()
    sep = st.sampled_from([" ", "  ", "\t", " \t "])
    line = st.tuples(st.lists(st.integers(0, 2**31 - 1), min_size=0, max_size=6), sep, st.sampled_from(["", " ", "\r"]))
    path = tmp_path / "f.txt"

    @settings(max_examples=60, deadline=None)
    @given(st.lists(line, min_size=0, max_size=30), st.sampled_from(["\n", ""]))
    def check(lines, tail):
        text = "\n".join(s.join(str(v) for v in vals) + end for vals, s, end in lines) + tail
        path.write_text(text)
        eu, ei = [], []
        for l in text.split("\n"):
            cols = l.split()
            if len(cols) >= 2:
                eu.extend([int(cols[0])] * (len(cols) - 1)); ei.extend(int(c) for c in cols[1:])
        u, i = dataloader._parse_interactions(str(path))
        assert u.tolist() == eu and i.tolist() == ei
        mu, mi = ctypes.c_int64(), ctypes.c_int64()
        n = lib.lgcn_parse_interactions(str(path).encode(), None, None, 0, ctypes.byref(mu), ctypes.byref(mi))
        assert n == len(eu) and mu.value == (max(eu) if eu else -1) and mi.value == (max(ei) if ei else -1)

    check()


@pytest.mark.parametrize("kind", ["random", "popular_adjacent_ids"])
def test_tensor_core_threshold_rule_on_the_host(kind):
    """The selection rule of the tensor-core ranking, replayed with numpy over the library's own item layout: tau = the
    28th largest maximum over the 64-item halves of every EVEN tile (the 50 % sample pass 1 scores).  Claims checked: at least 28 and on the order
    of 56 items reach tau, the k = 20 best items are comfortably above it, and no (split, half) event list of a row gets
    more than its 32 slots — for random embeddings AND for the trained-model picture (popular items = adjacent small
    ids, large norms, best for everybody), which in plain id order put > 100 hits into one list."""
    import lgcn_b200 as lg
    lib = lg._lib.load()
    rng = np.random.default_rng(5)
    nu, ni, d, k = 120, 24001, 64, 20
    if kind == "random":
        U = rng.normal(0, 0.2, (nu, d)).astype(np.float32); V = rng.normal(0, 0.2, (ni, d)).astype(np.float32)
    else:
        c = rng.normal(0, 1, d).astype(np.float32); c /= np.linalg.norm(c)
        pop = (4.0 / (1.0 + np.arange(ni) / 150.0)).astype(np.float32)
        V = pop[:, None] * (c[None, :] + 0.35 * rng.normal(0, 1, (ni, d)).astype(np.float32) / np.sqrt(d)) + 0.05 * rng.normal(0, 1, (ni, d)).astype(np.float32)
        U = (1.0 + rng.random(nu).astype(np.float32))[:, None] * (c[None, :] + 0.5 * rng.normal(0, 1, (nu, d)).astype(np.float32) / np.sqrt(d))
    T = (ni + 127) // 128
    pos = np.array([lib.lgcn_score_topk_tc_host_position(i, ni) for i in range(ni)], dtype=np.int64)
    S = U @ V.T
    n_splits = 5; per = -(-T // n_splits)
    worst_list, hits_all, in_order_worst = 0, [], 0
    for layout, positions in (("interleaved", pos), ("id order", np.arange(ni, dtype=np.int64))):
        for u in range(nu):
            sp = np.full(T * 128, -np.inf, dtype=np.float32); sp[positions] = S[u]
            tiles = sp.reshape(T, 128)
            tau = np.sort(tiles[0::2].reshape(-1, 2, 64).max(axis=2).ravel())[-28]
            hit_pos = np.nonzero(sp >= tau)[0]
            lists = np.bincount((hit_pos // 128 // per) * 2 + (hit_pos % 128) // 64, minlength=2 * n_splits)
            if layout == "interleaved":
                hits_all.append(hit_pos.size); worst_list = max(worst_list, int(lists.max()))
                assert hit_pos.size >= 28
                assert np.sort(S[u])[-k] > tau                        # the answer lies above the threshold
            else:
                in_order_worst = max(in_order_worst, int(lists.max()))
    assert 40 <= np.median(hits_all) <= 90, np.median(hits_all)
    assert max(hits_all) <= 160 and worst_list <= 32, (max(hits_all), worst_list)
    if kind == "popular_adjacent_ids":
        assert in_order_worst > 32                                    # what the layout is for


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver's contract for `bench.py --impl reference` (runs on the host cores, no GPU): stdout is ONE JSON line with the
    same metric/unit/config keys as the GPU arm, `impl: reference`, a cpu_baseline describing the run and a zero-copy e2e block;
    everything else (library banners, prints) goes to stderr."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'tiny', '--steps', '3', '--warmup', '3'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'samples/s' and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['steps'] == 3 and d['warmup'] == 3 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config'] and d['gpu_launches'] == 0

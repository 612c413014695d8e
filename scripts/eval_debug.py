import os, sys, numpy as np, torch
os.environ["LGCN_TC_KEEP_WORKSPACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import lgcn_b200 as lg
z = np.load(os.path.join(ROOT, "tests", "golden", "gowalla.npz"))
nu, ni = int(z['n_users']), int(z['m_items'])
tu = np.repeat(np.arange(nu), np.diff(z['train_indptr'])).astype(np.int64); ti = z['train_items'].astype(np.int64)
su = np.repeat(z['test_users'].astype(np.int64), np.diff(z['test_indptr'])); si = z['test_items'].astype(np.int64)
lg.world.configure(checkpoint_dir="/tmp/lgcn_report", topks=[20], seed=2020, device_sampler=False)
cfg = dict(lg.world.config)
ds = lg.InteractionDataset(nu, ni, tu, ti, su, si, config=cfg, name='gowalla')
lg.utils.set_seed(2020); lg.utils.sampler_seed(2020)
model = lg.LightGCN(cfg, ds); bpr = lg.utils.BPRLoss(model, cfg)
for ep in range(10): lg.Procedure.BPR_train_original(ds, model, bpr, ep)
users = ds.test_csr()[0]
model.rank_topk(users, 20)
fl = lg.ops.last_tc_flags
bad = users[fl == 3]
deg = torch.from_numpy(np.diff(z['train_indptr'])).cuda()
print("flagged", bad.numel(), "train degree of flagged: median", deg[bad].float().median().item(), "max", deg[bad].max().item(), "min", deg[bad].min().item(), "| all users median", deg.float().median().item())
import ctypes
ws, off0, Bt, mi = lg.ops.last_tc_workspace
lay = (ctypes.c_int64 * 8)(); lg._lib.load().lgcn_score_topk_tc_debug_layout(Bt, mi, lay)
lay = list(lay); print("layout", lay)
tau_k = ws[off0 + lay[0]: off0 + lay[0] + 4 * Bt].view(torch.float32)
cc = ws[off0 + lay[1]: off0 + lay[1] + 4 * Bt * lay[2]].view(torch.int32).view(Bt, lay[2])
rows_bad = torch.nonzero(fl == 3).flatten()
print("cand_cnt of flagged rows:", cc[rows_bad[:6]].tolist())
print("events per row (ok rows): mean", cc[fl == 0].clamp(min=0).sum(1).float().mean().item(), "max list", cc[fl == 0].max().item())
tau_of = {int(users[r]): float(tau_k[r]) for r in rows_bad[:6].tolist()}
U, V = model.computer()
g = model._csr
for u in bad[:6].tolist():
    s = (V @ U[u]).clone()
    tr = g.indices[g.indptr[u]:g.indptr[u + 1]].long() - nu
    s[tr] = -1e30
    top = torch.sort(s, descending=True).values
    T = (ni + 127) // 128
    pos_of = lambda i: (i % T) * 128 + ((i // T) >> 1) + 64 * ((i // T) & 1)
    ids = torch.arange(ni, device='cuda'); pos = (ids % T) * 128 + ((ids // T) >> 1) + 64 * ((ids // T) & 1)
    sampled = (pos % 128) < 64
    sp = torch.full((T * 128,), -1e30, device='cuda'); sp[pos] = s
    blk = sp.view(T, 128)[:, :64].max(dim=1).values
    tau = torch.sort(blk, descending=True).values[27]
    print("user", u, "deg", deg[u].item(), "|u|", U[u].norm().item(), "top1", top[0].item(), "s20", top[19].item(), "s56", top[55].item(), "tau", tau.item(), "tau_kernel", tau_of.get(u), "n>=tau", int((s >= tau).sum()))

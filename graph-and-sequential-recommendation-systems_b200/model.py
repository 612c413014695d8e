"""LightGCN with the reference's model API (code/model.py:37-231), running on liblgcn_b200 kernels.

Same constructor `LightGCN(config, dataset)`, same methods (`computer`, `getUsersRating`,
`getEmbedding`, `bpr_loss`, `forward`) and the same parameter names (`embedding_user.weight`,
`embedding_item.weight`, code/main.py:56-87 checkpoints load with strict=True), so it can be put in
the reference's registry:  register.MODELS['lgn'] = lgcn_b200.LightGCN.

Differences that are not visible through the API:
  * the two embedding tables are views into ONE contiguous (N,d) device buffer (no torch.cat);
  * propagation, BPR loss, their backward and Adam are CUDA kernels (engine.py); autograd sees two
    custom Functions whose backward is the closed form of SURVEY.md §3.2;
  * `fused_train_step` runs forward+backward+Adam without materialising gradients (used by this
    package's utils.BPRLoss);  the generic path (`bpr_loss(...)` -> `.backward()` -> any torch
    optimiser) stays available for the reference's own utils.BPRLoss.
The two optional variants of the reference are supported through the generic
autograd path: item-item smoothing (code/model.py:99-109,228-229) is one more K1 product with an explicit-value CSR
(and its transpose in the backward), the popularity gate (code/model.py:66-96,139-157,176-181) keeps its two tiny MLPs
as torch modules with the reference's parameter names.  Neither uses the fused training step.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops, world
from .engine import Engine


def as_csr_graph(dataset, graph, seg_len):
    """Find or derive the int32 CSR of the normalised adjacency for any dataset the reference accepts."""
    if hasattr(dataset, 'getCSRGraph'):
        return dataset.getCSRGraph()
    if isinstance(graph, ops.CSRGraph):
        return graph
    if isinstance(graph, torch.Tensor) and graph.layout == torch.sparse_csr:
        g = graph.to(world.device)
        return ops.CSRGraph(g.crow_indices().to(torch.int32).contiguous(), g.col_indices().to(torch.int32).contiguous(),
                            g.values().to(torch.float32).contiguous(), g.shape[1], seg_len=seg_len)
    if isinstance(graph, torch.Tensor) and graph.layout == torch.sparse_coo:
        g = graph.to(world.device).coalesce()           # reference layout (code/dataloader.py:244)
        idx = g.indices()
        return ops.coo_to_csr(idx[0].contiguous(), idx[1].contiguous(), g.values().contiguous(),
                              g.shape[0], g.shape[1], seg_len=seg_len)
    raise TypeError(f"cannot interpret graph of type {type(graph)}")


class _Propagate(torch.autograd.Function):
    """(user_w, item_w) -> out (N,d) = mean_k A^k E0.  Backward: g_0 = Horner chain over A (symmetric)."""

    @staticmethod
    def forward(ctx, user_w, item_w, model):
        eng = model._engine
        model._sync_params_into_engine(user_w, item_w)
        eng.forward()
        ctx.model = model
        # a fresh alias of the resident buffer: the buffer object itself never acquires autograd history
        return eng.out.detach()

    @staticmethod
    def backward(ctx, g_out):
        model = ctx.model
        eng = model._engine
        g = g_out.contiguous()
        if g.data_ptr() in (x.data_ptr() for x in eng.X):
            g = g.clone()
        grad = eng.backward_to(g, eng.grad_buffer())
        nu = model.n_users
        return grad[:nu].clone(), grad[nu:].clone(), None


class _BprLoss(torch.autograd.Function):
    """out, (u,p,n) -> (bpr, reg).  Backward scatters the closed-form gradient with K2."""

    @staticmethod
    def forward(ctx, out, users, pos, neg, model):
        eng = model._engine
        B = eng.stage_batch_now(users, pos, neg)
        out_c = out.contiguous()
        ops.bpr_fwd_bwd(out_c, eng.bu, eng.bp, eng.bn, eng.B_cap, eng.ctl, eng.nu, eng.ni, 1.0 / max(B, 1), 0.0,
                        0.0, 0.0, eng.loss_out, None, eng.bpr_ws)
        ctx.model = model
        ctx.B = B
        ctx.save_for_backward(out_c, eng.bu[:B].clone(), eng.bp[:B].clone(), eng.bn[:B].clone())
        res = eng.loss_out.clone()
        return res[0], res[1]

    @staticmethod
    def backward(ctx, g_bpr, g_reg):
        model = ctx.model
        eng = model._engine
        out_c, u, p, n = ctx.saved_tensors
        B = eng.stage_batch_now(u, p, n)
        G = eng.scratch()
        G.zero_()
        tmp = torch.empty(4, dtype=torch.float32, device=out_c.device)
        ops.bpr_fwd_bwd(out_c, eng.bu, eng.bp, eng.bn, eng.B_cap, eng.ctl, eng.nu, eng.ni, 1.0 / max(B, 1), 0.0,
                        float(g_bpr), float(g_reg), tmp, G, eng.bpr_ws, deterministic=eng.deterministic)
        return G, None, None, None, None


class _I2ISmooth(torch.autograd.Function):
    """items -> items + alpha * (I2I @ items)   (code/model.py:228-229); backward uses the transposed CSR."""

    @staticmethod
    def forward(ctx, items, model):
        x = items.contiguous()
        y = torch.empty_like(x)
        ops.spmm(model._i2i, x, y, model.i2i_alpha, 1.0, [x])
        ctx.model = model
        return y

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        gc = g.contiguous()
        gx = torch.empty_like(gc)
        ops.spmm(model._i2i_t, gc, gx, model.i2i_alpha, 1.0, [gc])
        return gx, None


def _csr_from_scipy(m, device, seg_len):
    m = m.tocsr().astype(np.float32)
    m.sort_indices()
    return ops.CSRGraph(torch.from_numpy(m.indptr.astype(np.int32)).to(device), torch.from_numpy(m.indices.astype(np.int32)).to(device),
                        torch.from_numpy(m.data.astype(np.float32)).to(device), m.shape[1], seg_len=seg_len)


class LightGCN(nn.Module):
    def __init__(self, config, dataset):
        super().__init__()
        self.config = config
        self.dataset = dataset
        self.device = world.device
        if self.device.type != 'cuda':
            raise RuntimeError("lgcn_b200.LightGCN needs a CUDA device (B200); there is no CPU path")
        self.n_users = dataset.n_users
        self.m_items = dataset.m_items
        self.latent_dim = config['latent_dim_rec']
        self.n_layers = config['lightGCN_n_layers']
        self.keep_prob = config.get('keep_prob', 0.6)
        self.use_pop_gate = bool(config.get('use_pop_gate', False))
        self.use_item_item = bool(config.get('use_item_item', False))
        self.i2i_alpha = float(config.get('i2i_alpha', 0.0))
        # Same RNG consumption as the reference (code/model.py:57-60): two default nn.Embedding inits on
        # the CPU generator, then normal_(std=0.1) on user then item table.
        self.embedding_user = nn.Embedding(self.n_users, self.latent_dim)
        self.embedding_item = nn.Embedding(self.m_items, self.latent_dim)
        nn.init.normal_(self.embedding_user.weight, std=0.1)
        nn.init.normal_(self.embedding_item.weight, std=0.1)

        N = self.n_users + self.m_items
        if (config.get('dist_mode') == 'rowpart' and config.get('rowpart_partition_memory', True)
                and hasattr(dataset, 'getRowBlockBuilder')):
            # row partition: this rank assembles and keeps ONLY its row block of the adjacency; the
            # whole-graph tensor the reference hands to torch.sparse.mm does not exist on any rank
            self.Graph = None
            self._csr = None
            graph_src = dataset.getRowBlockBuilder()
        else:
            self.Graph = dataset.getSparseGraph()
            self._csr = as_csr_graph(dataset, self.Graph, int(config.get('spmm_seg_len', ops.DEFAULT_SEG_LEN)))
            graph_src = self._csr
        if graph_src.n_rows != N or graph_src.n_cols != N:
            raise RuntimeError(f"adjacency is {graph_src.n_rows}x{graph_src.n_cols}, expected {N}x{N}")
        self._engine = Engine(graph_src, self.n_users, self.m_items, self.latent_dim, self.n_layers, self.device,
                              lr=config.get('lr', 1e-3), decay=config.get('decay', 1e-4),
                              B_cap=config.get('bpr_batch_size', 2048),
                              deterministic=config.get('deterministic', False),
                              use_graph=config.get('cuda_graph', True),
                              dist_mode=config.get('dist_mode', None), prune=config.get('prune_dead_rows', True),
                              p2p=config.get('rowpart_p2p', True), row_cost=config.get('rowpart_row_cost', None),
                              multicast=config.get('rowpart_multicast', True), rebalance=config.get('rowpart_rebalance', 3))
        self._cache_key = None
        self._pack_params()
        seg_len = int(config.get('spmm_seg_len', ops.DEFAULT_SEG_LEN))

        # popularity gate (code/model.py:66-96): same construction order, so the CPU RNG stream matches the reference
        self.pop_hidden = int(config.get('pop_hidden', 32))
        self.gate_hidden = int(config.get('gate_hidden', 64))
        self.gate_entropy_coeff = float(config.get('gate_entropy_coeff', 1e-4))
        self.pop_gate_temp = float(config.get('pop_gate_temp', 1.0))
        self.item_pop_scalar, self.pop_mlp, self.gate_mlp = None, None, None
        if self.use_pop_gate:
            pop = torch.log1p(torch.from_numpy(np.asarray(dataset.items_D)).float().clamp(min=0.0))
            self.item_pop_scalar = ((pop - pop.mean()) / (pop.std() + 1e-8)).to(self.device)
            self.pop_mlp = nn.Sequential(nn.Linear(1, self.pop_hidden), nn.ReLU(), nn.Linear(self.pop_hidden, self.latent_dim)).to(self.device)
            self.gate_mlp = nn.Sequential(nn.Linear(2 * self.latent_dim, self.gate_hidden), nn.ReLU(), nn.Linear(self.gate_hidden, 1)).to(self.device)
            self._pg_flat = None
            if (self.latent_dim in (32, 64, 128) and self.pop_hidden <= 64 and self.gate_hidden <= 128
                    and config.get('dist_mode') is None and config.get('popgate_kernel', True)):
                self._pack_popgate()

        # item-item smoothing (code/model.py:99-109): explicit-value CSR and its transpose, both served by K1
        self._i2i = self._i2i_t = None
        if self.use_item_item and config.get('i2i_path'):
            import scipy.sparse as sp
            try:
                m = sp.load_npz(config['i2i_path'])
                self._i2i = _csr_from_scipy(m, self.device, seg_len)
                self._i2i_t = _csr_from_scipy(m.T, self.device, seg_len)
                world.cprint(f"[I2I] loaded {config['i2i_path']}, nnz={m.nnz}")
            except Exception as e:      # the reference only warns and carries on without the item graph
                world.cprint(f"[I2I] WARNING: cannot load {config['i2i_path']}: {e}")
                self._i2i = self._i2i_t = None
        if (self._i2i is not None and self.i2i_alpha > 0.0 and not self.use_pop_gate and config.get('dist_mode') is None
                and config.get('i2i_kernel_step', True)):
            self._engine.enable_i2i(self._i2i, self._i2i_t, self.i2i_alpha)

    @property
    def plain(self):
        """True when the fused training step applies: the plain model, the popularity gate with its kernels enabled
        (csrc/popgate.cu), or item-item smoothing (two more K1 products).  Both variants at once train through
        bpr_loss().backward()."""
        if self._i2i is not None and self.i2i_alpha > 0.0:
            return self._engine.i2i is not None                      # item-item smoothing: two more K1 products in the step
        return (not self.use_pop_gate) or getattr(self, '_pg_flat', None) is not None

    def popgate_tensors(self):
        """The 8 MLP tensors in the order of the kernel's parameter block (include/lgcn_b200.h)."""
        return [self.pop_mlp[0].weight, self.pop_mlp[0].bias, self.pop_mlp[2].weight, self.pop_mlp[2].bias,
                self.gate_mlp[0].weight, self.gate_mlp[0].bias, self.gate_mlp[2].weight, self.gate_mlp[2].bias]

    def _pack_popgate(self):
        """Make the 8 MLP tensors views of ONE flat device block (what the kernels read and the fused Adam updates); names,
        shapes and values stay the reference's (state_dict keys pop_mlp.*, gate_mlp.*)."""
        n = ops.popgate_param_count(self.latent_dim, self.pop_hidden, self.gate_hidden)
        flat = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=self.device)      # padded: the dense Adam kernel works in float4
        off = 0
        with torch.no_grad():
            for t in self.popgate_tensors():
                k = t.numel()
                flat[off:off + k].copy_(t.data.reshape(-1))
                t.data = flat[off:off + k].view(t.shape)
                off += k
        assert off == n
        self._pg_flat = flat
        self._engine.enable_popgate(self.item_pop_scalar.contiguous(), flat, self.pop_hidden, self.gate_hidden,
                                    self.pop_gate_temp, self.gate_entropy_coeff)

    def _popgate_packed(self):
        if self._pg_flat is None:
            return False
        off = 0
        for t in self.popgate_tensors():
            if t.data_ptr() != self._pg_flat.data_ptr() + 4 * off:
                return False
            off += t.numel()
        return True

    def _fuse_item_embeddings(self, items_emb):
        """gate = sigmoid(gate_mlp([items, pop_vec]) / T); fused = gate*items + (1-gate)*pop_vec  (code/model.py:139-157)."""
        pop_vec = self.pop_mlp(self.item_pop_scalar.unsqueeze(1))
        logit = self.gate_mlp(torch.cat([items_emb, pop_vec], dim=1))
        if self.pop_gate_temp != 1.0:
            logit = logit / self.pop_gate_temp
        gate = torch.sigmoid(logit)
        self._last_item_gate = gate
        return gate * items_emb + (1.0 - gate) * pop_vec

    def _items_for_scoring(self, all_items):
        if not self.use_pop_gate:
            return all_items
        eng = self._engine
        if (getattr(self, '_pg_flat', None) is not None and not torch.is_grad_enabled()
                and all_items.data_ptr() == eng.out.data_ptr() + self.n_users * self.latent_dim * 4):
            if not self._popgate_packed():
                self._pack_popgate()
            # the fusion kernel on the resident propagated table (no M x 2d concat, no M x H intermediates)
            fused, gate = ops.popgate_fuse(eng.out, self.n_users, self.m_items, self.item_pop_scalar, self._pg_flat,
                                           self.pop_hidden, self.gate_hidden, self.pop_gate_temp, want_gate=True)
            self._last_item_gate = gate.unsqueeze(1)
            return fused
        return self._fuse_item_embeddings(all_items).contiguous()

    # ------------------------------------------------------------------ parameter storage
    @property
    def _feat(self):
        """Feature partition (dist_mode='featpart', > 1 rank): the engine holds d/P COLUMNS of every table; the two embedding
        tables of the API are views of a full (N,d) table kept by the model, pushed into the engine's slice when they were
        written from outside and refreshed from the ranks' slices (collective) when they are read after training steps."""
        eng = self._engine
        return eng.dist_mode == 'featpart' and eng.world > 1

    def _param_table(self):
        if not self._feat:
            return self._engine.E0
        if getattr(self, '_E0_full', None) is None:
            self._E0_full = torch.zeros((self.n_users + self.m_items, self.latent_dim), dtype=torch.float32, device=self.device)
            self._out_full = None
            self._engine.pre_step_hook = self._feat_push_params
            self._engine.full_param_sync = self._feat_pull_params
        return self._E0_full

    def _pack_params(self):
        """Make both embedding tables views of ONE contiguous (N,d) buffer: the engine's E0 (feature partition: the model's
        full table, whose column slice the engine trains)."""
        eng = self._engine
        nu = self.n_users
        tab = self._param_table()
        with torch.no_grad():
            tab[:nu].copy_(self.embedding_user.weight.data)
            tab[nu:].copy_(self.embedding_item.weight.data)
        self.embedding_user.weight.data = tab[:nu]
        self.embedding_item.weight.data = tab[nu:]
        self._cache_key = None
        self._engine._e0_synced = False
        if self._feat:
            self._feat_pushed = None
            self._feat_push_params()

    def _params_packed(self, user_w=None, item_w=None):
        tab = self._param_table()
        uw = self.embedding_user.weight if user_w is None else user_w
        iw = self.embedding_item.weight if item_w is None else item_w
        return (uw.data_ptr() == tab.data_ptr()
                and iw.data_ptr() == tab.data_ptr() + self.n_users * self.latent_dim * 4)

    def _sync_params_into_engine(self, user_w, item_w):
        if not self._params_packed(user_w, item_w):
            tab = self._param_table()
            with torch.no_grad():
                tab[:self.n_users].copy_(user_w)
                tab[self.n_users:].copy_(item_w)
            if self._feat:
                self._feat_pushed = None
        if self._feat:
            self._feat_push_params()

    def _feat_versions(self):
        return (self.embedding_user.weight._version, self.embedding_item.weight._version, self.embedding_user.weight.data_ptr())

    def _feat_push_params(self):
        """Full table -> the engine's column slice, when the parameters were written from outside since the last push
        (initialisation, load_state_dict, an external optimiser).  Engine.step / epoch_step / forward call it first."""
        eng = self._engine
        if not self._params_packed():
            return self._pack_params()
        key = self._feat_versions()
        if getattr(self, '_feat_pushed', None) == key:
            return
        with torch.no_grad():
            eng.E0.copy_(self._E0_full[:, eng.c0:eng.c0 + eng.d])
        self._feat_pushed = key
        self._feat_pulled_epoch = eng.param_epoch
        self._cache_key = None

    def _feat_pull_params(self):
        """The ranks' trained column slices -> the full table behind embedding_user/embedding_item (collective: every rank)."""
        eng = self._engine
        self._feat_push_params()
        if getattr(self, '_feat_pulled_epoch', None) == eng.param_epoch:
            return
        with torch.no_grad():
            eng.gather_columns(eng.E0, out_full=self._E0_full)
        self._feat_pulled_epoch = eng.param_epoch
        self._feat_pushed = self._feat_versions()

    def state_dict(self, *args, **kwargs):
        if self._feat:
            self._feat_pull_params()                    # collective under the feature partition: every rank calls state_dict()
        return super().state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        # .to(device)/.cuda()/.float() re-create parameter storage; re-pack afterwards
        super()._apply(fn, *args, **kwargs)
        if not self._params_packed():
            if self.embedding_user.weight.device.type != 'cuda':
                raise RuntimeError("lgcn_b200.LightGCN parameters must stay on the CUDA device")
            self._pack_params()
        if getattr(self, '_pg_flat', None) is not None and not self._popgate_packed():
            self._pack_popgate()
        return self

    def invalidate_cache(self):
        """Probed with hasattr by the reference driver (code/main.py:190-191)."""
        self._cache_key = None

    # ------------------------------------------------------------------ propagation
    def _param_key(self):
        return (self.embedding_user.weight._version, self.embedding_item.weight._version, self._engine.param_epoch)

    def computer(self):
        """LightGCN propagation (code/model.py:201-231) -> (all_users, all_items)."""
        uw, iw = self.embedding_user.weight, self.embedding_item.weight
        nu = self.n_users
        if torch.is_grad_enabled() and (uw.requires_grad or iw.requires_grad):
            if self._feat:
                raise NotImplementedError("dist_mode='featpart' trains through the fused step (utils.BPRLoss.stageOne / "
                                          "Procedure.BPR_train_original); call computer() under torch.no_grad()")
            out = _Propagate.apply(uw, iw, self)
            self._cache_key = None
            items = out[nu:]
            if self._i2i is not None and self.i2i_alpha > 0.0:
                items = _I2ISmooth.apply(items, self)
            return out[:nu], items
        key = self._param_key()
        if self._cache_key != key:
            self._sync_params_into_engine(uw, iw)
            self._engine.forward()
            if self._feat:      # every rank propagated its columns: all-gather them into the full table the scoring reads (collective)
                self._out_full = self._engine.gather_columns(self._engine.out, out_full=self._out_full)
            self._cache_key = self._param_key()
            self._items_smoothed = None
            if self._i2i is not None and self.i2i_alpha > 0.0:
                x = self._engine.out[nu:]
                self._items_smoothed = torch.empty_like(x)
                ops.spmm(self._i2i, x, self._items_smoothed, self.i2i_alpha, 1.0, [x])
        out = self._out_full if self._feat else self._engine.out
        items = out[nu:] if getattr(self, '_items_smoothed', None) is None else self._items_smoothed
        return out[:nu], items

    def getUsersRating(self, users):
        """Scores of every item for a batch of users (code/model.py:114-123) -> (B, m_items)."""
        with torch.no_grad():            # the reference only calls this under no_grad (code/Procedure.py:161,174)
            all_users, all_items = self.computer()
            users = users.to(self.device, dtype=torch.int64).contiguous()
            items = self._items_for_scoring(all_items)
            if self.config.get('score_tensor_core', True) and self.latent_dim == 64:
                return ops.score_dense_tc(all_users, items, users)       # 3xTF32 on tcgen05 (fp32-class accuracy)
            return ops.score_dense(all_users, items, users)

    def rank_topk(self, users, k, mask=True):
        """Fused getUsersRating + train-item mask (-1024) + top-k (code/Procedure.py:174-183):
        returns (item ids int64 [B,k], scores float32 [B,k]); the B x M matrix is never written."""
        with torch.no_grad():
            all_users, all_items = self.computer()
            all_items = self._items_for_scoring(all_items)
            users = users.to(self.device, dtype=torch.int64).contiguous()
            mi, mx, off = self._train_mask() if mask else (None, None, 0)
            if self.config.get('score_tensor_core', True):
                idx, val, self.last_rank_redone = ops.score_topk_tc(all_users, all_items, users, k, mi, mx, off)
                return idx, val
            return ops.score_topk(all_users, all_items, users, k, mi, mx, off)

    def _train_mask(self):
        """(indptr, indices, column offset) of the train-item mask over the user rows: the user rows of the adjacency
        when this rank holds them, else the dataset's own user->items CSR (memory-partitioned row partition)."""
        if self._csr is not None:
            return self._csr.indptr, self._csr.indices, self.n_users
        if not hasattr(self.dataset, 'train_mask_csr'):
            raise RuntimeError("this dataset cannot supply the train-item mask without the whole adjacency")
        ip, it = self.dataset.train_mask_csr()
        return ip, it, 0

    def getEmbedding(self, users, pos_items, neg_items):
        all_users, all_items = self.computer()
        i_emb = self._fuse_item_embeddings(all_items) if self.use_pop_gate else all_items
        u = all_users[users.long()]
        pos = i_emb[pos_items.long()]
        neg = i_emb[neg_items.long()]
        return u, pos, neg, all_users, all_items

    def bpr_loss(self, users, pos, neg):
        """(bpr, reg) as in code/model.py:162-183; both support .backward() through loss + decay*reg."""
        uw, iw = self.embedding_user.weight, self.embedding_item.weight
        users, pos, neg = (t.to(torch.int64).contiguous() for t in (users, pos, neg))
        if self.use_pop_gate or (self._i2i is not None and self.i2i_alpha > 0.0):
            # variants: the reference's arithmetic on top of the kernel-backed propagation (code/model.py:162-183)
            u, pos_e, neg_e, _, _ = self.getEmbedding(users.to(self.device), pos.to(self.device), neg.to(self.device))
            bpr = -torch.mean(F.logsigmoid((u * pos_e).sum(dim=1) - (u * neg_e).sum(dim=1)))
            reg = 0.5 * (u.norm(2).pow(2) + pos_e.norm(2).pow(2) + neg_e.norm(2).pow(2)) / float(u.shape[0])
            if self.use_pop_gate:
                gates = torch.cat([self._last_item_gate[pos.to(self.device)], self._last_item_gate[neg.to(self.device)]], dim=0)
                gates = torch.clamp(gates, 1e-6, 1.0 - 1e-6)
                entropy = -(gates * torch.log(gates) + (1 - gates) * torch.log(1 - gates)).mean()
                bpr = bpr - self.gate_entropy_coeff * entropy
            return bpr, reg
        if self._feat:
            raise NotImplementedError("dist_mode='featpart' has no autograd path: train through utils.BPRLoss.stageOne")
        if torch.is_grad_enabled() and (uw.requires_grad or iw.requires_grad):
            out = _Propagate.apply(uw, iw, self)
            self._cache_key = None
        else:
            au, _ = self.computer()
            out = self._engine.out
        return _BprLoss.apply(out, users, pos, neg, self)

    def forward(self, users, items):
        all_users, all_items = self.computer()
        i_emb = self._fuse_item_embeddings(all_items) if self.use_pop_gate else all_items
        return (all_users[users.long()] * i_emb[items.long()]).sum(dim=1)

    # ------------------------------------------------------------------ fused training step
    def fused_train_step(self, users, pos, neg, lr=None, B_global=0):
        """stageOne without autograd: forward, BPR, backward and Adam in one captured sequence.
        Returns the engine (loss in engine.loss_out on the device)."""
        if not self.plain:
            raise RuntimeError("the fused step covers the plain model and ONE variant (pop-gate or item-item); both at once train through bpr_loss().backward()")
        if not self._params_packed():
            self._pack_params()
        if self.use_pop_gate and not self._popgate_packed():
            self._pack_popgate()
        eng = self._engine
        if lr is not None:
            eng.set_lr(lr)
        eng.step(users, pos, neg, B_global)
        return eng

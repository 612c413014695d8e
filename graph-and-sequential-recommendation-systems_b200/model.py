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
Pop-gate (code/model.py:66-96,139-157) and item-item smoothing (:99-109,228-229) are outside the
accelerated path and raise NotImplementedError when switched on.
"""
import torch
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
        B = eng._stage_batch(users, pos, neg)
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
        B = eng._stage_batch(u, p, n)
        G = eng.scratch()
        G.zero_()
        tmp = torch.empty(4, dtype=torch.float32, device=out_c.device)
        ops.bpr_fwd_bwd(out_c, eng.bu, eng.bp, eng.bn, eng.B_cap, eng.ctl, eng.nu, eng.ni, 1.0 / max(B, 1), 0.0,
                        float(g_bpr), float(g_reg), tmp, G, eng.bpr_ws, deterministic=eng.deterministic)
        return G, None, None, None, None


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
        if self.use_pop_gate or (self.use_item_item and config.get('i2i_path')):
            raise NotImplementedError("pop-gate / item-item variants are outside the accelerated hot path "
                                      "(SURVEY.md §2 rows 1a/1b, §8f #4)")
        # Same RNG consumption as the reference (code/model.py:57-60): two default nn.Embedding inits on
        # the CPU generator, then normal_(std=0.1) on user then item table.
        self.embedding_user = nn.Embedding(self.n_users, self.latent_dim)
        self.embedding_item = nn.Embedding(self.m_items, self.latent_dim)
        nn.init.normal_(self.embedding_user.weight, std=0.1)
        nn.init.normal_(self.embedding_item.weight, std=0.1)

        self.Graph = dataset.getSparseGraph()
        self._csr = as_csr_graph(dataset, self.Graph, int(config.get('spmm_seg_len', ops.DEFAULT_SEG_LEN)))
        N = self.n_users + self.m_items
        if self._csr.n_rows != N or self._csr.n_cols != N:
            raise RuntimeError(f"adjacency is {self._csr.n_rows}x{self._csr.n_cols}, expected {N}x{N}")
        self._engine = Engine(self._csr, self.n_users, self.m_items, self.latent_dim, self.n_layers, self.device,
                              lr=config.get('lr', 1e-3), decay=config.get('decay', 1e-4),
                              B_cap=config.get('bpr_batch_size', 2048),
                              deterministic=config.get('deterministic', False),
                              use_graph=config.get('cuda_graph', True),
                              dist_mode=config.get('dist_mode', None), prune=config.get('prune_dead_rows', True),
                              p2p=config.get('rowpart_p2p', True))
        self._cache_key = None
        self._pack_params()

    # ------------------------------------------------------------------ parameter storage
    def _pack_params(self):
        """Make both embedding tables views of the engine's contiguous E0 buffer."""
        eng = self._engine
        nu = self.n_users
        with torch.no_grad():
            eng.E0[:nu].copy_(self.embedding_user.weight.data)
            eng.E0[nu:].copy_(self.embedding_item.weight.data)
        self.embedding_user.weight.data = eng.E0[:nu]
        self.embedding_item.weight.data = eng.E0[nu:]
        self._cache_key = None
        self._engine._e0_synced = False

    def _params_packed(self, user_w=None, item_w=None):
        eng = self._engine
        uw = self.embedding_user.weight if user_w is None else user_w
        iw = self.embedding_item.weight if item_w is None else item_w
        return (uw.data_ptr() == eng.E0.data_ptr()
                and iw.data_ptr() == eng.E0.data_ptr() + self.n_users * self.latent_dim * 4)

    def _sync_params_into_engine(self, user_w, item_w):
        if not self._params_packed(user_w, item_w):
            eng = self._engine
            with torch.no_grad():
                eng.E0[:self.n_users].copy_(user_w)
                eng.E0[self.n_users:].copy_(item_w)

    def _apply(self, fn, *args, **kwargs):
        # .to(device)/.cuda()/.float() re-create parameter storage; re-pack afterwards
        super()._apply(fn, *args, **kwargs)
        if not self._params_packed():
            if self.embedding_user.weight.device.type != 'cuda':
                raise RuntimeError("lgcn_b200.LightGCN parameters must stay on the CUDA device")
            self._pack_params()
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
            out = _Propagate.apply(uw, iw, self)
            self._cache_key = None
            return out[:nu], out[nu:]
        key = self._param_key()
        if self._cache_key != key:
            self._sync_params_into_engine(uw, iw)
            self._engine.forward()
            self._cache_key = key
        out = self._engine.out
        return out[:nu], out[nu:]

    def getUsersRating(self, users):
        """Scores of every item for a batch of users (code/model.py:114-123) -> (B, m_items)."""
        with torch.no_grad():            # the reference only calls this under no_grad (code/Procedure.py:161,174)
            all_users, all_items = self.computer()
            users = users.to(self.device, dtype=torch.int64).contiguous()
            return ops.score_dense(all_users, all_items, users)

    def rank_topk(self, users, k, mask=True):
        """Fused getUsersRating + train-item mask (-1024) + top-k (code/Procedure.py:174-183):
        returns (item ids int64 [B,k], scores float32 [B,k]); the B x M matrix is never written."""
        with torch.no_grad():
            all_users, all_items = self.computer()
            users = users.to(self.device, dtype=torch.int64).contiguous()
            g = self._csr
            mi, mx = (g.indptr, g.indices) if mask else (None, None)
            if self.config.get('score_tensor_core', True):
                idx, val, self.last_rank_redone = ops.score_topk_tc(all_users, all_items, users, k, mi, mx, self.n_users)
                return idx, val
            return ops.score_topk(all_users, all_items, users, k, mi, mx, self.n_users)

    def getEmbedding(self, users, pos_items, neg_items):
        all_users, all_items = self.computer()
        u = all_users[users.long()]
        pos = all_items[pos_items.long()]
        neg = all_items[neg_items.long()]
        return u, pos, neg, all_users, all_items

    def bpr_loss(self, users, pos, neg):
        """(bpr, reg) as in code/model.py:162-183; both support .backward() through loss + decay*reg."""
        uw, iw = self.embedding_user.weight, self.embedding_item.weight
        users, pos, neg = (t.to(torch.int64).contiguous() for t in (users, pos, neg))
        if torch.is_grad_enabled() and (uw.requires_grad or iw.requires_grad):
            out = _Propagate.apply(uw, iw, self)
            self._cache_key = None
        else:
            au, _ = self.computer()
            out = self._engine.out
        return _BprLoss.apply(out, users, pos, neg, self)

    def forward(self, users, items):
        all_users, all_items = self.computer()
        return (all_users[users.long()] * all_items[items.long()]).sum(dim=1)

    # ------------------------------------------------------------------ fused training step
    def fused_train_step(self, users, pos, neg, lr=None, B_global=0):
        """stageOne without autograd: forward, BPR, backward and Adam in one captured sequence.
        Returns the engine (loss in engine.loss_out on the device)."""
        if not self._params_packed():
            self._pack_params()
        eng = self._engine
        if lr is not None:
            eng.set_lr(lr)
        eng.step(users, pos, neg, B_global)
        return eng

"""TEST INFRASTRUCTURE — call-for-call torch-CPU port of the reference's hot path.

The reference's arithmetic lives in torch (sparse.mm, autograd, optim.Adam, matmul, topk) and scipy;
this module issues the same library calls in the same order as LightGCN_work/code, without the
reference's global `world` module, so it can run on the GPU box's host cores as the CPU baseline
(`cpu_baseline.kind = "port"`) — /root/reference itself does not travel.  oracle/gen_golden.py
asserts, in the build container, that this port reproduces the real reference bit for bit.
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F
from torch import nn, optim


def build_graph(train_user, train_item, n_users, m_items):
    """dataloader.py:133-136 + :223-234 + :183-190,244.  The dok/lil block assignment is replaced by
    sp.bmat, which yields the same CSR (bit-identical, 0.11 s instead of 83.8 s);
    the normalisation lines are the reference's."""
    R = sp.csr_matrix((np.ones(len(train_user), dtype=np.float32), (train_user, train_item)), shape=(n_users, m_items))
    adj_mat = sp.bmat([[None, R], [R.T, None]], format='csr', dtype=np.float32)
    rowsum = np.array(adj_mat.sum(axis=1)).flatten()
    d_inv = np.power(rowsum, -0.5, where=rowsum != 0)
    d_inv[np.isinf(d_inv)] = 0.
    d_inv[rowsum == 0] = 0.
    D_inv = sp.diags(d_inv)
    norm_adj = D_inv.dot(adj_mat).dot(D_inv).tocsr()
    coo = norm_adj.tocoo().astype(np.float32)
    index = torch.stack([torch.from_numpy(coo.row).long(), torch.from_numpy(coo.col).long()], dim=0)
    graph = torch.sparse_coo_tensor(index, torch.from_numpy(coo.data).float(), torch.Size(coo.shape)).coalesce()
    return graph, norm_adj, R


class RefLightGCN(nn.Module):
    """model.py:37-231 restricted to the plain path (use_pop_gate=False, use_item_item=False)."""

    def __init__(self, n_users, m_items, latent_dim, n_layers, graph):
        super().__init__()
        self.n_users, self.m_items, self.latent_dim, self.n_layers = n_users, m_items, latent_dim, n_layers
        self.embedding_user = nn.Embedding(n_users, latent_dim)           # model.py:57-60
        self.embedding_item = nn.Embedding(m_items, latent_dim)
        nn.init.normal_(self.embedding_user.weight, std=0.1)
        nn.init.normal_(self.embedding_item.weight, std=0.1)
        self.Graph = graph

    def computer(self):                                                   # model.py:201-225
        all_emb = torch.cat([self.embedding_user.weight, self.embedding_item.weight], dim=0)
        embs = [all_emb]
        x = all_emb
        for _ in range(self.n_layers):
            x = torch.sparse.mm(self.Graph, x)
            embs.append(x)
        out = torch.mean(torch.stack(embs, dim=1), dim=1)
        return out[:self.n_users, :], out[self.n_users:, :]

    def getUsersRating(self, users):                                      # model.py:114-123
        all_users, all_items = self.computer()
        return torch.matmul(all_users[users], all_items.t())

    def bpr_loss(self, users, pos, neg):                                  # model.py:125-134,162-173
        all_users, all_items = self.computer()
        u, pos_e, neg_e = all_users[users], all_items[pos], all_items[neg]
        pos_scores = torch.sum(u * pos_e, dim=1)
        neg_scores = torch.sum(u * neg_e, dim=1)
        bpr = -torch.mean(F.logsigmoid(pos_scores - neg_scores))
        reg_loss = (0.5 * (u.norm(2).pow(2) + pos_e.norm(2).pow(2) + neg_e.norm(2).pow(2))) / float(u.shape[0])
        return bpr, reg_loss


class RefBPRLoss:                                                         # utils.py:38-64
    def __init__(self, model, decay, lr):
        self.model, self.weight_decay = model, decay
        self.opt = optim.Adam(model.parameters(), lr=lr)

    def stageOne(self, users, pos, neg):
        loss, reg_loss = self.model.bpr_loss(users, pos, neg)
        loss = loss + reg_loss * self.weight_decay
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.cpu().item()


def ref_test(model, test_dict, all_pos, topks, u_batch_size=100, hoist_computer=False):
    """Procedure.py:127-206 + :89-121 + utils.py:173-217.  Recomputes computer() per user batch like
    the reference unless hoist_computer (used only to keep the gowalla known-answer test short)."""
    max_K = max(topks)
    users = list(test_dict.keys())
    per_user = []
    topk_all = []
    with torch.no_grad():
        cached = model.computer() if hoist_computer else None
        for lo in range(0, len(users), u_batch_size):
            batch_users = users[lo:lo + u_batch_size]
            bu = torch.Tensor(batch_users).long()
            if cached is None:
                rating = model.getUsersRating(bu)
            else:
                rating = torch.matmul(cached[0][bu], cached[1].t())
            exclude_idx, exclude_items = [], []
            for i, u in enumerate(batch_users):
                items = all_pos[u]
                exclude_idx.extend([i] * len(items))
                exclude_items.extend(items)
            rating[exclude_idx, exclude_items] = -(1 << 10)
            _, topk = torch.topk(rating, k=max_K)
            topk = topk.numpy()
            topk_all.append(topk)
            for i, u in enumerate(batch_users):
                gt = test_dict[u]
                gts = set(gt)
                r = np.array([[1.0 if x in gts else 0.0 for x in topk[i]]], dtype=np.float32)
                pre, rec, nd = [], [], []
                for k in topks:
                    right = r[:, :k].sum(1)
                    rec.append(np.sum(right / np.array([len(gt)])))
                    pre.append(np.sum(right) / k)
                    tm = np.zeros((1, k)); tm[0, :min(k, len(gt))] = 1
                    idcg = np.sum(tm * 1. / np.log2(np.arange(2, k + 2)), axis=1)
                    dcg = np.sum(r[:, :k] * (1. / np.log2(np.arange(2, k + 2))), axis=1)
                    idcg[idcg == 0.] = 1.
                    nd.append(np.sum(dcg / idcg))
                per_user.append((np.array(pre), np.array(rec), np.array(nd)))
    res = {'precision': np.mean([p[0] for p in per_user], axis=0),
           'recall': np.mean([p[1] for p in per_user], axis=0),
           'ndcg': np.mean([p[2] for p in per_user], axis=0)}
    return res, np.concatenate(topk_all, axis=0)

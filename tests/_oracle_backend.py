"""CPU stand-in for the C-ABI backend of gat-pytorch_b200/partition.py -- TEST INFRASTRUCTURE.

Implements each backend method with dense torch/numpy maths in the reference's formulation (per-edge
gather / scatter), so the gloo world_size-2 test can exercise the partition plan and the collective
choreography without a GPU.  Mirrors the semantics documented in include/gat_b200.h.
"""
import torch

SLOPE, EPS = 0.01, 1e-8


class OracleBackend:
    def build_structure(self, edges_local, n_global):
        return {"src": edges_local[0].long(), "dst": edges_local[1].long(), "n": n_global}

    def n_edges(self, st):
        return int(st["src"].numel())

    def gemm(self, ta, tb, m, n, k, a, lda, b, ldb, c, ldc, act_b=False, mul_elu_grad=None):
        a2 = a.reshape(-1)[: (k if ta else m) * lda].view(-1, lda)
        b2 = b.reshape(-1)[: (n if tb else k) * ldb].view(-1, ldb)
        A = a2[:k, :m].T if ta else a2[:m, :k]
        B = b2[:n, :k].T if tb else b2[:k, :n]
        if act_b:
            B = torch.nn.functional.elu(B)
        out = A.double() @ B.double()
        if mul_elu_grad is not None:
            xs = mul_elu_grad[:m, :n].double()
            out = out * torch.where(xs > 0, torch.ones_like(xs), torch.exp(xs))
        c.reshape(-1)[: m * ldc].view(-1, ldc)[:m, :n] = out.float()

    def project(self, x, rows, f_in, w_p, dp, a_src, a_tgt, nh, wh, s_src, s_tgt, x_act=False):
        if x_act:
            x = torch.nn.functional.elu(x)
        self.gemm(False, True, rows, dp, f_in, x, x.stride(0), w_p, w_p.stride(0), wh, dp)
        if a_src is not None:       # const_attention has no score terms
            self.scores(wh, rows, dp, a_src, a_tgt, nh, s_src, s_tgt)

    def scores(self, wh, rows, dp, a_src, a_tgt, nh, s_src, s_tgt):
        s_src[:rows] = (wh[:rows].double() @ a_src.double().T).float()
        s_tgt[:rows] = (wh[:rows].double() @ a_tgt.double().T).float()

    def _logits(self, st, plan, s_src_full, s_tgt_local):
        return s_src_full[st["src"]] + s_tgt_local[st["dst"] - plan.lo]

    def edge_max(self, st, plan, s_src_full, s_tgt_local, nh, gmax):
        l = self._logits(st, plan, s_src_full, s_tgt_local)
        if l.numel():
            gmax[0] = torch.maximum(gmax[0], l.max())

    def _alpha(self, st, plan, s_src_full, s_tgt_local, gmax, rows, nh):
        l = self._logits(st, plan, s_src_full, s_tgt_local)
        t = l - gmax
        p = torch.exp(torch.where(t >= 0, t, t * SLOPE))
        z = torch.zeros((rows, nh)).index_add_(0, st["dst"] - plan.lo, p)
        return l, p, z

    def edge_fwd(self, st, plan, wh_full, nh, fp, s_src_full, s_tgt_local, gmax, out_p, z, tie_dst, tie_src, tie_total, out_act=False,
                 p_drop=0.0, seed=0, alpha=None, const_attention=False):
        assert p_drop == 0.0, "the oracle backend has no Philox: dropout is covered by the GPU tests"
        rows = plan.rows
        if const_attention:         # gat_layer.py:89-92: e = 0 -> p = 1, Z = in-degree
            p = torch.ones((st["src"].numel(), nh))
            zz = torch.zeros((rows, nh)).index_add_(0, st["dst"] - plan.lo, p)
            l = None
        else:
            l, p, zz = self._alpha(st, plan, s_src_full, s_tgt_local, gmax, rows, nh)
        z[:rows] = zz
        al = p / (zz[st["dst"] - plan.lo] + EPS)
        if alpha is not None:
            alpha.copy_(al)
        msg = al[:, :, None] * wh_full[st["src"]].view(-1, nh, fp)
        out_p.zero_()
        out_p.view(-1, nh, fp).index_add_(0, st["dst"] - plan.lo, msg)
        if const_attention:
            return
        tie = (l == gmax).to(torch.int32)
        tie_dst.view(-1, nh).index_add_(0, st["dst"] - plan.lo, tie)
        tie_src.view(-1, nh).index_add_(0, st["src"], tie)
        tie_total.view(torch.int64)[0] = int(tie.sum())
        if out_act:
            out_p.copy_(torch.nn.functional.elu(out_p))

    def head_merge(self, out_p, rows, nh, f, fp, concat):
        o = out_p.view(-1, nh, fp)[:rows, :, :f]
        return o.reshape(rows, nh * f).clone() if concat else o.mean(dim=1)

    def head_mean_bwd_shared(self, go, rows, nh, f, fp):
        out = torch.zeros((max(rows, 1), fp))
        out[:rows, :f] = go[:rows] / nh
        return out

    @staticmethod
    def _per_head(go_p, nh, fp, go_shared):
        return go_p.view(-1, 1, fp).expand(-1, nh, fp).reshape(-1, nh * fp) if go_shared else go_p

    def edge_bwd_main(self, st, plan, wh_full, nh, fp, s_src_full, s_tgt_local, gmax, z_local, go_p, rec, d_wh, p_drop=0.0, seed=0,
                      const_attention=False):
        assert p_drop == 0.0
        rows = plan.rows
        dl = st["dst"] - plan.lo
        if const_attention:
            alpha = 1.0 / (z_local[:rows][dl] + EPS)
            go_e = go_p.view(-1, nh, fp)[dl]
            d_wh.zero_()
            d_wh.view(-1, nh, fp).index_add_(0, st["src"], alpha[:, :, None] * go_e)
            return
        l, p, _ = self._alpha(st, plan, s_src_full, s_tgt_local, gmax, rows, nh)
        alpha = p / (z_local[:rows][dl] + EPS)
        go_e = go_p.view(-1, nh, fp)[dl]
        d_alpha = (go_e * wh_full[st["src"]].view(-1, nh, fp)).sum(-1)
        e = alpha.size(0)
        rec[:e, :nh] = d_alpha
        rec[:e, nh:] = alpha
        d_wh.zero_()
        d_wh.view(-1, nh, fp).index_add_(0, st["src"], alpha[:, :, None] * go_e)

    def edge_bwd_fused(self, st, plan, wh_full, nh, fp, s_src_full, s_tgt_local, gmax, z_local, go_p, s_sum_local,
                       a_src, a_tgt, tie_dst, tie_src, corr, ds_src, ds_tgt, d_wh, push_ptrs=None, p_drop=0.0, seed=0, go_shared=False):
        assert p_drop == 0.0
        go_p = self._per_head(go_p, nh, fp, go_shared)
        assert push_ptrs is None   # peer-memory push is a CUDA-only path (the oracle backend has no recv_buffer)
        rec = torch.zeros((max(self.n_edges(st), 1), 2 * nh))
        self.edge_bwd_main(st, plan, wh_full, nh, fp, s_src_full, s_tgt_local, gmax, z_local, go_p, rec, d_wh)
        self.edge_bwd_finish(st, plan, nh, fp, rec, s_sum_local, a_src, a_tgt, tie_dst, tie_src, corr, ds_src, ds_tgt, d_wh)

    def edge_bwd_rowdot(self, plan, nh, fp, go_p, out_p, z_local, s_sum, ds_tgt, go_pre=None, s_tgt_local=None, go_shared=False):
        rows = plan.rows
        go_p = self._per_head(go_p, nh, fp, go_shared)
        if go_pre is not None:     # out_p holds h = ELU(out): recover out and apply ELU'
            h = out_p
            go_pre.copy_(go_p * torch.where(h > 0, torch.ones_like(h), h + 1.0))
            go_p, out_p = go_pre, torch.where(h > 0, h, torch.log1p(h.clamp(min=-1 + 1e-30)))
        s = (go_p.view(-1, nh, fp)[:rows] * out_p.view(-1, nh, fp)[:rows]).sum(-1)
        s_sum[:rows] = s
        ds_tgt.zero_()
        ds_tgt[:rows] = SLOPE * s * (EPS / (z_local[:rows] + EPS))
        return ds_tgt[:rows].double().sum().reshape(1)

    def edge_bwd_finish(self, st, plan, nh, fp, rec, s_sum_local, a_src, a_tgt, tie_dst, tie_src, corr, ds_src, ds_tgt, d_wh):
        e = st["src"].numel()
        dl = st["dst"] - plan.lo
        g = SLOPE * rec[:e, nh:] * (rec[:e, :nh] - s_sum_local[dl])
        ds_src.zero_()
        ds_src.index_add_(0, st["src"], g)
        ds_src -= tie_src.view(-1, nh).float() * corr
        ds_tgt[: plan.rows] -= tie_dst.view(-1, nh)[: plan.rows].float() * corr
        d_wh += ds_src @ a_src
        d_wh[plan.lo:plan.hi] += ds_tgt[: plan.rows] @ a_tgt

    def scores_bwd(self, wh, n, dp, nh, ds_src, ds_tgt, da_src, da_tgt):
        da_src.copy_((ds_src[:n].double().T @ wh[:n].double()).float())
        da_tgt.copy_((ds_tgt[:n].double().T @ wh[:n].double()).float())


"""B200 drop-in for the reference `MyBranchformerEncoderLayer`
(src/encoder/branchformer/encoder_layer.py:49-321): same constructor, attributes, parameter names
and forward contract; the arithmetic runs as the fused kernel sequence of engine.py."""
from __future__ import annotations

from typing import Optional

import torch

from ... import engine, ops
from ...espnet_compat import LayerNorm


class MyBranchformerEncoderLayer(torch.nn.Module):
    """Macaron-Branchformer block: 1/2 FFN -> {rel-pos MHSA || cgMLP} -> merge -> 1/2 FFN -> LN.

    Args mirror the reference (encoder_layer.py:67-79).  `attn` / `cgmlp` / `feed_forward*` are the
    parameter containers of espnet_compat.py.
    """

    def __init__(self, size: int, attn: Optional[torch.nn.Module], cgmlp: Optional[torch.nn.Module],
                 feed_forward_macaron: Optional[torch.nn.Module],
                 feed_forward: Optional[torch.nn.Module], dropout_rate: float, merge_method: str,
                 cgmlp_weight: float = 0.5, attn_branch_drop_rate: float = 0.0,
                 stochastic_depth_rate: float = 0.0):
        super().__init__()
        assert (attn is not None) or (cgmlp is not None), "At least one branch should be valid"
        self.size = size
        self.attn = attn
        self.cgmlp = cgmlp
        self.feed_forward_macaron = feed_forward_macaron
        self.feed_forward = feed_forward
        self.ff_scale = 1.0
        self.merge_method = merge_method
        self.cgmlp_weight = cgmlp_weight
        self.attn_branch_drop_rate = attn_branch_drop_rate
        self.stochastic_depth_rate = stochastic_depth_rate
        self.use_two_branches = (attn is not None) and (cgmlp is not None)

        if self.feed_forward_macaron is not None:
            self.ff_scale = 0.5
            self.norm_ff_macaron = LayerNorm(size)
        if attn is not None:
            self.norm_mha = LayerNorm(size)
        if cgmlp is not None:
            self.norm_mlp = LayerNorm(size)
        if self.feed_forward is not None:
            self.norm_ff = LayerNorm(size)
        self.norm_final = LayerNorm(size)
        self.dropout = torch.nn.Dropout(dropout_rate)

        if self.use_two_branches:
            if merge_method == "concat":
                self.merge_proj = torch.nn.Linear(size + size, size)
            elif merge_method == "learned_ave":
                self.pooling_proj1 = torch.nn.Linear(size, 1)
                self.pooling_proj2 = torch.nn.Linear(size, 1)
                self.weight_proj1 = torch.nn.Linear(size, 1)
                self.weight_proj2 = torch.nn.Linear(size, 1)
                self.merge_proj = torch.nn.Linear(size, size)
            elif merge_method == "fixed_ave":
                assert 0.0 <= cgmlp_weight <= 1.0, "cgmlp weight should be between 0.0 and 1.0"
                # a branch with zero weight is removed, like the reference (:135-142)
                if cgmlp_weight == 0.0:
                    self.use_two_branches = False
                    self.cgmlp = None
                    self.norm_mlp = None
                elif cgmlp_weight == 1.0:
                    self.use_two_branches = False
                    self.attn = None
                    self.norm_mha = None
                self.merge_proj = torch.nn.Linear(size, size)
            else:
                raise ValueError(f"unknown merge method: {merge_method}")
        else:
            self.merge_proj = torch.nn.Identity()

        self.weight_global = None
        self.weight_local = None
        self._packed = engine.PackedCache()

    # ---------------------------------------------------------------------------------------
    def _check_supported(self):
        if self.size != 256:
            raise NotImplementedError("the B200 row-complete GEMM epilogue is built for size=256")
        if self.attn is not None and (self.attn.d_k != 64 or self.attn.h * self.attn.d_k != self.size):
            raise NotImplementedError(
                f"the B200 attention kernel is built for head width d_k = 64 (attention_heads = "
                f"size / 64); got h={self.attn.h}, d_k={self.attn.d_k}")
        if self.feed_forward_macaron is None or self.feed_forward is None:
            raise NotImplementedError("macaron=False / missing FFN is not built (the reference "
                                      "itself crashes at encoder_layer.py:193 without macaron)")
        if self.training and not torch.is_grad_enabled() and (
                self.dropout.p > 0 or self.stochastic_depth_rate > 0 or self.attn_branch_drop_rate > 0):
            raise NotImplementedError("a no-grad call in train() mode with dropout / stochastic depth / "
                                      "branch drop enabled: the inference kernels have no random "
                                      "paths; use .eval() (or a grad-mode call: training.py)")

    def run(self, x: torch.Tensor, xn: torch.Tensor, pos_proj: Optional[torch.Tensor],
            lens: torch.Tensor, B: int, T: int, next_norm=None, next_norm_dtype=None):
        """Core of the block on 2-D activations.

        x: (B*T, d) block input (fp32 residual stream); xn: LN_ff_macaron(x) in operand storage
        (fp32, or bf16 in bf16 mode); pos_proj: linear_pos(pos_emb) (2T-1, d) with any row pitch.
        Returns (y, yn): y = block output (fp32), yn = next_norm(y) or None, stored as
        `next_norm_dtype` (default: operand storage; the encoder asks for fp32 when next_norm is
        its after_norm, i.e. the encoder output).
        """
        d = self.size
        M = B * T
        dev = x.device
        adt = engine.act_dtype()
        new = lambda dt=torch.float32: torch.empty((M, d), device=dev, dtype=dt)  # noqa: E731
        x_a = new()
        two = self.use_two_branches
        xa = new(adt) if self.attn is not None else None
        xm = new(adt) if self.cgmlp is not None else None
        lnA = (self.norm_mha.weight, self.norm_mha.bias) if self.attn is not None else None
        lnB = (self.norm_mlp.weight, self.norm_mlp.bias) if self.cgmlp is not None else None
        if lnA is None:  # cgMLP-only block: use slot A for norm_mlp
            engine.ffn_block(x, xn, self.feed_forward_macaron, out_main=x_a, lnA=lnB, out_lnA=xm,
                             cache=self._packed, key="ffm")
        else:
            engine.ffn_block(x, xn, self.feed_forward_macaron, out_main=x_a, lnA=lnA, out_lnA=xa,
                             lnB=lnB, out_lnB=xm, cache=self._packed, key="ffm")

        learned = two and self.merge_method == "learned_ave"
        concat = two and self.merge_method == "concat"
        x_b = new()
        xf = new(adt)
        lnF = (self.norm_ff.weight, self.norm_ff.bias)
        mp = self.merge_proj
        if pos_proj is None and self.attn is not None:
            raise NotImplementedError("attention without relative positional embedding "
                                      "(abs_pos / selfattn) is not built on the B200 path")
        fold = (engine.FOLD_MERGE and two and not concat and d == 256
                and self.cgmlp.channel_proj2.weight.shape[1] % 64 == 0)
        if fold:
            # ---- merge with the branch output projections folded in (:208-209, 220, 227-309):
            #   x + Wm (w1 (Wo ctx + bo) + w2 (W2 g + b2)) + bm
            # = x + w1 ((Wm Wo) ctx + Wm bo) + w2 ((Wm W2) g + Wm b2) + bm
            # and the pooling scores x1 . a = ctx . (Wo^T a) + bo . a, so x1 / x2 never exist.
            lo, p2 = self.attn.linear_out, self.cgmlp.channel_proj2
            f = self._packed.get(
                "fold", [mp.weight, lo.weight, lo.bias, p2.weight, p2.bias],
                lambda: (mp.weight.double() @ lo.weight.double(),
                         mp.weight.double() @ p2.weight.double(),
                         (mp.weight.double() @ lo.bias.double()).float().contiguous(),
                         (mp.weight.double() @ p2.bias.double()).float().contiguous()))
            fv = None
            if learned:
                pp1, wp1, pp2, wp2 = (self.pooling_proj1, self.weight_proj1, self.pooling_proj2,
                                      self.weight_proj2)
                fv = self._packed.get(
                    "foldv", [lo.weight, lo.bias, p2.weight, p2.bias, pp1.weight, pp1.bias, wp1.weight,
                              wp1.bias, pp2.weight, pp2.bias, wp2.weight, wp2.bias],
                    lambda: dict(
                        va1=(pp1.weight.double() @ lo.weight.double()).reshape(-1).float().contiguous(),
                        vb1=(wp1.weight.double() @ lo.weight.double()).reshape(-1).float().contiguous(),
                        va2=(pp2.weight.double() @ p2.weight.double()).reshape(-1).float().contiguous(),
                        vb2=(wp2.weight.double() @ p2.weight.double()).reshape(-1).float().contiguous(),
                        sc=[float(pp1.bias.double() + pp1.weight.double().reshape(-1) @ lo.bias.double()),
                            float(pp2.bias.double() + pp2.weight.double().reshape(-1) @ p2.bias.double()),
                            float(wp1.bias.double() + wp1.weight.double().reshape(-1) @ lo.bias.double()),
                            float(wp2.bias.double() + wp2.weight.double().reshape(-1) @ p2.bias.double())]))
            # the two branches are independent until the merge: with TAVSR_BRANCH_FORK=1 the cgMLP
            # chain runs on a side stream (a fork / join in the CUDA graph)
            pre = engine.branch_projections(xa, xm, self.attn, self.cgmlp, self._packed)
            qkv, g = pre if pre is not None else (None, None)
            with engine.branch_fork(dev) as side:
                with side:
                    u = engine.cgmlp_gated(xm, self.cgmlp, B, T, self._packed, "conv", g=g)
                ctx = engine.attention_ctx(xa, self.attn, pos_proj, lens, B, T, self._packed, "qkv",
                                           qkv=qkv)
            if learned:
                sc = fv["sc"]
                if engine.FUSE_SCORES and ctx.shape[1] == 256 and u.shape[1] == 1024 and T <= 2048:
                    # row dots + pooling softmax + 2-way softmax in one launch
                    w1, w2 = ops.merge_scores(ctx, u, fv["va1"], fv["vb1"], fv["va2"], fv["vb2"], lens,
                                              sc[0], sc[1], sc[2], sc[3], d, B, T)
                else:
                    d1, d2 = ops.row_dots(ctx, fv["va1"], fv["vb1"], u, fv["va2"], fv["vb2"])
                    w1, w2 = ops.merge_weights(d1, d2, lens, sc[0], sc[1], sc[2], sc[3], d, B, T)
                self.weight_global = w1.view(B, 1, 1)
                self.weight_local = w2.view(B, 1, 1)
            else:
                w1, w2 = self._packed.get(
                    "fixedw" + str((B, str(dev))), [mp.weight],
                    lambda: (torch.full((B,), 1.0 - self.cgmlp_weight, device=dev, dtype=torch.float32),
                             torch.full((B,), float(self.cgmlp_weight), device=dev, dtype=torch.float32)))
            mode = engine.compute_dtype()
            if mode == "tf32x3":
                wf = self._packed.get("foldw:x3", [f[0], f[1]], lambda: torch.cat(
                    [ops.split_tf32(f[0].float().contiguous(), "w"),
                     ops.split_tf32(f[1].float().contiguous(), "w")], 1).contiguous())
                xc, xu = ops.split_tf32(ctx, "x"), ops.split_tf32(u, "x")
            else:
                wf = self._packed.get("foldw:" + mode, [f[0], f[1]], lambda: torch.cat(
                    [f[0], f[1]], 1).to(adt).contiguous())
                xc, xu = ctx, u
            ops.gemm_rowln(xc, wf, mp.bias, x2=xu, k1=xc.shape[1], segbias=(f[2], f[3]),
                           rowscale=(w1, w2), rows_per_seg=T, residual=x_a, alpha=1.0, out_main=x_b,
                           lnA=lnF, out_lnA=xf)
        else:
            self._run_branches_unfolded(xa, xm, x_a, x_b, xf, pos_proj, lens, B, T, learned, concat)

        # ---- FFN + norm_final (+ the next block's first LayerNorm) ----
        y = new()
        yn = new(next_norm_dtype or adt) if next_norm is not None else None
        engine.ffn_block(x_b, xf, self.feed_forward, out_main=y,
                         ln0=(self.norm_final.weight, self.norm_final.bias),
                         lnA=next_norm, out_lnA=yn, round_lnA=False, cache=self._packed, key="ff")
        return y, yn

    def _run_branches_unfolded(self, xa, xm, x_a, x_b, xf, pos_proj, lens, B, T, learned, concat):
        """Branch output projections as their own row-complete GEMMs, then merge (the layout of the
        reference: encoder_layer.py:208-209, 220, 227-309).  Used for `concat`, single-branch
        blocks and with TAVSR_FOLD_MERGE=0."""
        d = self.size
        M = B * T
        dev = x_a.device
        two = self.use_two_branches
        adt = engine.act_dtype()
        # x1 / x2 only feed the merge projection: operand storage
        new = lambda: torch.empty((M, d), device=dev, dtype=adt)  # noqa: E731
        cat_buf = torch.empty((M, 2 * d), device=dev, dtype=adt) if two else None
        x1 = x2 = d1 = d2 = None
        if self.attn is not None:
            if pos_proj is None:
                raise NotImplementedError("attention without relative positional embedding "
                                          "(abs_pos / selfattn) is not built on the B200 path")
            ctx = engine.attention_ctx(xa, self.attn, pos_proj, lens, B, T, self._packed, "qkv")
            x1 = cat_buf[:, :d] if two else new()
            dots = None
            if learned:
                d1 = torch.empty((M, 2), device=dev, dtype=torch.float32)
                dots = (self.pooling_proj1.weight.reshape(-1), self.weight_proj1.weight.reshape(-1))
            lo = self.attn.linear_out
            engine.linear_rowln(ctx, lo.weight, lo.bias, self._packed, "lo", out_main=x1, dots=dots,
                                dots_out=d1)
        if self.cgmlp is not None:
            u = engine.cgmlp_gated(xm, self.cgmlp, B, T, self._packed, "conv")
            x2 = cat_buf[:, d:] if two else new()
            dots = None
            if learned:
                d2 = torch.empty((M, 2), device=dev, dtype=torch.float32)
                dots = (self.pooling_proj2.weight.reshape(-1), self.weight_proj2.weight.reshape(-1))
            p2 = self.cgmlp.channel_proj2
            engine.linear_rowln(u, p2.weight, p2.bias, self._packed, "p2", out_main=x2, dots=dots,
                                dots_out=d2)

        # ---- merge (:227-309) + norm_ff ----
        lnF = (self.norm_ff.weight, self.norm_ff.bias)
        mp = self.merge_proj
        if two and self.merge_method in ("learned_ave", "fixed_ave"):
            if learned:
                sc = self._packed.get("mscal", [self.pooling_proj1.bias, self.pooling_proj2.bias,
                                                self.weight_proj1.bias, self.weight_proj2.bias],
                                      lambda: [float(self.pooling_proj1.bias), float(self.pooling_proj2.bias),
                                               float(self.weight_proj1.bias), float(self.weight_proj2.bias)])
                w1, w2 = ops.merge_weights(d1, d2, lens, sc[0], sc[1], sc[2], sc[3], d, B, T)
                self.weight_global = w1.view(B, 1, 1)
                self.weight_local = w2.view(B, 1, 1)
            else:
                w1 = torch.full((B,), 1.0 - self.cgmlp_weight, device=dev, dtype=torch.float32)
                w2 = torch.full((B,), float(self.cgmlp_weight), device=dev, dtype=torch.float32)
            # (w1 x1 + w2 x2) Wm^T as the sequential dual product over [x1 | x2] . [Wm | Wm]^T
            mode = engine.compute_dtype()
            if mode == "tf32x3":
                wmm = self._packed.get("mp2:x3", [mp.weight], lambda: torch.cat(
                    [ops.split_tf32(mp.weight.detach(), "w")] * 2, 1).contiguous())
                xa1, xa2 = ops.split_tf32(x1.contiguous(), "x"), ops.split_tf32(x2.contiguous(), "x")
            else:
                wmm = self._packed.get("mp2:" + mode, [mp.weight], lambda: torch.cat(
                    [mp.weight.detach(), mp.weight.detach()], 1).to(adt).contiguous())
                xa1, xa2 = x1, x2
            ops.gemm_rowln(xa1, wmm, mp.bias, x2=xa2, k1=xa1.shape[1], rowscale=(w1, w2),
                           rows_per_seg=T, residual=x_a, alpha=1.0, out_main=x_b, lnA=lnF, out_lnA=xf)
        elif concat:
            engine.linear_rowln(cat_buf, mp.weight, mp.bias, self._packed, "mp", residual=x_a,
                                alpha=1.0, out_main=x_b, lnA=lnF, out_lnA=xf)
        else:
            xs = x2 if self.attn is None else x1
            if isinstance(mp, torch.nn.Identity):
                raise NotImplementedError("single-branch block built with merge_proj=Identity "
                                          "(use_attn/use_cgmlp=False) is not built on the B200 path")
            engine.linear_rowln(xs, mp.weight, mp.bias, self._packed, "mp", residual=x_a, alpha=1.0,
                                out_main=x_b, lnA=lnF, out_lnA=xf)

    # ---------------------------------------------------------------------------------------
    def forward(self, x_input, mask, cache=None):
        """Same contract as the reference forward (encoder_layer.py:153-166): takes `x` or
        `(x, pos_emb)` and the (B,1,T) mask, returns the same structure."""
        if cache is not None:
            raise NotImplementedError("cache is not None, which is not tested")
        if isinstance(x_input, tuple):
            x, pos_emb = x_input[0], x_input[1]
        else:
            x, pos_emb = x_input, None
        self._check_supported()
        engine.require_cuda(x)
        B, T, d = x.shape
        x2 = x.reshape(B * T, d).contiguous().float()
        lens = engine.lens_from_mask(mask, B, T, x.device)
        from ... import training
        if training.wants_grad(self, x):
            y = training.run_block(self, x2, B, T, lens, pos_emb).view(B, T, d)
            return ((y, pos_emb), mask) if pos_emb is not None else (y, mask)
        xn = ops.layernorm(x2, self.norm_ff_macaron.weight, self.norm_ff_macaron.bias, eps=1e-12,
                           out_dtype=engine.act_dtype())
        pos_proj = None
        if pos_emb is not None and self.attn is not None:
            pos_proj = engine.pos_projection(self.attn, pos_emb.float(), self._packed)
        y, _ = self.run(x2, xn, pos_proj, lens, B, T)
        y = y.view(B, T, d)
        if pos_emb is not None:
            return (y, pos_emb), mask
        return y, mask

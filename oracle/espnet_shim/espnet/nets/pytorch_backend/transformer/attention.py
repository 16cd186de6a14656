"""TEST INFRASTRUCTURE (oracle): espnet attention modules restated (Appendix A.4 / A.5)."""
import math

import torch


class MultiHeadedAttention(torch.nn.Module):
    def __init__(self, n_head, n_feat, dropout_rate):
        super().__init__()
        assert n_feat % n_head == 0
        self.d_k = n_feat // n_head
        self.h = n_head
        self.linear_q = torch.nn.Linear(n_feat, n_feat)
        self.linear_k = torch.nn.Linear(n_feat, n_feat)
        self.linear_v = torch.nn.Linear(n_feat, n_feat)
        self.linear_out = torch.nn.Linear(n_feat, n_feat)
        self.attn = None
        self.dropout = torch.nn.Dropout(p=dropout_rate)

    def forward_qkv(self, query, key, value):
        n = query.size(0)
        q = self.linear_q(query).view(n, -1, self.h, self.d_k).transpose(1, 2)
        k = self.linear_k(key).view(n, -1, self.h, self.d_k).transpose(1, 2)
        v = self.linear_v(value).view(n, -1, self.h, self.d_k).transpose(1, 2)
        return q, k, v

    def forward_attention(self, value, scores, mask):
        n = value.size(0)
        if mask is not None:
            mask = mask.unsqueeze(1).eq(0)
            scores = scores.masked_fill(mask, torch.finfo(scores.dtype).min)
            self.attn = torch.softmax(scores, dim=-1).masked_fill(mask, 0.0)
        else:
            self.attn = torch.softmax(scores, dim=-1)
        x = torch.matmul(self.dropout(self.attn), value)
        x = x.transpose(1, 2).contiguous().view(n, -1, self.h * self.d_k)
        return self.linear_out(x)

    def forward(self, query, key, value, mask):
        q, k, v = self.forward_qkv(query, key, value)
        scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(self.d_k)
        return self.forward_attention(v, scores, mask)


class RelPositionMultiHeadedAttention(MultiHeadedAttention):
    def __init__(self, n_head, n_feat, dropout_rate, zero_triu=False):
        super().__init__(n_head, n_feat, dropout_rate)
        self.zero_triu = zero_triu
        self.linear_pos = torch.nn.Linear(n_feat, n_feat, bias=False)
        self.pos_bias_u = torch.nn.Parameter(torch.Tensor(self.h, self.d_k))
        self.pos_bias_v = torch.nn.Parameter(torch.Tensor(self.h, self.d_k))
        torch.nn.init.xavier_uniform_(self.pos_bias_u)
        torch.nn.init.xavier_uniform_(self.pos_bias_v)

    def rel_shift(self, x):
        b, h, t, n = x.size()
        zero_pad = torch.zeros((b, h, t, 1), device=x.device, dtype=x.dtype)
        x_padded = torch.cat([zero_pad, x], dim=-1).view(b, h, n + 1, t)
        x = x_padded[:, :, 1:].view_as(x)[:, :, :, : n // 2 + 1]
        if self.zero_triu:
            ones = torch.ones((x.size(2), x.size(3)), device=x.device)
            x = x * torch.tril(ones, x.size(3) - x.size(2))[None, None, :, :]
        return x

    def forward(self, query, key, value, pos_emb, mask):
        q, k, v = self.forward_qkv(query, key, value)
        q = q.transpose(1, 2)  # (B, T, h, d_k)
        p = self.linear_pos(pos_emb).view(pos_emb.size(0), -1, self.h, self.d_k).transpose(1, 2)
        q_u = (q + self.pos_bias_u).transpose(1, 2)
        q_v = (q + self.pos_bias_v).transpose(1, 2)
        ac = torch.matmul(q_u, k.transpose(-2, -1))
        bd = self.rel_shift(torch.matmul(q_v, p.transpose(-2, -1)))
        scores = (ac + bd) / math.sqrt(self.d_k)
        return self.forward_attention(v, scores, mask)


class LegacyRelPositionMultiHeadedAttention(MultiHeadedAttention):
    """Dormant alternative; present so the reference imports resolve."""

    def forward(self, *a, **k):
        raise NotImplementedError("legacy_rel_selfattn is outside the restated surface")

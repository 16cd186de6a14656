"""Drop-in for `src/ctc/interctc_residual_module.py` (InterCTCResidualModule `:3-16`).

Same parameters (`proj_1` Linear(D,V), `proj_2` Linear(V,D)), same return `(x + proj_2(softmax(
proj_1(x))), logits)`.  Two launches on the B200 path: the fp32 CTC head kernel (logits +
softmax) and the vocabulary-residual kernel; inference only, CUDA only.
"""
import torch

from .. import engine, ops


class InterCTCResidualModule(torch.nn.Module):
    def __init__(self, dim_model, vocab_size):
        super().__init__()
        self.proj_1 = torch.nn.Linear(dim_model, vocab_size)
        self.proj_2 = torch.nn.Linear(vocab_size, dim_model)

    def forward(self, x):
        engine.require_inference(self, x)
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous().float()
        _, prob, _, logits = ops.ctc_head(x2, self.proj_1.weight, self.proj_1.bias, want_logp=False,
                                          want_prob=True, want_logits=True)
        out, _ = ops.vocab_residual(x2, prob, self.proj_2.weight.contiguous(), self.proj_2.bias)
        return out.view(shape), logits.view(*shape[:-1], -1)

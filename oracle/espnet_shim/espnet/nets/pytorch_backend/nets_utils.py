"""TEST INFRASTRUCTURE (oracle): restatement of the espnet==202402 helpers the reference imports at
src/encoder/branchformer/encoder.py:20,48.  Parity unpinned: espnet itself is not installable here
(SURVEY.md §8c); semantics follow SURVEY.md Appendix A.2 / A.8."""
import torch


def make_pad_mask(lengths, xs=None, length_dim=-1, maxlen=None):
    """Bool mask (B, Tmax), True at padded positions (Appendix A.8)."""
    if not isinstance(lengths, list):
        lengths = lengths.long().tolist()
    bs = len(lengths)
    if maxlen is None:
        maxlen = int(max(lengths))
    seq = torch.arange(0, maxlen, dtype=torch.int64).unsqueeze(0).expand(bs, maxlen)
    lens = torch.tensor(lengths, dtype=torch.int64).unsqueeze(-1)
    return seq >= lens


class Swish(torch.nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def get_activation(act):
    table = {
        "hardtanh": torch.nn.Hardtanh,
        "tanh": torch.nn.Tanh,
        "relu": torch.nn.ReLU,
        "selu": torch.nn.SELU,
        "swish": Swish,
        "gelu": torch.nn.GELU,
    }
    return table[act]()


def rename_state_dict(old_prefix, new_prefix, state_dict):
    for k in [k for k in state_dict if k.startswith(old_prefix)]:
        state_dict[k.replace(old_prefix, new_prefix)] = state_dict.pop(k)

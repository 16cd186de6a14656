"""TEST INFRASTRUCTURE (oracle): espnet Conv2dSubsampling restated (Appendix A.8)."""
import torch


class TooShortUttError(Exception):
    def __init__(self, message, actual_size, limit):
        super().__init__(message)
        self.actual_size = actual_size
        self.limit = limit


class Conv2dSubsampling(torch.nn.Module):
    """(B, T, idim) -> (B, ((T-1)//2-1)//2, odim): two 3x3 stride-2 convs + ReLU, Linear, pos-enc."""

    def __init__(self, idim, odim, dropout_rate, pos_enc=None):
        super().__init__()
        self.conv = torch.nn.Sequential(
            torch.nn.Conv2d(1, odim, 3, 2), torch.nn.ReLU(),
            torch.nn.Conv2d(odim, odim, 3, 2), torch.nn.ReLU())
        self.out = torch.nn.Sequential(
            torch.nn.Linear(odim * (((idim - 1) // 2 - 1) // 2), odim), pos_enc)

    def forward(self, x, x_mask):
        x = self.conv(x.unsqueeze(1))
        b, c, t, f = x.size()
        x = self.out(x.transpose(1, 2).contiguous().view(b, t, c * f))
        if x_mask is None:
            return x, None
        return x, x_mask[:, :, :-2:2][:, :, :-2:2]


class _Unrestated(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("this subsampling variant is not used by any shipped config")


class Conv1dSubsampling2(_Unrestated):
    pass


class Conv1dSubsampling3(_Unrestated):
    pass


class Conv2dSubsampling1(_Unrestated):
    pass


class Conv2dSubsampling2(_Unrestated):
    pass


class Conv2dSubsampling6(_Unrestated):
    pass


class Conv2dSubsampling8(_Unrestated):
    pass


def check_short_utt(ins, size):
    if isinstance(ins, Conv2dSubsampling) and size < 7:
        return True, 7
    return False, -1

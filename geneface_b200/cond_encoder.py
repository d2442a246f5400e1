"""Condition encoders + bias-free MLP (drop-in for modules/radnerfs/cond_encoder.py: identical
sub-module names so reference state_dicts load: encoder_conv.{0,2,4,6}, encoder_fc1.{0,2},
attentionConvNet.{0,2,4,6,8}, attentionNet.0, net.{i})."""
import torch
import torch.nn as nn
import torch.nn.functional as F

_STRIDES = {1: (1, 1, 1, 1), 2: (2, 1, 1, 1), 3: (2, 2, 1, 1), 4: (2, 2, 1, 1), 16: (2, 2, 2, 2)}


class AudioNet(nn.Module):
    """cond_encoder.py:7-52: 4x conv1d(k3,p1)+LeakyReLU(.02), then 64->64->dim_aud."""

    def __init__(self, dim_in=29, dim_aud=64, win_size=16):
        super().__init__()
        if win_size not in _STRIDES:
            raise ValueError("unsupported win_size")
        self.win_size, self.dim_aud = win_size, dim_aud
        chans = (dim_in, 32, 32, 64, 64)
        layers = []
        for k, s in enumerate(_STRIDES[win_size]):
            layers += [nn.Conv1d(chans[k], chans[k + 1], kernel_size=3, stride=s, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.encoder_conv = nn.Sequential(*layers)
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, dim_aud))

    def forward(self, x):
        x = self.encoder_conv(x.permute(0, 2, 1)).squeeze(-1)
        return self.encoder_fc1(x).squeeze()


class AudioAttNet(nn.Module):
    """cond_encoder.py:55-89: attention weights over the smoothing window."""

    def __init__(self, in_out_dim=64, seq_len=8):
        super().__init__()
        self.seq_len, self.in_out_dim = seq_len, in_out_dim
        chans = (in_out_dim, 16, 8, 4, 2, 1)
        layers = []
        for k in range(5):
            layers += [nn.Conv1d(chans[k], chans[k + 1], kernel_size=3, stride=1, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.attentionConvNet = nn.Sequential(*layers)
        self.attentionNet = nn.Sequential(nn.Linear(seq_len, seq_len, bias=True), nn.Softmax(dim=1))

    def forward(self, x):
        y = x[:, :self.in_out_dim].permute(1, 0).unsqueeze(0)
        y = self.attentionConvNet(y)
        y = self.attentionNet(y.view(1, self.seq_len)).view(self.seq_len, 1)
        return torch.sum(y * x, dim=0)


class MLP(nn.Module):
    """cond_encoder.py:92-111: bias-free Linear stack, ReLU between layers."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.net = nn.ModuleList([
            nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=False)
            for l in range(num_layers)])

    # 'tc': in training (grad enabled) run forward / grad_input / grad_weight on the gf_tl_* tcgen05 operators (tc_linear.py; fp16 operands, fp32
    # accumulation = the reference's `amp: true` arithmetic) when the layer widths are inside their envelope; 'torch': library GEMMs under autograd.
    # Set per model from hparams['train_mlp_backend'] (default: the GF_TRAIN_MLP environment variable, else 'torch').
    backend = 'torch'

    def _dims(self):
        return [self.dim_in] + [self.dim_hidden] * (self.num_layers - 1) + [self.dim_out]

    def forward(self, x):
        """x: [M, dim_in], or a list of tensors standing for torch.cat(x, dim=1) (the tensor-core backend packs the parts without materialising the
        concatenation; an expand()-ed row stays a broadcast)"""
        parts = list(x) if isinstance(x, (list, tuple)) else None
        x0 = parts[0] if parts else x
        if self.backend == 'tc' and x0.is_cuda and torch.is_grad_enabled() and x0.dim() == 2:
            from . import tc_linear
            if tc_linear.supported(self._dims()):
                return tc_linear.tc_mlp(parts if parts else x, [layer.weight for layer in self.net])
        if parts:
            x = torch.cat(parts, dim=1)
        for l, layer in enumerate(self.net):
            x = layer(x)
            if l != self.num_layers - 1:
                x = F.relu(x, inplace=True)
        return x

"""SaeConv — API shell of the reference's models/sae_conv.py (3x3 Conv+ReLU encoder / decoder pair).

The reference class is unreachable from its own training path (utils.py:2458-2459 rejects every SAE name other than
`sae_mlp` / `gated_sae`); the "Conv-SAE" of the hot path is SaeMLP's 4-D branch (pixels as tokens, see sae_mlp.py).
This shell keeps the constructor / forward signature and state_dict keys (`encoder.0.*`, `decoder.0.*`) so code that
instantiates it keeps working; its two convolutions run on cuDNN through torch.nn and are not part of libsvb.
"""
import torch.nn as nn


class SaeConv(nn.Module):
    def __init__(self, img_size, expansion_factor):
        super().__init__()
        c = img_size[0]
        self.encoder = nn.Sequential(nn.Conv2d(c, c * expansion_factor, kernel_size=3, stride=1, padding=1), nn.ReLU())
        self.decoder = nn.Sequential(nn.Conv2d(c * expansion_factor, c, kernel_size=3, stride=1, padding=1), nn.ReLU())

    def forward(self, x):
        encoded = self.encoder(x)
        return encoded, self.decoder(encoded)   # 2-tuple, as in the reference (sae_conv.py:34-39)

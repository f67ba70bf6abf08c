from .sae_mlp import SaeMLP  # noqa: F401
from .gated_sae import GatedSae  # noqa: F401
from .sae_conv import SaeConv  # noqa: F401

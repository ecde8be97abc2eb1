"""``Cosine`` activation with the reference's interface (``src/activations.py:5-13``).  An elementwise ``torch.cos``:
it is not on the hot path (no experiment of the reference uses it, ``src/evaluation.py:37`` builds ReLU networks), so it
is a plain tensor op here too; inside a ``WHVINetwork`` it sees the MC samples folded into the batch like any foreign
module.  The activation that IS on the hot path, ``nn.ReLU`` between WHVI layers, is folded into the layer kernels
(``WHVINetwork._run``)."""
import torch
import torch.nn as nn


class Cosine(nn.Module):
    def __init__(self):
        """Cosine activation function."""
        super().__init__()

    def forward(self, x):
        return torch.cos(x)

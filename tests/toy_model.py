"""A tiny encoder whose parameter names follow the HF layout the reference groups by
(``model.embeddings.*``, ``model.encoder.layer.N.*``, ``model.pooler.*``) -- test helper shared by
tests/golden/make_golden.py (reference side) and the parity tests (our side)."""
import torch
from torch import nn


class _Block(nn.Module):
    def __init__(self, h: int, ffn: int):
        super().__init__()
        self.attention = nn.Linear(h, h)
        self.intermediate = nn.Linear(h, ffn)
        self.output = nn.Linear(ffn, h)
        self.LayerNorm = nn.LayerNorm(h)

    def forward(self, x):
        x = x + torch.tanh(self.attention(x))
        return self.LayerNorm(x + self.output(torch.relu(self.intermediate(x))))


class _Encoder(nn.Module):
    def __init__(self, layers: int, h: int, ffn: int):
        super().__init__()
        self.layer = nn.ModuleList([_Block(h, ffn) for _ in range(layers)])

    def forward(self, x):
        for blk in self.layer:
            x = blk(x)
        return x


class _Embeddings(nn.Module):
    def __init__(self, vocab: int, h: int):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, h)
        self.LayerNorm = nn.LayerNorm(h)

    def forward(self, ids):
        return self.LayerNorm(self.word_embeddings(ids))


class _Inner(nn.Module):
    def __init__(self, vocab, h, ffn, layers):
        super().__init__()
        self.embeddings = _Embeddings(vocab, h)
        self.encoder = _Encoder(layers, h, ffn)
        self.pooler = nn.Linear(h, h)

    def forward(self, ids):
        x = self.encoder(self.embeddings(ids))
        return torch.tanh(self.pooler(x.mean(dim=1)))


class ToyEncoder(nn.Module):
    """forward(batch: LongTensor (B, L)) -> (B, h) sequence embedding."""

    def __init__(self, vocab: int = 37, h: int = 24, ffn: int = 40, layers: int = 2):
        super().__init__()
        self.model = _Inner(vocab, h, ffn, layers)

    def forward(self, batch):
        return self.model(batch)


def make_toy_state_dicts(K: int, seed: int = 0, sigma: float = 1e-2, **kw):
    torch.manual_seed(seed)
    base = ToyEncoder(**kw)
    pre = {k: v.detach().clone() for k, v in base.state_dict().items()}
    fts = []
    for k in range(K):
        g = torch.Generator().manual_seed(seed + 1 + k)
        fts.append({n: (v + sigma * torch.randn(v.shape, generator=g)).contiguous() for n, v in pre.items()})
    return pre, fts

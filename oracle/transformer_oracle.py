"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's Transformer condition encoder in eval mode.

Follows /root/reference/src/bcnf/models/feature_network.py statement by statement:
  MultiHeadAttention.forward :207-229, TransformerBlock.forward :255-260, Transformer.forward :284-307.
Pinned (tests/test_oracle_golden.py) to the outputs of the live reference recorded in tests/golden/feature_networks.npz
(both Transformer fixtures, one with positional embeddings) and tests/golden/transformer_grads.npz.  Only tests/ may
import this module; the package never does.
"""
from __future__ import annotations

import math

import numpy as np


def _erf(x: np.ndarray) -> np.ndarray:
    try:
        from scipy.special import erf
        return erf(x)
    except ImportError:                                   # pragma: no cover
        return np.vectorize(math.erf)(x)


def gelu(x: np.ndarray) -> np.ndarray:
    """nn.GELU() (exact erf form), the activation of the block's FFN (feature_network.py:240-244)."""
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def layer_norm(x: np.ndarray, weight: np.ndarray, bias: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm over the last axis: biased variance, eps inside the square root."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    return (x - mean) / np.sqrt(var + eps) * weight + bias


def linear(x: np.ndarray, sd: dict, prefix: str) -> np.ndarray:
    return x @ sd[prefix + ".weight"].T + sd[prefix + ".bias"]


def attention(x: np.ndarray, sd: dict, prefix: str, n_heads: int) -> np.ndarray:
    """MultiHeadAttention.forward(x, x, x), mask=None (:207-229)."""
    B, T, E = x.shape
    hd = E // n_heads

    def heads(name: str) -> np.ndarray:                   # :212-214  .view(B, -1, heads, hd).transpose(1, 2)
        return linear(x, sd, f"{prefix}.{name}").reshape(B, T, n_heads, hd).transpose(0, 2, 1, 3)
    q, k, v = heads("q_linear"), heads("k_linear"), heads("v_linear")
    scores = q @ k.transpose(0, 1, 3, 2) / np.sqrt(hd)     # :217
    scores = scores - scores.max(axis=-1, keepdims=True)
    w = np.exp(scores)
    w = w / w.sum(axis=-1, keepdims=True)                  # :223 softmax over the keys
    out = (w @ v).transpose(0, 2, 1, 3).reshape(B, T, E)   # :226-227
    return linear(out, sd, f"{prefix}.fc_out")             # :228


def positional_table(seq_len: int, trf_size: int, input_size: int) -> np.ndarray:
    """:291-299: only the first `input_size` channels are filled."""
    pe = np.zeros((seq_len, trf_size))
    for i in range(seq_len):
        for j in range(input_size):
            ang = i / 10000 ** (2 * j / input_size)
            pe[i, j] = np.sin(ang) if j % 2 == 0 else np.cos(ang)
    return pe


def transformer_forward(sd: dict, x: np.ndarray, *, n_heads: int, n_blocks: int, input_size: int,
                        add_positional_embeddings: bool = False, dtype=np.float64) -> np.ndarray:
    """Transformer.forward in eval mode (both nn.Dropout are identities): (B, T, input_size) -> (B, output_size)."""
    sd = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    h = linear(np.asarray(x, dtype=dtype), sd, "features")                     # :285
    if add_positional_embeddings:                                              # :288-301
        h = h + positional_table(h.shape[1], h.shape[2], input_size).astype(dtype)
    for l in range(n_blocks):                                                  # :303-304
        p = f"layers.{l}"
        a = attention(h, sd, f"{p}.attention", n_heads)                        # :256
        h = layer_norm(h + a, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"])  # :257
        f = linear(gelu(linear(h, sd, f"{p}.ffn.0")), sd, f"{p}.ffn.2")        # :258
        h = layer_norm(h + f, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"])  # :259
    return linear(h[:, 0, :], sd, "output")                                    # :309-311: first token

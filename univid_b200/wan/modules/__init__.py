from .attention import attention, flash_attention
from .model import (WanAttentionBlock, WanCrossAttention, WanLayerNorm, WanModel, WanRMSNorm,
                    WanSelfAttention, rope_apply, rope_params, sinusoidal_embedding_1d)

__all__ = [
    'WanModel', 'WanAttentionBlock', 'WanSelfAttention', 'WanCrossAttention', 'WanRMSNorm',
    'WanLayerNorm', 'rope_params', 'rope_apply', 'sinusoidal_embedding_1d', 'flash_attention',
    'attention',
]

from .model_animate import WanAnimateCrossAttention, WanAnimateSelfAttention

__all__ = ['WanAnimateSelfAttention', 'WanAnimateCrossAttention']

from .sequence_parallel import sp_attn_forward, sp_dit_forward  # noqa: F401
from .ulysses import distributed_attention  # noqa: F401
from .util import (all_gather, all_to_all, gather_forward, get_rank, get_world_size,  # noqa: F401
                   init_distributed_group)

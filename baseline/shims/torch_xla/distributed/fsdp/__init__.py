"""XlaFullyShardedDataParallel / checkpoint_module on stock PyTorch FSDP (bf16 mixed precision) + NCCL."""
import torch
from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import CheckpointImpl, checkpoint_wrapper
from torch.distributed.fsdp import BackwardPrefetch, FullyShardedDataParallel, MixedPrecision, ShardingStrategy


class XlaFullyShardedDataParallel(FullyShardedDataParallel):
    def __init__(self, module, reshard_after_forward=True, flatten_parameters=True, **kwargs):
        strategy = ShardingStrategy.FULL_SHARD if reshard_after_forward else ShardingStrategy.SHARD_GRAD_OP
        super().__init__(
            module,
            sharding_strategy=strategy,
            mixed_precision=MixedPrecision(param_dtype=torch.bfloat16, buffer_dtype=torch.bfloat16),
            backward_prefetch=BackwardPrefetch.BACKWARD_PRE,
            device_id=torch.cuda.current_device(),
            use_orig_params=False,
            limit_all_gathers=True,
        )

    def get_shard_metadata(self):
        return {"world_size": self.world_size, "rank": self.rank, "impl": "torch.distributed.fsdp"}


def checkpoint_module(module):
    return checkpoint_wrapper(module, checkpoint_impl=CheckpointImpl.NO_REENTRANT)

"""b200fa — B200-native (sm_100a) flash attention behind the launcher interface of
FSSRepo/ggml-cuda-experiments (flash_attn_ext_f16 / flash_attn_row + fa_reduce).

The product is the C-ABI library built from csrc/ (include/b200fa.h).  This package is the thin host
side used by tests and bench.py: ctypes bindings over that ABI, taking torch tensors only as owners of
device memory.  There is no CPU or PyTorch compute path: importing works anywhere, but every call
raises if libb200fa.so is missing or no sm_100 device is present.
"""
from .api import (  # noqa: F401
    FLAG_CAUSAL, FLAG_NO_TCGEN05, FLAG_WORKSPACE_ZEROED, TYPE_F16, TYPE_F32, TYPE_Q8_0, B200FAError, ExtParams, PlanInfo, PLAN_PREFILL, PLAN_ROWS16, PLAN_STREAM, plan, Workspace, dequantize_q8_0,
    flash_attn_ext, flash_attn_ext_raw, flash_attn_partial, flash_attn_partial_scatter, flash_attn_seqpar, kv_cache_append, merge_partials_wait, PeerExchange, last_dispatch, last_launch_count, lib, merge_partials,
    quantize_q8_0, workspace_size)
from .build import build, build_example  # noqa: F401
from .tensor_io import capture_paths, read_tensor, replay_capture, tensor_info, write_tensor  # noqa: F401
from .sharding import (  # noqa: F401
    HeadShard, SeqShard, flash_attn_ext_head_parallel, flash_attn_ext_seq_parallel, head_shard, seq_shard, slice_heads)

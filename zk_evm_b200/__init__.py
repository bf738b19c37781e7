"""zk_evm_b200 — B200-native STARK proving path for zk_evm's evm_arithmetization (host-side mirror over libzkgpu)."""
from ._lib import ZkGpuError, StarkConfig, KernelLabels, lib, declared_symbols  # noqa: F401
from .prover import (Context, PolynomialBatch, CtlData, StarkProof, table_info, get_ctl_data, prove_single_table,  # noqa: F401
                     set_debug, DeviceTrace, keccak_generate_trace, logic_generate_trace, upload_trace, arithmetic_generate_range_checks, memory_finish_trace, PinnedArray, host_register, host_unregister)
from .segment import (Challenger, AllProof, prove_with_traces, prove_with_commitments, upload_traces, SegmentUpload, prove_with_traces_sharded, ZkGpuBackend, TorchComm, LocalComm,  # noqa: F401
                      default_owner, segment_challenges, ShardPlan, shard_plan, SplitCommit, NUM_TABLES, TABLE_NAMES, OPTIONAL_TABLES)
from .public_values import PublicValues, flatten_public_values, memory_extra_looking_values, memory_extra_looking_sum  # noqa: F401,E402
from .scheduler import SegmentProver, SegmentAborted, estimate_segment_bytes  # noqa: F401,E402

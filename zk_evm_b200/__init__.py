"""zk_evm_b200 — B200-native STARK proving path for zk_evm's evm_arithmetization (host-side mirror over libzkgpu)."""
from ._lib import ZkGpuError, StarkConfig, KernelLabels, lib, declared_symbols  # noqa: F401
from .prover import Context, PolynomialBatch  # noqa: F401

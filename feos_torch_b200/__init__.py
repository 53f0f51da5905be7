"""feos_torch_b200 -- B200-native (sm_100a CUDA) implementation of feos-torch's batched
phase-equilibrium hot path, exporting the reference's public names
(`/root/reference/feos_torch/__init__.py:1-4`)."""
from .gc_pcsaft import GcPcSaft, GcPcSaftMix
from .native import PcSaft
from .pcsaft_mix import PcSaftMix
from .pcsaft_pure import PcSaftPure

__version__ = "0.1.0"
__all__ = ["PcSaft", "GcPcSaft", "PcSaftPure", "PcSaftMix", "GcPcSaftMix"]

"""starfish_b200 -- B200-native (sm_100a CUDA) replacement for the kinetic-particle hot path of
particleincell/Starfish: KineticMaterial.updateFields() = particle move + mesh deposit.

The compute lives in ``libstarfish_gpu.so`` (C ABI: ``include/sfgpu.h``).  This package is the thin
host-side mirror of the reference's Java interface for that path (``KineticMaterial``, ``UniformMesh``,
field access) used by the parity tests and ``bench.py``.  There is no CPU fallback: importing works
anywhere, but creating a :class:`KineticMaterial` needs the built library and a CUDA device.
"""
from .domain import DomainType, Face, DomainBoundaryType, UniformMesh  # noqa: F401
from .kinetic_material import KineticMaterial, Particles, SfgpuError  # noqa: F401

__all__ = ["DomainType", "Face", "DomainBoundaryType", "UniformMesh", "KineticMaterial", "Particles", "SfgpuError"]

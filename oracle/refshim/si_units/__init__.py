"""TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for the `si_units` package.

The reference (`/root/reference/feos_torch/pcsaft_pure.py:4`, `pcsaft_mix.py:4`,
`gc_pcsaft.py:5`) only ever forms dimensionless combinations of these constants, so
plain floats in SI base units reproduce its arithmetic exactly (CODATA-2018 exact values).
"""
KELVIN = 1.0
PASCAL = 1.0
MOL = 1.0
METER = 1.0
JOULE = 1.0
KILO = 1.0e3
KB = 1.380649e-23
NAV = 6.02214076e23
ANGSTROM = 1.0e-10

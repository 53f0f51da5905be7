"""Input sets of the reference's own tests, restated as data (inputs only -- the reference tests
hold no expected numbers; those come from ``tests/golden/*.json``, see ``oracle/make_golden.py``).

Sources (all under /root/reference):
  README_PARAMS / README_T      README.md:13-17 (golden output README.md:26-29)
  PURE_PARAMS                   tests/test_pcsaft_pure.py:10-17 (T = 300 K, rho = 0.001, p = 1 bar)
  MIX_PARAMS / MIX_KIJ          tests/test_pcsaft_mix.py:17-36 (T = 300 K, rho = [0.001, 0.002])
  MIX_BUBBLE / MIX_DEW          tests/test_pcsaft_mix.py:127-151, :195-215
  GC_SEGMENTS / GC_BONDS / ...  tests/test_gc_pcsaft.py:17-44, :130-141, :177-188
"""

README_PARAMS = [1.5, 3.5, 250.0, 0.0, 0.03, 1500.0, 1.0, 1.0]
README_T = [250.0, 300.0, 350.0, 400.0, 450.0]
README_VP = [20693.5960, 216164.6184, 1049770.6187, 3281855.9640, 7875531.7021]
README_GRAD = [-6.7923e04, -1.7737e04, -7.0413e02, 0.0, -5.7458e05, -6.9122e01, -3.6892e04, -3.6892e04]

PURE_PARAMS = [
    [1.5, 3.2, 350, 0, 0, 0, 0, 0],
    [1.5, 3.2, 150, 2.5, 0.03, 2500, 2, 1],
    [1.5, 3.2, 150, 2.5, 0, 2500, 1, 1],
    [1.5, 3.2, 150, 2.5, 0.03, 0, 1, 1],
    [1.5, 3.2, 150, 2.5, 0, 0, 0, 0],
    [1.5, 3.2, 150, 2.5, 0.03, 2500, 0, 2],
]

_A = [1.5, 3.2, 150]
_B = [2.5, 3.5, 250]
MIX_PARAMS = [
    [_A + [0, 0, 0, 0, 0], _B + [0, 0, 0, 0, 0]],
    [_A + [2.5, 0, 0, 0, 0], _B + [0, 0, 0, 0, 0]],
    [_A + [0, 0, 0, 0, 0], _B + [2, 0, 0, 0, 0]],
    [_A + [2.5, 0, 0, 0, 0], _B + [2, 0, 0, 0, 0]],
    [_A + [0, 0.03, 2500, 2, 1], _B + [0, 0, 0, 0, 0]],
    [_A + [0, 0, 0, 0, 0], _B + [0, 0.025, 1500, 1, 2]],
    [_A + [0, 0.03, 2500, 1, 1], _B + [0, 0.025, 1500, 1, 1]],
    [_A + [2.5, 0.03, 2500, 1, 1], _B + [2, 0.025, 1500, 1, 1]],
    [_A + [0, 0.03, 2500, 1, 1], _B + [0, 0.025, 1500, 0, 1]],
    [_A + [0, 0.03, -500, 0, 2], _B + [0, 0.025, 1500, 1, 1]],
    [_A + [0, 0, 0, 0, 0], _B + [0, 0.025, 1500, 0, 1]],
    [_A + [0, 0.03, 2500, 2, 2], _B + [0, 0.025, 1500, 1, 1]],
    [_A + [0, 0.03, 2500, 2, 2], _B + [0, 0.025, 1500, 1, 1]],
    [_A + [0, 0.03, 2500, 1, 2], _B + [0, 0.025, 1500, 2, 1]],
]
MIX_KIJ = [[-0.05, 0.0] for _ in MIX_PARAMS]
MIX_KIJ[12][1] = 3000.0
MIX_LABELS = [
    "np/np", "p/np", "np/p", "p/p", "a/np", "np/a", "a/a", "pa/pa", "a/ia", "ia/a", "np/ia", "aa/a", "aa/a eps_aibj", "aa/aa",
]

MIX_BUBBLE = dict(
    params=[[1, 3.5, 150, 0, 0.02, 1500, 1, 1], [1, 3.5, 200, 0, 0.03, 2500, 1, 1]],
    kij=[-0.15, 1000.0], T=150.0, z=0.5, p0=1e5,
)
MIX_DEW = dict(
    params=[[1, 3.5, 150, 0, 0, 0, 0, 0], [1, 3.5, 200, 0, 0, 0, 0, 0]],
    kij=[-0.15, 0.0], T=150.0, z=0.5, p0=1e5,
)

_ISO = ["CH3", ">CH", "CH3"]
_NEO = ["CH3", ">C<", "CH3"]
GC_SEGMENTS = [
    [["CH3", "CH2", "CH2", "CH3"], ["CH3", "CH2", "CH3"]],
    [_ISO + ["CH3"], _NEO + ["CH3", "CH3"]],
    [_ISO + ["CH=O"], _NEO + ["CH3", "CH3"]],
    [_ISO + ["CH3"], _NEO + ["CH3", "HCOO"]],
    [_ISO + ["CH=O"], _NEO + ["CH3", "HCOO"]],
    [_ISO + ["OH"], _NEO + ["CH3", "CH3"]],
    [_ISO + ["CH3"], _NEO + ["CH3", "NH2"]],
    [_ISO + ["OH"], _NEO + ["CH3", "NH2"]],
    [["CH3", ">CH", "CH=O", "OH"], _NEO + ["HCOO", "NH2"]],
    [["CH3", ">CH", "CH=O", "OH"], _NEO + ["CH2", "IA"]],
    [["CH3", ">CH", "CH=O", "IA"], _NEO + ["CH2", "IA"]],
]
_STAR3 = [[0, 1], [1, 2], [1, 3]]
_STAR4 = [[0, 1], [1, 2], [1, 3], [1, 4]]
GC_BONDS = [[[[0, 1], [1, 2], [2, 3]], [[0, 1], [1, 2]]]] + [[_STAR3, _STAR4] for _ in range(10)]
GC_KAB = [("CH3", "CH=O", 0.03), (">CH", "HCOO", -0.01)]
GC_PHI = [1.1, 0.98]

GC_POINT = dict(
    segments=[["CH3", "CH2", "CH2", "CH3"], ["CH3", "CH2", "CH3"]],
    bonds=[[[0, 1], [1, 2], [2, 3]], [[0, 1], [1, 2]]],
    kab=[("CH3", "CH2", -0.15)], phi=[1.1, 0.98], T=150.0, z=0.5, p0=1e5,
)

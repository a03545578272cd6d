"""Material ids, classes and constitutive parameters of the FluidEngine hot path.

Data restated from fluidlab/configs/macros.py:1-17 (ids), :37-41 (classes), :65-83 (class map),
:131-201 (friction / mu / lambda / rho), :207-216 (dtype, EPS, NOWHERE).  One table row per
material instead of five parallel dicts; the dict views below keep the reference names.
"""
import numpy as np

MAT_LIQUID, MAT_PLASTO_ELASTIC, MAT_ELASTIC, MAT_RIGID, MAT_PLASTO_ELASTIC_DEMO = 200, 201, 202, 203, 204

#  id  symbol             name               class                       mu      lambda  rho
_MATERIALS = [
    (0,  "WATER",          "water",          MAT_LIQUID,                 0.0,    277.78, 1.0),
    (1,  "MILK",           "milk",           MAT_LIQUID,                 0.0,    277.78, 0.5),
    (2,  "COFFEE",         "coffee",         MAT_LIQUID,                 0.0,    277.78, 1.0),
    (3,  "ELASTIC",        "elastic",        MAT_ELASTIC,                416.67, 277.78, 1.0),
    (4,  "ICECREAM",       "ice-cream",      MAT_PLASTO_ELASTIC,         416.67, 277.78, 0.5),
    (5,  "RIGID",          "rigid",          MAT_RIGID,                  416.67, 277.78, 1.0),
    (6,  "RIGID_HEAVY",    "rigid-heavy",    MAT_RIGID,                  416.67, 277.78, 10.0),
    (7,  "RIGID_LIGHT",    "rigid-light",    MAT_RIGID,                  416.67, 277.78, 0.5),
    (8,  "MILK_VIS",       "milk-viscous",   MAT_LIQUID,                 200.0,  277.78, 1.0),
    (9,  "COFFEE_VIS",     "coffee-viscous", MAT_LIQUID,                 200.0,  277.78, 1.0),
    (10, "ELASTIC_DEMO",   "elastic-demo",   MAT_ELASTIC,                10.0,   100.0,  1.0),
    (11, "PLASTIC_DEMO",   "plastic-demo",   MAT_PLASTO_ELASTIC_DEMO,    160.0,  277.78, 1.0),
    (12, "INVISCID_DEMO",  "inviscid-demo",  MAT_LIQUID,                 0.0,    277.78, 5.0),
    (13, "VISCOUS_DEMO",   "viscous-demo",   MAT_LIQUID,                 800.0,  277.78, 5.0),
    (14, "INVISCID_DEMO2", "inviscid-demo2", MAT_LIQUID,                 0.0,    277.78, 1.0),
    (15, "INVISCID_DEMO3", "inviscid-demo3", MAT_LIQUID,                 0.0,    277.78, 3.0),
    (16, "ICECREAM1",      "ice-cream1",     MAT_PLASTO_ELASTIC,         216.67, 277.78, 0.5),
]
for _id, _sym, *_ in _MATERIALS:
    globals()[_sym] = _id

MAT_NAME = {m[0]: m[2] for m in _MATERIALS}
MAT_CLASS = {m[0]: m[3] for m in _MATERIALS}
MU = {m[0]: m[4] for m in _MATERIALS}
LAMDA = {m[0]: m[5] for m in _MATERIALS}
RHO = {m[0]: m[6] for m in _MATERIALS}

# collider materials (mesh statics / effectors), macros.py:19-30,131-141
CUP, TANK, LADDLE, POURER, DISPENSER, CONE, ROBOT, BOTTLE, PILLAR, STIRRER, PLATE, BOWL = range(50, 62)
FRICTION = {CUP: 0.5, TANK: 0.5, BOWL: 0.0, LADDLE: 0.1, CONE: 8.0, BOTTLE: 0.1, PILLAR: 0.0, STIRRER: 8.0, PLATE: 0.1}

DTYPE_NP = np.float32
EPS = 1e-12
NOWHERE = [-100.0, -100.0, -100.0]

"""Joint naming and skeleton topology shared with the reference (ref:cs_vit/constants.py).

Only what the hot path reads is restated: the 21-joint target order and the 20 bones used by
``mean_connection_length`` to de-normalise the root translation (ref:cs_vit/net/ti_poser.py:591).
"""
_FINGERS = ("Thumb", "Index", "Middle", "Ring", "Pinky")

# wrist first, then each finger from base (1) to tip (4)            ref:cs_vit/constants.py:72-94
TARGET_JOINTS_ORDER = ("Wrist",) + tuple(f"{f}_{k}" for f in _FINGERS for k in range(1, 5))

# wrist -> five finger bases, then the three bones of each finger    ref:cs_vit/constants.py:96-121
TARGET_JOINTS_CONNECTION = [(0, 1 + 4 * f) for f in range(5)] + [
    (1 + 4 * f + k, 2 + 4 * f + k) for f in range(5) for k in range(3)
]

# MANO's own 16-joint order (ref:cs_vit/constants.py:53-70)
MANO_JOINTS_ORDER = ("Wrist",) + tuple(f"{f}_{k}" for f in ("Index", "Middle", "Pinky", "Ring", "Thumb") for k in range(1, 4))

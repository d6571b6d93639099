"""Integer ranges of the fields the structures store: counters are C int32 (array('i') in the reference's
Count-Min sketch), the elements-added footer is int64.  Python ints do not wrap, so the shims clamp to
these before anything crosses the C ABI (the reference clamps at the same values: probables/constants.py)."""

INT32_T_MIN, INT32_T_MAX = -(1 << 31), (1 << 31) - 1
INT64_T_MIN, INT64_T_MAX = -(1 << 63), (1 << 63) - 1
UINT32_T_MAX, UINT64_T_MAX = (1 << 32) - 1, (1 << 64) - 1

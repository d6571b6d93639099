"""integer limits of the counter / footer fields (same names as probables/constants.py:3-8)"""

INT32_T_MIN = -(2**31)
INT32_T_MAX = 2**31 - 1
INT64_T_MIN = -(2**63)
INT64_T_MAX = 2**63 - 1
UINT32_T_MAX = 2**32 - 1
UINT64_T_MAX = 2**64 - 1

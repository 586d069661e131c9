"""
safelife_b200 -- B200-native batched step engine behind SafeLife's own API.

    safelife_b200.speedups   drop-in mirror of the reference `safelife.speedups`
    safelife_b200.env        BatchedSafeLifeEnv (SafeLifeEnv reset/step, batched)
    safelife_b200.levels     level files -> device level pool
    safelife_b200._lib       ctypes loader / nvcc build of libsafelife_b200.so

Importing the package does not need a GPU; calling into the library does.
"""
__version__ = "0.1.0"

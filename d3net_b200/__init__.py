"""B200-native PointGroup proposal ops: a drop-in for D3Net's lib/pointgroup_ops.

    from d3net_b200 import pointgroup_ops        # the reference's Python operator API
    import d3net_b200.PG_OP as PG_OP             # the reference's native-module surface

The kernels live in d3net_b200/csrc (sm_100a only) behind the C ABI of include/pg_b200.h and are
loaded from d3net_b200/libpg_b200.so; there is no CPU or Triton fallback.
"""
__version__ = "0.1.0"

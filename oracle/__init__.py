"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's (BienLuky/EDA-DM) quantized-UNet algorithm.  Only `tests/`,
`__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs of `bench.py` may import
this package, and only as the checker or the timed CPU baseline -- never on the product path.
"""

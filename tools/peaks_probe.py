"""Pipe peaks incl. the register-operand FFMA / FFMA2 patterns and FFMA2 vs resident warps (developer tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import engine  # noqa: E402

ctx = engine.Context(0)
for n, w in (("ffma const-operand", 0), ("ffma2 const-operand", 1), ("ffma register outer product", 5),
             ("ffma2 register outer product", 6)):
    print(f"{n:32s} {ctx.measure_peak(w):7.2f} TFLOP/s")
for k in range(1, 9):
    print(f"ffma2 outer, {k} warps/scheduler   {ctx.measure_peak(9 + k):7.2f} TFLOP/s")

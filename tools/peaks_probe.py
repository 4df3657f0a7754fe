import sys; sys.path.insert(0,'/root/repo')
from strugatzki_b200 import engine
ctx=engine.Context(0)
for n,w in (("ffma",0),("ffma2",1),("ffma_outer",5),("ffma2_outer",6)):
    print(n, round(ctx.measure_peak(w),2))

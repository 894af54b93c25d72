# E9: four queries per batched full-search round instead of three (a block of <= 64 queries then needs ONE round per warp).
p='csrc/lm.cuh'
s=open(p).read()
a="constexpr int LM_KNN_BATCH = 3;"
assert s.count(a)==1
s=s.replace(a,"constexpr int LM_KNN_BATCH = 4;")
open(p,'w').write(s)

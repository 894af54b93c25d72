# E1: merge / centroid sub-tiles of up to 2560 entries (five per thread), a block's tile cut into EQUAL sub-tiles
# (0.62 M entries on 148 blocks = 4190 per block: two rounds of 2095 instead of 2048 + 2048 + 94).
p='csrc/map_incremental.cuh'
s=open(p).read()
a="constexpr int INC_SUB = 2048;"
assert s.count(a)==1
s=s.replace(a,"constexpr int INC_SUB = 2560;")
a="""  for (int a = a0; a < b0 || (a == a0 && a0 >= b0); a += INC_SUB) {  // (one pass with an empty range when the window is empty)
    const int b = min(b0, a + INC_SUB), cnt = max(0, b - a);"""
assert s.count(a)==1
s=s.replace(a,"""  const int m_nsub = max(1, (b0 - a0 + INC_SUB - 1) / INC_SUB), m_step = max(1, (b0 - a0 + m_nsub - 1) / m_nsub);
  for (int a = a0; a < b0 || (a == a0 && a0 >= b0); a += m_step) {  // (one pass with an empty range when the window is empty)
    const int b = min(b0, a + m_step), cnt = max(0, b - a);""")
a="LL_TL_SUB(11 + min(4, (a - a0) / INC_SUB));"
assert s.count(a)==1
s=s.replace(a,"LL_TL_SUB(11 + min(4, (a - a0) / m_step));")
a="""  for (int base = tile_begin; base < tile_end; base += INC_SUB) {
    const int sub_n = min(INC_SUB, tile_end - base);"""
assert s.count(a)==1
s=s.replace(a,"""  const int r_nsub = max(1, (tile_end - tile_begin + INC_SUB - 1) / INC_SUB), r_step = max(1, (tile_end - tile_begin + r_nsub - 1) / r_nsub);
  for (int base = tile_begin; base < tile_end; base += r_step) {
    const int sub_n = min(r_step, tile_end - base);""")
a="LL_TL_SUB(21 + min(4, (base - tile_begin) / INC_SUB));"
assert s.count(a)==1
s=s.replace(a,"LL_TL_SUB(21 + min(4, (base - tile_begin) / r_step));")
open(p,'w').write(s)

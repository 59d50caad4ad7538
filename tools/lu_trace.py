"""Summarise an LB200_TRACE_LU timeline (stderr of a traced run): per step, GEMM time, panel time and the idle gaps of
the update stream.  usage: LB200_TRACE_LU=1 python tools/prof_fact.py getrf 32768 2 2> trace.txt; python tools/lu_trace.py trace.txt"""
import sys, collections
recs = collections.defaultdict(dict)
runs = []
for line in open(sys.argv[1]):
    if line.startswith("LU_TRACE step"):
        recs = collections.defaultdict(dict); runs.append(recs); continue
    if not line.startswith("LU_TRACE"): continue
    _, step, kind, a, b = line.split()
    recs[int(step)][int(kind)] = (float(a), float(b))
recs = runs[-1]
tot_gemm = tot_gap = 0.0
prev_end = None
print("step  g0      g1      g2      g3      g4     panel  prep0  gap_before_g0 gap_g0_g1 gap_g1_g2 gap_g2_g3 gap_g3_g4  (ms)")
for st in sorted(recs):
    r = recs[st]
    g = [r.get(q) for q in range(5)]
    dur = lambda x: (x[1] - x[0]) if x else 0.0
    gap0 = (g[0][0] - prev_end) if (prev_end is not None and g[0]) else 0.0
    gap1 = (g[1][0] - g[0][1]) if g[1] else 0.0
    gap2 = (g[2][0] - g[1][1]) if g[2] else 0.0
    gap3 = (g[3][0] - g[2][1]) if g[3] else 0.0
    gap4 = (g[4][0] - g[3][1]) if g[4] else 0.0
    last = [x for x in g if x][-1]
    prev_end = last[1]
    tot_gemm += sum(dur(x) for x in g); tot_gap += gap0 + gap1 + gap2 + gap3 + gap4
    if st % 4 == 0 or st > 56 or len(recs) < 20:
        print(f"{st:3d} {dur(g[0]):7.3f} {dur(g[1]):7.3f} {dur(g[2]):7.3f} {dur(g[3]):7.3f} {dur(g[4]):7.3f} {dur(r.get(20)):6.3f} {dur(r.get(10)):6.3f} {gap0:9.3f} {gap1:9.3f} {gap2:9.3f} {gap3:9.3f} {gap4:9.3f}")
print(f"total GEMM {tot_gemm:.1f} ms, total update-stream gaps {tot_gap:.1f} ms, end {prev_end:.1f} ms")

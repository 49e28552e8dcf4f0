import csv, sys, subprocess
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
def g(name):
    return r[hdr.index(name)] if name in hdr else "n/a"
print("kernel:", g("Kernel Name")[:70])
for m in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
          "sm__cycles_elapsed.max", "smsp__average_warp_latency_per_inst_issued.ratio"]:
    print("  %-70s %s" % (m, g(m)))
print("  stalls per issue:")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(r[i])
        if v > 0.1: print("     %-30s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
iS, iW, iE = h2.index("Source"), h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed")
seen = set(); data = []
for rr in rows[2:]:
    if len(rr) > iE and rr[iW].isdigit():
        key = rr[0]
        if key in seen: continue
        seen.add(key); data.append((int(rr[iW]), int(rr[iE] or 0), rr[iS].strip()))
tot = sum(d[0] for d in data); totE = sum(d[1] for d in data)
cs, ce = Counter(), Counter()
for w, e, s in data:
    t = s.split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    cs[op] += w; ce[op] += e
print("  by opcode: stall%% | exec%% (total warp-inst %.3g)" % totE)
for op, w in cs.most_common(14): print("     %-10s %5.1f%% | %5.1f%%" % (op, 100*w/tot, 100*ce[op]/totE))
print("  top stall lines:")
for w, e, s in sorted(data, reverse=True)[:14]: print("     %5.2f%% %10d  %s" % (100*w/tot, e, s[:80]))

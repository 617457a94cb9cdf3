"""Digest an .ncu-rep (read here, no GPU): headline raw metrics, stall breakdown, dynamic SASS mix, hottest source lines."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]


def page(*args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("--page", "raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = dict(zip(hdr, zip(vals, units)))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "sm__cycles_elapsed.max"]
print("## raw metrics")
for k in keys:
    if k in m:
        print("* `%s` = %s %s" % (k, m[k][0], m[k][1]))
st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for h, v in zip(hdr, vals)
      if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v.replace(".", "").isdigit()}
tot = sum(st.values()) or 1
print("\n## warp stall sampling (all samples)")
for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]:
    print("* %s: %.1f %%" % (k, 100 * v / tot))
src = page("--page", "source", "--print-source", "cuda,sass")
dyn, lines = collections.Counter(), []
total = 0
for r in src[3:]:
    if len(r) < 8:
        continue
    if r[0].strip().isdigit():
        try:
            lines.append((int(r[0]), r[1].strip()[:100], int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
    elif r[2].startswith("0x"):
        mm = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[3].strip())
        if mm:
            n = int(r[7] or 0)
            dyn[mm.group(2)] += n
            total += n
print("\n## dynamic SASS mix (% of warp instructions executed)")
for op, n in dyn.most_common(14):
    print("* %s: %.2f %%" % (op, 100 * n / max(total, 1)))
ts = sum(l[2] for l in lines) or 1
print("\n## hottest source lines (share of stall samples | share of instructions)")
for ln, s, sm, ex in sorted(lines, key=lambda x: -x[2])[:14]:
    print("* L%d %.1f %% | %.1f %% : `%s`" % (ln, 100 * sm / ts, 100 * ex / max(total, 1), s))

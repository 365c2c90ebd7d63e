"""Summarises an ncu report (here, without a GPU): key raw metrics, SASS opcode histogram weighted by executions, the
shared-memory wavefront budget per opcode.  usage: python tools/ncu_summary.py <report.ncu-rep> <out prefix> [units]
Writes <out prefix>.txt (human) and <out prefix>.json (bench.py reads `traffic` from it).  `units` = work units per launch
(e.g. evaluator batches) used for the per-unit columns."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, unit_row, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit_row, vals)}
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "derived__memory_l1_wavefronts_shared_excessive", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier"]
lines = ["ncu summary of %s" % rep, ""]
js = {"report": rep}
WANT += [k for k in hdr if "pipe_tensor" in k and k not in WANT]
for k in WANT:
    if k in m and m[k][0] != "":
        lines.append("%-86s %s %s" % (k, m[k][0], m[k][1]))
        js[k] = m[k][0]


def to_bytes(key):
    v, u = m.get(key, ("", ""))
    x = num(v)
    if x is None:
        return None
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
js["dram_bytes_read_per_launch"], js["dram_bytes_write_per_launch"] = rd, wr
js["traffic_bytes_per_launch"] = None if rd is None or wr is None else rd + wr
lines += ["", "DRAM traffic per launch: read %s B + write %s B" % (rd, wr)]

src = list(csv.reader(io.StringIO(ncu("source"))))
h = src[1]
ix = {k: i for i, k in enumerate(h)}
ops, smem = {}, {}
total = 0.0
for r in src[2:]:
    if len(r) < len(h):
        continue
    s = r[ix["Source"]].strip().split()
    if not s:
        continue
    op = (s[1] if s[0].startswith("@") and len(s) > 1 else s[0]).rstrip(";")
    ex = num(r[ix["Instructions Executed"]]) or 0.0
    w = (num(r[ix["L1 Wavefronts Shared"]]) or 0.0) if "L1 Wavefronts Shared" in ix else 0.0
    base = op.split(".")[0]
    ops.setdefault(base, [0.0, 0])
    ops[base][0] += ex
    ops[base][1] += 1
    total += ex
    if w:
        smem[op] = smem.get(op, 0.0) + w
lines += ["", "SASS opcode histogram (warp-level executions; per unit = / %g)" % units,
          "%-12s %16s %8s %12s %8s" % ("opcode", "executed", "share", "per unit", "static")]
for k, (ex, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:48]:
    lines.append("%-12s %16.0f %7.2f%% %12.1f %8d" % (k, ex, 100 * ex / max(total, 1), ex / units, n))
js["sass_static_counts"] = {k: ops[k][1] for k in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "MEMBAR", "USETMAXREG") if k in ops}
js["sass_executed"] = {k: ops[k][0] for k in js["sass_static_counts"]}
lines += ["", "proof-of-path opcodes (static instances / executions): " +
          ", ".join("%s %d / %.0f" % (k, ops[k][1], ops[k][0]) for k in js["sass_static_counts"])]
lines += ["", "shared-memory wavefronts by LSU opcode (per unit):"]
for k, w in sorted(smem.items(), key=lambda kv: -kv[1]):
    lines.append("%-12s %16.0f %12.1f" % (k, w, w / units))
open(out + ".txt", "w").write("\n".join(lines) + "\n")
json.dump(js, open(out + ".json", "w"), indent=1)
print("\n".join(lines[:40]))

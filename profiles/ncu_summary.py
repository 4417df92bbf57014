"""Summarise an .ncu-rep: key raw metrics per kernel and the hottest SASS lines with stall reasons.
Usage: python profiles/ncu_summary.py report.ncu-rep [kernel-regex] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
       "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = rows[0], rows[1]
print("== raw metrics (units in brackets) ==")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(name[:90])
    for k in RAW:
        if k in hdr:
            print(f"    {k:70s} {r[hdr.index(k)]:>16s} [{units[hdr.index(k)]}]")
args = ["--page", "source", "--csv"] + (["--kernel-name", f"regex:{kre}"] if kre else [])
rows = list(csv.reader(io.StringIO(run(args))))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = [r[1]]
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
seen = set()
for b in blocks:
    name, hdr, data = b[0], b[1], b[2:]
    if name in seen:
        continue
    seen.add(name)
    si = hdr.index("# Samples")
    ie = hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in data)
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall}
    print(f"\n== {name[:100]} ==\ninstructions {len(data)}, executed {sum(int(r[ie] or 0) for r in data)}, samples {tot}")
    print("stall totals:", sorted(((k, v) for k, v in agg.items() if v), key=lambda x: -x[1])[:8])
    for r in sorted(data, key=lambda r: -int(r[si] or 0))[:22]:
        st = sorted(((hdr[i], int(r[i] or 0)) for i in stall if int(r[i] or 0)), key=lambda x: -x[1])[:3]
        print(f"  {int(r[si] or 0):6d}  {r[1].strip()[:72]:72s} {st}")

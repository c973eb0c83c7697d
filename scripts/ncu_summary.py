"""Summarise an ncu report: headline raw metrics + opcode mix + hottest source regions.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg']
for r in rows[2:]:
    print("== kernel:", r[hdr.index('Kernel Name')][:100])
    for k in keys:
        if k in hdr:
            print("  %-70s %s %s" % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
    if units:
        ins = float(r[hdr.index('smsp__inst_executed.sum')].replace(',', ''))
        print("  thread-instructions per unit: %.1f" % (ins * 32 / units))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si, ii, wi = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
swi = h.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in h else None
swx = h.index('L1 Wavefronts Shared Excessive') if 'L1 Wavefronts Shared Excessive' in h else None
smem_w = smem_x = 0
data = []
for r in rows[2:]:
    if len(r) < 10 or r[0] in ('Kernel Name', 'Address'):
        if len(data) > 50: break
        continue
    data.append((r[si].strip(), int(r[ii]), int(r[wi])))
    if swi is not None and r[swi].isdigit():
        smem_w += int(r[swi]); smem_x += int(r[swx]) if (swx is not None and r[swx].isdigit()) else 0
tot = sum(d[1] for d in data); samp = sum(d[2] for d in data)
if smem_w:
    print("shared-memory wavefronts per launch: %d (excessive: %d)" % (smem_w, smem_x) + ("  per unit: %.3f" % (smem_w / units) if units else ""))
opc = collections.Counter(); ops = collections.Counter()
for s, i, w in data:
    t = s.split()
    op = t[1] if t[0].startswith('@') else t[0]
    opc[op.split('.')[0]] += i; ops[op.split('.')[0]] += w
print("opcode mix (executed %, stall-sample %" + (", per unit" if units else "") + "):")
for k, v in opc.most_common(22):
    print("  %-10s %6.2f%% %6.2f%%" % (k, 100 * v / tot, 100 * ops[k] / max(samp, 1)) + ("  %8.1f" % (v * 32 / units) if units else ""))

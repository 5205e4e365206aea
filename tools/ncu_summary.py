"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [kernel-regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None
def page(p):
    cmd = ["ncu", "-i", rep, "--page", p, "--csv"] + (["--kernel-name", "regex:" + kre] if kre else [])
    return list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
rows = page("raw"); hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum']
for r in rows[2:]:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w: print(f"{w} = {r[i]} {units[i]}")
    print("---")
rows = page("source")
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith('0x')]
first = data[0][0]; ends = [i for i, r in enumerate(data) if r[0] == first] + [len(data)]
d = data[ends[0]:ends[1]]
tot = sum(int(r[ix['# Samples']]) for r in d)
print('SASS instructions', len(d), 'samples', tot, 'warp-instr executed', sum(int(r[ix['Instructions Executed']]) for r in d))
agg = {c: sum(int(r[ix[c]] or 0) for r in d) for c in stall}
for c, v in sorted(agg.items(), key=lambda x: -x[1])[:9]: print(f"  {c:28s} {100*v/tot:5.1f}%")
print("top instructions by samples:")
for r in sorted(d, key=lambda r: -int(r[ix['# Samples']]))[:28]:
    st = max(stall, key=lambda c: int(r[ix[c]] or 0))
    print(f"  {100*int(r[ix['# Samples']])/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s} {st:22s} {r[ix['Source']].strip()[:80]}")

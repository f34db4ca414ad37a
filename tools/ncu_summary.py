#!/usr/bin/env python
"""Summarise an ncu report (raw page + per-instruction warp-state samples): python tools/ncu_summary.py report.ncu-rep "title" out.txt"""
import csv,sys,re,subprocess
f,title,out=sys.argv[1],sys.argv[2],sys.argv[3]
raw=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; 
lines=['# '+title]
want=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__cluster_dim_x','launch__registers_per_thread','sm__cycles_active.avg','sm__cycles_elapsed.avg.per_second','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum.per_second','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_tc_wavefronts_mem_shared.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes.sum.per_second']
for vals in rows[2:]:
    name=vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    lines.append('## kernel: '+name[:110])
    for h,v,u in zip(hdr,vals,rows[1]):
        if h in want: lines.append(f'{h:75s} {v:>16s} {u}')
try:
    src=subprocess.run(['ncu','-i',f,'--page','source','--csv'],capture_output=True,text=True).stdout
    r=list(csv.reader(src.splitlines()))
    if len(r)>3:
        hdr=r[1]; data=r[2:]
        isrc=hdr.index('Source'); isamp=hdr.index('# Samples'); iexec=hdr.index('Instructions Executed')
        stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        tot=sum(int(x[isamp]) for x in data)
        lines.append(f'## warp-state samples (last kernel): total {tot}; instructions with >= 0.6 % of the samples')
        for idx,x in enumerate(data):
            s=int(x[isamp])
            if tot and s>tot*0.006:
                st=sorted([(int(x[i]),h) for i,h in stall_cols if int(x[i])>0],reverse=True)[:2]
                lines.append(f'{idx:6d} {x[isrc].strip()[:64]:64s} {s:6d} {100*s/tot:5.1f}%  exec {x[iexec]:>9s}  {st}')
except Exception as ex:
    lines.append('## (source page: several kernels in this report; see the raw metrics above)')
open(out,'w').write('\n'.join(lines)+'\n')
print(out, len(lines))

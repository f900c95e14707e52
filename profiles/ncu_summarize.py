import csv, sys, subprocess, collections, re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__grid_size','launch__block_size','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fp64.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum','lts__t_sector_hit_rate.pct']
stalls=[h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')]
for r in rows[2:]:
    print('=====')
    for w in want:
        for i,h in enumerate(hdr):
            if h==w: print(' ',w,'=',r[i],units[i])
    st=[]
    for i,h in enumerate(hdr):
        if h in stalls:
            try: st.append((float(r[i]),h.replace('smsp__pcsamp_warps_issue_stalled_','')))
            except: pass
    st.sort(reverse=True); tot=sum(x for x,_ in st)
    print('  stalls:', ', '.join(f'{n} {100*x/tot:.0f}%' for x,n in st[:9]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
kern=None; data=collections.defaultdict(list)
for r in csv.reader(src.splitlines()):
    if r and r[0]=='Kernel Name': kern=r[1][:70]; continue
    if kern and len(r)>6 and re.match(r'^(0x)?[0-9a-f]+$', r[0] or 'x'): data[kern].append(r)
for k,v in data.items():
    ops=collections.Counter(); stall=collections.Counter(); tot=0
    for r in v:
        m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1].strip()); op=m.group(2).split('.')[0] if m else r[1][:8]
        try: n=int(r[5])
        except: n=0
        ops[op]+=n; tot+=n
        try: stall[op]+=int(r[2])
        except: pass
    print('==',k,len(v),'sass lines; inst',tot)
    print('  '+', '.join(f'{op} {100*n/tot:.1f}%' for op,n in ops.most_common(16)))
    print('  stall-by-op: '+', '.join(f'{op} {n}' for op,n in stall.most_common(10)))

import csv, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
for h,v in zip(hdr,vals):
    if h in want: print(h,'=',v)
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Instructions Executed"); ta=hdr.index("Thread Instructions Executed"); sa=hdr.index("# Samples")
tot=sum(int(r[ia]) for r in data); ts=sum(int(r[sa]) for r in data)
print("total inst", tot, "nsass", len(data))
regions=[]; cur=None
for i,r in enumerate(data):
    c=int(r[ia]); t=int(r[ta]); s=int(r[sa])
    if cur and c>0 and 0.7<c/max(cur['c0'],1)<1.4:
        cur['n']+=1; cur['inst']+=c; cur['thr']+=t; cur['smp']+=s; cur['end']=i
    else:
        if cur: regions.append(cur)
        cur=dict(start=i,end=i,n=1,c0=c,inst=c,thr=t,smp=s)
regions.append(cur)
for r in sorted([r for r in regions if r['inst']>tot*0.008],key=lambda r:-r['inst'])[:22]:
    print(f"sass[{r['start']:4d}..{r['end']:4d}] n={r['n']:3d} exec={r['c0']:>10d} inst={r['inst']/tot*100:5.1f}% avgthr={r['thr']/max(r['inst'],1):5.1f} smp={r['smp']/ts*100:5.1f}%  {data[r['start']][1].strip()[:60]}")
open('/tmp/sass.txt','w').write('\n'.join(f"{i} {r[ia]} {r[ta]} {r[sa]} {r[1].strip()}" for i,r in enumerate(data)))

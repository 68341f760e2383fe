import csv,sys,subprocess,io,collections
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr,units=rows[0],rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__occupancy_limit_warps','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','lts__t_sector_hit_rate.pct','launch__shared_mem_per_block_dynamic','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active']
for v in rows[2:]:
    print('==',v[hdr.index('Kernel Name')][:90])
    for i,h in enumerate(hdr):
        if h in want: print('  %-70s %-14s %s'%(h,units[i],v[i]))
    for i,h in enumerate(hdr):
        if 'smsp__average_warp' in h and 'issue_stalled' in h and 'not_issued' not in h:
            try:
                x=float(v[i].replace(',',''))
                if x>0.25: print('  stall %-40s %.2f'%(h.split('issue_stalled_')[1].split('_per_')[0],x))
            except: pass
if len(sys.argv)>2:
    src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(src)))
    h=rows[1]; data=[r for r in rows[2:] if len(r)==len(h)]
    ia=h.index('Instructions Executed'); isamp=h.index('# Samples'); isrc=h.index('Source')
    tot=sum(int(r[ia]) for r in data); ts=sum(int(r[isamp]) for r in data)
    byop=collections.Counter(); bys=collections.Counter()
    for r in data:
        t=r[isrc].split()
        op=t[1] if t[0].startswith('@') else t[0]
        byop[op.split('.')[0]]+=int(r[ia]); bys[op.split('.')[0]]+=int(r[isamp])
    print('total inst',tot,'samples',ts)
    for k,v in byop.most_common(18): print('  %-10s inst %5.1f%%  samples %5.1f%%'%(k,100*v/tot,100*bys[k]/ts))
    print(' top sampled:')
    for i,r in sorted(enumerate(data),key=lambda t:-int(t[1][isamp]))[:int(sys.argv[2])]:
        print('  %5d %-70s inst=%s samp=%s'%(i,r[isrc][:70],r[ia],r[isamp]))

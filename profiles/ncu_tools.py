import csv, subprocess, sys
def launches(path):
    rows=list(csv.reader(open(path)))
    hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    hdr=rows[hi]; ci={h:i for i,h in enumerate(hdr)}
    agg={}
    for r in rows[hi+1:]:
        if len(r)<len(hdr): continue
        k=r[ci['Kernel Name']].split('(')[0]; v=float(r[ci['Metric Value']]); u=r[ci['Metric Unit']]
        if u=='us': v/=1e3
        elif u=='ns': v/=1e6
        agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
    tot=sum(v[1] for v in agg.values())
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:50s} n={n:3d} total={t:9.3f} ms avg={t/n:8.3f} share={t/tot*100:5.1f}%")
def raw(rep):
    out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr=rows[0]; units=rows[1]
    want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','lts__t_bytes.sum']
    for r in rows[2:]:
        name=r[hdr.index('Kernel Name')]
        print("==",name[:60])
        for i,h in enumerate(hdr):
            if h in want: print(f"   {h:70s} {units[i]:12s} {r[i]}")
def src(rep, kernel, top=25):
    out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","-k",kernel],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    his=[i for i,r in enumerate(rows) if r and r[0]=='Line No']
    for hi in his[:1]:
        hdr=rows[hi]; ci={h:i for i,h in enumerate(hdr)}
        tot_inst=0; tot_samp=0; lines=[]
        for r in rows[hi+1:]:
            if not r or r[0]=='' : continue
            try: ln=int(r[0])
            except: continue
            try:
                inst=int(r[ci["Instructions Executed"]].replace("-","0") or 0); samp=int(r[ci["# Samples"]].replace("-","0") or 0)
                st={k:int(r[ci[k]].replace("-","0") or 0) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
            except (ValueError, IndexError):
                continue
            lines.append((ln,r[1].strip()[:80],inst,samp,st)); tot_inst+=inst; tot_samp+=samp
        print("total inst",tot_inst,"samples",tot_samp)
        for ln,s,inst,samp,st in sorted(lines,key=lambda x:-x[3])[:top]:
            t3=sorted(st.items(),key=lambda kv:-kv[1])[:2]
            print(f"{ln:4d} inst={inst/tot_inst*100:5.1f}% samp={samp/tot_samp*100:5.1f}% {s}  {[(k[6:],v) for k,v in t3 if v]}")
if __name__=="__main__":
    cmd=sys.argv[1]
    if cmd=="launches": launches(sys.argv[2])
    elif cmd=="raw": raw(sys.argv[2])
    elif cmd=="src": src(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv)>4 else 25)

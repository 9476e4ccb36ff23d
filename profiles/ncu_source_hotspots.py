import csv,sys,subprocess
rep=sys.argv[1]
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
agg={}; cur_file=None; hdr=None
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)>3 and r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0] not in ('',):
        try: line=int(r[0])
        except: continue
        ie=int(r[hdr.index('Instructions Executed')] or 0); samp=int(r[hdr.index('# Samples')] or 0); thr=int(r[hdr.index('Thread Instructions Executed')] or 0)
        a=agg.setdefault((cur_file,line),[0,0,0,r[1][:95]]); a[0]+=ie; a[1]+=samp; a[2]+=thr
tot=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
print('total inst',tot,'samples',ts)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f'{k[0]:14s}:{k[1]:4d} inst {a[0]/1e6:7.1f}M ({100*a[0]/tot:4.1f}%) samp {100*a[1]/ts:4.1f}% thr/inst {a[2]/max(1,a[0]):4.1f} | {a[3]}')

import csv,collections,sys,re
path=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 50
rows=list(csv.reader(open(path)))
cur=None; agg=collections.Counter(); thr=collections.Counter(); src={}; smp=collections.Counter()
hdr=None
ops=collections.Counter()
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; iI=hdr.index('Instructions Executed'); iT=hdr.index('Thread Instructions Executed'); iS=hdr.index('# Samples'); continue
    if hdr is None or len(r)<=iT: continue
    try: n=int(r[iI]); t=int(r[iT]); s=int(r[iS])
    except: continue
    if r[0]=='':
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)',r[3]); 
        if m: ops[m.group(2)]+=n
        continue
    key=(cur,int(r[0])); agg[key]+=n; thr[key]+=t; src[key]=r[1]; smp[key]+=s
tot=sum(agg.values()); ts=sum(smp.values()); print('tot warp inst',tot,'samples',ts)
byfile=collections.Counter()
for (f,l),n in agg.items(): byfile[f]+=n
for f,n in byfile.most_common(): print(f'{f:30s} {n/tot*100:6.2f}%')
for k,n in agg.most_common(top): print(f'{k[0]:24s}:{k[1]:>5d} {n/tot*100:5.2f}% smp {smp[k]/ts*100:5.2f}% lanes {thr[k]/max(n,1):4.1f} | {src[k].strip()[:100]}')
to=sum(ops.values())
print([(o,round(n/to*100,1)) for o,n in ops.most_common(25)])

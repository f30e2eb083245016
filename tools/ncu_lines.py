import csv,sys,subprocess,collections
rep=sys.argv[1]; nruns=float(sys.argv[2]) if len(sys.argv)>2 else 100000; top=int(sys.argv[3]) if len(sys.argv)>3 else 45
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
agg=collections.Counter(); samp=collections.Counter(); text={}
cur=None
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]
    if r and r[0].isdigit():
        try: n=int(r[7])
        except: continue
        k=(cur,int(r[0])); agg[k]+=n; samp[k]+=int(r[6]) if r[6].isdigit() else 0; text[k]=r[1].strip()[:100]
S=sum(samp.values()); T=sum(agg.values())
print('total/run', T/nruns)
for k,n in sorted(agg.items(), key=lambda kv:-kv[1])[:top]:
    print('%-20s %4d %6.1f/run s%%=%4.1f  %s'%(k[0][:20],k[1],n/nruns,100*samp[k]/S,text[k]))

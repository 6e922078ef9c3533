# usage: python ncu_opcode_histogram.py report.ncu-rep kernel_name launch_index lo hi  -> executed instructions per warp by opcode inside [lo, hi] (hex offsets)
import csv, subprocess, sys, io, re, collections
rep, kname, idx, lo, hi_ = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4],16), int(sys.argv[5],16)
cmd=['ncu','-i',rep,'--page','source','--csv','--kernel-name',kname,'--launch-skip',idx,'--launch-count','1']
raw=subprocess.run(cmd,capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; isrc=hdr.index('Source'); ie=hdr.index('Instructions Executed')
data=[]
for r in rows[hi+1:]:
    try: data.append((int(r[0],16),r[isrc].strip(),int(r[ie])))
    except: break
base=data[0][0]; nw=data[0][2]
c=collections.Counter()
for a,s,e in data:
    if lo<=a-base<=hi_:
        op=re.sub(r'^@!?U?P\d+\s+','',s).split()[0].split('.')[0]
        c[op]+=e
tot=sum(c.values())
print('total',tot,'per warp',tot/nw)
for op,v in c.most_common(40): print(f'{op:12s} {v/nw:8.1f}')

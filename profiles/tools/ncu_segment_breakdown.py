# usage: python ncu_segment_breakdown.py report.ncu-rep kernel_name [launch_index [lo hi]]  -> executed warp instructions and stall samples between barriers
import csv, subprocess, sys, io
rep, kname = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv)>3 else None
cmd=['ncu','-i',rep,'--page','source','--csv','--kernel-name',kname]
if idx: cmd+=['--launch-skip',idx,'--launch-count','1']
raw=subprocess.run(cmd,capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; ia=0; isrc=hdr.index('Source'); ie=hdr.index('Instructions Executed'); iss=hdr.index('# Samples')
data=[]
for r in rows[hi+1:]:
    if len(r)<=ie or r[0]=='Address' : break
    try: data.append((int(r[ia],16),r[isrc].strip(),int(r[ie]),int(r[iss])))
    except: break
base=data[0][0]; tot=sum(d[2] for d in data); ts=sum(d[3] for d in data)
print('total warp instr',tot,'samples',ts,'warps(first instr)',data[0][2])
acc=0;sacc=0;start=0
for a,s,e,n in data:
    acc+=e; sacc+=n
    if 'BAR.SYNC' in s or s.startswith('EXIT') or 'EXIT' in s.split()[-2:] :
        print(f'{start:5x}-{a-base:5x} {s[:34]:34s} instr {acc:>11d} {acc/tot*100:5.1f}%  per-warp {acc/data[0][2]:7.1f}  samples {sacc/ts*100:5.1f}%')
        acc=0;sacc=0;start=a-base+16
print('tail',acc)
if len(sys.argv)>4:
    lo,hi_=int(sys.argv[4],16),int(sys.argv[5],16)
    for a,s,e,n in data:
        if lo<=a-base<=hi_: print(f'{a-base:5x} {e:>10d} {n:>5d} {s}')

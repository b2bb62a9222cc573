import csv, subprocess, sys
rep=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 18
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--launch-skip","0","--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=next(r for r in rows if r and r[0]=="Line No")
si=hdr.index("# Samples"); ii=hdr.index("Instructions Executed")
data=[]
for r in rows:
    if r and r[0].isdigit() and len(r)>si:
        try: data.append((int(r[si] or 0), int(r[ii] or 0), int(r[0]), r[1][:110]))
        except ValueError: pass
tot=sum(d[0] for d in data); ti=sum(d[1] for d in data)
print("total samples",tot,"instr",ti)
for d in sorted(data,reverse=True)[:n]:
    print(f"{100*d[0]/tot:5.1f}% {d[0]:6d} inst {d[1]:9d} L{d[2]:<4} {d[3]}")

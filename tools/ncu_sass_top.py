#!/usr/bin/env python3
"""Top stalled SASS instructions of an .ncu-rep (needs --set full --import-source on): share of all stall
samples, average active lanes and the two dominant stall reasons per instruction.
usage: ncu_sass_top.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep=sys.argv[1]; top=int(sys.argv[2])
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hd=rows[1]; ix={h:i for i,h in enumerate(hd)}
ins=rows[2:]
tot=sum(float(r[ix['# Samples']] or 0) for r in ins)
reasons=[h for h in hd if h.startswith('stall_') and 'Not Issued' not in h]
order=sorted(range(len(ins)), key=lambda i:-float(ins[i][ix['# Samples']] or 0))[:top]
for i in sorted(order):
    r=ins[i]; s=float(r[ix['# Samples']])
    rs=sorted(((float(r[ix[h]] or 0),h) for h in reasons), reverse=True)[:2]
    prev=ins[i-1][1].strip()[:40] if i>0 else ''
    print(f"{i:5d} {100*s/tot:5.2f}% lanes {r[ix['Avg. Threads Executed']]:>5} {r[1].strip()[:58]:58s} | {rs[0][1][6:]}:{rs[0][0]/s:.2f} {rs[1][1][6:]}:{rs[1][0]/s:.2f}")

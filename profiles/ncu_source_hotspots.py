import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
iS=hdr.index('Source'); iSt=hdr.index('Warp Stall Sampling (All Samples)'); iEx=hdr.index('Instructions Executed')
iW=hdr.index('L1 Wavefronts Shared'); iWi=hdr.index('L1 Wavefronts Shared Ideal')
data=rows[2:]
tot=sum(int(r[iSt]) for r in data)
print('total samples',tot,'instr',len(data))
# cumulative regions: print instrs with >1.2% samples, with index
for k,r in enumerate(data):
    st=int(r[iSt])
    if st>tot*float(sys.argv[2]) :
        print(f"{k:4d} {st/tot:6.1%} ex={r[iEx]:>8s} shW={r[iW]:>8s}/{r[iWi]:>8s}  {r[iS].strip()[:90]}")
# region summary by barrier positions
print('--- regions split at BAR/SYNCS')
acc=0; start=0; shw=0
for k,r in enumerate(data):
    acc+=int(r[iSt]); shw+=int(r[iW] or 0)
    src=r[iS]
    if 'BAR.SYNC' in src or 'PHASECHK' in src or k==len(data)-1:
        print(f"instr {start:4d}-{k:4d}: {acc/tot:6.1%} samples, shared wavefronts {shw}   ends with {src.strip()[:50]}")
        acc=0; start=k+1; shw=0

import csv, subprocess, sys, io, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h=None; sect=0
agg=collections.Counter(); inst=collections.Counter(); st=collections.defaultdict(collections.Counter)
ranges=[(283,305,'emit'),(306,330,'begin'),(331,385,'tile_layer'),(386,450,'finalize'),(451,500,'texel helpers'),(501,530,'cand/texidx'),(531,590,'ordered/continue'),(591,636,'base'),(637,668,'post blit'),(669,715,'store/raster loop'),(716,740,'init')]
for r in rows:
    if r and r[0]=="Line No": h=r; sect+=1; continue
    if h and len(r)>8 and r[0].isdigit():
        s=h.index("# Samples"); i=h.index("Instructions Executed")
        try: n=int(r[s] or 0); ni=int(r[i] or 0)
        except ValueError: continue
        line=int(r[0]); key="sect%d"%sect
        if sect==6:
            key='render:?'
            for a,b,nm in ranges:
                if a<=line<=b: key='render:'+nm
        agg[key]+=n; inst[key]+=ni
        for j,nm in enumerate(h):
            if nm.startswith("stall_") and "Not Issued" not in nm:
                try: x=int(r[j] or 0)
                except ValueError: x=0
                if x: st[key][nm[6:]]+=x
tot=sum(agg.values()); ti=sum(inst.values())
for k,v in agg.most_common():
    print("%-26s smp %5.1f%% inst %5.1f%%  %s"%(k,100*v/tot,100*inst[k]/ti, st[k].most_common(4)))

import csv, subprocess, sys, io
rep=sys.argv[1]
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg={c:0 for c in stall_cols}
for r in data:
    for c in stall_cols:
        try: agg[c]+=int(r[ix[c]] or 0)
        except: pass
tot=sum(agg.values())
print(rep, "instructions", len(data), "samples", tot)
print({k:round(100*v/tot,1) for k,v in sorted(agg.items(), key=lambda x:-x[1])[:8]})
top=sorted(data, key=lambda r:-int(r[ix["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 12]
for r in top:
    st={c:int(r[ix[c]] or 0) for c in stall_cols if (r[ix[c]] or "0")!="0"}
    st=dict(sorted(st.items(), key=lambda x:-x[1])[:2])
    print("  ", r[ix["Address"]][-5:], r[ix["# Samples"]], r[ix["Source"]][:64], st)

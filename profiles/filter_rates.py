"""Candidate rates of wider n-gram pre-filters on the bench URL column (CPU experiment, 24 entries of synth.url_entry):
share of dictionary values that pass the reference fingerprint gate AND the filter, per needle. Run: PYTHONPATH=. python profiles/filter_rates.py"""
import numpy as np, pyarrow as pa, pyarrow.compute as pc, synth
synth.lib().lcs_init(synth.URL_POOL)
def fp32(b):
    x=0
    for c in b: x|=1<<(c&31)
    return x
def bigram64(b):
    x=0
    for a,c in zip(b,b[1:]): x|=1<<((((a<<8)|c)*0x9E3779B1 & 0xFFFFFFFF)>>26)
    return x
def tri(b,bits):
    x=0; sh=32-int(np.log2(bits))
    for a,c,d in zip(b,b[1:],b[2:]): x|=1<<((((a<<16)|(c<<8)|d)*0x9E3779B1 & 0xFFFFFFFF)>>sh)
    return x
def bi(b,bits):
    x=0; sh=32-int(np.log2(bits))
    for a,c in zip(b,b[1:]): x|=1<<((((a<<8)|c)*0x9E3779B1 & 0xFFFFFFFF)>>sh)
    return x
needles=[b"google", b".google.", b"tours", b"Google"]
tot=0; stats={n:dict(fp=0,b64=0,t128=0,t256=0,b256=0,true=0) for n in needles}
for e in range(24):
    arr=synth.url_entry(e); u=pc.unique(arr).to_pylist(); tot+=len(u)
    ub=[s.encode() for s in u]
    F=[(fp32(b),bigram64(b),tri(b,128),tri(b,256),bi(b,256)) for b in ub]
    for n in needles:
        nf=(fp32(n),bigram64(n),tri(n,128),tri(n,256),bi(n,256))
        for b,(f,b64,t128,t256,b256) in zip(ub,F):
            g=(f&nf[0])==nf[0]
            stats[n]['fp']+=g
            stats[n]['b64']+= g and (b64&nf[1])==nf[1]
            stats[n]['t128']+= g and (t128&nf[2])==nf[2]
            stats[n]['t256']+= g and (t256&nf[3])==nf[3]
            stats[n]['b256']+= g and (b256&nf[4])==nf[4]
            stats[n]['true']+= n in b
print("uniques", tot)
for n in needles: print(n, {k: round(v/tot*100,3) for k,v in stats[n].items()})

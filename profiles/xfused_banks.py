"""Shared-memory wavefront model used to pick the swizzle of csrc/xfused_kernel.cuh (run: python profiles/xfused_banks.py)."""
# Shared-memory wavefront model for the x-fused kernel.  16-B accesses: quarter-warp (8 lanes) per wavefront
# group, conflict degree = max multiplicity of (unit mod 8) with distinct addresses.  8-B accesses: half-warp.
import itertools
def wf128(units):   # units: list of 16B-unit indices for 32 lanes (None = inactive)
    tot=0
    for q in range(4):
        lanes=[u for u in units[8*q:8*q+8] if u is not None]
        if not lanes: continue
        cnt={}
        for u in set(lanes): cnt[u%8]=cnt.get(u%8,0)+1
        tot+=max(cnt.values())
    return tot
def wf64(units8):   # 8-B unit indices
    tot=0
    for h in range(2):
        lanes=[u for u in units8[16*h:16*h+16] if u is not None]
        if not lanes: continue
        cnt={}
        for u in set(lanes): cnt[u%16]=cnt.get(u%16,0)+1
        tot+=max(cnt.values())
    return tot

def analyse(N, radices, NP, NT, sw, verbose=True):
    S=len(radices)
    P=[1]
    for r in radices: P.append(P[-1]*r)
    total=0; ideal=0
    res={}
    for s,R in enumerate(radices):
        Ps=P[s]; M=N//Ps; Q=M//R
        items=NP*(N//R)
        def decode(i):
            if s==0:
                pen=i//Q; b=i%Q; q=0
            else:
                b=i//(NP*Ps); rem=i%(NP*Ps); pen=rem//Ps; q=rem%Ps
            return pen,q,b
        w=0; wi=0
        for i0 in range(0,items,32):
            for j in range(R):
                units=[]
                for l in range(32):
                    i=i0+l
                    if i>=items: units.append(None); continue
                    pen,q,b=decode(i)
                    e=q*M+b+j*Q
                    units.append(pen*N+sw(e))
                w+=wf128(units); wi+=sum(1 for u in units if u is not None)/8.0
        res['stage%d(R=%d)'%(s,R)]=(w,wi)
    # middle: items (g=0, w', line) with line fastest? or w' fastest
    for name,order in (('mid_linefast',0),('mid_wfast',1)):
        items=N//2*2
        w=0; wi=0
        for f in range(1):
            for i0 in range(0,items,32):
                for half in range(2):
                    units=[]
                    for l in range(32):
                        i=i0+l
                        if order==0: wp=i//2; c=i%2
                        else: c=i//(N//2); wp=i%(N//2)
                        e=2*wp+half
                        units.append(sw(e)*2+c)
                    w+=wf64(units); wi+=2
        res[name]=(w,wi)
    return res

if __name__ == "__main__":
  N=512
  cands={
   'none': lambda e:e,
   'q5': lambda e: e^((e>>5)&7),
   'q5b3': lambda e: e^((e>>5)&7)^((e>>3)&1),
   'b3b6 (adopted)': lambda e: e^((e>>3)&7)^((e>>6)&7),
  }
  for nm,sw in cands.items():
    print(nm, analyse(512,[16,16],6,288,sw))


def report(N, radices):
    sw = lambda e: e ^ ((e >> 3) & 7) ^ ((e >> 6) & 7)
    res = analyse_tp(N, radices, sw)
    print(N, radices, {k: "%d/%d" % v for k, v in res.items()})


def analyse_tp(N, radices, sw):
    """Wavefronts (actual/ideal) per pencil and stage with the kernel's thread mapping: TP = N/R0 threads per pencil,
    butterfly w = lane + it*TP, b = w // P, q = w % P."""
    TP = N // radices[0]
    P = 1
    out = {}
    for s, R in enumerate(radices):
        M = N // P
        Q = M // R
        items = N // R
        w_act = w_id = 0
        for w0 in range(0, items, TP):
            lanes = list(range(w0, min(w0 + TP, items)))
            for warp0 in range(0, len(lanes), 32):
                grp = lanes[warp0:warp0 + 32]
                for j in range(R):
                    units = []
                    for w in grp:
                        b, q = w // P, w % P
                        units.append(sw(q * M + b + j * Q))
                    units += [None] * (32 - len(units))
                    w_act += wf128(units)
                    w_id += len(grp) / 8.0
        out["stage%d(R=%d)" % (s, R)] = (w_act, int(w_id))
        P *= R
    return out


if __name__ == "__main__":
    for N, rad in ((512, [16, 16]), (256, [16, 8]), (128, [8, 8]), (1024, [16, 8, 4]), (1024, [16, 16, 2]), (1024, [8, 8, 8]),
                   (1024, [16, 4, 8]), (1024, [4, 16, 8]), (1024, [32, 16])):
        try:
            report(N, rad)
        except Exception as e:
            print(N, rad, "n/a", e)

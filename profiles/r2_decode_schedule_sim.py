# free-running groups over several steps; resources G and A (serial each); dynamic greedy or static alternating order
import statistics
def items_for(t, layers=6, first=False):
    S = 0.3 + 8.1 * (t + 1) / 501.0; C = 5.0
    layer = [("qkv","G",2.2),("self","A",S),("oproj","G",1.7),("cq","G",1.8),("cross","A",C),("coproj","G",1.6),("ffn1","G",2.9),("ffn2","G",2.6),("red","G",1.2)]
    return [("pre1","G",1.7),("pre2","G",1.8)] + layer*layers + [("final","G",1.8)]
def run(t, steps=6, L=1.9, ngroups=2, merged=False, gs=1.0, As=1.0, static_off=None, barrier=False):
    its = items_for(t); n = len(its)
    seq = its * steps
    N = len(seq)
    nxt = [0]*ngroups; ready=[0.0]*ngroups; free={"G":0.0,"A":0.0}
    fin_step = [[0.0]*ngroups for _ in range(steps)]
    done=0
    # static order for G: precomputed merged list
    while done < N*ngroups:
        best=None
        for g in range(ngroups):
            if nxt[g]>=N: continue
            nm,k,w = seq[nxt[g]]
            if barrier and nxt[g] % n == 0 and nxt[g] > 0:
                s = nxt[g]//n - 1
                if any(nxt[h] < (s+1)*n for h in range(ngroups)): continue
                r = max(fin_step[s]) + L
            else: r = ready[g]
            res = "G" if merged else k
            w = w*(gs if k=="G" else As)
            start=max(r, free[res])
            key=(start, nxt[g], g)
            if best is None or key<best[0]: best=(key,g,k,w,res,start)
        _,g,k,w,res,start=best
        end=start+w; free[res]=end; ready[g]=end+L
        if (nxt[g]+1) % n == 0: fin_step[nxt[g]//n][g]=end
        nxt[g]+=1; done+=1
    # steady-state step time: between completion of step 1 and last step
    return (max(fin_step[-1]) - max(fin_step[1]))/(steps-2)
for t in (0,250,500,750,999):
    print(t, "now-like(merged+barrier) %.1f"%run(t,merged=True,barrier=True), "A|G barrier %.1f"%run(t,barrier=True), "A|G free %.1f"%run(t),
          "A|G free L=1.4 gs=.85 %.1f"%run(t,L=1.4,gs=.85), "A|G free As=1.15 gs=1.1 %.1f"%run(t,As=1.15,gs=1.1))
for kw in (dict(merged=True,barrier=True), dict(barrier=True), dict(), dict(L=1.4,gs=.85), dict(As=1.15,gs=1.1), dict(ngroups=3), dict(ngroups=4)):
    print(kw, "avg %.1f"%statistics.mean(run(t,**kw) for t in range(0,1000,20)))

"""Host-side replay of the pair kernel's weight-ring placement (mlp_tc3.cuh, producer loop): where every weight chunk
of a tile lands in the byte FIFO, which earlier chunks it must evict first, and whether data or operand windows overrun
the ring.  Needs the schedule dumper:

    nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ibeetroots_b200/csrc -Iinclude -o /tmp/sd/sched_dump scripts/sched_dump.cu
    cd /tmp/sd && python /root/repo/scripts/ring_sim.py

Used to look for a structural dead-lock of the shortened (pre-pass) schedule on the 32-line network: none found -- the
eviction sets only ever point backwards and every window fits."""
import subprocess, sys, re
NT=16; MISC=4096
def sched(n_out, skip):
    out=subprocess.run(['./sched_dump',str(n_out),str(skip)],capture_output=True,text=True).stdout.splitlines()
    hdr=out[0]; ring_bytes=int(re.search(r'ring_bytes=(\d+)',hdr).group(1))
    U=[]
    for l in out[2:]:
        m=re.match(r'U(\d+)\s+k16 \[\s*(\d+),\s*(\d+)\) new\s+(\d+) dep (\d+) sp (\d+) flags\s+(\d+) ready_idx\s+(-?\d+) two (\d) ring_lo\s+(\d+) cw (\d) rep (\d) groups\s+(\d+)/\s*(\d+)',l)
        g=list(map(int,m.groups()))
        U.append(dict(kb=g[1],ke=g[2],new=g[3],dep=g[4],sp=g[5],flags=g[6],ridx=g[7],ring_lo=g[9],cw=g[10],g0=g[12],g1=g[13]))
    return ring_bytes,U
def simulate(n_out, skip, tiles=3):
    ring,U=sched(n_out,skip)
    # enumerate chunks
    chunks=[]  # dict: tile,u,ci,pos,vpos,adv(bytes),needs_dep
    pos=0; vpos=0; i=0
    for t in range(tiles):
        for u,X in enumerate(U):
            advkb=max(X['g0'],X['g1'])*1024
            kb0=X['kb']>>2; kb1=(X['ke']+3)>>2; cw=X['cw']
            win=(cw-1)*advkb+16384-MISC
            ci=0
            kb=kb0
            while kb<kb1:
                nk=min(cw,kb1-kb); adv=advkb*nk
                if pos+win>ring or pos+adv>ring: vpos+=ring-pos; pos=0
                if pos<X['ring_lo']: vpos+=X['ring_lo']-pos; pos=X['ring_lo']
                # does this chunk contain new features?
                last_kb=kb+nk-1
                ks_hi = (X['ke']-4*last_kb) if last_kb==kb1-1 else 4
                needs = 4*last_kb+ks_hi > X['new']
                chunks.append(dict(i=i,tile=t,u=u,ci=ci,pos=pos,vpos=vpos,adv=adv,needs=needs,win=win,endwin=pos+(nk-1)*advkb+16384))
                pos+=adv; vpos+=adv; kb+=cw; i+=1; ci+=1
    n=len(chunks)
    # wait sets for producing chunk i (own-chunk rule, as executed by its producer which also applies ticket rule for others)
    # simulate each producer's tail pointer
    need=[set() for _ in range(n)]
    for pid in (0,1):
        tail=0
        for c in chunks:
            i=c['i']
            if i%2!=pid:
                if tail+NT<=i: need[i].add(('ticket',tail)); tail+=1   # producer pid blocked here (affects its later chunks)
                continue
            lo=c['vpos']+c['adv']-ring
            while tail<i and (tail+NT<=i or lo-chunks[tail]['vpos']>0):
                need[i].add(tail); tail+=1
    # overflow check: chunk larger than the window available
    for c in chunks:
        X=U[c['u']]
        if c['pos']+c['adv']>ring: print('OVERFLOW data',c)
        if c['endwin']>ring+MISC: print('OVERFLOW window',c, ring+MISC)
    return ring,U,chunks,need
for a in ((10,0),(10,4),(32,0),(32,4)):
    ring,U,chunks,need=simulate(*a)
    per_tile=len(chunks)//3
    print('== n_out,skip',a,'ring',ring,'chunks/tile',per_tile)
    for c in chunks[per_tile:2*per_tile]:
        nd=sorted(x for x in need[c['i']] if isinstance(x,int))
        print(f"  i={c['i']:3d} U{c['u']} ci={c['ci']} pos={c['pos']:6d} adv={c['adv']:5d} endwin={c['endwin']:6d} evicts={[x for x in nd][-4:] if nd else []} (n={len(nd)})")

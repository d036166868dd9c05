#!/usr/bin/env python3
"""Shared-memory bank model of k_dense's table lookups (no GPU needed).

Replays the C2 corpus through the Aho-Corasick automaton of (hello|world|foo) with the kernel's thread /
chain geometry (256 threads x 144 bytes, chain A = bytes [0, 80), chain B = [80, 144)) and counts, for every
warp-wide lookup, max over banks of the number of distinct 32-bit words addressed = the wavefronts the LSU
needs.  For the committed layout (uint16 entries, rows 520 B apart, accept flag +4) the model gives 2.042
wavefronts per lookup; ncu measured 2.04 (profiles/r01_k_dense_h_ncu.txt).  The other layouts are the ones
weighed for round 2 (known weak spots): uint8 entries roughly halve a row's bank footprint but only
reach ~1.75; duplicated tables are worse.  usage: python profiles/sim_dense_banks.py  (about 5 minutes)
"""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_regex_b200 import corpora
N = 4<<20
text = corpora.corpus_c2(N).numpy()
lits=[b"hello",b"world",b"foo"]
# AC automaton, BFS numbering like build_dense: nodes created in insertion order
nxt=[[-1]*256]; out=[0]
for w in lits:
    cur=0
    for ch in w:
        if nxt[cur][ch]<0:
            nxt[cur][ch]=len(nxt); nxt.append([-1]*256); out.append(0)
        cur=nxt[cur][ch]
    out[cur]=len(w)
fail=[0]*len(nxt)
from collections import deque
q=deque()
for b in range(256):
    t=nxt[0][b]
    if t<0: nxt[0][b]=0
    else: q.append(t)
while q:
    u=q.popleft()
    for b in range(256):
        t=nxt[u][b]
        if t<0: nxt[u][b]=nxt[fail[u]][b]
        else: fail[t]=nxt[fail[u]][b]; q.append(t)
nst=len(nxt)
T=np.array(nxt,dtype=np.int32)
# placement: plain rows first then accept rows
plain=[i for i in range(nst) if not out[i]]; acc=[i for i in range(nst) if out[i]]
order=plain+acc
rowidx={s:k for k,s in enumerate(order)}
print("states",nst,"plain",len(plain))
# state sequence over the text (global walk; warm-up makes per-thread states exact)
st=np.zeros(N,dtype=np.int32)  # state BEFORE consuming byte i
s=0
tl=text.tolist()
for i,b in enumerate(tl):
    st[i]=s; s=nxt[s][b]
hist=collections.Counter(st.tolist()); print("state hist (top)",[(k,round(v/N,4)) for k,v in hist.most_common(6)])
kTile=36864; kChunk=144; kA=80
def wavefronts(addr_fn, name, base_off=0xc1d8):
    # addr_fn(rowidx array, byte array) -> byte address in shared
    total=0; cnt=0
    ntiles=N//kTile
    ridx=np.vectorize(rowidx.get)(st).astype(np.int64)
    for t in range(min(ntiles,40)):
        base=t*kTile
        for warp in range(8):
            th0=warp*32
            # positions: lane l, step i (chain A: i in 0..79; chain B: 80+i, i in 0..63)
            for chain,(o,L) in enumerate(((0,80),(80,64))):
                pos = base + (th0+np.arange(32))[:,None]*kChunk + o + np.arange(L)[None,:]   # [32,L]
                a = addr_fn(ridx[pos], text[pos].astype(np.int64)) + base_off
                word = a>>2; bank = word & 31
                for i in range(L):
                    w=word[:,i]; bk=bank[:,i]
                    # distinct words per bank
                    uniq=np.unique(w)
                    c=np.bincount(uniq&31,minlength=32).max()
                    total+=c; cnt+=1
    print(name, "avg wavefronts/lookup", round(total/cnt,3))
wavefronts(lambda r,b: r*520 + (np.isin(r, [rowidx[a] for a in acc])*4) + 2*b, "u16 stride 520 (current)")
wavefronts(lambda r,b: r*512 + 2*b, "u16 stride 512")
for stride in (256,260,264,272,288,320):
    wavefronts(lambda r,b: r*stride + b, "u8 stride %d"%stride)
wavefronts(lambda r,b: r*1024+4*b, "u32 stride 1024")
wavefronts(lambda r,b: r*1028+4*b, "u32 stride 1028")
print("---- BFS-depth ordering")
depth={0:0}
dq=deque([0]); 
# recompute trie depth via goto edges from original trie: depth = length of path; use fail-free BFS on trie edges
trie=[[-1]*256]; 
for w in lits:
    cur=0
    for ch in w:
        if trie[cur][ch]<0:
            trie[cur][ch]=len(trie); trie.append([-1]*256)
        cur=trie[cur][ch]
while dq:
    u=dq.popleft()
    for b in range(256):
        t=trie[u][b]
        if t>=0: depth[t]=depth[u]+1; dq.append(t)
order2=sorted(range(nst), key=lambda s:(depth[s], s))
print(order2)
rowidx.clear(); rowidx.update({s:k for k,s in enumerate(order2)})
for stride in (260,264,272,288,296,304,320,352):
    wavefronts(lambda r,b: r*stride + b, "u8 bfs stride %d"%stride)
wavefronts(lambda r,b: r*520 + 2*b, "u16 bfs stride 520")
wavefronts(lambda r,b: r*544 + 2*b, "u16 bfs stride 544")
wavefronts(lambda r,b: r*576 + 2*b, "u16 bfs stride 576")
print("---- two copies (lane parity)")
rowidx.clear(); rowidx.update({s:k for k,s in enumerate(order)})
accrows=[rowidx[a] for a in acc]
lanes=np.arange(32)[:,None]
TB=520*14+16
for shift in (0,32,64,96):
    wavefronts(lambda r,b: r*520 + (np.isin(r, accrows)*4) + 2*b + (lanes&1)*(((TB+127)//128)*128+shift), "u16 520 two copies shift %d"%shift)
for shift in (0,32,64):
    wavefronts(lambda r,b: r*296 + b + (lanes&1)*(((296*14+127)//128)*128+shift), "u8 296 two copies shift %d"%shift)
    wavefronts(lambda r,b: r*296 + b + (lanes&3)*(((296*14+127)//128)*128+shift), "u8 296 four copies shift %d"%shift)

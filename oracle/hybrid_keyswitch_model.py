# Toy model of level-aware hybrid key switching with idle primes as extra special moduli.
import random
from math import prod
N=16
def is_prime(p):
    if p<2: return False
    i=2
    while i*i<=p:
        if p%i==0: return False
        i+=1
    return True
def gen_primes(bits,count,avoid=()):
    ps=[];p=(1<<bits)+1
    while len(ps)<count:
        if p%(2*N)==1 and is_prime(p) and p not in avoid: ps.append(p)
        p+=2*N
    return ps
L=6                       # data primes q_0..q_5, special P
q=gen_primes(20,L)
P=gen_primes(21,1,q)[0]
def polymul(a,b,m):
    r=[0]*N
    for i,x in enumerate(a):
        if x==0: continue
        for j,y in enumerate(b):
            k=i+j
            if k<N: r[k]=(r[k]+x*y)%m
            else: r[k-N]=(r[k-N]-x*y)%m
    return r
random.seed(1)
s=[random.choice([-1,0,1]) for _ in range(N)]
s2=[random.choice([-1,0,1]) for _ in range(N)]   # "old" key s' (e.g. s^2 or sigma(s))
def small(): return [random.choice([-2,-1,0,1,2]) for _ in range(N)]

def keygen(l,alpha):
    """key for switching from s2 to s at level l, digits of alpha primes, specials = {P} + q[l..l+alpha-2]"""
    assert l+alpha-1<=L
    S=[q[l+i] for i in range(alpha-1)]+[P]
    E=q[:l]+S
    PS=prod(S)
    dnum=-(-l//alpha)
    keys=[]
    for d in range(dnum):
        D=range(d*alpha,min((d+1)*alpha,l))
        e=small()
        kb=[];ka=[]
        for idx,m in enumerate(E):
            a=[random.randrange(m) for _ in range(N)]
            b=[(-x+y)%m for x,y in zip(polymul(a,[v%m for v in s],m),e)]
            if idx<l and idx in D:
                f=PS%m
                b=[(x+f*(v%m))%m for x,v in zip(b,s2)]
            kb.append(b);ka.append(a)
        keys.append((kb,ka))
    return keys,E,S,PS,dnum

def bconv(res,src,dst):
    """fast basis conversion: res[i] poly mod src[i] -> polys mod each dst"""
    Qs=prod(src)
    y=[[ (c*pow(Qs//m,-1,m))%m for c in r] for r,m in zip(res,src)]
    out=[]
    for t in dst:
        out.append([sum(y[i][j]*((Qs//src[i])%t) for i in range(len(src)))%t for j in range(N)])
    return out

def keyswitch(c,l,alpha):
    """c: residues of a polynomial mod q[:l] (coefficient form). returns (r0,r1) mod q[:l] with r0+r1*s ~ c*s2"""
    keys,E,S,PS,dnum=keygen(l,alpha)
    acc0=[[0]*N for _ in E]; acc1=[[0]*N for _ in E]
    for d in range(dnum):
        Didx=list(range(d*alpha,min((d+1)*alpha,l)))
        src=[q[i] for i in Didx]
        ext=[]
        conv=bconv([c[i] for i in Didx],src,E)
        for idx,m in enumerate(E):
            if idx in Didx: ext.append(c[idx])
            else: ext.append(conv[idx])
        kb,ka=keys[d]
        for idx,m in enumerate(E):
            p0=polymul(ext[idx],kb[idx],m);p1=polymul(ext[idx],ka[idx],m)
            acc0[idx]=[(x+y)%m for x,y in zip(acc0[idx],p0)]
            acc1[idx]=[(x+y)%m for x,y in zip(acc1[idx],p1)]
    outs=[]
    for acc in (acc0,acc1):
        conv=bconv(acc[l:],S,q[:l])
        r=[]
        for i in range(l):
            m=q[i];inv=pow(PS%m,-1,m)
            r.append([((acc[i][j]-conv[i][j])*inv)%m for j in range(N)])
        outs.append(r)
    return outs

def crt(res,mods):
    M=prod(mods);out=[]
    for j in range(N):
        x=0
        for r,m in zip(res,mods):
            x+=r[j]*(M//m)*pow(M//m,-1,m)
        x%=M
        if x>M//2:x-=M
        out.append(x)
    return out
for l,alpha in [(6,1),(5,2),(4,3),(3,2),(3,3),(2,2),(5,1),(1,1),(4,2)]:
    Q=prod(q[:l])
    cpoly=[random.randrange(Q) for _ in range(N)]
    c=[[x%m for x in cpoly] for m in q[:l]]
    r0,r1=keyswitch(c,l,alpha)
    # check r0 + r1*s - c*s2  (mod Q) is small
    err=[]
    for i in range(l):
        m=q[i]
        lhs=[(x+y)%m for x,y in zip(r0[i],polymul(r1[i],[v%m for v in s],m))]
        rhs=polymul(c[i],[v%m for v in s2],m)
        err.append([(a-b)%m for a,b in zip(lhs,rhs)])
    e=crt(err,q[:l])
    print(l,alpha,"max |err| =",max(abs(x) for x in e),"log2 Q",Q.bit_length())

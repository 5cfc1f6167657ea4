# Model of candidate reduce sequences with 32-bit limbs and carry flags; check vs big ints on crafted limbs.
import itertools, random
W=1<<32; P=(1<<64)-(1<<32)+1; M=W-1
def addcc(a,b,c=0): s=a+b+c; return s&M, s>>32
def subcc(a,b,bw=0): s=a-b-bw; return s&M, 1 if s<0 else 0
def red_wide(r0,r1,r2,r3):
    # T,cy = r2*EPS + (r1:r0)
    T = r2*M + (r1<<32|r0); cy=T>>64; T&=(1<<64)-1
    t0,t1=T&M,T>>32
    u0,b=subcc(t0,r3); u1,b=subcc(t1,0,b); k,_=subcc(cy,0,b)   # k = cy - b  (mod W)
    # fix: + k*EPS
    khi = M if k>>31 else 0
    lo,b=subcc(u0,k); hi,_=subcc(u1,khi,b); hi=(hi+k)&M
    return hi<<32|lo
def red_alu(r0,r1,r2,r3):   # variant R1 from the lab
    t0,b=subcc(r0,r3); t1,b=subcc(r1,0,b); t2,_=subcc(0,0,b)
    t0,b=subcc(t0,r2); t1,b=subcc(t1,0,b); t2,_=subcc(t2,0,b)
    t1,c=addcc(t1,r2); t2,_=addcc(t2,0,c)
    t0,b=subcc(t0,t2); u1,_=subcc(t1,0,b); t1=(u1+t2)&M
    return t1<<32|t0
vals=[0,1,2,3,0x7FFFFFFF,0x80000000,0xFFFFFFFD,0xFFFFFFFE,0xFFFFFFFF]
bad={'wide':0,'alu':0}; n=0
for r in itertools.product(vals,repeat=4):
    x=r[0]+r[1]*W+r[2]*W**2+r[3]*W**3
    if x > (2**64-1)**2+ (2**64-1): continue   # not a reachable a*b+c
    n+=1
    for name,f in (('wide',red_wide),('alu',red_alu)):
        y=f(*r)
        if y>=1<<64 or y%P!=x%P: bad[name]+=1
print(n,bad)

def red_v2(r0,r1,r2,r3,r4=0):
    a0,b=subcc(r0,r3); a1,b=subcc(r1,r4,b); r2p,b2=subcc(r2,0,b)
    T=r2p*M + (a1<<32|a0) + b2; cy=T>>64; T&=(1<<64)-1
    assert cy<=1
    t0,t1=T&M,T>>32
    lo,be=subcc(t0,cy); h2,_=subcc(t1,0,be); hi=(h2+cy)&M
    return hi<<32|lo
bad=0;n=0
vals5=[0,1,2,7,15]
for r in itertools.product(vals,repeat=4):
    for r4 in vals5:
        x=r[0]+r[1]*W+r[2]*W**2+r[3]*W**3+r4*W**4
        if r4==0 and x > (2**64-1)**2+(2**64-1): continue
        n+=1
        y=red_v2(*r,r4)
        if y>=1<<64 or y%P!=x%P: bad+=1
print("v2",n,bad)
random.seed(1)
for _ in range(200000):
    r=[random.choice(vals+[random.getrandbits(32)]) for _ in range(4)]; r4=random.choice(vals5)
    x=r[0]+r[1]*W+r[2]*W**2+r[3]*W**3+r4*W**4
    y=red_v2(*r,r4); assert y<1<<64 and y%P==x%P
print("random ok")

"""CPU model (oracle labels): how many points MUST follow their own trajectory (26-neighbour edge points of the final
labelling) against how many the last level of the FAST algorithm walks (points of non-uniform stride-2 cubes).
160^3, 8 atoms (80 points per atom spacing): edge points 8.3 % of the grid, last-level walkers 11.6 % -> the walker
COUNT is within 1.4x of its floor; the gap to the roofline is cost per step and steps per walk.  usage: python tools/model_walkers.py"""
import sys,time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, systems as S
from oracle import oracle as orc
N=160; side=2   # 80 points per atom spacing ~ like 1024^3 with 8^3 atoms -> 128 per atom; use 80
n=(N,N,N)
x2c=S.cell_x2c(5.0*side,5.0*side,5.0*side)
at,z,al=S.jittered_lattice(side,5)
at=S.snap_to_grid(at,n)
f=orc.promolecular(n,x2c,at,z,al,nimg=1)
t=time.time()
term,_=orc.bader_canonical(f,x2c)
print("canonical in",round(time.time()-t,1),"s; basins",len(np.unique(term)))
lab=term
# edge points: any 26-neighbour with a different label
edge=np.zeros(n,bool)
for dx in (-1,0,1):
  for dy in (-1,0,1):
    for dz in (-1,0,1):
      if dx==dy==dz==0: continue
      edge|= np.roll(lab,(dx,dy,dz),(0,1,2))!=lab
print("edge points fraction", edge.mean())
# stride-2 cubes: corners lab[::2,::2,::2]; cube uniform if 8 corners agree
c=lab[::2,::2,::2]
uni=np.ones(c.shape,bool)
for dx in (0,1):
  for dy in (0,1):
    for dz in (0,1):
      uni&= np.roll(c,(-dx,-dy,-dz),(0,1,2))==c
nonuni=~uni
# points walked at l1: non-lattice points belonging to at least one non-uniform cube... (a point is filled if it lies in a uniform cube? use: walks if ALL cubes containing it are non-uniform -> lower bound; or ANY -> upper bound)
big_any=np.zeros(n,bool); big_all=np.ones(n,bool)
# cube (i,j,k) covers points 2i..2i+2 in each dim
cubes_any=np.zeros(n,bool)
for ox in (0,1,2):
  for oy in (0,1,2):
    for oz in (0,1,2):
      m=np.zeros(n,bool); m[::2,::2,::2]=nonuni
      cubes_any|=np.roll(m,(ox,oy,oz),(0,1,2))
latt=np.zeros(n,bool); latt[::2,::2,::2]=True
print("points in some non-uniform stride-2 cube (excl. lattice points):",(cubes_any&~latt).mean())
print("edge & not lattice:",(edge&~latt).mean(), " edge within nonuniform:",(edge&cubes_any).sum()/edge.sum())

import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import sized_cases as Z
from oracle import oracle as orc
name = sys.argv[1]
c = Z.CASES[name]()
f = orc.promolecular(c['n'], c['x2c'], c['atoms'], c['z'], c['alpha'], nimg=c['nimg'], rc=c['rc'])
term, st = orc.bader_canonical(f, c['x2c'])
_, c2l, lid = orc.bader_metrics(c['x2c'], c['n'])
f.ravel(order='F').tofile(f'/tmp/{name}_f.raw'); term.ravel(order='F').tofile(f'/tmp/{name}_t.raw')
meta = np.concatenate([np.asarray(c2l).ravel(order='F'), np.asarray(lid).ravel()])
meta.tofile(f'/tmp/{name}_m.raw')
print(name, c['n'], st)

import sys, numpy as np
sys.path.insert(0,'.')
from oracle import pyoracle as po
from tests import util
import phasetype_b200 as pb
n=3; mhit=1
rng = np.random.default_rng(100 + n)
R, s = util.dense_rates(n, rng)
l = 3000
y = rng.exponential(1.0, l) * (1.0 + 0.2 * n) / 2 + 0.01
cens = np.zeros(l, dtype=np.int32)
T, C, theta = util.general_model(R, s)
m=theta.shape[0]
eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method=1, mhit=mhit, seed=0xABCDEF12345, mhrs_cap=256)
eng.set_theta(theta, next_iter=7)
B,N,z = eng.paths()
mdl = eng.model()
np.savez('gpurun_out/dbg1.npz', S=mdl['S'], s=mdl['s'], y=y, B=B, N=N, z=z, Pfull=mdl['Pfull'])
Bo,No,zo,co=po.mhrs_paths("oracle",0xABCDEF12345,7,y,cens,mdl['S'],mdl['s'],mhit=mhit)
Br,Nr,zr,cr=po.mhrs_paths("ref",0xABCDEF12345,7,y,cens,mdl['S'],mdl['s'],mhit=mhit)
bad=np.where((No!=Nr).any(1))[0]
print('mismatch count', len(bad), bad[:10])
for b in bad[:3]:
    print(y[b], No[b], Nr[b], zo[b], zr[b])
print(co, cr)
print('gpu==oracle', np.array_equal(N,No), np.array_equal(z,zo))
print(mdl['S'], mdl['s'])

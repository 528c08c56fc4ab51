import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
from cases import DensityCurrentCase, rel_l2
case = DensityCurrentCase(p=7, NeX=4, NeY=2, NeZ=3, perturb=2.0)
o = case.make_oracle()
d = case.make_driver(o)
Ne, Np = case.mesh.Ne, case.elem.Np
# tendency parity
o.piece('exchange'); o.piece('pressure'); o.piece('bc'); o.piece('tend_ex')
t = d.cal_tend_ex()
te = o.arr('tend_ex').reshape(5, -1)[:, :Ne*Np]  # oracle var order DENS, RHOT, MOMZ, MOMX, MOMY
for nm, iv in (('DENS_dt',0),('RHOT_dt',1),('MOMZ_dt',2),('MOMX_dt',3),('MOMY_dt',4)):
    print(nm, rel_l2(t[nm], te[iv]), np.abs(te[iv]).max())
# steps
for k, v in case.fields.items(): o.arr(k)[:] = v.reshape(-1)
o.prepare()
o.update(10); d.Update(10)
g = d.get_prog()
for nm in ('DDENS','MOMX','MOMY','MOMZ','DRHOT'):
    print(nm, rel_l2(g[nm][:Ne*Np], o.arr(nm)[:Ne*Np]))
print(d.monitor(), o.monitor())
print(d.last_timing())

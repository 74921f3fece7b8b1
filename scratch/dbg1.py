import sys, numpy as np, torch
sys.path.insert(0, '.')
def load_golden(name):
    z = np.load("tests/golden/" + name + ".npz"); return {k: z[k] for k in z.files}
from oracle import pywfo_oracle as orc
from pywfo_b200 import _lib, excitations as ex
ctx = _lib.context(0)
g = load_golden("r_n17"); occ = int(g["occ"])
bt = ex.build_tables(g["bra_ci"], 0.05); kt = ex.build_tables(g["ket_ci"], 0.05, order="locality")
dev = torch.device("cuda:0"); f64 = dict(dtype=torch.float64, device=dev)
bra, ket = torch.tensor(g["bra_mos"], **f64), torch.tensor(g["ket_mos"], **f64)
invh = np.linalg.inv(g["ket_mos"]); inv = torch.tensor(invh, **f64)
n_mo = bra.shape[0]
s_ao, s_mo = torch.empty((n_mo, n_mo), **f64), torch.empty((n_mo, n_mo), **f64)
st = torch.cuda.current_stream().cuda_stream
print("stream", st)
ctx.sao_dev(inv, s_ao, st); torch.cuda.synchronize()
print("sao err", np.abs(s_ao.cpu().numpy() - invh @ invh.T).max())
ctx.smo_dev(bra, s_ao, ket, s_mo, st); torch.cuda.synchronize()
S = orc.mo_overlap(g["bra_mos"], g["ket_mos"], invh @ invh.T)
print("smo err", np.abs(s_mo.cpu().numpy() - S).max())
be = torch.tensor(bt.exc, dtype=torch.int32, device=dev); ke = torch.tensor(kt.exc, dtype=torch.int32, device=dev)
cb, ck = torch.tensor(bt.coeff, **f64), torch.tensor(kt.coeff, **f64)
out = torch.zeros((cb.shape[0], ck.shape[0]), **f64)
want = ctx.state_overlaps_host(S, occ, bt.exc, bt.coeff, kt.exc, kt.coeff, tier="update")
print("want", want)
for tier in ("update", "lu"):
    out.zero_()
    ctx.state_overlaps_dev(s_mo, occ, be, cb, ke, ck, out, tier=tier, stream=st); torch.cuda.synchronize()
    print(tier, out.cpu().numpy())
s2 = torch.tensor(S, **f64)
ctx.state_overlaps_dev(s2, occ, be, cb, ke, ck, out, tier="update", stream=st); torch.cuda.synchronize()
print("with uploaded S", out.cpu().numpy())

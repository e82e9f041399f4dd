"""Where does the device RK45 separate from scipy + oracle?  python tools/dbg_rk45.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy import integrate
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.model import VFModel
from flowmse_b200.sampling import get_black_box_solver, _rk45_on_device
from oracle import ncsnpp_oracle as orc

sd = synthetic_state_dict(0)
model = VFModel(backbone="ncsnpp", ode="flowmatching"); model.dnn.load_state_dict(sd, strict=True); model.eval()
g = torch.Generator().manual_seed(111)
Y = torch.view_as_complex(0.3 * torch.randn(1, 1, 256, 64, 2, generator=g))
torch.manual_seed(4321); z = torch.randn_like(Y.cuda()).cpu()
x0 = orc.prior_sample(Y, z)
sd_cuda = {k: v.cuda() for k, v in sd.items()}
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def vf_oracle_cuda(x, t, y):
    with torch.device("cuda"):
        return orc.vf_forward(sd_cuda, x, t, y)
def sep(a, b):
    a, b = torch.view_as_real(a.cpu().to(torch.complex64)), torch.view_as_real(b.cpu().to(torch.complex64))
    d = (a - b).abs(); return float((d > 1e-4 + 1e-3 * b.abs()).float().mean()), float(d.max())
# single VF evaluations at a few t
for t in (1.0, 0.5, 0.031):
    tt = torch.full((1,), t)
    a = model(x0.cuda(), tt.cuda(), Y.cuda()); b = vf_oracle_cuda(x0.cuda(), tt.cuda(), Y.cuda()); c = orc.vf_forward(sd, x0, tt, Y)
    print(f"VF t={t}: model vs cuda-oracle {sep(a, b)}, cuda-oracle vs cpu-oracle {sep(b, c)}")
for tol in (1e-3,):
    s1, n1 = _rk45_on_device(model, x0.cuda(), Y.cuda(), 1.0, 0.03, tol, tol)
    s2, n2 = _rk45_on_device(vf_oracle_cuda, x0.cuda(), Y.cuda(), 1.0, 0.03, tol, tol)
    def ode_func(t, flat):
        xt = torch.from_numpy(flat.reshape(tuple(Y.shape))).type(torch.complex64).cuda()
        return vf_oracle_cuda(xt, torch.ones(1, device="cuda") * t, Y.cuda()).cpu().numpy().reshape(-1)
    sol = integrate.solve_ivp(ode_func, (1.0, 0.03), x0.numpy().reshape(-1), rtol=tol, atol=tol, method="RK45")
    s3 = torch.tensor(sol.y[:, -1]).reshape(Y.shape)
    print(f"tol={tol}: nfe device(model)={n1} device(oracle-cuda)={n2} scipy(oracle-cuda)={sol.nfev}")
    print("  device(model) vs device(oracle-cuda):", sep(s1, s2))
    print("  device(oracle-cuda) vs scipy(oracle-cuda):", sep(s2, s3))
    # sensitivity: scipy with x0 moved by 1 ulp
    x1 = torch.view_as_complex(torch.nextafter(torch.view_as_real(x0), torch.full((), float("inf"))))
    sol2 = integrate.solve_ivp(ode_func, (1.0, 0.03), x1.numpy().reshape(-1), rtol=tol, atol=tol, method="RK45")
    print("  scipy(oracle-cuda) vs itself with x0 + 1 ulp:", sep(torch.tensor(sol2.y[:, -1]).reshape(Y.shape), s3), "nfe", sol2.nfev)
